"""Differentiable forward of DeMFI-Net for the training step (`main.py:402`, `DeMFInet.forward(..., is_training=True)`,
`DeMFInet.py:46-179`) -- the reverse-mode counterpart of `engine.py`, first version.

The inference engine fuses, reuses buffers and never keeps what a backward pass needs; training wants the opposite.  This
module therefore re-expresses the forward as a torch autograd graph whose HEAVY nodes are this repository's kernels:

  ops.conv2d        every convolution          forward + dx on the tcgen05 kernel, dW / db on the wgrad kernel   (grad.py)
  ops.bwarp_blend   bwarp + Eq.(2), 3 sites    demfi_bwarp_blend / demfi_bwarp_blend_backward
  ops.cfr           complementary flow reversal demfi_cfr_splat + finalize / demfi_cfr_backward
  ops.fgac_sample   FGAC bilinear sampling     demfi_fgac_sample / demfi_fgac_sample_backward

and whose glue (channel concatenation and slicing, space-to-depth / pixel shuffle / nearest up-sampling, tanh / sigmoid, the
GRU gate arithmetic, Eq.(4)) is torch tensor arithmetic for now -- < 1 % of the FLOPs; fusing it into epilogues is the later
optimisation, exactly the path the inference engine took.  `ops` is a parameter so that the wiring can be checked on a
machine without a GPU against autograd through the unmodified reference (tests/test_train_net.py injects torch
implementations of the four operators); the default, `KernelOps`, has no CPU path and raises on CPU tensors.
"""
from __future__ import annotations

from typing import List

import torch

from . import _abi as A
from . import grad as G

NF = 64


# ---------------------------------------------------------------------------------------------- kernel-backed operators
def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


class _BwarpBlend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, flow, occ, t):
        if not a.is_cuda:
            raise A.DemfiError("demfi_b200.train_net.KernelOps runs on the GPU only (no CPU fallback)")
        N, C_, H, W = a.shape
        ld = G._ru(C_, 4)
        dev = a.device
        with torch.cuda.device(dev):
            ab, bb = G._to_nhwc(a.detach().float(), ld), G._to_nhwc(b.detach().float(), ld)
            fb, ob = G._to_nhwc(flow.detach().float(), 4), G._to_nhwc(occ.detach().float(), 4)
            tv = t.detach().reshape(-1).float().contiguous()
            out = torch.zeros(N, H, W, ld, dtype=torch.float32, device=dev)
            A.check(A.lib().demfi_bwarp_blend(ab.data_ptr(), ld, bb.data_ptr(), ld, fb.data_ptr(), 4, ob.data_ptr(), 4, tv.data_ptr(),
                                              N, H, W, C_, out.data_ptr(), ld, None, 0, _stream(dev)), "demfi_bwarp_blend")
            y = G._to_nchw(out, C_)
        ctx.bufs, ctx.shape = (ab, bb, fb, ob, tv), (N, C_, H, W, ld)
        return y

    @staticmethod
    def backward(ctx, gy):
        ab, bb, fb, ob, tv = ctx.bufs
        N, C_, H, W, ld = ctx.shape
        dev = gy.device
        with torch.cuda.device(dev):
            gb = G._to_nhwc(gy.detach().float(), ld)
            da, db = torch.zeros_like(ab), torch.zeros_like(bb)
            dfl = torch.empty(N, H, W, 4, dtype=torch.float32, device=dev)
            doc = torch.empty(N, H, W, 4, dtype=torch.float32, device=dev)
            A.check(A.lib().demfi_bwarp_blend_backward(ab.data_ptr(), ld, bb.data_ptr(), ld, fb.data_ptr(), 4, ob.data_ptr(), 4, tv.data_ptr(),
                                                       gb.data_ptr(), ld, N, H, W, C_, da.data_ptr(), ld, db.data_ptr(), ld,
                                                       dfl.data_ptr(), 4, doc.data_ptr(), 4, _stream(dev)), "demfi_bwarp_blend_backward")
            return G._to_nchw(da, C_), G._to_nchw(db, C_), G._to_nchw(dfl, 4), G._to_nchw(doc, 1), None


class _Cfr(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f01, f10, t):
        if not f01.is_cuda:
            raise A.DemfiError("demfi_b200.train_net.KernelOps runs on the GPU only (no CPU fallback)")
        N, _, H, W = f01.shape
        dev = f01.device
        with torch.cuda.device(dev):
            fo = G._to_nhwc(torch.cat([f01, f10], 1).detach().float(), 4)
            tv = t.detach().reshape(-1).float().contiguous()
            acc = torch.zeros(N, H, W, 8, dtype=torch.float32, device=dev)
            out = torch.empty(N, H, W, 4, dtype=torch.float32, device=dev)
            lib = A.lib()
            A.check(lib.demfi_cfr_splat(fo.data_ptr(), 4, tv.data_ptr(), N, H, W, acc.data_ptr(), _stream(dev)), "demfi_cfr_splat")
            A.check(lib.demfi_cfr_finalize(acc.data_ptr(), tv.data_ptr(), N, H, W, out.data_ptr(), 4, _stream(dev)), "demfi_cfr_finalize")
            y = G._to_nchw(out, 4)
        ctx.bufs, ctx.shape = (fo, tv, acc), (N, H, W)
        return y[:, 0:2], y[:, 2:4]

    @staticmethod
    def backward(ctx, g0, g1):
        fo, tv, acc = ctx.bufs
        N, H, W = ctx.shape
        dev = fo.device
        with torch.cuda.device(dev):
            gb = G._to_nhwc(torch.cat([g0, g1], 1).detach().float(), 4)
            gacc = torch.empty(N, H, W, 8, dtype=torch.float32, device=dev)
            dfo = torch.empty(N, H, W, 4, dtype=torch.float32, device=dev)
            A.check(A.lib().demfi_cfr_backward(fo.data_ptr(), 4, tv.data_ptr(), acc.data_ptr(), gb.data_ptr(), 4, N, H, W, gacc.data_ptr(),
                                               dfo.data_ptr(), 4, _stream(dev)), "demfi_cfr_backward")
            d = G._to_nchw(dfo, 4)
        return d[:, 0:2], d[:, 2:4], None


class _FgacSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, refk, flow):
        if not refk.is_cuda:
            raise A.DemfiError("demfi_b200.train_net.KernelOps runs on the GPU only (no CPU fallback)")
        N, C_, H, W = refk.shape
        dev = refk.device
        with torch.cuda.device(dev):
            rb, fb = G._to_nhwc(refk.detach().float(), C_), G._to_nhwc(flow.detach().float(), 4)
            out = torch.empty(N, H, W, C_, dtype=torch.float32, device=dev)
            A.check(A.lib().demfi_fgac_sample(rb.data_ptr(), C_, fb.data_ptr(), 4, N, H, W, C_, out.data_ptr(), C_, _stream(dev)), "demfi_fgac_sample")
            y = G._to_nchw(out, C_)
        ctx.bufs, ctx.shape = (rb, fb), (N, C_, H, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        rb, fb = ctx.bufs
        N, C_, H, W = ctx.shape
        dev = gy.device
        with torch.cuda.device(dev):
            gb = G._to_nhwc(gy.detach().float(), C_)
            dr = torch.zeros_like(rb)
            dfl = torch.empty(N, H, W, 2, dtype=torch.float32, device=dev)
            A.check(A.lib().demfi_fgac_sample_backward(rb.data_ptr(), C_, fb.data_ptr(), 4, gb.data_ptr(), C_, N, H, W, C_, dr.data_ptr(), C_,
                                                       dfl.data_ptr(), 2, _stream(dev)), "demfi_fgac_sample_backward")
            return G._to_nchw(dr, C_), G._to_nchw(dfl, 2)


class KernelOps:
    """The four operator families on this repository's sm_100a kernels (no CPU path)."""

    @staticmethod
    def conv2d(x, w, b, act="none", stride=1):
        return G.conv2d(x, w, b, act, stride)

    @staticmethod
    def bwarp_blend(a, b, flow, occ_logit, t):
        return _BwarpBlend.apply(a, b, flow, occ_logit, t)

    @staticmethod
    def cfr(flow_01, flow_10, t):
        return _Cfr.apply(flow_01, flow_10, t)

    @staticmethod
    def fgac_sample(refk, flow):
        return _FgacSample.apply(refk, flow)


# ---------------------------------------------------------------------------------------------- the graph
def _w(conv):
    w = conv.weight
    return w.squeeze(2) if w.dim() == 5 else w        # Conv3d [Co,Ci,1,k,k] = the same 2-D kernel on every frame


def _space_to_depth(x, r=2):                            # pixel_reshuffle, DeMFInet.py:290-316: channel = c*r*r + dy*r + dx
    b, c, h, w = x.shape
    return x.reshape(b, c, h // r, r, w // r, r).permute(0, 1, 3, 5, 2, 4).reshape(b, c * r * r, h // r, w // r)


def _up2(z):
    return z.repeat_interleave(2, 2).repeat_interleave(2, 3)   # nn.UpsamplingNearest2d(scale_factor=2)


def _resblocks(ops, x, blocks):
    for blk in blocks:
        x = x + ops.conv2d(ops.conv2d(x, _w(blk.conv1), blk.conv1.bias, "relu"), _w(blk.conv2), blk.conv2.bias, "none")
    return x


def forward_train(model, x: torch.Tensor, t_value: torch.Tensor, num_update: int, ops=KernelOps):
    """The training 7-tuple of `DeMFInet.forward` (`DeMFInet.py:170-172`) as a differentiable function of `model`'s parameters.
    x: [B,3,4,H,W] in [-1,1] (frames B0, B1, B-1, B2); t_value: [B,1]."""
    cv = lambda m, z, act="none", stride=1: ops.conv2d(z, _w(m), m.bias, act, stride)
    B = x.shape[0]
    B0, B1, Bm1, B2 = x[:, :, 0], x[:, :, 1], x[:, :, 2], x[:, :, 3]
    frames12 = torch.cat((B0, B1, Bm1, B2), 1)
    t = t_value.reshape(B)

    # ---- FF_RDB (DeMFInet.py:233-287)
    ff = model.FF_RDB_Module
    f1 = cv(ff.SFENet1, _space_to_depth(frames12, 2))
    z = cv(ff.SFENet2, f1)
    rdb_outs = []
    for rdb in ff.RDBs:
        d = z
        for c in rdb.convs:
            d = torch.cat((d, cv(c.conv[0], d, "relu")), 1)
        z = cv(rdb.LFF, d) + z
        rdb_outs.append(z)
    z = cv(ff.GFF[1], cv(ff.GFF[0], torch.cat(rdb_outs, 1))) + f1
    s = cv(ff.UPNet[2], torch.nn.functional.pixel_shuffle(cv(ff.UPNet[0], z), 2))
    F0, F1 = torch.tanh(s[:, :NF]), torch.tanh(s[:, NF:2 * NF])
    flow_01, flow_10, occ_logit = s[:, 2 * NF:2 * NF + 2], s[:, 2 * NF + 2:2 * NF + 4], s[:, 2 * NF + 4:2 * NF + 5]

    # ---- CFR + Eq.(2) on the features (DeMFInet.py:60-71)
    flow_t0, flow_t1 = ops.cfr(flow_01, flow_10, t)
    Ft = ops.bwarp_blend(F0, F1, torch.cat((flow_t0, flow_t1), 1), occ_logit, t)

    # ---- FAC-FB with the shared FGAC (DeMFInet.py:335-358, 386-452)
    fb = model.FAC_FB_Module
    e = _resblocks(ops, cv(fb.conv_first, torch.cat([F0, F1], 0), "relu"), fb.feature_extraction)
    e0, e1 = e[:B], e[B:]
    fg = fb.shared_FGAC

    def fgac(ref, src, flow_s2r):
        e_s = cv(fg.fusion, ops.fgac_sample(cv(fg.conv_ref_k, ref), flow_s2r))
        w_sr = cv(fg.w_gen_2, cv(fg.w_gen, torch.cat([src, e_s], 1), "relu"), "sigmoid")
        res = w_sr * src + (1 - w_sr) * e_s                                  # Eq.(4)
        with torch.no_grad():                                                # difference map, DeMFInet.py:456-462
            dm = torch.mean(torch.abs(res - src), 1, keepdim=True).reshape(B, -1)
            dm = dm - dm.min(1, keepdim=True)[0]
            dm = (dm / dm.max(1, keepdim=True)[0]).reshape(B, 1, *res.shape[2:])
        return res, dm

    aF0, d10 = fgac(e1, e0, flow_01)
    aF1, d01 = fgac(e0, e1, flow_10)

    # ---- UNet refinement (DeMFInet.py:76-93, 586-603)
    rm = model.Refine_Module
    agg1 = torch.cat([aF0, aF1, Ft, flow_t0, flow_t1, flow_01, flow_10, occ_logit], 1)
    c1 = cv(rm.enc1, agg1, "relu", 2)
    c2 = cv(rm.enc2, c1, "relu", 2)
    o = cv(rm.dec0, cv(rm.enc3, c2, "relu", 2), "relu")
    o = cv(rm.dec1, torch.cat((_up2(o), c2), 1), "relu")
    o = cv(rm.dec2, torch.cat((_up2(o), c1), 1), "relu")
    agg1 = cv(rm.dec3, _up2(o)) + torch.cat([flow_t0, flow_t1, occ_logit, aF0, aF1], 1)
    rflow_t0, rflow_t1, occ_logit_r = agg1[:, 0:2], agg1[:, 2:4], agg1[:, 4:5]
    occ_0 = torch.sigmoid(occ_logit_r)
    rF0, rF1 = torch.tanh(agg1[:, 5:5 + NF]), torch.tanh(agg1[:, 5 + NF:5 + 2 * NF])
    flow_init = torch.cat((rflow_t0, rflow_t1), 1)
    rFt = ops.bwarp_blend(rF0, rF1, flow_init, occ_logit_r, t)

    # ---- D1 on the three frames (DeMFInet.py:95-111)
    d = _resblocks(ops, cv(model.Dec_first, torch.cat([rF0, rF1, rFt], 0), "relu"), model.Decoder_res)
    d = cv(model.Dec_last2, cv(model.Dec_last1, d, "relu"))
    S0p, S1p, Stp = d[:B], d[B:2 * B], d[2 * B:]

    # ---- recursive boosting (DeMFInet.py:113-165)
    f_rec = torch.tanh(cv(model.Ch_Reducer, torch.cat((rF0, rF1, rFt), 1)))
    ref30 = torch.cat((S0p, S1p, Stp, B0, B1, Bm1, B2, flow_10, flow_01, flow_init, occ_logit_r), 1)
    bm = model.Booster_Module
    mx, gb, fo = bm.Mixer, bm.GB, bm.flow_occ
    flows, occs, sharps_final = [flow_init], [occ_0], []
    dflow, docc = flow_init, occ_logit_r
    for _ in range(num_update):
        r = cv(mx.conv_ref2, cv(mx.conv_ref1, ref30, "relu"), "relu")
        dl = cv(mx.conv_delta2, cv(mx.conv_delta1, torch.cat([dflow, docc], 1), "relu"), "relu")
        xm = cv(mx.conv_blend2, cv(mx.conv_blend1, torch.cat([r, dl], 1), "relu"), "relu")
        h = f_rec
        for cz, cr, cq in ((gb.convz1, gb.convr1, gb.convq1), (gb.convz2, gb.convr2, gb.convq2)):   # SepConvGRU, :838-857
            hx = torch.cat([h, xm], 1)
            zg, rg = cv(cz, hx, "sigmoid"), cv(cr, hx, "sigmoid")
            q = cv(cq, torch.cat([rg * h, xm], 1), "tanh")
            h = (1 - zg) * h + zg * q
        f_rec = h
        dfo = cv(fo.conv2, cv(fo.conv1, h, "relu"))
        dflow = dflow + dfo[:, :4]
        docc = docc + dfo[:, 4:5]
        occ_f = torch.sigmoid(docc)
        occs.append(occ_f)
        flows.append(dflow)
        st_new = ops.bwarp_blend(S0p, S1p, dflow, docc, t)
        agg3 = torch.cat([S0p, S1p, st_new, f_rec, occ_0, rflow_t0, rflow_t1, flow_10, flow_01, dflow[:, :2], dflow[:, 2:4], occ_f,
                          B0, B1, Bm1, B2], 1)
        o = _resblocks(ops, cv(model.Dec_first_2, agg3, "relu"), model.Decoder_res_2)
        o = cv(model.Dec_last2_2, cv(model.Dec_last1_2, o, "relu"))
        sharps_final.append([o[:, 0:3] + S0p, o[:, 3:6] + S1p, o[:, 6:9] + st_new])
    two_blurry = torch.mean(x[:, :, 0:2], dim=2)
    return ([S0p, S1p, Stp], sharps_final, flows, occs, two_blurry, [d10, d01, d10, d01], [[rflow_t0, rflow_t1]])
