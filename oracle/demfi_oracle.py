"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU (torch fp32) restatement of the DeMFI-Net forward / recursive-boosting path,
`/root/reference/DeMFInet.py:46-179`, written as pure functions over a state_dict.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this module, and only as the checker / CPU baseline.

Pinning: `oracle/gen_golden.py` runs the UNMODIFIED reference module (imported from
/root/reference in the build container) on seeded inputs/weights and commits its
outputs under `tests/golden/`; `tests/test_oracle.py` checks this restatement
against those vectors (CPU, no GPU needed).  The reference itself ships no golden
vectors, tests or checkpoint: trained-weight parity is
UNPINNED, and results are pinned to the reference source as executed by torch
2.11 CPU (the reference pins torch 1.7.1; `README.md:63-64` warns about
`align_corners` semantics -- `align_corners=True` is restated explicitly below).

Third-party arithmetic: convolutions use `torch.nn.functional.conv2d` (the library
op the reference's `nn.Conv2d/Conv3d` call); bilinear sampling and the Gaussian
splat are restated in closed form (no `grid_sample`, no `put_`) so that the axis
conventions the CUDA kernels follow are written down once:
  * flow channel 0 = x / columns, channel 1 = y / rows (`DeMFInet.py:744-754`);
  * in `fwarp` the variable called `x` is the ROW index (`DeMFInet.py:647-648,712-719`).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

NF = 64


# ----------------------------------------------------------------------------- primitives
def _conv(x, sd, name, stride=1, pad=None):
    """nn.Conv2d / nn.Conv3d([1,k,k]) application.  A Conv3d with a [1,kh,kw] kernel over
    [B,C,T,H,W] is a per-frame Conv2d (`DeMFInet.py:30-34,532-533`); callers batch frames."""
    w = sd[name + ".weight"]
    if w.dim() == 5:
        w = w[:, :, 0]
    if pad is None:
        pad = (w.shape[2] // 2, w.shape[3] // 2)
    return F.conv2d(x, w, sd[name + ".bias"], stride=stride, padding=pad)


def space_to_depth(x, r=2):
    """`pixel_reshuffle`, `DeMFInet.py:290-316`: out channel = c*r*r + dy*r + dx."""
    b, c, h, w = x.shape
    x = x.reshape(b, c, h // r, r, w // r, r).permute(0, 1, 3, 5, 2, 4)
    return x.reshape(b, c * r * r, h // r, w // r)


def bilinear_gather(img, px, py):
    """Bilinear sample of img[B,C,H,W] at absolute pixel coords (px = column, py = row),
    zero outside the image -- the arithmetic of `F.grid_sample(..., mode='bilinear',
    padding_mode='zeros', align_corners=True)` after un-normalisation.
    Returns (sample[B,C,Ho,Wo], inbounds_weight_sum[B,1,Ho,Wo])."""
    B, C, H, W = img.shape
    x0 = torch.floor(px)
    y0 = torch.floor(py)
    wx1 = px - x0
    wy1 = py - y0
    wx0 = 1.0 - wx1
    wy0 = 1.0 - wy1
    flat = img.reshape(B, C, H * W)
    out = torch.zeros((B, C) + tuple(px.shape[1:]), dtype=img.dtype, device=img.device)
    wsum = torch.zeros((B, 1) + tuple(px.shape[1:]), dtype=img.dtype, device=img.device)
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xi = x0 + dx
            yi = y0 + dy
            ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
            wgt = (wx * wy) * ok.to(img.dtype)
            idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).long()
            g = torch.gather(flat, 2, idx.reshape(B, 1, -1).expand(B, C, -1)).reshape(out.shape)
            out = out + g * wgt.unsqueeze(1)
            wsum = wsum + wgt.unsqueeze(1)
    return out, wsum


def _unnormalised_grid(flo):
    """The coordinate `bwarp` hands to grid_sample, including its fp32 normalise ->
    un-normalise round trip (`DeMFInet.py:750-757` then align_corners=True)."""
    B, _, H, W = flo.shape
    xx = torch.arange(W, dtype=torch.float32, device=flo.device).view(1, 1, W).expand(B, H, W)
    yy = torch.arange(H, dtype=torch.float32, device=flo.device).view(1, H, 1).expand(B, H, W)
    gx = 2.0 * (xx + flo[:, 0]) / max(W - 1, 1) - 1.0
    gy = 2.0 * (yy + flo[:, 1]) / max(H - 1, 1) - 1.0
    px = ((gx + 1.0) / 2.0) * (W - 1)
    py = ((gy + 1.0) / 2.0) * (H - 1)
    return px, py


def bwarp(x, flo):
    """`bwarp`, `DeMFInet.py:732-766`: backward bilinear warp, then zero every pixel whose
    in-bounds bilinear weight sum (the warped all-ones image) is < 0.999."""
    px, py = _unnormalised_grid(flo)
    out, wsum = bilinear_gather(x, px, py)
    return out * (wsum >= 0.999).to(x.dtype)


def eq2_blend(a, fa, b, fb, occ_logit, t):
    """Eq.(2), `DeMFInet.py:66-71, 90-93, 146-149`.  t is [B,1,1,1]."""
    o0 = torch.sigmoid(occ_logit)
    o1 = 1 - o0
    num = (1 - t) * o0 * bwarp(a, fa) + t * o1 * bwarp(b, fb)
    return num / ((1 - t) * o0 + t * o1)


def gaussian_splat(img, flo):
    """`fwarp` + `get_gaussian_weights` + `sample_one`, `DeMFInet.py:625-729`.
    Source pixel (r, c) with displacement (dx = flo ch0 along columns, dy = flo ch1 along
    rows) adds img*w and w to the four targets (r + floor(dy) + {0,1}, c + floor(dx) + {0,1})
    with w = exp(-((dy - cy)^2 + (dx - cx)^2)), (cy, cx) the integer corner offsets; targets
    outside the image are dropped.  `.long()` on the already-floored value is exact."""
    B, C, H, W = img.shape
    dx = flo[:, 0:1]
    dy = flo[:, 1:2]
    fy = torch.floor(dy)
    fx = torch.floor(dx)
    rows = torch.arange(H, device=img.device).view(1, 1, H, 1)
    cols = torch.arange(W, device=img.device).view(1, 1, 1, W)
    acc = torch.zeros(B, C, H * W, dtype=img.dtype, device=img.device)
    nrm = torch.zeros(B, 1, H * W, dtype=img.dtype, device=img.device)
    for oy in (0.0, 1.0):
        for ox in (0.0, 1.0):
            cy = fy + oy
            cx = fx + ox
            wgt = torch.exp(-((dy - cy) ** 2 + (dx - cx) ** 2))
            tr = cy.long() + rows
            tc = cx.long() + cols
            ok = (tr >= 0) & (tr < H) & (tc >= 0) & (tc < W)
            wgt = wgt * ok.to(img.dtype)
            idx = (tr.clamp(0, H - 1) * W + tc.clamp(0, W - 1)).reshape(B, 1, H * W)
            acc.scatter_add_(2, idx.expand(B, C, -1), (img * wgt).reshape(B, C, H * W))
            nrm.scatter_add_(2, idx, wgt.reshape(B, 1, H * W))
    return acc.reshape(B, C, H, W), nrm.reshape(B, 1, H, W)


def cfr_flow_t_align(flow_01, flow_10, t):
    """Complementary flow reversal, `DeMFInet.py:606-622`.  The reference replicates the
    weight map per channel (`:650-651`); one shared map is the same numbers."""
    a01, n0 = gaussian_splat(flow_01, t * flow_01)
    a10, n1 = gaussian_splat(flow_10, (1 - t) * flow_10)
    ft0 = -(1 - t) * t * a01 + t * t * a10
    ft1 = (1 - t) * (1 - t) * a01 - t * (1 - t) * a10
    norm = (1 - t) * n0 + t * n1
    m = (norm > 0).to(norm.dtype)
    ft0 = (1 - m) * ft0 + m * (ft0 / (norm + (1 - m)))
    ft1 = (1 - m) * ft1 + m * (ft1 / (norm + (1 - m)))
    return ft0, ft1


# ----------------------------------------------------------------------------- modules
def ff_rdb(sd, frames12, out):
    """`FF_RDB.forward`, `DeMFInet.py:233-253` (+ RDB `:256-287`)."""
    p = "FF_RDB_Module."
    x = space_to_depth(frames12, 2)
    f1 = _conv(x, sd, p + "SFENet1")
    x = _conv(f1, sd, p + "SFENet2")
    rdb_outs = []
    for i in range(12):
        d = x
        for c in range(4):
            d = torch.cat((d, F.relu(_conv(d, sd, f"{p}RDBs.{i}.convs.{c}.conv.0"))), 1)
        x = _conv(d, sd, f"{p}RDBs.{i}.LFF") + x
        rdb_outs.append(x)
    x = _conv(_conv(torch.cat(rdb_outs, 1), sd, p + "GFF.0"), sd, p + "GFF.1") + f1
    out["ff_trunk"] = x
    s = _conv(F.pixel_shuffle(_conv(x, sd, p + "UPNet.0"), 2), sd, p + "UPNet.2")
    ff = torch.tanh(s[:, :2 * NF])
    return ff[:, :NF], ff[:, NF:], s[:, 2 * NF:2 * NF + 2], s[:, 2 * NF + 2:2 * NF + 4], s[:, 2 * NF + 4:2 * NF + 5]


def _resblocks(x, sd, prefix, n=5):
    """`ResidualBlock_noBN(_3D)`, `DeMFInet.py:524-563`."""
    for i in range(n):
        x = x + _conv(F.relu(_conv(x, sd, f"{prefix}.{i}.conv1")), sd, f"{prefix}.{i}.conv2")
    return x


def _minmax_map(x):
    """channel mean of |x|, min-max normalised per sample (`DeMFInet.py:465-491`, the same five lines four times)"""
    m = torch.mean(torch.abs(x), 1, keepdim=True)
    b = m.shape[0]
    f = m.reshape(b, -1)
    f = f - f.min(1, keepdim=True)[0]
    f = f / f.max(1, keepdim=True)[0]
    return f.reshape(m.shape)


def fgac(sd, ref, src, flow_s2r, out=None, tag="", visualization=False):
    """`FGAC.forward` with rr = sr = 0, `DeMFInet.py:386-452`: the key conv of `ref` is
    bilinearly sampled at ABSOLUTE position (x, y) = flow value (no base grid is added,
    `:413-419`, `bilinear_sampler` `:499-514`); the correlation softmax runs over one
    element and is identically 1, so `conv_source_k` never reaches the output."""
    p = "FAC_FB_Module.shared_FGAC."
    B, C, H, W = ref.shape
    ref_k = _conv(ref, sd, p + "conv_ref_k")
    gx = 2 * flow_s2r[:, 0] / (W - 1) - 1
    gy = 2 * flow_s2r[:, 1] / (H - 1) - 1
    px = ((gx + 1.0) / 2.0) * (W - 1)
    py = ((gy + 1.0) / 2.0) * (H - 1)
    sampled, _ = bilinear_gather(ref_k, px, py)
    e_s = _conv(sampled, sd, p + "fusion")
    w_sr = torch.sigmoid(_conv(F.relu(_conv(torch.cat([src, e_s], 1), sd, p + "w_gen")), sd, p + "w_gen_2"))
    res = w_sr * src + (1 - w_sr) * e_s
    diff = torch.mean(torch.abs(res - src), 1, keepdim=True)
    b = diff.shape[0]
    d = diff.reshape(b, -1)
    d = d - d.min(1, keepdim=True)[0]
    d = d / d.max(1, keepdim=True)[0]
    if out is not None:
        out["fgac_sampled" + tag] = sampled
        out["fgac_w" + tag] = w_sr
    if visualization:  # `DeMFInet.py:464-493`: [w, 1-w, source, key conv of ref (before sampling), E_s, result]
        return res, [w_sr, 1 - w_sr, _minmax_map(src), _minmax_map(ref_k), _minmax_map(e_s), _minmax_map(res)], d.reshape(diff.shape)
    return res, w_sr, d.reshape(diff.shape)


def fac_fb(sd, F0, F1, flow_10, flow_01, out, visualization=False):
    """`FAC_FB.forward`, `DeMFInet.py:335-358` (shared FGAC)."""
    p = "FAC_FB_Module."
    B = F0.shape[0]
    x = torch.cat([F0, F1], 0)  # frames batched: index f*B + b
    e = _resblocks(F.relu(_conv(x, sd, p + "conv_first")), sd, p + "feature_extraction")
    e0, e1 = e[:B], e[B:]
    out["enc0"], out["enc1"] = e0, e1
    a0, w0, d10 = fgac(sd, e1, e0, flow_01, out, "0", visualization)
    a1, w1, d01 = fgac(sd, e0, e1, flow_10, out, "1", visualization)
    return a0, a1, [w0, w1, w0, w1], [d10, d01, d10, d01]


def unet(sd, x):
    """`UNet.forward`, `DeMFInet.py:586-603`."""
    p = "Refine_Module."
    e1 = F.relu(_conv(x, sd, p + "enc1", 2, 1))
    e2 = F.relu(_conv(e1, sd, p + "enc2", 2, 1))
    o = F.relu(_conv(e2, sd, p + "enc3", 2, 1))
    o = F.relu(_conv(o, sd, p + "dec0"))
    up = lambda z: z.repeat_interleave(2, 2).repeat_interleave(2, 3)  # nearest x2
    o = F.relu(_conv(torch.cat((up(o), e2), 1), sd, p + "dec1"))
    o = F.relu(_conv(torch.cat((up(o), e1), 1), sd, p + "dec2"))
    return _conv(up(o), sd, p + "dec3")


def booster(sd, f_rec, ref30, delta5, out=None, tag=""):
    """`Booster.forward` = Mixer + SepConvGRU + FlowOcc, `DeMFInet.py:779-868`."""
    p = "Booster_Module."
    r = F.relu(_conv(F.relu(_conv(ref30, sd, p + "Mixer.conv_ref1")), sd, p + "Mixer.conv_ref2"))
    d = F.relu(_conv(F.relu(_conv(delta5, sd, p + "Mixer.conv_delta1")), sd, p + "Mixer.conv_delta2"))
    x = F.relu(_conv(F.relu(_conv(torch.cat([r, d], 1), sd, p + "Mixer.conv_blend1")), sd, p + "Mixer.conv_blend2"))
    h = f_rec
    for s in ("1", "2"):  # horizontal (1x5) then vertical (5x1)
        hx = torch.cat([h, x], 1)
        z = torch.sigmoid(_conv(hx, sd, p + "GB.convz" + s))
        rr = torch.sigmoid(_conv(hx, sd, p + "GB.convr" + s))
        q = torch.tanh(_conv(torch.cat([rr * h, x], 1), sd, p + "GB.convq" + s))
        h = (1 - z) * h + z * q
    dfo = _conv(F.relu(_conv(h, sd, p + "flow_occ.conv1")), sd, p + "flow_occ.conv2")
    if out is not None:
        out["blend_enc" + tag] = x
    return h, dfo[:, :4], dfo[:, 4:5]


# ----------------------------------------------------------------------------- forward
@torch.no_grad()
def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, t_value: torch.Tensor,
            num_update: Optional[int] = None, is_training=None, intermediates: Optional[dict] = None,
            visualization_flag: bool = False):
    """`DeMFInet.forward`, `DeMFInet.py:46-179`.  Returns the eval 5-tuple, the training 7-tuple when is_training, or the
    visualisation 7-tuple (`args.visualization_flag`, eval only).  `intermediates`, when a dict, receives named tensors (NCHW)."""
    out = intermediates if intermediates is not None else {}
    B = x.shape[0]
    B0, B1, Bm1, B2 = x[:, :, 0], x[:, :, 1], x[:, :, 2], x[:, :, 3]
    frames12 = torch.cat((B0, B1, Bm1, B2), 1)
    F0, F1, flow_01, flow_10, occ_logit = ff_rdb(sd, frames12, out)
    out.update(F0=F0, F1=F1, flow_01=flow_01, flow_10=flow_10, occ_logit_ff=occ_logit)

    t = t_value.reshape(B, 1, 1, 1)
    flow_t0, flow_t1 = cfr_flow_t_align(flow_01, flow_10, t)
    out.update(flow_t0=flow_t0, flow_t1=flow_t1)
    Ft = eq2_blend(F0, flow_t0, F1, flow_t1, occ_logit, t)
    out["Ft"] = Ft

    aF0, aF1, blending_weights, difference_maps = fac_fb(sd, F0, F1, flow_10, flow_01, out, visualization_flag)
    out.update(aF0=aF0, aF1=aF1)

    agg1 = torch.cat([aF0, aF1, Ft, flow_t0, flow_t1, flow_01, flow_10, occ_logit], 1)
    agg1 = unet(sd, agg1) + torch.cat([flow_t0, flow_t1, occ_logit, aF0, aF1], 1)
    rflow_t0, rflow_t1, occ_logit = agg1[:, 0:2], agg1[:, 2:4], agg1[:, 4:5]
    occ_0 = torch.sigmoid(occ_logit)
    rF0 = torch.tanh(agg1[:, 5:5 + NF])
    rF1 = torch.tanh(agg1[:, 5 + NF:5 + 2 * NF])
    rFt = eq2_blend(rF0, rflow_t0, rF1, rflow_t1, occ_logit, t)
    out.update(rflow_t0=rflow_t0, rflow_t1=rflow_t1, occ_logit_ref=occ_logit, rF0=rF0, rF1=rF1, rFt=rFt)

    d = torch.cat([rF0, rF1, rFt], 0)  # D1 runs per frame (Conv3d [1,3,3])
    d = _resblocks(F.relu(_conv(d, sd, "Dec_first")), sd, "Decoder_res")
    d = _conv(F.relu(_conv(d, sd, "Dec_last1")), sd, "Dec_last2")
    S0p, S1p, Stp = d[:B], d[B:2 * B], d[2 * B:]
    sharps_dec1 = [S0p, S1p, Stp]

    flow_init = torch.cat((rflow_t0, rflow_t1), 1)
    flow_predictions = [flow_init]
    occ0_predictions = [occ_0]
    f_rec = torch.tanh(_conv(torch.cat((rF0, rF1, rFt), 1), sd, "Ch_Reducer"))
    out["F_rec0"] = f_rec
    ref30 = torch.cat((S0p, S1p, Stp, B0, B1, Bm1, B2, flow_10, flow_01, flow_init, occ_logit), 1)
    dflow, docc = flow_init, occ_logit
    sharps_final: List[list] = []
    if num_update is None:
        num_update = 1
    for itr in range(num_update):
        f_rec, delta_flow, delta_occ = booster(sd, f_rec, ref30, torch.cat([dflow, docc], 1), out, str(itr))
        dflow = dflow + delta_flow
        docc = docc + delta_occ
        ft0f, ft1f = dflow[:, :2], dflow[:, 2:4]
        occ_f = torch.sigmoid(docc)
        occ0_predictions.append(occ_f)
        flow_predictions.append(torch.cat((ft0f, ft1f), 1))
        st_new = eq2_blend(S0p, ft0f, S1p, ft1f, docc, t)
        agg3 = torch.cat([S0p, S1p, st_new, f_rec, occ_0, rflow_t0, rflow_t1, flow_10, flow_01,
                          ft0f, ft1f, occ_f, B0, B1, Bm1, B2], 1)
        o = _resblocks(F.relu(_conv(agg3, sd, "Dec_first_2")), sd, "Decoder_res_2")
        o = _conv(F.relu(_conv(o, sd, "Dec_last1_2")), sd, "Dec_last2_2")
        sharps_final.append([o[:, 0:3] + S0p, o[:, 3:6] + S1p, o[:, 6:9] + st_new])
        out[f"F_rec{itr + 1}"] = f_rec
        out[f"St_new{itr}"] = st_new
    two_blurry = torch.mean(x[:, :, 0:2], dim=2)
    if is_training:
        return (sharps_dec1, sharps_final, flow_predictions, occ0_predictions, two_blurry,
                difference_maps, [[rflow_t0, rflow_t1]])
    if visualization_flag:  # `DeMFInet.py:167-168, 174-176`
        return (sharps_dec1, sharps_final, flow_predictions, occ0_predictions, two_blurry,
                blending_weights + [[flow_01, flow_10]], difference_maps)
    return sharps_dec1, sharps_final, flow_predictions, occ0_predictions, two_blurry


def flatten_outputs(res) -> Dict[str, torch.Tensor]:
    """Name the tensors of the eval 5-tuple (same names as tests/golden/*.npz)."""
    s1, sf, fl, oc, tb = res[:5]
    d = {"S0p": s1[0], "S1p": s1[1], "Stp": s1[2], "two_blurry": tb}
    for i, tri in enumerate(sf):
        d[f"S0_final{i}"], d[f"S1_final{i}"], d[f"St_final{i}"] = tri
    for i, f in enumerate(fl):
        d[f"flow{i}"] = f
    for i, o in enumerate(oc):
        d[f"occ{i}"] = o
    return d
