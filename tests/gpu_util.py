"""Helpers for the -m gpu parity tests: NHWC staging and direct C-ABI calls."""
import ctypes as C

import numpy as np
import torch

from demfi_b200 import _abi as A

DEV = torch.device("cuda:0")
CONV_TC16_H3 = 102  # test-only pseudo kind: DEMFI_CONV_TC16 forced onto the second-generation kernel (tc_gen = 2)


def nhwc(t_nchw, ld=None):
    """NCHW cpu tensor -> flat NHWC cuda buffer with pixel stride ld (zero padded)"""
    n, c, h, w = t_nchw.shape
    ld = ld or ((c + 3) // 4 * 4)
    buf = torch.zeros(n, h, w, ld, dtype=torch.float32)
    buf[..., :c] = t_nchw.permute(0, 2, 3, 1)
    return buf.to(DEV).contiguous(), ld


def s16_encode(buf):
    """fp32 NHWC buffer [..., C] (C % 32 == 0) -> the S16 storage format (include/demfi_b200.h): per 32-channel group
    32 fp16 hi then 32 fp16 lo, hi = fp16(v), lo = fp16((v - hi) * 2048); returned as a float32-typed buffer of the same shape"""
    t = buf.float().cpu()
    assert t.shape[-1] % 32 == 0
    g = t.reshape(*t.shape[:-1], t.shape[-1] // 32, 32)
    hi = g.half()
    lo = ((g - hi.float()) * 2048.0).half()
    packed = torch.cat([hi, lo], dim=-1).contiguous()            # [..., G, 64] fp16
    return packed.view(torch.float32).reshape(t.shape).to(buf.device).contiguous()


def s16_decode(buf):
    t = buf.cpu().contiguous()
    g = t.reshape(*t.shape[:-1], t.shape[-1] // 32, 32).view(torch.float16)  # [..., G, 64]
    val = g[..., :32].float() + g[..., 32:].float() / 2048.0
    return val.reshape(t.shape).to(buf.device)


def from_nhwc(buf, c):
    return buf[..., :c].permute(0, 3, 1, 2).contiguous().cpu()


def stream():
    return torch.cuda.current_stream(DEV).cuda_stream


def run_conv(w, b, srcs, out_hw, kind, segs_spec, stride=1, pad=None, in_map=None, out_map=None, cout_pad=None):
    """Direct demfi_conv2d call.  srcs: [(nhwc cuda buffer [N,Hs,Ws,ld], C, up)].
    segs_spec: [dict(ch0, nch, dst=buffer [N,Ho',Wo',ld], act, res=buffer, res2=buffer, store)].
    Returns nothing (dst buffers are written)."""
    lib = A.lib()
    A.set_option("tc_gen", 2 if kind == CONV_TC16_H3 else 3)
    if kind == CONV_TC16_H3:
        kind = A.CONV_TC16
    Co, Ci, KH, KW = w.shape
    if pad is None:
        pad = (KH // 2, KW // 2)
    src_C = [s_[1] for s_ in srcs]
    kt = sum(src_C)
    cout_pad = cout_pad or (Co + 15) // 16 * 16
    in_map = in_map if in_map is not None else list(range(Ci)) + [-1] * (kt - Ci)
    out_map = out_map if out_map is not None else list(range(Co)) + [-1] * (cout_pad - Co)
    sC = (A.i32 * len(src_C))(*src_C)
    nfl = lib.demfi_packed_weight_floats(kind, KH, KW, sC, len(src_C), cout_pad)
    packed = np.empty(nfl, dtype=np.float32)
    wn = np.ascontiguousarray(w.numpy())
    A.check(lib.demfi_pack_weights(kind, wn.ctypes.data, Co, Ci, KH, KW, (A.i32 * kt)(*in_map), sC, len(src_C),
                                   (A.i32 * cout_pad)(*out_map), cout_pad, packed.ctypes.data), "pack")
    wdev = torch.from_numpy(packed).to(DEV)
    bias = np.zeros(cout_pad, dtype=np.float32)
    for n, m in enumerate(out_map):
        if m >= 0:
            bias[n] = float(b[m])
    bdev = torch.from_numpy(bias).to(DEV)
    d = A.Conv()
    N = srcs[0][0].shape[0]
    Ho, Wo = out_hw
    d.N, d.H, d.W = N, Ho, Wo
    up0 = srcs[0][2]
    d.Hi, d.Wi = srcs[0][0].shape[1] << up0, srcs[0][0].shape[2] << up0
    d.KH, d.KW, d.stride, d.pad_h, d.pad_w = KH, KW, stride, pad[0], pad[1]
    d.nsrc, d.nseg, d.cout_pad, d.kind = len(srcs), len(segs_spec), cout_pad, kind
    for i, src in enumerate(srcs):
        buf, c, up = src[:3]
        d.src[i].ptr, d.src[i].C, d.src[i].ld, d.src[i].up = buf.data_ptr(), c, buf.shape[3], up
        d.src[i].fmt = src[3] if len(src) > 3 else 0
    for i, sg in enumerate(segs_spec):
        s = d.seg[i]
        s.dst, s.dst_ld = sg["dst"].data_ptr() + 4 * sg.get("dst_c0", 0), sg["dst"].shape[3]
        s.ch0, s.nch, s.act, s.store = sg["ch0"], sg["nch"], sg.get("act", 0), sg.get("store", 0)
        s.fmt = sg.get("fmt", 0)
        if sg.get("res") is not None:
            s.res, s.res_ld = sg["res"].data_ptr() + 4 * sg.get("res_c0", 0), sg["res"].shape[3]
        if sg.get("res2") is not None:
            s.res2, s.res2_ld = sg["res2"].data_ptr(), sg["res2"].shape[3]
    d.wpack, d.bias = wdev.data_ptr(), bdev.data_ptr()
    try:
        A.check(lib.demfi_conv2d(C.byref(d), stream()), "conv2d")
        torch.cuda.synchronize()
    finally:
        A.set_option("tc_gen", 3)
    return wdev, bdev
