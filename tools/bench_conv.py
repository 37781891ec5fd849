"""Per-kernel timing of demfi_conv2d at the north-star resolution (736x1280) with CUDA events.
Reports algorithmic TFLOP/s (2*MAC, unpadded channels, the 3x split NOT credited)."""
import argparse
import ctypes as C
import json
import math
import sys, os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demfi_b200 import _abi as A

DEV = torch.device("cuda:0")


def make_conv(kind, n, h, w, srcC, co, k, act=A.ACT_RELU, seed=0, res=False, s16=False):
    lib = A.lib()
    g = torch.Generator().manual_seed(seed)
    ci = sum(srcC)
    kh, kw = k
    wt = (torch.randn(co, ci, kh, kw, generator=g) * math.sqrt(2.0 / ((ci + co) * kh * kw))).numpy()
    cout_pad = (co + 15) // 16 * 16
    sC = (A.i32 * len(srcC))(*srcC)
    nfl = lib.demfi_packed_weight_floats(kind, kh, kw, sC, len(srcC), cout_pad)
    packed = np.empty(nfl, dtype=np.float32)
    A.check(lib.demfi_pack_weights(kind, wt.ctypes.data, co, ci, kh, kw, (A.i32 * ci)(*range(ci)), sC, len(srcC),
                                   (A.i32 * cout_pad)(*(list(range(co)) + [-1] * (cout_pad - co))), cout_pad,
                                   packed.ctypes.data), "pack")
    keep = [torch.from_numpy(packed).to(DEV), torch.zeros(cout_pad, device=DEV)]
    d = A.Conv()
    d.N, d.H, d.W, d.Hi, d.Wi = n, h, w, h, w
    d.KH, d.KW, d.stride, d.pad_h, d.pad_w = kh, kw, 1, kh // 2, kw // 2
    d.nsrc, d.nseg, d.cout_pad, d.kind = len(srcC), 1, cout_pad, kind
    for i, c in enumerate(srcC):
        buf = torch.randn(n, h, w, c, device=DEV)
        keep.append(buf)
        d.src[i].ptr, d.src[i].C, d.src[i].ld, d.src[i].up = buf.data_ptr(), c, c, 0
        if s16 and c % 32 == 0:
            buf.copy_(torch.randn(n, h, w, c * 2, device=DEV).half().view(torch.float32))  # valid fp16 bit patterns
            d.src[i].fmt = A.FMT_S16
    ld = (co + 3) // 4 * 4
    out = torch.zeros(n, h, w, ld, device=DEV)
    keep.append(out)
    d.seg[0].dst, d.seg[0].dst_ld, d.seg[0].ch0, d.seg[0].nch, d.seg[0].act = out.data_ptr(), ld, 0, ld, act
    if s16 and co % 32 == 0:
        d.seg[0].fmt = A.SEG_DST_S16
    if res:
        rb = torch.randn(n, h, w, ld, device=DEV)
        keep.append(rb)
        d.seg[0].res, d.seg[0].res_ld = rb.data_ptr(), ld
        if s16 and co % 32 == 0:
            rb.copy_(torch.randn(n, h, w, ld * 2, device=DEV).half().view(torch.float32))
            d.seg[0].fmt |= A.SEG_RES_S16
    d.wpack, d.bias = keep[0].data_ptr(), keep[1].data_ptr()
    return d, keep


def time_conv(d, iters=5, warm=2):
    lib = A.lib()
    st = torch.cuda.current_stream(DEV).cuda_stream
    for _ in range(warm):
        A.check(lib.demfi_conv2d(C.byref(d), st), "conv")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        A.check(lib.demfi_conv2d(C.byref(d), st), "conv")
        ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))


SHAPES = [  # name, batch, res divisor, srcC, cout, k
    ("resblock 64->64 3x3 (x1 frame)", 1, 1, [64], 64, (3, 3)),
    ("resblock 64->64 3x3 (x3 frames, D1)", 3, 1, [64], 64, (3, 3)),
    ("resblock conv2 64->64 3x3 +res (x3 frames)", 3, 1, [64], 64, (3, 3)),
    ("conv_delta1 8->32 7x7", 1, 1, [8], 32, (7, 7)),
    ("blend2 32->64 3x3", 1, 1, [32], 64, (3, 3)),
    ("blend1 64->32 3x3", 1, 1, [64], 32, (3, 3)),
    ("delta2 32->32 3x3", 1, 1, [32], 32, (3, 3)),
    ("RDB conv 96->32 3x3 @1/2", 1, 2, [96], 32, (3, 3)),
    ("SFENet1 48->96 5x5 @1/2", 1, 2, [48], 96, (5, 5)),
    ("Ch_Reducer 192->64 7x7", 1, 1, [64, 64, 64], 64, (7, 7)),
    ("GRU zr 128->128 1x5", 1, 1, [64, 64], 128, (1, 5)),
    ("GRU q 128->64 5x1", 1, 1, [64, 64], 64, (5, 1)),
    ("GRU z 128->64 1x5", 1, 1, [64, 64], 64, (1, 5)),
    ("GRU z 128->64 1x5 sigmoid", 1, 1, [64, 64], 64, (1, 5)),
    ("GRU r 128->64 1x5 sigmoid_mul +res", 1, 1, [64, 64], 64, (1, 5)),
    ("one source 128->64 1x5", 1, 1, [128], 64, (1, 5)),
    ("resblock 64->64 5x5 (x1 frame)", 1, 1, [64], 64, (5, 5)),
    ("RDB conv 192->32 3x3 @1/2", 1, 2, [192], 32, (3, 3)),
    ("LFF 224->96 1x1 @1/2", 1, 2, [224], 96, (1, 1)),
    ("UPNet.0 96->256 3x3 @1/2", 1, 2, [96], 256, (3, 3)),
    ("UPNet.2 64->133 3x3", 1, 1, [64], 133, (3, 3)),
    ("w_gen 128->64 3x3 (x2)", 2, 1, [128], 64, (3, 3)),
    ("conv_ref1 32->32 7x7", 1, 1, [32], 32, (7, 7)),
    ("Dec_last2 64->3 3x3 (x3)", 3, 1, [64], 3, (3, 3)),
]

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=736)
    ap.add_argument("--w", type=int, default=1280)
    ap.add_argument("--kinds", default="tc16,ffma")
    ap.add_argument("--opts", default="")  # e.g. tc_mask_hi=0,tc_split=1
    ap.add_argument("--only", default="")
    ap.add_argument("--s16", action="store_true", help="sources / destination / residual in the S16 storage format where channel counts allow")
    a = ap.parse_args()
    for kv in filter(None, a.opts.split(",")):
        k, v = kv.split("=")
        A.set_option(k, int(v))
    rows = []
    for name, n, div, srcC, co, k in SHAPES:
        if a.only and a.only not in name:
            continue
        h, w = a.h // div, a.w // div
        macs = n * h * w * sum(srcC) * co * k[0] * k[1]
        row = {"conv": name, "GMAC": round(macs / 1e9, 2)}
        for kind_name in a.kinds.split(","):
            kind = {"tc16": A.CONV_TC16, "tc16h3": A.CONV_TC16, "ffma": A.CONV_FFMA, "tc16p": A.CONV_TC16P}[kind_name]
            if kind == A.CONV_TC16P and (co + 15) // 16 * 16 not in (32, 64):
                continue
            A.set_option("tc_gen", 2 if kind_name == "tc16h3" else 3)
            d, keep = make_conv(kind, n, h, w, srcC, co, k, res="+res" in name, act=(A.ACT_SIGMOID_MUL if "sigmoid_mul" in name else A.ACT_SIGMOID if "sigmoid" in name else A.ACT_NONE if "+res" in name else A.ACT_RELU), s16=a.s16)
            ms = time_conv(d)
            row[kind_name + "_ms"] = round(ms, 3)
            row[kind_name + "_TFLOPs"] = round(2 * macs / ms / 1e9, 1)
            del keep
            torch.cuda.empty_cache()
        rows.append(row)
        print(json.dumps(row), flush=True)
