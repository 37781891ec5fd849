"""Per-role cycle breakdown of one tcgen05 conv launch (tc_diag & 128): where each warp role waits."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from demfi_b200 import _abi as A
from tools.bench_conv import make_conv, time_conv

NAMES = {0: "split.total", 1: "split.wait_full|epi.top(wait_read+operands)", 2: "epi.store.wait_staging", 3: "epi.store.fence+bar", 4: "epi.total", 5: "epi.wait_acc",
         6: "epi.store", 7: "split.convert|epi.store.tma_issue", 8: "prod.total", 9: "prod.wait_empty", 10: "epi.tmem_ld", 11: "epi.arrive", 12: "mma.total", 13: "mma.wait_acc_free", 14: "mma.wait_A", 15: "mma.wait_peer"}


def run(shape, kind=A.CONV_TC16, s16=False, **opts):
    base = dict(tc_flush=10, tc_stages=0, tc_grid=0, tc_split=3, tc_mask_hi=1, tc_a_tmem=1, tc_diag=0, tc_comp_milli=270, tc_gen=3)
    base.update(opts)
    base["tc_diag"] |= 128
    for k, v in base.items():
        A.set_option(k, v)
    d, keep = make_conv(kind, s16=s16, **shape)
    ms = time_conv(d)
    buf = np.zeros((148, 16), dtype=np.int64)
    A.check(A.lib().demfi_tc_debug_read(buf.ctypes.data_as(C.POINTER(C.c_int64)), 148), "debug_read")
    med = np.median(buf[0::2], axis=0) if kind == A.CONV_TC16P else np.median(buf, axis=0)   # pairs: the leaders (even CTAs)
    print(json.dumps({"kind": kind, "s16": s16, "shape": {k: v for k, v in shape.items()}, **opts, "ms": round(ms, 3),
                      "kclk": {NAMES[i]: round(float(med[i]) / 1e3, 1) for i in NAMES}}), flush=True)
    A.set_option("tc_diag", 0)


if __name__ == "__main__":
    s64 = dict(n=1, h=736, w=1280, srcC=[64], co=64, k=(3, 3))
    rdb = dict(n=1, h=368, w=640, srcC=[192], co=32, k=(3, 3))
    chr_ = dict(n=1, h=736, w=1280, srcC=[64, 64, 64], co=64, k=(7, 7))
    gru = dict(n=1, h=736, w=1280, srcC=[64, 64], co=128, k=(1, 5))
    run(s64, tc_gen=2)
    run(s64)
    run(s64, s16=True)
    run(s64, s16=True, tc_diag=4)   # weights streamed instead of resident
    run(s64, s16=True, tc_diag=1)   # no stores
    run(rdb)
    run(rdb, s16=True)
    run(chr_)
    run(gru)
    run(gru, s16=True)
