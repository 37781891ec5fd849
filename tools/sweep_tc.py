"""Diagnostic sweeps of the tcgen05 conv (one shape) over runtime knobs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demfi_b200 import _abi as A
from tools.bench_conv import make_conv, time_conv

H, W = 736, 1280
shape = dict(n=1, h=H, w=W, srcC=[64], co=64, k=(3, 3))
d, keep = make_conv(A.CONV_TC, **shape)
macs = H * W * 64 * 64 * 9
base = dict(tc_flush=10, tc_stages=0, tc_grid=0, tc_split=3, tc_mask_hi=1, tc_a_tmem=1, tc_diag=0, tc_comp_milli=270)
def run(**kw):
    o = dict(base); o.update(kw)
    for k, v in o.items():
        A.set_option(k, v)
    ms = time_conv(d)
    print(json.dumps({**kw, "ms": round(ms, 3), "TFLOPs": round(2 * macs / ms / 1e9, 1)}), flush=True)
for args in sys.argv[1:]:
    run(**{k: int(v) for k, v in (kv.split("=") for kv in args.split(","))})
for k, v in base.items():
    A.set_option(k, v)
