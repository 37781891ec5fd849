"""conv_s3 with parts switched off (tc_diag bits): which client of the SM slows the MMA stream?
1 = no TMA stores, 32 = no epilogue arithmetic / staging, 256 = no activation TMA loads after the first tile,
512 = no tcgen05.ld in the epilogue.  Timing only: results are garbage with any bit set."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demfi_b200 import _abi as A
from tools.bench_conv import make_conv, time_conv

SHAPES = {"D1 64->64 3x3 x3": dict(n=3, h=736, w=1280, srcC=[64], co=64, k=(3, 3)),
          "Ch_Reducer 192->64 7x7": dict(n=1, h=736, w=1280, srcC=[64, 64, 64], co=64, k=(7, 7)),
          "RDB 192->32 3x3 @1/2": dict(n=1, h=368, w=640, srcC=[192], co=32, k=(3, 3)),
          "GRU zr 128->128 1x5": dict(n=1, h=736, w=1280, srcC=[64, 64], co=128, k=(1, 5))}
for name, sh in SHAPES.items():
    for diag in (0, 1, 33, 256, 512, 256 | 33, 512 | 33, 256 | 512 | 33):
        A.set_option("tc_diag", diag)
        d, keep = make_conv(A.CONV_TC16, s16=True, **sh)
        ms = time_conv(d)
        macs = sh["n"] * sh["h"] * sh["w"] * sum(sh["srcC"]) * sh["co"] * sh["k"][0] * sh["k"][1]
        print(json.dumps({"conv": name, "tc_diag": diag, "ms": round(ms, 4), "TFLOPs": round(2 * macs / ms / 1e9, 1)}), flush=True)
        del keep
A.set_option("tc_diag", 0)
