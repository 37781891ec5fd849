"""Role timers (clock64 cycles) of conv_s3 with parts switched off: cycles, not milliseconds (the SM clock moves with power).
tc_diag 1024 = the bare MMA issue loop (no waits, no commits, no other role)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.role_timers import run
s64 = dict(n=1, h=736, w=1280, srcC=[64], co=64, k=(3, 3))
s64x3 = dict(n=3, h=736, w=1280, srcC=[64], co=64, k=(3, 3))
rdb = dict(n=1, h=368, w=640, srcC=[192], co=32, k=(3, 3))
chr_ = dict(n=1, h=736, w=1280, srcC=[64, 64, 64], co=64, k=(7, 7))
gru = dict(n=1, h=736, w=1280, srcC=[64, 64], co=128, k=(1, 5))
n16 = dict(n=3, h=736, w=1280, srcC=[64], co=3, k=(3, 3))
n96 = dict(n=1, h=368, w=640, srcC=[96], co=96, k=(3, 3))
which = sys.argv[1] if len(sys.argv) > 1 else "all"
shapes = {"all": (s64, s64x3, rdb, chr_, gru), "n": (n16, rdb, s64x3, n96)}[which]
for sh in shapes:
    for diag in (0, 1024):
        run(sh, s16=True, tc_diag=diag)
