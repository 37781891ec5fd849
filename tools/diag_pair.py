"""Role timers of the CTA-pair kernel beside the single-CTA kernel (clock64 cycles per role, medians over the CTAs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demfi_b200 import _abi as A
from tools.role_timers import run
s64x3 = dict(n=3, h=736, w=1280, srcC=[64], co=64, k=(3, 3))
chr_ = dict(n=1, h=736, w=1280, srcC=[64, 64, 64], co=64, k=(7, 7))
rdb = dict(n=1, h=368, w=640, srcC=[192], co=32, k=(3, 3))
for sh in (s64x3,):
    for kind in (A.CONV_TC16, A.CONV_TC16P):
        run(sh, kind=kind, s16=True)
