"""conv_s3 descriptor-variant check: max-abs error of a few shapes against float64 conv2d for tc_diag variants."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, torch.nn.functional as F
from demfi_b200 import _abi as A
from gpu_util import DEV, from_nhwc, nhwc, run_conv

def rnd(*shape, seed=0, scale=1.0):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((g.standard_normal(shape) * scale).astype(np.float32))

for diag in [int(a) for a in sys.argv[1:]] or [0, 8]:
    A.set_option("tc_diag", diag)
    for (n, h, w_, ci, co, k) in [(1, 16, 8, 32, 32, (1, 1)), (1, 16, 8, 32, 32, (3, 3)), (2, 24, 40, 64, 64, (3, 3)), (1, 40, 56, 128, 64, (1, 5)),
                                  (1, 40, 56, 128, 64, (5, 1)), (1, 24, 24, 32, 32, (7, 7))]:
        x = rnd(n, ci, h, w_, seed=3)
        w = rnd(co, ci, *k, seed=1, scale=math.sqrt(2.0 / ((ci + co) * k[0] * k[1])))
        b = rnd(co, seed=2, scale=0.1)
        xb, _ = nhwc(x)
        out = torch.zeros(n, h, w_, co, device=DEV)
        run_conv(w, b, [(xb, ci, 0)], (h, w_), A.CONV_TC16, [dict(ch0=0, nch=co, dst=out, act=A.ACT_NONE)])
        want = F.conv2d(x.double(), w.double(), b.double(), padding=(k[0] // 2, k[1] // 2))
        err = float((from_nhwc(out, co).double() - want).abs().max())
        print(f"diag {diag} {ci}->{co} {k} {h}x{w_}: max-abs err {err:.3e}", flush=True)
A.set_option("tc_diag", 0)
