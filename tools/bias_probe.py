"""Is the tensor-core accumulation error a predictable gain error?  Regress (got - exact) on exact for the
tcgen05 conv at several chain lengths (flush=0: one chain of 4*stages MMAs)."""
import json, math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from demfi_b200 import _abi as A
from gpu_util import DEV, from_nhwc, nhwc, run_conv

def rnd(*shape, seed=0, scale=1.0):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((g.standard_normal(shape) * scale).astype(np.float32))

cases = [("64->64 3x3", 64, 64, (3, 3)), ("128->64 3x3", 128, 64, (3, 3)), ("192->64 7x7", 192, 64, (7, 7)), ("64->64 1x1", 64, 64, (1, 1)), ("96->32 3x3", 96, 32, (3, 3))]
for relu_in in (False, True):
  for name, ci, co, k in cases:
    n, h, w = 1, 48, 64
    x = rnd(n, ci, h, w, seed=3)
    if relu_in:
        x = F.relu(x)
    wt = rnd(co, ci, *k, seed=1, scale=math.sqrt(2.0 / ((ci + co) * k[0] * k[1])))
    b = torch.zeros(co)
    want = F.conv2d(x.double(), wt.double(), None, padding=(k[0] // 2, k[1] // 2))
    xb, _ = nhwc(x)
    stages = ((ci + 31) // 32) * k[0] * k[1]
    for flush, comp in ((0, 0), (2, 0), (0, int(sys.argv[1]) if len(sys.argv) > 1 else 0)):
        A.set_option("tc_flush", flush)
        A.set_option("tc_comp_milli", comp)
        out = torch.zeros(n, h, w, co, device=DEV)
        run_conv(wt, b, [(xb, ci, 0)], (h, w), A.CONV_TC16, [dict(ch0=0, nch=co, dst=out)])
        got = from_nhwc(out, co).double()
        err = (got - want).flatten(); wv = want.flatten()
        slope = float((err * wv).sum() / (wv * wv).sum())
        nm = 4 * (stages if flush == 0 else min(flush, stages))
        print(json.dumps({"conv": name, "relu_in": relu_in, "flush": flush, "comp_milli": comp, "mmas_per_chain": nm, "slope": slope,
                          "slope_per_mma_in_2^-24": slope / nm * 2 ** 24, "max_abs": float(err.abs().max()),
                          "rms_after_removing_slope": float((err - slope * wv).pow(2).mean().sqrt())}), flush=True)
A.set_option("tc_flush", 10); A.set_option("tc_comp_milli", 270)
