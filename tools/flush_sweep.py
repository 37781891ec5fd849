"""End-to-end error vs the reference goldens and single-conv error vs float64 as a function of the accumulation-segment cap
(tc_flush: stages per drained segment), for the 3xFP16 kernel.  Decides how long a chain of MMAs may run between drains."""
import json, math, os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from demfi_b200 import synth, _abi as A
from demfi_b200.engine import Engine
from oracle import demfi_oracle as O  # checker only
from gpu_util import nhwc, from_nhwc, run_conv

dev = torch.device("cuda:0")
sd = synth.make_state_dict(0)
meta = json.load(open(os.path.join(ROOT, "tests/golden/meta.json")))
g = torch.Generator().manual_seed(5)
x = torch.randn(2, 64, 64, 96, generator=g).relu()
w = torch.randn(64, 64, 3, 3, generator=g) * math.sqrt(2.0 / (128 * 9))
b = torch.randn(64, generator=g) * 0.1
want = F.conv2d(x.double(), w.double(), b.double(), padding=1)
for flush in (10, 12, 18, 20, 24, 32):
    A.set_option("tc_flush", flush)
    row = {"flush": flush}
    xb, _ = nhwc(x)
    out = torch.zeros(2, 64, 96, 64, device=dev)
    run_conv(w, b, [(xb, 64, 0)], (64, 96), A.CONV_TC16, [dict(ch0=0, nch=64, dst=out)])
    e = from_nhwc(out, 64).double() - want
    row["conv64_3x3_max_abs"] = float(e.abs().max())
    row["conv64_3x3_gain_slope"] = float((e * want).sum() / (want * want).sum())
    for case in ("c32x32_n1_noise", "c64x96_n3", "c256x256_n1"):
        cfg = meta["cases"][case]["cfg"]
        gold = dict(np.load(os.path.join(ROOT, "tests/golden", case + ".npz")))
        xx = synth.make_frames(cfg["h"], cfg["w"], 0, cfg["batch"], cfg["smooth"]).to(dev)
        t = torch.tensor(cfg["t"]).reshape(-1, 1).to(dev)
        eng = Engine(sd, cfg["batch"], cfg["h"], cfg["w"], dev)
        got = O.flatten_outputs(eng.forward(xx, t, cfg["n"]))
        torch.cuda.synchronize()
        row[case] = max(float((got[k].cpu() - torch.from_numpy(v)).abs().max()) for k, v in gold.items() if k in got)
        del eng
    print(json.dumps(row), flush=True)
A.set_option("tc_flush", 10)
