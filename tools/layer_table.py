"""Per-layer table of one forward at 736x1280, N_tst=3 (CUDA events around every C-ABI call, mean of 3 forwards)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demfi_b200 import _abi as A, synth
from demfi_b200.engine import Engine
dev = torch.device("cuda:0")
eng = Engine(synth.make_state_dict(0), 1, 736, 1280, dev)
x = synth.make_frames(736, 1280, 0).to(dev)
t = torch.tensor([[0.375]], device=dev)
for _ in range(2):
    eng.forward(x, t, 3)
eng.profile = []
R = 3
for _ in range(R):
    eng.forward(x, t, 3)
torch.cuda.synchronize()
rows = {}
for op, e0, e1 in eng.profile:
    key = op[2] if op[0] == "conv" else op[0]
    d = rows.setdefault(key, {"n": 0, "ms": 0.0, "macs": 0})
    d["n"] += 1
    d["ms"] += e0.elapsed_time(e1)
    d["macs"] += op[4] if op[0] == "conv" else 0
tot = sum(d["ms"] for d in rows.values()) / R
print(f"total {tot:.2f} ms per forward")
for k, d in sorted(rows.items(), key=lambda kv: -kv[1]["ms"]):
    ms = d["ms"] / R
    tf = 2 * d["macs"] / R / (ms / 1e3) / 1e12 if d["macs"] else 0
    print(f"{k:55s} x{d['n'] // R:3d} {ms:7.3f} ms {100 * ms / tot:5.1f} %  {tf:6.1f} TFLOP/s")
