"""A/B of Dec_first_2 with its loop-invariant part hoisted out of the boosting loop (DEMFI_HOIST_D2): ms per forward at 736x1280,
N_tst = 3, two engines on the same box, interleaved; max-abs difference of the first D2 convolution's output and of St_final."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demfi_b200 import synth
from demfi_b200.engine import Engine
dev = torch.device("cuda:0")
x = synth.make_frames(736, 1280, 0).to(dev)
t = torch.tensor([[0.375]], device=dev)
engs, outs = {}, {}
for h in ("0", "1"):
    os.environ["DEMFI_HOIST_D2"] = h
    engs[h] = Engine(synth.make_state_dict(0), 1, 736, 1280, dev)
    for _ in range(2):
        outs[h] = engs[h].forward(x, t, 3)[1][-1][2].clone()
print("max-abs St_final hoisted vs not:", float((outs["0"] - outs["1"]).abs().max()), "p99.9:",
      float(torch.quantile((outs["0"] - outs["1"]).abs().flatten()[::7], 0.999)), flush=True)
for rep in range(3):
    for h in ("0", "1"):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            engs[h].forward(x, t, 3)
        e1.record()
        torch.cuda.synchronize()
        print(f"DEMFI_HOIST_D2={h}: {e0.elapsed_time(e1) / 8:.3f} ms per forward", flush=True)
