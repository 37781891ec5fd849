"""A/B of CUDA-graph replay against eager launches of one forward at 736x1280, N_tst = 3 (same engine, interleaved)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demfi_b200 import synth
from demfi_b200.engine import Engine
dev = torch.device("cuda:0")
eng = Engine(synth.make_state_dict(0), 1, 736, 1280, dev)
x = synth.make_frames(736, 1280, 0).to(dev)
t = torch.tensor([[0.375]], device=dev)
for g in (False, True):
    for _ in range(2):
        eng.forward(x, t, 3, graph=g)
for rep in range(3):
    for g in (False, True):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            eng.forward(x, t, 3, graph=g)
        e1.record()
        torch.cuda.synchronize()
        print(f"graph={int(g)}: {e0.elapsed_time(e1) / 8:.3f} ms per forward", flush=True)
