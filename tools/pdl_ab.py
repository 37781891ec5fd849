"""A/B of programmatic dependent launch between the conv kernels: ms per forward at 736x1280, N_tst=3."""
import os, sys
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
for rep in range(2):
    for pdl in (0, 1):
        A.set_option("tc_pdl", pdl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6):
            eng.forward(x, t, 3)
        e1.record()
        torch.cuda.synchronize()
        print(f"tc_pdl={pdl}: {e0.elapsed_time(e1) / 6:.3f} ms per forward", flush=True)
