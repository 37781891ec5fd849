"""One warm-up forward + one forward at the north-star shape (for ncu launch lists / captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demfi_b200 import synth
from demfi_b200.engine import Engine
dev = torch.device("cuda:0")
eng = Engine(synth.make_state_dict(0), 1, 736, 1280, dev)
x = synth.make_frames(736, 1280, 0).to(dev)
t = torch.tensor([[0.375]], device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    eng.forward(x, t, 3)
torch.cuda.synchronize()
