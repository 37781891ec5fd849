"""Same-engine A/B of a run-time option (default: tc_diag 32768 = ONE operand tile for the ResBlock conv2 against 0 = two tiles
in turn): ms per forward at 736x1280, N_tst = 3, interleaved.  Usage: python tools/opt_ab.py [option value_a value_b]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demfi_b200 import _abi as A, synth
from demfi_b200.engine import Engine
name, va, vb = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("tc_diag", 32768, 0)
dev = torch.device("cuda:0")
eng = Engine(synth.make_state_dict(0), 1, 736, 1280, dev)
x = synth.make_frames(736, 1280, 0).to(dev)
t = torch.tensor([[0.375]], device=dev)
outs = {}
for v in (va, vb):
    A.set_option(name, v)
    for _ in range(2):
        outs[v] = eng.forward(x, t, 3)[1][-1][2].clone()
print("max-abs St_final between the two settings:", float((outs[va] - outs[vb]).abs().max()), flush=True)
for rep in range(3):
    for v in (va, vb):
        A.set_option(name, v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            eng.forward(x, t, 3)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}={v}: {e0.elapsed_time(e1) / 8:.3f} ms per forward", flush=True)
A.set_option(name, vb)
