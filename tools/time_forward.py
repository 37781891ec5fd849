"""Times one full forward (default: 736x1280, N_tst=3) through the engine, per conv kind."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demfi_b200 import synth, _abi as A
from demfi_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--h", type=int, default=736)
ap.add_argument("--w", type=int, default=1280)
ap.add_argument("--n", type=int, default=3)
ap.add_argument("--kinds", default="auto,ffma")
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
sd = synth.make_state_dict(0)
x = synth.make_frames(a.h, a.w, 0).to(dev)
t = torch.tensor([[0.375]], device=dev)
outs = {}
for kind in a.kinds.split(","):
    t0 = time.time()
    eng = Engine(sd, 1, a.h, a.w, dev, conv_kind=kind)
    build_s = time.time() - t0
    for _ in range(2):
        r = eng.forward(x, t, a.n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = A.launch_count()
    e0.record()
    for _ in range(a.iters):
        r = eng.forward(x, t, a.n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    macs = eng.conv_macs(a.n)
    outs[kind] = r[1][-1][2].cpu()
    print(json.dumps({"kind": kind, "HxW": [a.h, a.w], "N": a.n, "ms_per_forward": round(ms, 2), "fps": round(1000 / ms, 3),
                      "conv_TFLOPs": round(2 * macs / ms / 1e9, 1), "launches_per_forward": (A.launch_count() - l0) // a.iters,
                      "workspace_GB": round(eng.workspace_bytes() / 1e9, 2), "build_s": round(build_s, 1),
                      "mem_alloc_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}), flush=True)
    del eng
    torch.cuda.empty_cache()
ks = list(outs)
if len(ks) == 2:
    print("max-abs St_final between kinds:", float((outs[ks[0]] - outs[ks[1]]).abs().max()))
