"""ORACLE TOOLING -- runs only in the build container (needs /root/reference).

How far does the UNMODIFIED reference module drift from ITSELF when only the CPU thread count changes (different summation
order inside the library convolutions)?  The network contains discontinuous operators (fwarp's floor, bwarp's 0.999 validity
mask, DeMFInet.py:606-766): a 1e-6 difference upstream flips isolated pixels by 1e-2..1e-1 downstream.  This is the yardstick for
the "fraction > 5e-4" figures of the full-size GPU property test (tests/test_forward_gpu.py) and profiles/r1_4k_whole_frame.json.

    python oracle/self_noise.py [H W]        # default 368 640, N_tst = 3; writes profiles/r1_reference_self_noise.json
"""
import json, os, sys, time
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from demfi_b200 import synth  # noqa: E402
from gen_golden import load_reference  # noqa: E402

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (368, 640)
ref = load_reference()
net = ref.DeMFInet(synth.default_args(gpu=0)).eval()
net.load_state_dict(synth.make_state_dict(0), strict=True)
x = synth.make_frames(h, w, seed=11)
t = torch.tensor([[0.625]])
outs, secs = [], []
for nt in (1, 8):
    torch.set_num_threads(nt)
    t0 = time.time()
    with torch.no_grad():
        r = net(x, t, 3)
    secs.append(round(time.time() - t0, 1))
    outs.append({"St_final": r[1][-1][2], "flow_N": r[2][-1]})
res = {"what": "unmodified reference DeMFInet (torch CPU fp32), same input and weights, 1 thread vs 8 threads", "shape": [h, w], "N_tst": 3,
       "seconds": secs}
for k in outs[0]:
    e = (outs[0][k] - outs[1][k]).abs()
    res[k] = {"max_abs": float(e.max()), "frac_gt_5e-4": float((e > 5e-4).float().mean()), "frac_gt_5e-5": float((e > 5e-5).float().mean())}
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "profiles", "r1_reference_self_noise.json"), "w"), indent=1)
