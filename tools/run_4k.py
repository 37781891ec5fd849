"""BASELINE config 5 shape on ONE B200, whole frame, no tiler: 3840x2160 (padded 3840x2176), N_tst=5.  The reference tiles
4K because of memory; 180 GB of HBM hold the whole-frame workspace (~122 GB).  Checks the tensor-core path against the
CUDA-core fp32 path on the same input (size-dependent indexing) and reports seconds per interpolated frame."""
import gc, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demfi_b200 import synth
from demfi_b200.engine import Engine

H, W, N = 2176, 3840, int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
sd = synth.make_state_dict(0)
x = synth.make_frames(H, W, seed=0).to(dev)
t = torch.tensor([[7.0 / 16.0]], device=dev)
out = {}
for kind in ("auto", "ffma"):
    eng = Engine(sd, 1, H, W, dev, conv_kind=kind)
    ws = eng.workspace_bytes() / 1e9
    eng.forward(x, t, N, final_only=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = eng.forward(x, t, N, final_only=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[kind] = {"seconds_per_frame": round(dt, 3), "workspace_GB": round(ws, 1),
                 "St": res[1][-1][2].cpu(), "flow": res[2][-1].cpu(), "S1p": res[0][1].cpu(),
                 # pure conv stacks (no discontinuous operator upstream): FF_RDB features / flows, FAC-FB encoder output
                 "F01": eng.views["F01"].to_nchw().cpu(), "FO": eng.views["FO"].to_nchw().cpu(),
                 "SE": eng.views["SE"].to_nchw()[:, :64].cpu()}
    print(kind, out[kind]["seconds_per_frame"], "s per 4K interpolated frame, workspace", out[kind]["workspace_GB"], "GB", flush=True)
    del eng, res
    gc.collect()
    torch.cuda.empty_cache()
def stats(a, b):
    e = (a - b).abs().flatten()
    big = e[e > 5e-4]
    return {"max_abs": float(e.max()), "p99.99": float(e.float().kthvalue(int(0.9999 * e.numel())).values), "frac_gt_5e-4": float(big.numel() / e.numel())}
d = {k: stats(out["auto"][k], out["ffma"][k]) for k in ("F01", "FO", "SE", "S1p", "flow", "St")}
print(json.dumps({"shape": [H, W], "N_tst": N, "tensor_core_seconds_per_frame": out["auto"]["seconds_per_frame"],
                  "cuda_core_seconds_per_frame": out["ffma"]["seconds_per_frame"], "workspace_GB": out["auto"]["workspace_GB"],
                  "max_abs_tc_vs_cuda_core": d}))
