"""ORACLE TOOLING -- runs only in the build container (needs /root/reference).

Golden at the size BASELINE.json's metric is quoted on: the UNMODIFIED reference module (`/root/reference/DeMFInet.py:13-179`)
run once on the 1280x720 frame reflect-padded to 1280x736 (`utils.py:1357-1366`), N_tst = 3, seeded synthetic input / weights
(`demfi_b200/synth.py`).  A full-resolution dump of the 17 returned tensors would be ~190 MB, so the committed file
`tests/golden/full_736x1280_n3.npz` keeps, per tensor,

  * `<name>/pts`  -- point samples on the stride-4 lattice `[..., 1::4, 2::4]` (fp32, exact values: pointwise max-abs / p99 /
                     fraction beyond 5e-4 / PSNR are computed on these 1/16 of the pixels), and
  * `<name>/blk`  -- the sums over every 8x8 block of the full-resolution tensor (accumulated in float64): every pixel of the
                     frame contributes, so a wrong tile / seam / edge column anywhere shows up as a block-mean error.

The block sums of `flow0` (the complementary-flow-reversal output, `DeMFInet.py:606-651`) also locate the sites where the
splat's floor() went the other way in the implementation under test (a block whose sum moved by > 1e-3 * 1 px): the
discontinuous operators make such flips unavoidable -- the reference disagrees with ITSELF there when only its CPU thread
count changes, which this script measures at the same size (1 thread against all threads) and stores beside the golden
(`self_noise` in tests/golden/full_736x1280_n3.json) as the yardstick of tests/test_full_size_parity.py.

    python oracle/gen_golden_full.py          # ~2 min with 8 threads + ~5 min for the 1-thread self-noise run
    python oracle/gen_golden_full.py 4k       # the same for the BASELINE config 5 shape scaled by 1/4 per side: 3840x2176 -> 960x544,
                                              # N_tst = 5, t = 7/16 (x16 MFI) -> tests/golden/full_544x960_n5.{npz,json}
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from demfi_b200 import synth  # noqa: E402
from gen_golden import load_reference, name_outputs  # noqa: E402

H, W, N, T, SEED = 736, 1280, 3, 0.375, 0
if len(sys.argv) > 1 and sys.argv[1] == "4k":
    H, W, N, T = 544, 960, 5, 0.4375
STEM = "full_%dx%d_n%d" % (H, W, N)
L = N - 1  # last boosting iteration
PTS = ["Stp", "St_final0", f"St_final{L}", f"S0_final{L}", "flow0", f"flow{N}", f"occ{N}"]
BLK = ["S0p", "S1p", "Stp", "St_final0", "St_final1", f"St_final{L}", f"S0_final{L}", f"S1_final{L}", "flow0", f"flow{N}", "occ0", f"occ{N}"]
TOL = 5e-4


def pts(v):
    return np.ascontiguousarray(v[..., 1::4, 2::4])


def blk(v):
    b, c, h, w = v.shape
    return v.astype(np.float64).reshape(b, c, h // 8, 8, w // 8, 8).sum(axis=(3, 5))


def run(net, x, t, threads):
    torch.set_num_threads(threads)
    t0 = time.time()
    with torch.no_grad():
        res = net(x, t, N)
    return {k: v.detach().numpy().astype(np.float32) for k, v in name_outputs(res).items()}, round(time.time() - t0, 1)


def main():
    ref_mod = load_reference()
    net = ref_mod.DeMFInet(synth.default_args(gpu=0)).eval()
    net.load_state_dict(synth.make_state_dict(0), strict=True)
    x = synth.make_frames(H, W, seed=SEED)
    t = torch.tensor([[T]])
    nt = os.cpu_count() or 1
    full, sec_all = run(net, x, t, nt)
    out = {}
    for k in PTS:
        out[k + "/pts"] = pts(full[k])
    for k in BLK:
        out[k + "/blk"] = blk(full[k]).astype(np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", STEM + ".npz"), **out)
    meta = {"what": "unmodified reference DeMFInet.forward (torch CPU fp32, %d threads) on synth.make_frames(%d, %d, seed=0), "
                    "synth.make_state_dict(0), t = %g, N_tst = %d" % (nt, H, W, T, N),
            "torch": torch.__version__, "shape": [H, W], "N_tst": N, "t": T, "seconds": sec_all,
            "pts": "v[..., 1::4, 2::4]", "blk": "8x8 block sums (float64 accumulation)"}
    # the reference against itself: one thread vs all threads, the same figures the GPU test computes
    one, sec_one = run(net, x, t, 1)
    site = np.abs(blk(one["flow0"]) - blk(full["flow0"])).max(axis=1) > 1e-3
    noise = {"threads": [1, nt], "seconds": [sec_one, sec_all], "splat_flip_blocks": int(site.sum()), "tensors": {}}
    for k in sorted(set(PTS + BLK)):
        e = np.abs(one[k] - full[k])
        ep = pts(e)
        noise["tensors"][k] = {"max_abs": float(e.max()), "frac_gt_5e-4": float((e > TOL).mean()),
                               "pts_frac_gt_5e-4": float((ep > TOL).mean()), "pts_max_abs": float(ep.max()),
                               "blk_mean_max_abs": float(np.abs(blk(one[k]) - blk(full[k])).max() / 64)}
    meta["self_noise"] = noise
    with open(os.path.join(ROOT, "tests", "golden", STEM + ".json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
