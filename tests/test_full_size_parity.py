"""Opt-in (DEMFI_FULL_PARITY=1): the B200 forward against the ORACLE ITSELF at BASELINE.json's full size (1280x720 padded to
1280x736, N_tst=3) -- one to two minutes of CPU work on the GPU box, so it is not part of the default `-m gpu` run, which shows
full-size parity through properties (tests/test_forward_gpu.py).  Writes the figures SURVEY.md section 8(d) asks to be
reported beside throughput (max-abs, p99.99, fraction > 5e-4, PSNR / SSIM(ours, reference) on rounded 0..255 images) to
gpurun_out/parity_full_size.json; the committed copy is profiles/r1_parity_full_size.json.

Yardstick: the unmodified reference against itself (1 vs 8 CPU threads, half this size) differs by up to 4.5e-2 on 1.1e-3 of
the St samples (profiles/r1_reference_self_noise.json) -- isolated flips of the discontinuous operators (floor() in the
splat, the 0.999 validity threshold of bwarp) that the decoders' receptive fields spread over a neighbourhood.  The report
also gives, per tensor, max-abs over the pixels farther than r = 32 / 64 / 96 px from any site where the splat's floor() went
the other way (|flow0 - oracle| > 1e-3; 156 pixels of 942 080): that is where the 5e-4 bound applies, and it holds on every
returned tensor already at r = 32 (frames <= 9e-5, St_final <= 3e-4, flows 1e-5, occlusion 2e-6).  (The oracle with 4 threads against the oracle
with all 16 threads was measured bit-identical on the B200 box, so thread-count noise is not the yardstick there.)
The assertion is on the bulk (p99 < 5e-4;
measured 3e-5), the flip fraction (< 5e-3; measured 1.1-1.5e-3, the reference's own figure) and PSNR(ours, oracle) > 60 dB
(measured 78-81 dB), not on max-abs."""
import json
import os
import time

import pytest
import torch

from demfi_b200 import metrics, synth
from demfi_b200.DeMFInet import DeMFInet
from oracle import demfi_oracle as O

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("DEMFI_FULL_PARITY") != "1", reason="opt-in: DEMFI_FULL_PARITY=1 (minutes of CPU)")]
H, W, N, T = 736, 1280, 3, 0.375
TOL = 5e-4


def test_full_size_forward_against_oracle(state_dict):
    dev = torch.device("cuda:0")
    x = synth.make_frames(H, W, seed=0)
    t = torch.tensor([[T]])
    net = DeMFInet(synth.default_args()).to(dev).eval()
    net.load_state_dict(state_dict, strict=True)
    with torch.no_grad():
        res = net(x.to(dev), t.to(dev), N)
    torch.cuda.synchronize()
    got = O.flatten_outputs(res)
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    with torch.no_grad():
        ref = O.flatten_outputs(O.forward(state_dict, x, t, N))
    oracle_s = time.time() - t0
    # where the complementary-flow-reversal splat took a different floor() branch than the oracle's: everything downstream of
    # the UNet (three stride-2 levels) inherits those sites over a neighbourhood; away from them the 5e-4 bound must hold
    site = ((got["flow0"] - ref["flow0"].to(dev)).abs().amax(dim=1, keepdim=True) > 1e-3).float()
    masks = {r: torch.nn.functional.max_pool2d(site, 2 * r + 1, 1, r) > 0 for r in (32, 64, 96)}
    report = {"splat_flip_sites": int(site.sum()), "what": "demfi_b200 forward on B200 vs oracle (torch CPU fp32 restatement pinned to the reference) on the same input/weights",
              "shape": [H, W], "N_tst": N, "t": T, "oracle_seconds": round(oracle_s, 1), "cpu_threads": torch.get_num_threads(),
              "tolerance": TOL, "tensors": {}}
    for k, r in ref.items():
        g = got[k]
        rd = r.to(dev)
        e = (g - rd).abs().flatten()
        ent = {"max_abs": float(e.max()), "p99": float(torch.quantile(e[:: max(1, e.numel() // 4_000_000)], 0.99)),
               "p99_9": float(torch.quantile(e[:: max(1, e.numel() // 4_000_000)], 0.999)),
               "frac_gt_5e-4": float((e > TOL).float().mean()), "frac_gt_1e-2": float((e > 1e-2).float().mean()),
               "max_abs_ref": float(rd.abs().max())}
        if k.startswith("S") and g.shape[1] == 3:
            p, s = metrics.frame_metrics(g.contiguous(), rd.contiguous(), target_is_prediction=True)
            ent["psnr_ours_vs_ref_db"], ent["ssim_ours_vs_ref"] = p, s
        if g.shape[-2:] == (H, W) and k != "two_blurry":
            ea = (g - rd).abs().amax(dim=1, keepdim=True)
            ent["away_from_splat_flips"] = {f"r{r}": {"masked_frac": float(m.float().mean()), "max_abs": float(ea[~m].max())}
                                            for r, m in masks.items()}
        report["tensors"][k] = ent
        print(k, json.dumps(ent))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_full_size.json", "w") as f:
        json.dump(report, f, indent=1)
    for k, ent in report["tensors"].items():
        assert ent["p99"] < TOL, (k, ent)
        assert ent["frac_gt_5e-4"] < 5e-3, (k, ent)   # flips: same order as the oracle's own thread-count noise (reported beside it)
        if "away_from_splat_flips" in ent:   # measured: <= 3e-4 on every tensor with only r = 32 px (2.9 % of the image) excluded
            assert ent["away_from_splat_flips"]["r32"]["max_abs"] < TOL, (k, ent)
        if "psnr_ours_vs_ref_db" in ent:
            assert ent["psnr_ours_vs_ref_db"] > 60.0, (k, ent)
