"""Parity at the size BASELINE.json's metric is quoted on (1280x720 reflect-padded to 1280x736, N_tst = 3).

test_full_size_forward_against_reference_golden (default under -m gpu, no CPU oracle time): the B200 forward against
tests/golden/full_736x1280_n3.npz -- outputs of the UNMODIFIED reference module run once at this size by
oracle/gen_golden_full.py: point samples on a stride-4 lattice (exact fp32 values) and 8x8 block sums of the full-resolution
tensors (every pixel contributes).  The network contains discontinuous operators (floor() in the flow-reversal splat,
bwarp's 0.999 validity threshold, DeMFInet.py:606-766): a 1e-6 difference upstream flips isolated pixels by 1e-2..1e-1
downstream, and the decoders' receptive fields spread each flip over a neighbourhood.  The reference does this to ITSELF
when only its CPU thread count changes; the golden's .json records that yardstick at this very size (1 thread vs all:
St_final max-abs 0.13, 1.1e-3 of the samples beyond 5e-4, 68 blocks of flow0 moved).  Hence:
  * reported, not asserted: the literal max-abs per tensor (it equals a flip's size, as in the reference's self-noise);
  * asserted: away (>= 32 px) from the blocks where the splat took the other floor() branch, max-abs <= 5e-4 -- the north
    star's tolerance -- on every sampled tensor and block-mean error <= 2e-4 on every block; over ALL samples p99 <= 5e-4,
    the fraction beyond 5e-4 and the number of flipped blocks no larger than 2x the reference's own; PSNR(ours, reference)
    >= 60 dB on the frames (the north star asks for 0.01 dB of PSNR against ground truth).
The figures are written to gpurun_out/parity_full_size_golden.json (committed copy: profiles/r2_parity_full_size.json).

test_full_size_forward_against_oracle: opt-in (DEMFI_FULL_PARITY=1), the same forward against the ORACLE run on the GPU
box's host cores at full resolution (minutes of CPU) -- every pixel pointwise; report profiles/r1_parity_full_size.json."""
import json
import os
import time

import pytest
import torch

from demfi_b200 import metrics, synth
from demfi_b200.DeMFInet import DeMFInet
from oracle import demfi_oracle as O

pytestmark = [pytest.mark.gpu]
H, W, N, T = 736, 1280, 3, 0.375
TOL = 5e-4
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# (H, W, N_tst, t): the headline size, and BASELINE config 5's 3840x2176 / N_tst = 5 / x16 MFI scaled by 1/4 per side (the 4K frame
# itself is run whole by `bench.py --workload 4k`; a reference golden at that size would be half an hour of CPU and 1.7 GB)
CASES = [pytest.param(736, 1280, 3, 0.375, id="736x1280_n3"), pytest.param(544, 960, 5, 0.4375, id="4k_quarter_544x960_n5")]


@pytest.mark.parametrize("H,W,N,T", CASES)
def test_full_size_forward_against_reference_golden(state_dict, H, W, N, T):
    import numpy as np
    dev = torch.device("cuda:0")
    stem = f"full_{H}x{W}_n{N}"
    gold = np.load(os.path.join(GOLD, stem + ".npz"))
    with open(os.path.join(GOLD, stem + ".json")) as f:
        meta = json.load(f)
    assert meta["shape"] == [H, W] and meta["N_tst"] == N and meta["t"] == T
    noise = meta["self_noise"]
    net = DeMFInet(synth.default_args()).to(dev).eval()
    net.load_state_dict(state_dict, strict=True)
    with torch.no_grad():
        res = net(synth.make_frames(H, W, seed=0).to(dev), torch.tensor([[T]], device=dev), N)
    torch.cuda.synchronize()
    got = O.flatten_outputs(res)

    def blk(v):
        b, c, h, w = v.shape
        return v.double().reshape(b, c, h // 8, 8, w // 8, 8).sum(dim=(3, 5))

    def g(name):
        return torch.from_numpy(gold[name]).to(dev)

    # blocks where the complementary-flow-reversal splat took the other floor() branch than the reference run
    site = ((blk(got["flow0"]) - g("flow0/blk").double()).abs().amax(dim=1, keepdim=True) > 1e-3).float()   # [1,1,H/8,W/8]
    near_blk = torch.nn.functional.max_pool2d(site, 9, 1, 4) > 0                                           # within 32 px
    near_px = near_blk.repeat_interleave(8, dim=2).repeat_interleave(8, dim=3)[..., 1::4, 2::4]
    report = {"what": f"demfi_b200 forward on B200 vs the unmodified reference (tests/golden/{stem}.npz)", "shape": [H, W],
              "N_tst": N, "t": T, "tolerance": TOL, "splat_flip_blocks": int(site.sum()),
              "reference_self_noise_flip_blocks": noise["splat_flip_blocks"], "masked_fraction_r32": float(near_blk.float().mean()),
              "tensors": {}}
    for key in sorted(gold.files):
        name, kind = key.split("/")
        ent = report["tensors"].setdefault(name, {})
        if kind == "pts":
            e = (got[name][..., 1::4, 2::4] - g(key)).abs()
            ea = e.amax(dim=1, keepdim=True)
            ent.update({"pts_max_abs": float(e.max()), "pts_p99": float(torch.quantile(e.flatten()[:4_000_000], 0.99)),
                        "pts_frac_gt_5e-4": float((e > TOL).float().mean()),
                        "pts_max_abs_away_from_flips_r32": float(ea[~near_px].max()),
                        "reference_self_noise": {k: noise["tensors"][name][k] for k in ("pts_max_abs", "pts_frac_gt_5e-4")}})
            if name.startswith("S"):
                mse = float((((got[name][..., 1::4, 2::4] - g(key)) * 0.5) ** 2).mean())
                ent["psnr_ours_vs_ref_db"] = 99.0 if mse == 0 else -10.0 * float(np.log10(mse))
        else:
            eb = (blk(got[name]) - g(key).double()).abs().amax(dim=1, keepdim=True) / 64.0
            ent.update({"blk_mean_max_abs": float(eb.max()), "blk_mean_max_abs_away_from_flips_r32": float(eb[~near_blk].max())})
    for name, ent in report["tensors"].items():
        print(name, json.dumps(ent))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_full_size_golden.json" if (H, W) == (736, 1280) else f"gpurun_out/parity_{stem}_golden.json", "w") as f:
        json.dump(report, f, indent=1)
    # yardstick: the reference against itself (1 thread vs all) at this size; at 544x960 that pair of runs happened to take the
    # same floor() branch everywhere (0 flipped blocks), so the bound has a floor of 0.5 % of the blocks / 1e-3 of the samples --
    # the reference's own figures at 736x1280 (68 of 14720 blocks, 1.4e-3 of the samples)
    nblocks = (H // 8) * (W // 8)
    assert report["splat_flip_blocks"] <= max(2 * noise["splat_flip_blocks"], nblocks // 200), report["splat_flip_blocks"]
    assert report["masked_fraction_r32"] < 0.25
    for name, ent in report["tensors"].items():
        if "pts_p99" in ent:
            assert ent["pts_p99"] < TOL, (name, ent)
            assert ent["pts_max_abs_away_from_flips_r32"] < TOL, (name, ent)
            assert ent["pts_frac_gt_5e-4"] <= max(2 * ent["reference_self_noise"]["pts_frac_gt_5e-4"], 1e-3), (name, ent)
        if "blk_mean_max_abs" in ent:
            assert ent["blk_mean_max_abs_away_from_flips_r32"] < 2e-4, (name, ent)
        if "psnr_ours_vs_ref_db" in ent:
            assert ent["psnr_ours_vs_ref_db"] > 60.0, (name, ent)


@pytest.mark.skipif(os.environ.get("DEMFI_FULL_PARITY") != "1", reason="opt-in: DEMFI_FULL_PARITY=1 (minutes of CPU)")

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
