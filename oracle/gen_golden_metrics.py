"""ORACLE TOOLING -- runs only in the build container (needs /root/reference and OpenCV).

Executes the reference's OWN `psnr`, `ssim`, `ssim_matlab_func`, `denorm255_np` and `crop_8x8` (utils.py:628-721) on seeded
synthetic frames, driven exactly like the call site main.py:763-771, and stores inputs + results in tests/golden/metrics.npz.
`utils.py` cannot be imported as a module (TabError at utils.py:271/273, and it imports matplotlib / skimage which are not
installed), so only the source lines of those five functions are compiled -- text read from the read-only reference at run
time, never copied into this repository.

    python oracle/gen_golden_metrics.py
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_UTILS = "/root/reference/utils.py"
WANTED = ("crop_8x8", "psnr", "ssim_matlab_func", "ssim", "denorm255_np")


def load_reference_functions():
    import cv2
    lines = open(REF_UTILS).read().split("\n")
    ns = {"np": np, "math": math, "cv2": cv2}
    for name in WANTED:
        start = next(i for i, l in enumerate(lines) if l.startswith("def %s(" % name))
        end = next((i for i in range(start + 1, len(lines))
                    if lines[i] and not lines[i][0].isspace() and not lines[i].startswith("#")), len(lines))
        exec(compile("\n".join(lines[start:end]).replace("\t", "    "), REF_UTILS + ":" + name, "exec"), ns)
    return ns


def reference_call_site(ns, pred_chw_f64, gt_chw_f32):
    """main.py:763-771 verbatim in meaning: prediction rounded, target not, both BGR->RGB, crop_8x8, psnr(target, output)."""
    output_img = np.around(ns["denorm255_np"](np.transpose(pred_chw_f64, [1, 2, 0])[:, :, ::-1]))
    target_img = ns["denorm255_np"](np.transpose(gt_chw_f32, [1, 2, 0])[:, :, ::-1])
    o, _, _ = ns["crop_8x8"](output_img)
    t, _, _ = ns["crop_8x8"](target_img)
    return ns["psnr"](t, o), ns["ssim"](t, o)


def make_case(h, w, seed, noise):
    """A smooth 'ground truth' in [-1,1] (fp32) and a prediction = truth + noise, with out-of-range values and exact .5 ties."""
    from demfi_b200 import synth
    rng = np.random.default_rng(seed)
    gt = synth.make_frames(h, w, seed).numpy()[0, :, 0].astype(np.float32)          # [3,H,W]
    gt = (gt * 1.2).astype(np.float32)                                               # some values beyond [-1,1]: clip path
    pred = (gt + noise * rng.standard_normal(gt.shape)).astype(np.float32)
    ties = rng.integers(0, 255, size=16)
    pred.reshape(-1)[:16] = ((ties + 0.5) / 255.0 * 2.0 - 1.0).astype(np.float32)    # near-tie values for np.around
    return pred, gt


def main():
    ns = load_reference_functions()
    out = {}
    for name, (h, w, seed, noise) in {"a": (48, 64, 1, 0.02), "b": (33, 45, 2, 0.2), "c": (96, 80, 3, 0.003),
                                      "d": (11, 11, 4, 0.05), "same": (24, 40, 5, 0.0)}.items():
        pred, gt = make_case(h, w, seed, noise)
        if name == "same":   # truth already on the 8-bit grid: only the fp32 rounding of the target's scaling is left
            gt = ((np.around((gt.astype(np.float64) + 1) / 2 * 255)) / 255 * 2 - 1).astype(np.float32)
            pred = gt.copy()
        p, s = reference_call_site(ns, pred.astype(np.float64), gt)
        out[name + "_pred"], out[name + "_gt"] = pred, gt
        out[name + "_psnr"], out[name + "_ssim"] = np.float64(p), np.float64(s)
        print(name, pred.shape, p, s)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **out)


if __name__ == "__main__":
    main()
