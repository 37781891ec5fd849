"""PSNR / SSIM of the frames `DeMFInet.forward` returns, computed on the GPU -- the host-side mirror of the metric code
of the reference's evaluation loop:

  * `psnr`, `ssim`                      utils.py:652-660, 663-705 (same names, same argument meaning: 0..255 images)
  * `AverageClass`                      utils.py:113-137 (val / avg / sum / count bookkeeping of the test loop)
  * `frame_metrics`, `score_forward`    the call sites main.py:757-838: D1 ("prime") and final-stage PSNR / SSIM of
                                        St, S0, S1 against their ground truths, BGR tensors in [-1,1]

The images never leave the device: `demfi_frame_metrics` (csrc/metrics.cu) scales, rounds, filters and reduces in fp64 and
only 16 bytes per frame come back.  No CPU fallback: without the CUDA library these functions raise.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence, Tuple

import torch

from . import _abi as A

_WS: Dict[Tuple[int, int], torch.Tensor] = {}


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), 0)
    ws = _WS.get(key)
    if ws is None or ws.numel() * 8 < nbytes:
        ws = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
        _WS[key] = ws
    return ws


def metric_sums(pred: torch.Tensor, target: torch.Tensor, target_is_prediction: bool = False) -> torch.Tensor:
    """[B,2] float64 on the device: per image the sum of squared errors and the sum of the SSIM map (see include/demfi_b200.h).
    pred, target: [B,C,H,W] (or [C,H,W]) fp32 CUDA tensors in [-1,1]."""
    if pred.dim() == 3:
        pred, target = pred[None], target[None]
    if pred.shape != target.shape:
        raise ValueError("Input images must have the same dimensions.")  # utils.py:692-693
    if not (pred.is_cuda and target.is_cuda):
        raise A.DemfiError("demfi_b200.metrics runs on the GPU only (no CPU fallback)")
    if pred.dtype != torch.float32 or target.dtype != torch.float32:
        raise TypeError("metrics expect float32 tensors in [-1,1]")
    pred, target = pred.contiguous(), target.contiguous()
    B, C_, H, W = pred.shape
    need = A.lib().demfi_frame_metrics_workspace(B, C_, H, W)
    if need < 0:
        raise ValueError(f"images must be at least 11 x 11 (the SSIM window), got {H} x {W}")
    ws = _workspace(pred.device, need)
    out = torch.empty(B, 2, dtype=torch.float64, device=pred.device)
    with torch.cuda.device(pred.device):
        A.check(A.lib().demfi_frame_metrics(pred.data_ptr(), target.data_ptr(), B, C_, H, W, 1 if target_is_prediction else 0,
                                            ws.data_ptr(), ws.numel() * 8, out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "demfi_frame_metrics")
    return out


def _finish(sums: Sequence[float], C_: int, H: int, W: int) -> Tuple[float, float]:
    mse = sums[0] / (C_ * H * W)
    p = float("inf") if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))   # utils.py:657-660
    return p, sums[1] / (C_ * (H - 10) * (W - 10))


def frame_metrics(pred: torch.Tensor, target: torch.Tensor, target_is_prediction: bool = False):
    """(psnr, ssim) of one frame [C,H,W] / [1,C,H,W], or a list of such pairs for a batch [B,C,H,W] (B > 1)."""
    sums = metric_sums(pred, target, target_is_prediction).cpu().tolist()
    C_, H, W = pred.shape[-3:]
    res = [_finish(s, C_, H, W) for s in sums]
    return res[0] if len(res) == 1 else res


def psnr(img1: torch.Tensor, img2: torch.Tensor) -> float:
    """utils.py:652 `psnr(target, output)` on [-1,1] device tensors (the 0..255 scaling of main.py:763-766 happens inside)."""
    return frame_metrics(img2, img1)[0]


def ssim(img1: torch.Tensor, img2: torch.Tensor) -> float:
    """utils.py:686 `ssim(target, output)` on [-1,1] device tensors."""
    return frame_metrics(img2, img1)[1]


class AverageClass:
    """Running value / average, the bookkeeping object of the reference's loops (`utils.AverageClass`, utils.py:113-137;
    prints as "name val (avg:avg)")."""

    def __init__(self, name: str, fmt: str = ":f"):
        self.name, self.fmt = name, fmt
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = 0.0
        self.count = 0

    def update(self, val, n: int = 1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __str__(self):
        return ("{name} {val" + self.fmt + "} (avg:{avg" + self.fmt + "})").format(name=self.name, val=self.val, avg=self.avg)


def score_forward(result, S0_GT: torch.Tensor, S1_GT: torch.Tensor, St_GT: torch.Tensor) -> Dict[str, float]:
    """The twelve numbers main.py:757-838 computes per (pair, t) from one `DeMFInet.forward` result: PSNR / SSIM of St, S0, S1
    after D1 (`*_prime`) and after the last boosting iteration, under the reference's variable names.  All six comparisons go
    through ONE kernel launch pair (a batch of six frames) and one 96-byte read-back."""
    s0p, s1p, stp = result[0]
    s0, s1, st = result[1][-1]
    preds = torch.cat([stp, s0p, s1p, st, s0, s1], dim=0)
    gts = torch.cat([St_GT, S0_GT, S1_GT, St_GT, S0_GT, S1_GT], dim=0).to(preds.dtype)
    if preds.shape[0] != 6:
        raise ValueError("score_forward scores one sample at a time (batch 1), like the reference's test loop")
    vals = frame_metrics(preds, gts)
    names = ["intp_test_{}_prime", "test_{}_S0_prime", "test_{}_S1_prime", "intp_test_{}", "test_{}_S0", "test_{}_S1"]
    out = {}
    for n, (p, s) in zip(names, vals):
        out[n.format("psnr")] = p
        out[n.format("ssim")] = s
    return out
