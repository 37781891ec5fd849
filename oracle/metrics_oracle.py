"""ORACLE (test infrastructure, never on the product path): CPU restatement of the frame metrics of the reference's
evaluation loop -- the step right after the hot path.

Follows, in numpy float64 and without OpenCV:
  * `psnr`               utils.py:652-660   20*log10(255/sqrt(mean((a-b)^2)))
  * `ssim_matlab_func`   utils.py:663-683   11x11 Gaussian window (sigma 1.5), 'valid' region, C1=(0.01*255)^2, C2=(0.03*255)^2
  * `ssim`               utils.py:686-705   for an [H,W,3] image the 3-channel map is averaged as a whole (the loop calls the
                                             same function on the same 3-channel arrays three times: mean of three equal numbers)
  * `denorm255_np`       utils.py:718-721   (x+1)/2, clip to [0,1], *255, in the dtype of x
  * the call site        main.py:763-771    prediction: float64, denormalised and rounded half-to-even (`np.around`);
                                             target: float32 arithmetic, NOT rounded; BGR->RGB flip and the (identity) `crop_8x8`

The window weights follow `cv2.getGaussianKernel(11, 1.5)` (OpenCV imgproc/smooth: exp(-0.5*(i-5)^2/sigma^2), normalised to
sum 1, float64), which is third-party code absent from /root/reference; `cv2.filter2D` on the cropped 'valid' region is a plain
correlation with the outer-product window.  Pinned against the reference's own functions executed here with OpenCV 4.13
(`oracle/gen_golden_metrics.py` -> `tests/golden/metrics.npz`, checked by `tests/test_metrics.py`).
"""
from __future__ import annotations

import math

import numpy as np


def gaussian_window_1d(n: int = 11, sigma: float = 1.5) -> np.ndarray:
    x = np.arange(n, dtype=np.float64) - (n - 1) * 0.5
    g = np.exp(-0.5 * x * x / (sigma * sigma))
    return g / g.sum()


def denorm255(x: np.ndarray) -> np.ndarray:
    out = (x + 1) / 2
    return out.clip(0, 1) * 255


def quantise_prediction(pred_chw: np.ndarray) -> np.ndarray:
    """main.py:763-764: float64 network output [3,H,W] in [-1,1] -> rounded 0..255 [H,W,3] RGB (float64)."""
    p = np.asarray(pred_chw, dtype=np.float64)
    return np.around(denorm255(np.transpose(p, [1, 2, 0])[:, :, ::-1]))


def scale_target(target_chw: np.ndarray) -> np.ndarray:
    """main.py:765-766: float32 ground truth [3,H,W] in [-1,1] -> 0..255 [H,W,3] RGB, float32 arithmetic, not rounded."""
    t = np.asarray(target_chw, dtype=np.float32)
    return denorm255(np.transpose(t, [1, 2, 0])[:, :, ::-1])


def psnr(img1: np.ndarray, img2: np.ndarray) -> float:
    a = img1.astype(np.float64)
    b = img2.astype(np.float64)
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return float("inf")
    return 20 * math.log10(255.0 / math.sqrt(mse))


def _valid_filter(img: np.ndarray, win: np.ndarray) -> np.ndarray:
    """Correlation with `win` over the positions where the window fits ([5:-5, 5:-5] of the reference's filter2D output)."""
    k = win.shape[0]
    h, w = img.shape[0] - k + 1, img.shape[1] - k + 1
    out = np.zeros((h, w) + img.shape[2:], dtype=np.float64)
    for i in range(k):
        for j in range(k):
            out += win[i, j] * img[i:i + h, j:j + w]
    return out


def ssim(img1: np.ndarray, img2: np.ndarray) -> float:
    if img1.shape != img2.shape:
        raise ValueError("Input images must have the same dimensions.")
    c1 = (0.01 * 255) ** 2
    c2 = (0.03 * 255) ** 2
    a = img1.astype(np.float64)
    b = img2.astype(np.float64)
    g = gaussian_window_1d()
    win = np.outer(g, g)
    mu1, mu2 = _valid_filter(a, win), _valid_filter(b, win)
    s11 = _valid_filter(a * a, win) - mu1 * mu1
    s22 = _valid_filter(b * b, win) - mu2 * mu2
    s12 = _valid_filter(a * b, win) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2))
    return float(m.mean())


def frame_metrics(pred_chw: np.ndarray, target_chw: np.ndarray):
    """(PSNR, SSIM) of one predicted frame against its ground truth exactly as main.py:763-771 computes them."""
    out_img = quantise_prediction(pred_chw)
    tgt_img = scale_target(target_chw)
    return psnr(tgt_img, out_img), ssim(tgt_img, out_img)
