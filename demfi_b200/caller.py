"""The caller side of the hot path.

`patch_forward_DeFInet_itr` mirrors the reference's only inference caller of the model
(`utils.py:1339-1477`): reflect-pad H, W up to a multiple of `patch_boundary`, run the model,
copy every output into float64 numpy arrays, crop back.  As in the reference only
`patch=(1, 1)` is supported (the reference's `trim_patch_boundary` slices Python lists and
raises for any other tiling); other values raise here too.

`interpolate` is the lean call the clip runner and bench.py use: same padding, result stays a
torch tensor (no 13 float64 host copies per call).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def pad_to_multiple(input_frames: torch.Tensor, m: int):
    """utils.py:1351-1365: reflect padding on the right / bottom only."""
    B, C, T, h, w = input_frames.shape
    if m == 0:
        return input_frames, h, w
    ph, pw = (m - h % m) % m, (m - w % m) % m
    if ph == 0 and pw == 0:
        return input_frames, h, w
    x = input_frames.contiguous().view(B, C * T, h, w)
    x = F.pad(x, pad=[0, pw, 0, ph], mode="reflect")
    return x.view(B, C, T, h + ph, w + pw).contiguous(), h, w


def patch_forward_DeFInet_itr(model_net, input_frames, frameT, t_value, num_update, patch, patch_boundary=0):
    """Drop-in for utils.py:1339-1477.  Returns (two_blurry_inputs, Sharps_prime[3], Sharps_final[3], St_GT,
    [[ft0_init, ft0_final], [ft1_init, ft1_final]], [occ0_init, occ0_final]) as float64 numpy arrays."""
    if tuple(patch) != (1, 1):
        raise TypeError("only test_patch=(1,1) works in the reference (utils.py:1777-1798 slices lists); same here")
    x, oh, ow = pad_to_multiple(input_frames, patch_boundary)
    sharps_prime, sharps_final, flows, occs, two_blurry = model_net(x, t_value, num_update)
    np64 = lambda t: np.squeeze(t.detach().cpu().numpy()).astype(np.float64)[..., :oh, :ow]
    Sp = [np64(s) for s in sharps_prime]
    Sf = [np64(s) for s in sharps_final[-1]]
    St_GT = np.squeeze(frameT) if frameT is not None else 0
    f0, f1 = flows[0], flows[-1]
    flows_pred = [[np64(f0[:, :2]), np64(f1[:, :2])], [np64(f0[:, 2:]), np64(f1[:, 2:])]]
    occs_pred = [np64(occs[0])[None] if occs[0].shape[1] == 1 else np64(occs[0]),
                 np64(occs[-1])[None] if occs[-1].shape[1] == 1 else np64(occs[-1])]
    return np64(two_blurry), Sp, Sf, St_GT, flows_pred, occs_pred


@torch.no_grad()
def interpolate(model_net, input_frames: torch.Tensor, t_value: torch.Tensor, num_update: int,
                patch_boundary: int = 32, reuse_prefix: bool = False):
    """(S0_final, S1_final, St_final) of the last boosting iteration, cropped to the input size (torch, on device)."""
    x, oh, ow = pad_to_multiple(input_frames, patch_boundary)
    res = model_net(x, t_value, num_update, reuse_prefix=reuse_prefix)
    return tuple(s[..., :oh, :ow] for s in res[1][-1])
