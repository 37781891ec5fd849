"""Deterministic synthetic inputs and weights for tests, smoke() and bench.py.

Nothing here touches the GPU or the oracle.  Everything is generated from
numpy's PCG64 generator so that fixtures made in the build container are
reproduced bit-for-bit on the GPU box (torch's CPU generator is not used).

Weight regime: `xavier_normal_` + zero bias is what the
reference applies at construction (`utils.py:173-180`, used at `main.py:176`).
With those weights the predicted flows reach tens of pixels and the reference's
Gaussian forward splat (`DeMFInet.py:654-680`, floor() corner selection) makes the
reference disagree with itself between thread counts.  The "tamed" regime scales
the flow/occlusion head rows by 0.3 and biases the flow head so that flows are a
few pixels: warps are exercised and the reference is self-consistent to ~4e-5.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import torch

NF = 64


def default_args(**over):
    """The model-relevant flags of `main.py:22-127` (only those `DeMFInet.py` reads)."""
    a = dict(gpu=0, nf=NF, scale_factor=2, num_ResB_FACFB=5, num_ResB_Dec=5,
             shared_FGAC_flag=True, visualization_flag=False)
    a.update(over)
    return SimpleNamespace(**a)


def param_shapes(nf: int = NF) -> "OrderedDict[str, tuple]":
    """Names and shapes of the 260 parameter tensors, in the reference's
    registration order (`DeMFInet.py:15-44` and sub-modules).  Tests check this
    table against `reference DeMFInet(args).state_dict()` (tests/golden/state_dict_keys.json)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()

    def conv(name, co, ci, kh, kw, d3=False):
        s[name + ".weight"] = (co, ci, 1, kh, kw) if d3 else (co, ci, kh, kw)
        s[name + ".bias"] = (co,)

    G0, G, C, R = 96, 32, 4, 12
    conv("FF_RDB_Module.SFENet1", G0, 48, 5, 5)
    conv("FF_RDB_Module.SFENet2", G0, G0, 3, 3)
    for i in range(R):
        for c in range(C):
            conv(f"FF_RDB_Module.RDBs.{i}.convs.{c}.conv.0", G, G0 + c * G, 3, 3)
        conv(f"FF_RDB_Module.RDBs.{i}.LFF", G0, G0 + C * G, 1, 1)
    conv("FF_RDB_Module.GFF.0", G0, R * G0, 1, 1)
    conv("FF_RDB_Module.GFF.1", G0, G0, 3, 3)
    conv("FF_RDB_Module.UPNet.0", 256, G0, 3, 3)
    conv("FF_RDB_Module.UPNet.2", 2 * nf + 5, 64, 3, 3)
    conv("FAC_FB_Module.conv_first", nf, nf, 3, 3)
    for i in range(5):
        conv(f"FAC_FB_Module.feature_extraction.{i}.conv1", nf, nf, 3, 3)
        conv(f"FAC_FB_Module.feature_extraction.{i}.conv2", nf, nf, 3, 3)
    conv("FAC_FB_Module.shared_FGAC.conv_ref_k", nf, nf, 1, 1)
    conv("FAC_FB_Module.shared_FGAC.conv_source_k", nf, nf, 1, 1)
    conv("FAC_FB_Module.shared_FGAC.w_gen", nf, 2 * nf, 3, 3)
    conv("FAC_FB_Module.shared_FGAC.w_gen_2", 1, nf, 3, 3)
    conv("FAC_FB_Module.shared_FGAC.fusion", nf, nf, 1, 1)
    conv("Refine_Module.enc1", nf, 3 * nf + 9, 4, 4)
    conv("Refine_Module.enc2", 2 * nf, nf, 4, 4)
    conv("Refine_Module.enc3", 4 * nf, 2 * nf, 4, 4)
    conv("Refine_Module.dec0", 4 * nf, 4 * nf, 3, 3)
    conv("Refine_Module.dec1", 2 * nf, 6 * nf, 3, 3)
    conv("Refine_Module.dec2", nf, 3 * nf, 3, 3)
    conv("Refine_Module.dec3", 2 * nf + 5, nf, 3, 3)
    conv("Dec_first", nf, nf, 3, 3, True)
    for i in range(5):
        conv(f"Decoder_res.{i}.conv1", nf, nf, 3, 3, True)
        conv(f"Decoder_res.{i}.conv2", nf, nf, 3, 3, True)
    conv("Dec_last1", nf, nf, 3, 3, True)
    conv("Dec_last2", 3, nf, 3, 3, True)
    conv("Ch_Reducer", nf, 3 * nf, 7, 7)
    conv("Booster_Module.Mixer.conv_ref1", nf // 2, 30, 7, 7)
    conv("Booster_Module.Mixer.conv_ref2", nf // 2, nf // 2, 3, 3)
    conv("Booster_Module.Mixer.conv_delta1", nf // 2, 5, 7, 7)
    conv("Booster_Module.Mixer.conv_delta2", nf // 2, nf // 2, 3, 3)
    conv("Booster_Module.Mixer.conv_blend1", nf // 2, nf, 3, 3)
    conv("Booster_Module.Mixer.conv_blend2", nf, nf // 2, 3, 3)
    for g in ("z", "r", "q"):
        conv(f"Booster_Module.GB.conv{g}1", nf, 2 * nf, 1, 5)
    for g in ("z", "r", "q"):
        conv(f"Booster_Module.GB.conv{g}2", nf, 2 * nf, 5, 1)
    conv("Booster_Module.flow_occ.conv1", nf // 2, nf, 3, 3)
    conv("Booster_Module.flow_occ.conv2", 5, nf // 2, 3, 3)
    conv("Dec_first_2", nf, 9 + nf + 9 + 5 + 12, 3, 3)
    for i in range(5):
        conv(f"Decoder_res_2.{i}.conv1", nf, nf, 3, 3)
        conv(f"Decoder_res_2.{i}.conv2", nf, nf, 3, 3)
    conv("Dec_last1_2", nf, nf, 3, 3)
    conv("Dec_last2_2", 9, nf, 3, 3)
    return s


def make_state_dict(seed: int = 0, tame_scale: float | None = 0.3, tame_bias: float = 2.3,
                    nf: int = NF) -> "OrderedDict[str, torch.Tensor]":
    """xavier_normal weights, zero biases (semantics of `utils.py:173-180`), then the
    SURVEY 7.4 taming of the flow / occlusion heads.  `tame_scale=None` leaves the
    plain xavier weights."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shp in param_shapes(nf).items():
        if name.endswith(".weight"):
            rf = int(np.prod(shp[2:]))
            fan_in, fan_out = shp[1] * rf, shp[0] * rf
            std = math.sqrt(2.0 / (fan_in + fan_out))
            w = (rng.standard_normal(shp) * std).astype(np.float32)
            sd[name] = torch.from_numpy(w)
        else:
            sd[name] = torch.zeros(shp, dtype=torch.float32)
    if tame_scale is not None:
        sd["FF_RDB_Module.UPNet.2.weight"][2 * nf:2 * nf + 5] *= tame_scale
        sd["Refine_Module.dec3.weight"][0:5] *= tame_scale
        sd["Booster_Module.flow_occ.conv2.weight"] *= tame_scale
        b = tame_bias
        sd["FF_RDB_Module.UPNet.2.bias"][2 * nf:2 * nf + 4] += torch.tensor(
            [b, -0.6 * b, -b, 0.6 * b], dtype=torch.float32)
    return sd


def _box_blur(a: np.ndarray, k: int) -> np.ndarray:
    """k x k box blur with edge replication along the last two axes (float64)."""
    r = k // 2
    for ax in (-2, -1):
        pad = [(0, 0)] * a.ndim
        pad[ax] = (r, r)
        p = np.pad(a, pad, mode="edge")
        c = np.cumsum(p, axis=ax)
        z = np.zeros_like(np.take(c, [0], axis=ax))
        c = np.concatenate([z, c], axis=ax)
        n = a.shape[ax]
        hi = np.take(c, np.arange(k, k + n), axis=ax)
        lo = np.take(c, np.arange(0, n), axis=ax)
        a = (hi - lo) / k
    return a


def make_frames(h: int, w: int, seed: int = 0, batch: int = 1, smooth: bool = True) -> torch.Tensor:
    """x[B,3,4,H,W] fp32 in [-1,1], frame order (B0,B1,B-1,B2) as `utils.py:568-571`.

    smooth=True: Gaussian noise -> two 9x9 box blurs -> normalise; the four frames are
    integer-shifted crops of one larger canvas (a slowly panning scene).
    smooth=False: uniform noise (config-1 plumbing input)."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    if not smooth:
        x = rng.random((batch, 3, 4, h, w), dtype=np.float32) * 2 - 1
        return torch.from_numpy(x)
    m = 8
    out = np.empty((batch, 3, 4, h, w), dtype=np.float32)
    for b in range(batch):
        canvas = rng.standard_normal((3, h + 2 * m, w + 2 * m))
        canvas = _box_blur(_box_blur(canvas, 9), 9)
        canvas = canvas / (np.abs(canvas).max() + 1e-12)
        # time order B-1, B0, B1, B2 pans by (dy,dx) = (1,2) px per frame
        for slot, tt in ((0, 1), (1, 2), (2, 0), (3, 3)):
            oy, ox = m + (tt - 1) * 1, m + (tt - 1) * 2
            out[b, :, slot] = canvas[:, oy:oy + h, ox:ox + w]
    return torch.from_numpy(out)


def mfi_t_values(multiple: int) -> list:
    """t = 1/M ... (M-1)/M as `utils.py:556-566` enumerates for x M interpolation."""
    return [i / multiple for i in range(1, multiple)]
