"""ORACLE TOOLING -- runs only in the build container (needs /root/reference).

Gradient goldens for the bilinear-sampling operators of the hot path, from torch autograd through the UNMODIFIED reference
functions `bwarp` (DeMFInet.py:732-766) + the Eq.(2) expression of `DeMFInet.forward` (:64-71) and `bilinear_sampler`
(:499-514) and `CFR_flow_t_align` (:606-729), on seeded inputs that include integer displacements, out-of-image targets and the 0.999 validity band.
Writes tests/golden/warp_grads.npz (inputs, upstream gradient, and the gradients w.r.t. every input).

    python oracle/gen_golden_grads.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.gen_golden import GOLD, load_reference  # noqa: E402


def flows(rng, n, c, h, w, scale):
    f = (rng.standard_normal((n, c, h, w)) * scale).astype(np.float32)
    f[:, :, 0:2, :] = np.round(f[:, :, 0:2, :])      # integer displacements
    f[:, :, 2, :] = 0.0                               # identity
    f[:, :, 3, 0:3] = 500.0                           # far outside
    f[:, 0, 4, :] = -np.arange(w, dtype=np.float32) - 0.0005   # lands in the 0.999 band at the left border
    return f


def main():
    ref = load_reference()
    dev = torch.device("cpu")
    rng = np.random.default_rng(0)
    out = {}
    for tag, C_ in (("c64", 64), ("c3", 3)):
        n, h, w = 2, 10, 14
        a = torch.tensor(rng.standard_normal((n, C_, h, w)).astype(np.float32), requires_grad=True)
        b = torch.tensor(rng.standard_normal((n, C_, h, w)).astype(np.float32), requires_grad=True)
        fl = torch.tensor(flows(rng, n, 4, h, w, 2.0), requires_grad=True)
        occ = torch.tensor(rng.standard_normal((n, 1, h, w)).astype(np.float32), requires_grad=True)
        t = torch.tensor([[0.375], [0.75]])
        gy = torch.tensor(rng.standard_normal((n, C_, h, w)).astype(np.float32))
        tv = t[:, :, None, None]                                                # DeMFInet.py:62
        o0 = torch.sigmoid(occ)
        o1 = 1 - o0
        res = (1 - tv) * o0 * ref.bwarp(dev, a, fl[:, 0:2]) + tv * o1 * ref.bwarp(dev, b, fl[:, 2:4])
        res = res / ((1 - tv) * o0 + tv * o1)                                   # Eq.(2), DeMFInet.py:68-71
        (res * gy).sum().backward()
        for k, v in dict(a=a, b=b, flow=fl, occ=occ, t=t, gy=gy, out=res, da=a.grad, db=b.grad, dflow=fl.grad, docc=occ.grad).items():
            out[f"blend_{tag}_{k}"] = v.detach().numpy()
    n, C_, h, w = 2, 64, 9, 12
    refk = torch.tensor(rng.standard_normal((n, C_, h, w)).astype(np.float32), requires_grad=True)
    fl = (rng.standard_normal((n, 2, h, w)) * 4.0 + 5.0).astype(np.float32)     # absolute positions
    fl[:, :, 0, :] = -3.0
    fl[:, :, 1, :] = np.round(fl[:, :, 1, :])
    fl = torch.tensor(fl, requires_grad=True)
    gy = torch.tensor(rng.standard_normal((n, C_, h, w)).astype(np.float32))
    smp = ref.bilinear_sampler(refk, fl.permute(0, 2, 3, 1))                    # DeMFInet.py:394,419
    (smp * gy).sum().backward()
    for k, v in dict(refk=refk, flow=fl, gy=gy, out=smp, drefk=refk.grad, dflow=fl.grad).items():
        out[f"sample_{k}"] = v.detach().numpy()
    # complementary flow reversal (Gaussian forward splat), DeMFInet.py:606-729
    n, h, w = 2, 12, 16
    f01 = (rng.standard_normal((n, 2, h, w)) * 2.5).astype(np.float32)
    f10 = (rng.standard_normal((n, 2, h, w)) * 2.5).astype(np.float32)
    f01[:, :, 0:2, :] = np.round(f01[:, :, 0:2, :]) * 8.0 / 3.0    # t * f lands on integers for t = 0.375: the floor() edge
    f10[:, :, 2, :] = 0.0
    f01[:, :, 3, 0:3] = 300.0                                        # splats outside the image
    f01, f10 = torch.tensor(f01, requires_grad=True), torch.tensor(f10, requires_grad=True)
    t = torch.tensor([[0.375], [0.75]])
    g0 = torch.tensor(rng.standard_normal((n, 2, h, w)).astype(np.float32))
    g1 = torch.tensor(rng.standard_normal((n, 2, h, w)).astype(np.float32))
    ft0, ft1 = ref.CFR_flow_t_align(dev, f01, f10, t[:, :, None, None])
    ((ft0 * g0).sum() + (ft1 * g1).sum()).backward()
    for k, v in dict(f01=f01, f10=f10, t=t, g0=g0, g1=g1, ft0=ft0, ft1=ft1, df01=f01.grad, df10=f10.grad).items():
        out[f"cfr_{k}"] = v.detach().numpy()
    np.savez_compressed(os.path.join(GOLD, "warp_grads.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))


if __name__ == "__main__":
    main()
