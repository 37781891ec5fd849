"""ORACLE TOOLING -- runs only in the build container (needs /root/reference).

Gradient golden for the whole training forward: the UNMODIFIED reference `DeMFInet` (`is_training=True`, N_trn = 2) on seeded
synthetic inputs / weights, the L1 losses of main.py:404-440 (oracle/train_oracle.py), `total_loss.backward()` (main.py:443).
Stores, for each of the 260 parameters in state_dict order: [L2 norm, sum, projection on a seeded random direction] of its
gradient, a few gradients in full, the loss values and output statistics -> tests/golden/train_grads.npz.

    python oracle/gen_golden_train.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demfi_b200 import synth  # noqa: E402
from oracle import train_oracle as TO  # noqa: E402
from oracle.gen_golden import GOLD, load_reference  # noqa: E402

CFG = dict(h=32, w=32, batch=2, n=2, t=[0.25, 0.625])
FULL = ["Dec_last2.bias", "Booster_Module.flow_occ.conv2.bias", "FF_RDB_Module.UPNet.2.bias", "Refine_Module.enc1.bias",
        "FAC_FB_Module.shared_FGAC.w_gen_2.weight", "Booster_Module.GB.convq2.bias", "Dec_last2_2.weight"]


def case_tensors():
    x = synth.make_frames(CFG["h"], CFG["w"], seed=0, batch=CFG["batch"])
    t = torch.tensor(CFG["t"], dtype=torch.float32).reshape(-1, 1)
    gt = synth.make_frames(CFG["h"], CFG["w"], seed=7, batch=CFG["batch"])          # stand-ins for the sharp ground truths
    return x, t, [gt[:, :, 0].contiguous(), gt[:, :, 1].contiguous(), gt[:, :, 2].contiguous()]


def training_frames():
    """the [B, C, 9, H, W] sample of main.py:387-390 built from case_tensors(): 4 blurry inputs, frameT, then sharp 0, 1, -1, 2
    (only the first two of the last four are read by the loss)"""
    x, t, gts = case_tensors()
    return torch.cat([x, gts[2][:, :, None], gts[0][:, :, None], gts[1][:, :, None], x[:, :, 2:4]], 2), t


def summarise(named_grads):
    rows = []
    for i, (_, g) in enumerate(named_grads):
        v = g.detach().double().reshape(-1)
        r = torch.randn(v.numel(), generator=torch.Generator().manual_seed(1000 + i), dtype=torch.float64)
        rows.append([float(v.norm()), float(v.sum()), float((v * r).sum() / r.norm())])
    return np.asarray(rows, dtype=np.float64)


def main():
    ref_mod = load_reference()
    net = ref_mod.DeMFInet(synth.default_args()).train()
    net.load_state_dict(synth.make_state_dict(seed=0), strict=True)
    x, t, gts = case_tensors()
    res = net(x, t, CFG["n"], is_training=True)
    total, d1, d2 = TO.rec_losses(res[0], res[1], *gts)
    total.backward()
    named = [(n, p.grad) for n, p in net.named_parameters()]
    assert all(g is not None for n, g in named if "conv_source_k" not in n)
    out = {"summary": summarise([(n, g if g is not None else torch.zeros(1)) for n, g in named]),
           "names": np.asarray([n for n, _ in named]),
           "losses": np.asarray([float(total), float(d1), float(d2)]),
           "St_final_last": res[1][-1][2].detach().numpy(), "flow_last": res[2][-1].detach().numpy()}
    for n, g in named:
        if n in FULL:
            out["full:" + n] = g.numpy()
    # two full iterations of the loop body of train() (main.py:386-448) with the reference's optimizer settings (main.py:179-180)
    net2 = ref_mod.DeMFInet(synth.default_args()).train()
    net2.load_state_dict(synth.make_state_dict(seed=0), strict=True)
    opt = torch.optim.Adam(net2.parameters(), lr=1e-4, betas=(0.9, 0.999), weight_decay=0)
    frames, t = training_frames()
    step_losses = []
    for _ in range(2):
        input_frames, frameT, input_frames_GT = frames[:, :, :4], frames[:, :, 4], frames[:, :, -4:]
        opt.zero_grad()
        r = net2(input_frames, t, CFG["n"], is_training=True)
        tot, a, b = TO.rec_losses(r[0], r[1], input_frames_GT[:, :, 0], input_frames_GT[:, :, 1], frameT)
        tot.backward()
        opt.step()
        step_losses.append([float(tot), float(a), float(b)])
    out["step_losses"] = np.asarray(step_losses)
    out["params_after_2_steps"] = summarise([(n, p) for n, p in net2.named_parameters()])
    out["params_before"] = summarise([(n, p) for n, p in net.named_parameters()])
    np.savez_compressed(os.path.join(GOLD, "train_grads.npz"), **out)
    print("losses", out["losses"], "params", len(named), "grad norm range", out["summary"][:, 0].min(), out["summary"][:, 0].max())
    print("no-grad params:", [n for n, g in named if g is None])


if __name__ == "__main__":
    main()
