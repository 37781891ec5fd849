"""ORACLE TOOLING -- runs only in the build container (needs /root/reference).

Golden vectors for the visualisation return tuple of `DeMFInet.forward` (`DeMFInet.py:167-176`, FGAC maps `:454-495`): the
UNMODIFIED reference module with `visualization_flag=True`, eval mode, on the seeded synthetic inputs / weights of
`demfi_b200/synth.py`.  Writes tests/golden/c48x64_vis_b2.npz.

    python oracle/gen_golden_vis.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demfi_b200 import synth  # noqa: E402
from oracle.gen_golden import GOLD, load_reference  # noqa: E402

CFG = dict(h=48, w=64, n=1, batch=2, t=[0.25, 0.625])


def name_vis(res):
    """items 5 and 6 of the 7-tuple: blending_weights = [maps of FGAC(F1->F0), maps of FGAC(F0->F1), the same two again,
    [flow_01, flow_10]], each `maps` = [w, 1-w, source, ref_k, E_s, bolstered] (min-max normalised channel means);
    difference_maps = [d10, d01, d10, d01]"""
    bw, dm = res[5], res[6]
    assert len(res) == 7 and len(bw) == 5 and len(dm) == 4 and bw[2] is bw[0] and bw[3] is bw[1] and dm[2] is dm[0]
    d = {}
    for i in range(2):
        for j in range(6):
            d[f"bw{i}_{j}"] = bw[i][j]
        d[f"diff{i}"] = dm[i]
    d["flow_01"], d["flow_10"] = bw[4]
    return d


def main():
    ref_mod = load_reference()
    net = ref_mod.DeMFInet(synth.default_args(visualization_flag=True)).eval()
    net.load_state_dict(synth.make_state_dict(seed=0), strict=True)
    x = synth.make_frames(CFG["h"], CFG["w"], seed=0, batch=CFG["batch"])
    t = torch.tensor(CFG["t"], dtype=torch.float32).reshape(CFG["batch"], 1)
    with torch.no_grad():
        res = net(x, t, CFG["n"])
    out = {k: v.detach().numpy().astype(np.float32) for k, v in name_vis(res).items()}
    out["St_final0"] = res[1][0][2].numpy()
    np.savez_compressed(os.path.join(GOLD, "c48x64_vis_b2.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(v.min()), float(v.max()))


if __name__ == "__main__":
    main()
