"""CPU: pins oracle/demfi_oracle.py to the golden vectors generated from the unmodified reference
(oracle/gen_golden.py) and its closed-form primitives to the torch library ops the reference calls."""
import numpy as np
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import case_inputs, load_golden
from oracle import demfi_oracle as O

# the reference's own 1-thread vs 8-thread noise on these cases is up to 3.8e-5 (tests/golden/meta.json)
ORACLE_TOL = 1e-4


@pytest.mark.parametrize("case", ["c32x32_n1_noise", "c64x96_n3", "c48x40_n2_b2", "c256x256_n1"])
def test_oracle_matches_reference_golden(state_dict, golden_meta, case):
    cfg = golden_meta["cases"][case]["cfg"]
    x, t = case_inputs(cfg)
    inter = {}
    res = O.forward(state_dict, x, t, cfg["n"], intermediates=inter)
    got = O.flatten_outputs(res)
    got.update({"F0_c8": inter["F0"][:, ::8], "aF0_c8": inter["aF0"][:, ::8], "aF1_c8": inter["aF1"][:, ::8],
                "F_rec0_c8": inter["F_rec0"][:, ::8], "flow_01": inter["flow_01"], "flow_10": inter["flow_10"],
                "occ_logit_ff": inter["occ_logit_ff"]})
    gold = load_golden(case)
    assert len(gold) >= 4
    for k, g in gold.items():
        d = float((got[k] - torch.from_numpy(g)).abs().max())
        assert d < ORACLE_TOL, f"{case}:{k} max-abs {d}"
    assert len(res[1]) == cfg["n"] and len(res[2]) == cfg["n"] + 1


def test_oracle_visualisation_tuple_matches_reference_golden(state_dict):
    """eval + args.visualization_flag (DeMFInet.py:167-176, FGAC maps :464-493): reference output in
    tests/golden/c48x64_vis_b2.npz (oracle/gen_golden_vis.py)"""
    import numpy as np
    from demfi_b200 import synth
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c48x64_vis_b2.npz"))
    x = synth.make_frames(48, 64, seed=0, batch=2)
    t = torch.tensor([[0.25], [0.625]])
    res = O.forward(state_dict, x, t, 1, visualization_flag=True)
    assert len(res) == 7 and len(res[5]) == 5 and len(res[6]) == 4 and res[5][2] is res[5][0] and res[6][3] is res[6][1]
    for i in range(2):
        for j in range(6):
            assert float((res[5][i][j] - torch.from_numpy(g[f"bw{i}_{j}"])).abs().max()) < ORACLE_TOL, (i, j)
        assert float((res[6][i] - torch.from_numpy(g[f"diff{i}"])).abs().max()) < ORACLE_TOL
    assert float((res[5][4][0] - torch.from_numpy(g["flow_01"])).abs().max()) < ORACLE_TOL
    assert float((res[5][4][1] - torch.from_numpy(g["flow_10"])).abs().max()) < ORACLE_TOL


def test_oracle_autograd_matches_reference_autograd_on_the_warps():
    """Gradients through the oracle's bwarp + Eq.(2) blend and bilinear gather equal autograd through the reference's own
    `bwarp` / `bilinear_sampler` (tests/golden/warp_grads.npz, oracle/gen_golden_grads.py): this is what the CUDA backward
    operators are checked against."""
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "warp_grads.npz"))
    for tag in ("c64", "c3"):
        T = lambda k, rg=False: torch.tensor(g[f"blend_{tag}_{k}"], requires_grad=rg)
        a, b, fl, occ = T("a", True), T("b", True), T("flow", True), T("occ", True)
        out = O.eq2_blend(a, fl[:, 0:2], b, fl[:, 2:4], occ, T("t").view(-1, 1, 1, 1))
        (out * T("gy")).sum().backward()
        assert float((out.detach() - T("out")).abs().max()) < 2e-6
        for name, got in (("da", a.grad), ("db", b.grad), ("dflow", fl.grad), ("docc", occ.grad)):
            want = T(name)
            assert float((got - want).abs().max()) < 2e-6 * max(1.0, float(want.abs().max())), (tag, name)
    T = lambda k, rg=False: torch.tensor(g[f"sample_{k}"], requires_grad=rg)
    refk, fl = T("refk", True), T("flow", True)
    H, W = refk.shape[-2:]
    px = ((2 * fl[:, 0] / (W - 1) - 1) + 1.0) / 2.0 * (W - 1)
    py = ((2 * fl[:, 1] / (H - 1) - 1) + 1.0) / 2.0 * (H - 1)
    out, _ = O.bilinear_gather(refk, px, py)
    (out * T("gy")).sum().backward()
    assert float((refk.grad - T("drefk")).abs().max()) < 2e-6
    assert float((fl.grad - T("dflow")).abs().max()) < 2e-6 * float(T("dflow").abs().max())
    # complementary flow reversal: put_(accumulate=True) splat, floor() without gradient, detached norm mask
    T = lambda k, rg=False: torch.tensor(g[f"cfr_{k}"], requires_grad=rg)
    f01, f10 = T("f01", True), T("f10", True)
    ft0, ft1 = O.cfr_flow_t_align(f01, f10, T("t").view(-1, 1, 1, 1))
    ((ft0 * T("g0")).sum() + (ft1 * T("g1")).sum()).backward()
    assert float((ft0.detach() - T("ft0")).abs().max()) < 2e-6 and float((ft1.detach() - T("ft1")).abs().max()) < 2e-6
    assert float((f01.grad - T("df01")).abs().max()) < 5e-6 and float((f10.grad - T("df10")).abs().max()) < 5e-6


def test_bilinear_gather_is_grid_sample_align_corners_true():
    g = torch.Generator().manual_seed(0)
    img = torch.randn(2, 5, 9, 13, generator=g)
    px = torch.rand(2, 9, 13, generator=g) * 18 - 3
    py = torch.rand(2, 9, 13, generator=g) * 14 - 3
    out, wsum = O.bilinear_gather(img, px, py)
    grid = torch.stack([2 * px / 12 - 1, 2 * py / 8 - 1], -1)
    want = F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    ones = F.grid_sample(torch.ones_like(img), grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    assert float((out - want).abs().max()) < 1e-5
    assert float((wsum - ones[:, :1]).abs().max()) < 1e-5


def test_bwarp_identity_and_border_mask():
    img = torch.arange(2 * 3 * 6 * 7, dtype=torch.float32).reshape(2, 3, 6, 7)
    assert float((O.bwarp(img, torch.zeros(2, 2, 6, 7)) - img).abs().max()) < 1e-4
    flo = torch.zeros(2, 2, 6, 7)
    flo[:, 0] = 0.5  # half a pixel to the right: the last column mixes an out-of-image corner -> masked to 0
    out = O.bwarp(img, flo)
    assert float(out[..., -1].abs().max()) == 0.0
    assert float((out[..., :-1] - 0.5 * (img[..., :-1] + img[..., 1:])).abs().max()) < 1e-4


def test_gaussian_splat_against_scalar_loops():
    """restates sample_one (DeMFInet.py:683-729) with python loops on a tiny case"""
    g = torch.Generator().manual_seed(1)
    H, W = 5, 6
    img = torch.randn(1, 2, H, W, generator=g)
    flo = torch.randn(1, 2, H, W, generator=g) * 2
    flo[0, :, 0, 0] = torch.tensor([1.0, -2.0])  # integer displacement
    acc, nrm = O.gaussian_splat(img, flo)
    ea = np.zeros((2, H, W))
    en = np.zeros((H, W))
    for r in range(H):
        for c in range(W):
            dx, dy = float(flo[0, 0, r, c]), float(flo[0, 1, r, c])
            for oy in (0, 1):
                for ox in (0, 1):
                    cy, cx = np.floor(dy) + oy, np.floor(dx) + ox
                    wgt = np.exp(-((dy - cy) ** 2 + (dx - cx) ** 2))
                    tr, tc = r + int(cy), c + int(cx)
                    if 0 <= tr < H and 0 <= tc < W:
                        ea[:, tr, tc] += img[0, :, r, c].numpy() * wgt
                        en[tr, tc] += wgt
    assert np.abs(acc[0].numpy() - ea).max() < 1e-5 and np.abs(nrm[0, 0].numpy() - en).max() < 1e-5


def test_space_to_depth_channel_order():
    x = torch.arange(1 * 2 * 4 * 6, dtype=torch.float32).reshape(1, 2, 4, 6)
    y = O.space_to_depth(x, 2)
    for c in range(2):
        for dy in range(2):
            for dx in range(2):
                assert torch.equal(y[0, c * 4 + dy * 2 + dx], x[0, c, dy::2, dx::2])
