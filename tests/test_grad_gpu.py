"""Backward of a convolution layer through the C ABI (training row f-2, first slice): demfi_b200.grad.conv2d against torch
autograd on the same layer in float64.  Forward and dx run on the tcgen05 kernel (fp32 parity), dW / db on the mma.sync 3xTF32
wgrad kernel and on the CUDA-core one (both with fp32 atomics): tolerances relative to the gradient's own scale."""
import pytest
import torch
import torch.nn.functional as F

from demfi_b200 import _abi as A, grad

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
ACT = {"none": lambda v: v, "relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}


@pytest.mark.parametrize("ci,co,k,act,n,h,w", [
    (64, 64, (3, 3), "relu", 2, 40, 56),       # ResBlock conv
    (5, 32, (7, 7), "relu", 1, 32, 48),        # Mixer.conv_delta1: ragged Cin
    (64, 3, (3, 3), "none", 3, 24, 40),        # Dec_last2: ragged Cout
    (128, 64, (1, 5), "tanh", 1, 32, 64),      # GRU q-like, asymmetric kernel
    (96, 133, (3, 3), "none", 1, 24, 24),      # more output channels than one 64-wide block, ragged
    (224, 96, (1, 1), "sigmoid", 1, 20, 36),   # 1x1 (LFF-like)
])
@pytest.mark.parametrize("wgrad_kind", [pytest.param(1, id="wgrad-mma-3xtf32"), pytest.param(0, id="wgrad-cuda-core")])
def test_conv2d_layer_gradients(ci, co, k, act, n, h, w, wgrad_kind):
    A.set_option("wgrad_kind", wgrad_kind)
    try:
        _layer_gradients(ci, co, k, act, n, h, w)
    finally:
        A.set_option("wgrad_kind", 1)


def _layer_gradients(ci, co, k, act, n, h, w):
    g = torch.Generator().manual_seed(ci * 1000 + co)
    x = torch.randn(n, ci, h, w, generator=g)
    wt = torch.randn(co, ci, *k, generator=g) / (ci * k[0] * k[1]) ** 0.5
    b = torch.randn(co, generator=g) * 0.1
    gy = torch.randn(n, co, h, w, generator=g)
    # float64 reference
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, wt, b))
    yr = ACT[act](F.conv2d(xr, wr, br, padding=(k[0] // 2, k[1] // 2)))
    yr.backward(gy.double())
    # ours
    xd, wd, bd = (t.to(DEV).requires_grad_(True) for t in (x, wt, b))
    y = grad.conv2d(xd, wd, bd, act)
    y.backward(gy.to(DEV))
    torch.cuda.synchronize()

    def rel(a, r):
        return float((a.double().cpu() - r).abs().max() / r.abs().max())
    print(f"{ci}->{co} {k} {act}: y {rel(y.detach(), yr.detach()):.2e} dx {rel(xd.grad, xr.grad):.2e} "
          f"dW {rel(wd.grad, wr.grad):.2e} db {rel(bd.grad, br.grad):.2e}")
    assert rel(y.detach(), yr.detach()) < 1e-5
    assert rel(xd.grad, xr.grad) < 1e-5
    assert rel(wd.grad, wr.grad) < 2e-5
    assert rel(bd.grad, br.grad) < 2e-5


def test_conv2d_without_bias_and_partial_grads():
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 32, 16, 24, generator=g).to(DEV)
    wt = (torch.randn(32, 32, 3, 3, generator=g) / 17.0).to(DEV).requires_grad_(True)
    y = grad.conv2d(x, wt, None, "relu")          # x does not require grad: no dx conv is run
    y.sum().backward()
    ref_w = wt.detach().double().cpu().requires_grad_(True)
    torch.relu(F.conv2d(x.double().cpu(), ref_w, None, padding=1)).sum().backward()
    assert float((wt.grad.double().cpu() - ref_w.grad).abs().max() / ref_w.grad.abs().max()) < 2e-5


def test_conv2d_refuses_cpu_and_even_kernels():
    from demfi_b200._abi import DemfiError
    with pytest.raises(DemfiError):
        grad.conv2d(torch.zeros(1, 8, 8, 8), torch.zeros(8, 8, 3, 3))
    with pytest.raises(ValueError):
        grad.conv2d(torch.zeros(1, 8, 8, 8, device=DEV), torch.zeros(8, 8, 4, 4, device=DEV))


# ---------------------------------------------------------------------------------------------- warps
import os

import numpy as np

from demfi_b200 import _abi as A
from gpu_util import from_nhwc, nhwc, stream

WARP_GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "warp_grads.npz"))


def _close(got, want, what, tol=5e-6):
    d = float((got - want).abs().max())
    scale = max(1.0, float(want.abs().max()))
    print(f"{what}: max-abs {d:.2e} (scale {scale:.1f})")
    assert d < tol * scale, (what, d)


@pytest.mark.parametrize("tag,C,ld", [("c64", 64, 64), ("c64", 64, 72), ("c3", 3, 4), ("c3", 3, 36)])
def test_bwarp_blend_backward_matches_reference_autograd(tag, C, ld):
    """gradients of bwarp + Eq.(2) w.r.t. both sources, the four flow channels and the occlusion logit against autograd
    through the reference's own functions (tests/golden/warp_grads.npz): integer displacements, out-of-image targets and the
    0.999 validity band included; vector (C % 4 == 0) and scalar (3-channel pixel warp) paths, dense and strided rows"""
    T = lambda k: torch.from_numpy(WARP_GOLD[f"blend_{tag}_{k}"])
    a, b, fl, occ, gy = T("a"), T("b"), T("flow"), T("occ"), T("gy")
    n, _, h, w = a.shape
    ab, _ = nhwc(a, ld)
    bb, _ = nhwc(b, ld)
    gb, _ = nhwc(gy, ld)
    fb, _ = nhwc(fl, 8)
    ob, _ = nhwc(occ, 4)
    da = torch.zeros(n, h, w, ld, device=DEV)
    db = torch.zeros(n, h, w, ld, device=DEV)
    dfl = torch.full((n, h, w, 4), 7.0, device=DEV)
    doc = torch.full((n, h, w, 4), 7.0, device=DEV)
    tv = T("t").reshape(-1).to(DEV)
    A.check(A.lib().demfi_bwarp_blend_backward(ab.data_ptr(), ld, bb.data_ptr(), ld, fb.data_ptr(), 8, ob.data_ptr(), 4, tv.data_ptr(),
                                               gb.data_ptr(), ld, n, h, w, C, da.data_ptr(), ld, db.data_ptr(), ld, dfl.data_ptr(), 4,
                                               doc.data_ptr(), 4, stream()), "bwarp_blend_backward")
    torch.cuda.synchronize()
    _close(from_nhwc(da, C), T("da"), "da")
    _close(from_nhwc(db, C), T("db"), "db")
    _close(from_nhwc(dfl, 4), T("dflow"), "dflow", 1e-5)
    _close(from_nhwc(doc, 1), T("docc"), "docc", 1e-5)
    assert float(da[..., C:].abs().max() if ld > C else 0.0) == 0.0      # nothing scattered into the padding channels
    assert float((doc[..., 1:] - 7.0).abs().max()) == 0.0


def test_fgac_sample_backward_matches_reference_autograd():
    T = lambda k: torch.from_numpy(WARP_GOLD[f"sample_{k}"])
    refk, fl, gy = T("refk"), T("flow"), T("gy")
    n, C, h, w = refk.shape
    rb, _ = nhwc(refk)
    gb, _ = nhwc(gy)
    fb, _ = nhwc(fl, 8)
    dr = torch.zeros(n, h, w, C, device=DEV)
    dfl = torch.zeros(n, h, w, 2, device=DEV)
    A.check(A.lib().demfi_fgac_sample_backward(rb.data_ptr(), C, fb.data_ptr(), 8, gb.data_ptr(), C, n, h, w, C, dr.data_ptr(), C,
                                               dfl.data_ptr(), 2, stream()), "fgac_sample_backward")
    torch.cuda.synchronize()
    _close(from_nhwc(dr, C), T("drefk"), "drefk")
    _close(from_nhwc(dfl, 2), T("dflow"), "dflow", 1e-5)
    # value gradients only (no flow gradient requested), accumulated on top of what is there
    A.check(A.lib().demfi_fgac_sample_backward(rb.data_ptr(), C, fb.data_ptr(), 8, gb.data_ptr(), C, n, h, w, C, dr.data_ptr(), C,
                                               None, 0, stream()), "fgac_sample_backward")
    torch.cuda.synchronize()
    _close(from_nhwc(dr, C), 2 * T("drefk"), "drefk accumulated", 1e-5)


def test_cfr_backward_matches_reference_autograd():
    """gradient of the complementary flow reversal w.r.t. flow_01 / flow_10 against autograd through the reference's
    CFR_flow_t_align (put_ with accumulate, floor without gradient): displacements landing exactly on integers, splats leaving
    the image, two different t in the batch"""
    T = lambda k: torch.from_numpy(WARP_GOLD[f"cfr_{k}"])
    f01, f10 = T("f01"), T("f10")
    n, _, h, w = f01.shape
    fo, _ = nhwc(torch.cat([f01, f10], 1), 8)
    gb, _ = nhwc(torch.cat([T("g0"), T("g1")], 1), 4)
    tv = T("t").reshape(-1).to(DEV)
    acc = torch.zeros(n, h, w, 8, device=DEV)
    out = torch.zeros(n, h, w, 4, device=DEV)
    lib = A.lib()
    A.check(lib.demfi_cfr_splat(fo.data_ptr(), 8, tv.data_ptr(), n, h, w, acc.data_ptr(), stream()), "splat")
    A.check(lib.demfi_cfr_finalize(acc.data_ptr(), tv.data_ptr(), n, h, w, out.data_ptr(), 4, stream()), "finalize")
    gacc = torch.empty(n, h, w, 8, device=DEV)
    dfo = torch.full((n, h, w, 8), 7.0, device=DEV)
    A.check(lib.demfi_cfr_backward(fo.data_ptr(), 8, tv.data_ptr(), acc.data_ptr(), gb.data_ptr(), 4, n, h, w, gacc.data_ptr(),
                                   dfo.data_ptr(), 8, stream()), "cfr_backward")
    torch.cuda.synchronize()
    _close(from_nhwc(out, 4), torch.cat([T("ft0"), T("ft1")], 1), "forward flow_t0|flow_t1", 2e-6)
    _close(from_nhwc(dfo, 4), torch.cat([T("df01"), T("df10")], 1), "d flow_01|flow_10", 5e-6)
    assert float((dfo[..., 4:] - 7.0).abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------- losses, optimiser, small ops
def test_rec_losses_and_their_gradients():
    """Eq.(9)-(10) of main.py:404-440 against the oracle restatement (torch L1Loss) and its autograd"""
    from demfi_b200 import train
    from oracle import train_oracle as TO
    g = torch.Generator().manual_seed(11)
    B, H, W, N = 2, 24, 40, 3
    gts = [torch.randn(B, 3, H, W, generator=g) for _ in range(3)]
    prime = [(gt + 0.1 * torch.randn(B, 3, H, W, generator=g)).requires_grad_(True) for gt in gts]
    final = [[(gt + 0.05 * torch.randn(B, 3, H, W, generator=g)).requires_grad_(True) for gt in gts] for _ in range(N)]
    final[1][2].data[0, 0, 0, :5] = gts[2][0, 0, 0, :5]          # exact zeros of the difference: sign(0) = 0
    want = TO.rec_losses(prime, final, *gts, rec_D1_lambda=1.0, rec_D2_lambda=0.5)
    want[0].backward()
    d = lambda t: t.detach().to(DEV)
    got = train.rec_losses([d(t) for t in prime], [[d(t) for t in tri] for tri in final], *[d(t) for t in gts],
                           rec_D1_lambda=1.0, rec_D2_lambda=0.5, with_grads=True)
    for a, b in zip(got[:3], want):
        assert abs(a - float(b.detach())) < 1e-6 * max(1.0, abs(float(b.detach()))), (a, float(b.detach()))
    for i in range(3):
        assert float((got[3][i].cpu() - prime[i].grad).abs().max()) < 1e-9
        for j in range(N):
            assert float((got[4][j][i].cpu() - final[j][i].grad).abs().max()) < 1e-9
    # bit-reproducible sums
    again = train.rec_losses([d(t) for t in prime], [[d(t) for t in tri] for tri in final], *[d(t) for t in gts], 1.0, 0.5)
    assert again[0] == got[0]
    with pytest.raises(A.DemfiError):
        train.rec_losses(prime, final, *gts)


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_adam_matches_torch_optim(wd):
    from demfi_b200 import train
    g = torch.Generator().manual_seed(5)
    shapes = [(64, 64, 3, 3), (64,), (3, 64, 1, 3, 3), (1000, 7)]
    ref = [torch.randn(*s, generator=g).to(DEV).requires_grad_(True) for s in shapes]
    mine = [r.detach().clone().requires_grad_(True) for r in ref]
    opt_ref = torch.optim.Adam(ref, lr=1e-2, betas=(0.9, 0.999), weight_decay=wd)
    opt = train.Adam(mine, lr=1e-2, betas=(0.9, 0.999), weight_decay=wd)
    for step in range(4):
        for r, m in zip(ref, mine):
            gr = torch.randn(r.shape, generator=g).to(DEV) * (0.1 + step)
            r.grad, m.grad = gr.clone(), gr.clone()
        opt_ref.step()
        opt.step()
    torch.cuda.synchronize()
    for r, m in zip(ref, mine):
        assert float((r.detach() - m.detach()).abs().max()) < 2e-6
    assert opt.state[0]["step"] == 4 and opt.param_groups[0]["lr"] == 1e-2
    assert all(m._version >= 4 for m in mine)     # in-place updates are visible to torch (the inference engine repacks on it)
    opt.zero_grad()
    assert all(p.grad is None for p in mine)


def test_fgac_blend_and_upsample_backward():
    g = torch.Generator().manual_seed(9)
    n, C, h, w = 2, 64, 12, 20
    wgt = torch.sigmoid(torch.randn(n, 1, h, w, generator=g)).double().requires_grad_(True)
    src = torch.randn(n, C, h, w, generator=g).double().requires_grad_(True)
    e = torch.randn(n, C, h, w, generator=g).double().requires_grad_(True)
    gy = torch.randn(n, C, h, w, generator=g)
    ((wgt * src + (1 - wgt) * e) * gy.double()).sum().backward()
    wb, _ = nhwc(wgt.detach().float(), 4)
    sb, _ = nhwc(src.detach().float(), 128)
    eb, _ = nhwc(e.detach().float(), 64)
    gb, _ = nhwc(gy, 64)
    dw = torch.zeros(n, h, w, 4, device=DEV)
    ds = torch.zeros(n, h, w, 64, device=DEV)
    de = torch.zeros(n, h, w, 72, device=DEV)
    A.check(A.lib().demfi_fgac_blend_backward(wb.data_ptr(), 4, sb.data_ptr(), 128, eb.data_ptr(), 64, gb.data_ptr(), 64, n * h * w, C,
                                              dw.data_ptr(), 4, ds.data_ptr(), 64, de.data_ptr(), 72, stream()), "fgac_blend_backward")
    torch.cuda.synchronize()
    _close(from_nhwc(dw, 1).double(), wgt.grad, "dw", 2e-6)
    _close(from_nhwc(ds, C).double(), src.grad, "dsrc", 1e-6)
    _close(from_nhwc(de, C).double(), e.grad, "de", 1e-6)
    # nearest x2 up-sampling
    x = torch.randn(n, C, 5, 7, generator=g).double().requires_grad_(True)
    gu = torch.randn(n, C, 10, 14, generator=g)
    (torch.nn.functional.interpolate(x, scale_factor=2, mode="nearest") * gu.double()).sum().backward()
    gub, _ = nhwc(gu, 68)
    gx = torch.zeros(n, 5, 7, 64, device=DEV)
    A.check(A.lib().demfi_upsample2x_backward(gub.data_ptr(), 68, n, 5, 7, C, gx.data_ptr(), 64, stream()), "upsample2x_backward")
    torch.cuda.synchronize()
    _close(from_nhwc(gx, C).double(), x.grad, "upsample2x backward", 1e-6)


# ---------------------------------------------------------------------------------------------- a multi-layer training slice
def test_three_training_steps_of_the_d1_decoder_follow_torch():
    """forward (13 tcgen05 convs with residual adds) -> L1 losses with fused gradients -> backward (dx on the tcgen05 kernel,
    dW/db on the wgrad kernel) -> demfi Adam, three steps on the module's own D1 parameters, against the same loop written
    with F.conv2d / autograd / torch.optim.Adam in float64: per-step loss and first-step gradients"""
    import copy
    from demfi_b200 import synth, train
    from demfi_b200.DeMFInet import DeMFInet
    torch.backends.cudnn.allow_tf32 = False
    model = DeMFInet(synth.default_args()).to(DEV)
    model.load_state_dict(synth.make_state_dict(0))
    names = ["Dec_first", "Dec_last1", "Dec_last2"] + [f"Decoder_res.{i}.conv{j}" for i in range(5) for j in (1, 2)]
    mods = {n: model.get_submodule(n) for n in names}
    ref = {n: (copy.deepcopy(m.weight.detach()).double().cpu().squeeze(2).requires_grad_(True),
               copy.deepcopy(m.bias.detach()).double().cpu().requires_grad_(True)) for n, m in mods.items()}
    g = torch.Generator().manual_seed(3)
    B, H, W = 1, 24, 40
    feats = torch.tanh(torch.randn(3 * B, 64, H, W, generator=g))
    gts = [torch.randn(B, 3, H, W, generator=g) * 0.5 for _ in range(3)]

    def ref_forward(x):
        c = lambda n, t: torch.nn.functional.conv2d(t, ref[n][0], ref[n][1], padding=1)
        x = torch.relu(c("Dec_first", x))
        for i in range(5):
            x = x + c(f"Decoder_res.{i}.conv2", torch.relu(c(f"Decoder_res.{i}.conv1", x)))
        return c("Dec_last2", torch.relu(c("Dec_last1", x)))

    params = [p for m in mods.values() for p in (m.weight, m.bias)]
    opt = train.Adam(params, lr=1e-4)
    opt_ref = torch.optim.Adam([t for pair in ref.values() for t in pair], lr=1e-4)
    fd, gd = feats.to(DEV), [t.to(DEV) for t in gts]
    for step in range(3):
        # ours
        opt.zero_grad()
        out = grad.decoder_d1(model, fd)
        sharps = [out[0:B], out[B:2 * B], out[2 * B:3 * B]]
        total, d1, _, g_prime, _ = train.rec_losses(sharps, [], *gd, with_grads=True)
        torch.autograd.backward(sharps, g_prime)
        # reference
        opt_ref.zero_grad()
        o = ref_forward(feats.double())
        loss = sum(torch.nn.functional.l1_loss(gts[i].double(), o[i * B:(i + 1) * B]) for i in range(3)) / 3
        loss.backward()
        print(f"step {step}: loss ours {total:.8f} torch {float(loss.detach()):.8f}")
        assert abs(total - float(loss.detach())) < 5e-6 * max(1.0, abs(float(loss.detach())))
        if step == 0:
            for n, m in mods.items():
                gw, gr = m.weight.grad.squeeze(2).double().cpu(), ref[n][0].grad
                assert float((gw - gr).abs().max()) < 5e-5 * float(gr.abs().max()), n
                assert float((m.bias.grad.double().cpu() - ref[n][1].grad).abs().max()) < 5e-5 * float(ref[n][1].grad.abs().max()), n
        opt.step()
        opt_ref.step()
    torch.cuda.synchronize()
    for n, m in mods.items():   # Adam normalises by |g|: parameters whose gradient is ~0 may step differently, so compare in bulk
        d = (m.weight.detach().squeeze(2).double().cpu() - ref[n][0].detach()).abs()
        assert float(d.mean()) < 2e-6, (n, float(d.mean()))
