"""Backward of a convolution layer through the C ABI (training row f-2, first slice): demfi_b200.grad.conv2d against torch
autograd on the same layer in float64.  Forward and dx run on the tcgen05 kernel (fp32 parity), dW / db on the CUDA-core
wgrad kernel (fp32 atomics): tolerances relative to the gradient's own scale."""
import pytest
import torch
import torch.nn.functional as F

from demfi_b200 import grad

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
def test_conv2d_layer_gradients(ci, co, k, act, n, h, w):
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
