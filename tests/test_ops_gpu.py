"""Op-level parity (identical inputs, tight tolerance) of the HBM-bound kernels against the oracle's
closed-form restatements: bwarp + Eq.(2), CFR Gaussian splat, FGAC sampling / blend, input unpack."""
import numpy as np
import pytest
import torch

from demfi_b200 import _abi as A
from gpu_util import DEV, from_nhwc, nhwc, stream
from oracle import demfi_oracle as O

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((g.standard_normal(shape) * scale).astype(np.float32))


def flows_with_edge_cases(n, h, w, seed):
    """smooth-ish random flows of a few px plus: exact integers, exact zeros, far out-of-image values"""
    f = rnd(n, 4, h, w, seed=seed, scale=2.5)
    f[:, :, 0:2, :] = torch.round(f[:, :, 0:2, :])       # integer displacements (fwarp floor edge)
    f[:, :, 2, :] = 0.0                                    # identity
    f[:, :, 3, 0:4] = 1000.0                               # far outside
    f[:, :, 4, 0:4] = -1000.0
    f[:, 0, 5, :] = -torch.arange(w, dtype=torch.float32) - 0.0005  # lands in the 0.999 band at the left border
    return f


@pytest.mark.parametrize("C", [64, 3])
@pytest.mark.parametrize("n,h,w", [(1, 24, 40), (2, 16, 16)])
def test_bwarp_blend(C, n, h, w):
    a, b = rnd(n, C, h, w, seed=1), rnd(n, C, h, w, seed=2)
    fl = flows_with_edge_cases(n, h, w, 3)
    occ = rnd(n, 1, h, w, seed=4)
    t = torch.tensor([0.375, 0.75][:n])
    want = O.eq2_blend(a, fl[:, 0:2], b, fl[:, 2:4], occ, t.view(n, 1, 1, 1))
    ld = 64 if C == 64 else 36
    ab, _ = nhwc(a, ld)
    bb, _ = nhwc(b, ld)
    fb, _ = nhwc(fl, 8)
    ob, _ = nhwc(occ, 8)
    out = torch.zeros(n, h, w, ld, device=DEV)
    oo = torch.zeros(n, h, w, 4, device=DEV)
    A.check(A.lib().demfi_bwarp_blend(ab.data_ptr(), ld, bb.data_ptr(), ld, fb.data_ptr(), 8, ob.data_ptr(), 8,
                                      t.to(DEV).data_ptr(), n, h, w, C, out.data_ptr(), ld, oo.data_ptr(), 4, stream()), "bwarp")
    torch.cuda.synchronize()
    got = from_nhwc(out, C)
    diff = (got - want).abs()
    nflip = int((diff > 1e-3).sum())
    print(f"bwarp_blend C={C}: max-abs {float(diff.max()):.3e}, elements off by >1e-3 (mask flips): {nflip}")
    assert float(diff.max()) < 1e-5
    assert float((from_nhwc(oo, 1) - torch.sigmoid(occ)).abs().max()) < 1e-6


@pytest.mark.parametrize("n,h,w", [(1, 24, 40), (2, 16, 16), (1, 33, 51)])
def test_pwb_fused(n, h, w):
    """PWB of the boosting loop (DeMFInet.py:146-149) on the padded 8-channel row [S0' pad | S1' pad] -> [St occ | flows],
    against the oracle's Eq.(2) blend; integer flows, out-of-image targets and the 0.999 band included"""
    a, b = rnd(n, 3, h, w, seed=1), rnd(n, 3, h, w, seed=2)
    fl = flows_with_edge_cases(n, h, w, 3)
    occ = rnd(n, 1, h, w, seed=4)
    t = torch.tensor([0.375, 0.75][:n])
    want = O.eq2_blend(a, fl[:, 0:2], b, fl[:, 2:4], occ, t.view(n, 1, 1, 1))
    z1 = torch.zeros(n, 1, h, w)
    row, ld = nhwc(torch.cat([a, z1, b, z1, torch.full((n, 8, h, w), 7.0)], 1), 40)   # out slot pre-filled: every channel is written
    fo, _ = nhwc(torch.cat([fl, occ], 1), 8)
    img = row.view(-1)
    A.check(A.lib().demfi_pwb(row.data_ptr(), ld, fo.data_ptr(), 8, t.to(DEV).data_ptr(), n, h, w, row.data_ptr() + 4 * 8, ld, stream()), "pwb")
    torch.cuda.synchronize()
    got = from_nhwc(row, 16)
    assert float((got[:, 8:11] - want).abs().max()) < 1e-5
    assert float((got[:, 11:12] - torch.sigmoid(occ)).abs().max()) < 1e-6
    assert torch.equal(got[:, 12:16], fl)                       # the flows are copied bit for bit
    assert torch.equal(got[:, 0:8], torch.cat([a, z1, b, z1], 1))  # the inputs are untouched


@pytest.mark.parametrize("n,h,w", [(1, 24, 40), (2, 16, 24)])
def test_cfr_splat_finalize(n, h, w):
    fl = flows_with_edge_cases(n, h, w, 5)
    fl[:, :, 8:] += 0.37  # keep a region of generic non-integer flows
    t = torch.tensor([0.125, 0.625][:n])
    ft0, ft1 = O.cfr_flow_t_align(fl[:, 0:2], fl[:, 2:4], t.view(n, 1, 1, 1))
    fb, _ = nhwc(fl, 8)
    acc = torch.zeros(n, h, w, 8, device=DEV)
    out = torch.zeros(n, h, w, 4, device=DEV)
    td = t.to(DEV)
    A.check(A.lib().demfi_cfr_splat(fb.data_ptr(), 8, td.data_ptr(), n, h, w, acc.data_ptr(), stream()), "splat")
    A.check(A.lib().demfi_cfr_finalize(acc.data_ptr(), td.data_ptr(), n, h, w, out.data_ptr(), 4, stream()), "finalize")
    torch.cuda.synchronize()
    got = from_nhwc(out, 4)
    want = torch.cat([ft0, ft1], 1)
    diff = (got - want).abs()
    print(f"cfr: max-abs {float(diff.max()):.3e}; max|flow_t| {float(want.abs().max()):.2f}")
    # atomics reorder fp32 adds; values up to 1e3 px (the far-out test rows) -> relative tolerance
    assert float((diff / (1 + want.abs())).max()) < 2e-5


def test_fgac_sample_and_blend():
    n, h, w, C = 2, 24, 40, 64
    ref = rnd(n, C, h, w, seed=6)
    fl = rnd(n, 2, h, w, seed=7, scale=8.0) + 10.0  # absolute coordinates near the top-left corner
    fl[:, :, 0, :] = -5.0
    fl[:, :, 1, :] = torch.round(fl[:, :, 1, :])
    px = ((2 * fl[:, 0] / (w - 1) - 1) + 1.0) / 2.0 * (w - 1)
    py = ((2 * fl[:, 1] / (h - 1) - 1) + 1.0) / 2.0 * (h - 1)
    want, _ = O.bilinear_gather(ref, px, py)
    rb, _ = nhwc(ref)
    fb, _ = nhwc(fl, 8)
    out = torch.zeros(n, h, w, C, device=DEV)
    A.check(A.lib().demfi_fgac_sample(rb.data_ptr(), C, fb.data_ptr(), 8, n, h, w, C, out.data_ptr(), C, stream()), "fgac_sample")
    torch.cuda.synchronize()
    d = float((from_nhwc(out, C) - want).abs().max())
    print("fgac_sample max-abs", d)
    assert d < 1e-5
    wgt = torch.sigmoid(rnd(n, 1, h, w, seed=8))
    src, e = rnd(n, C, h, w, seed=9), rnd(n, C, h, w, seed=10)
    wb_, _ = nhwc(wgt, 4)
    se = torch.zeros(n, h, w, 128)
    se[..., :64] = src.permute(0, 2, 3, 1)
    se[..., 64:] = e.permute(0, 2, 3, 1)
    se = se.to(DEV)
    o2 = torch.zeros(n, h, w, 204, device=DEV)
    A.check(A.lib().demfi_fgac_blend(wb_.data_ptr(), 4, se.data_ptr(), 128, se.data_ptr() + 256, 128, n * h * w, C,
                                     o2.data_ptr() + 256, 204, stream()), "fgac_blend")
    torch.cuda.synchronize()
    got = o2[..., 64:128].permute(0, 3, 1, 2).cpu()
    assert float((got - (wgt * src + (1 - wgt) * e)).abs().max()) < 1e-6
    assert float(o2[..., :64].abs().max()) == 0 and float(o2[..., 128:].abs().max()) == 0


def test_pack_input_and_layout_converters():
    n, h, w = 2, 16, 24
    x = rnd(n, 3, 4, h, w, seed=11)
    xd = x.to(DEV)
    s2d = torch.zeros(n, h // 2, w // 2, 48, device=DEV)
    fa = torch.zeros(n, h, w, 32, device=DEV)
    fb = torch.zeros(n, h, w, 36, device=DEV)
    mean = torch.zeros(n, 3, h, w, device=DEV)
    A.check(A.lib().demfi_pack_input(xd.data_ptr(), n, h, w, s2d.data_ptr(), fa.data_ptr() + 4 * 9, 32,
                                     fb.data_ptr() + 4 * 23, 36, mean.data_ptr(), stream()), "pack_input")
    torch.cuda.synchronize()
    f12 = torch.cat([x[:, :, 0], x[:, :, 1], x[:, :, 2], x[:, :, 3]], 1)
    assert torch.equal(from_nhwc(s2d, 48), O.space_to_depth(f12, 2))
    assert torch.equal(fa[..., 9:21].permute(0, 3, 1, 2).cpu(), f12)
    assert torch.equal(fb[..., 23:35].permute(0, 3, 1, 2).cpu(), f12)
    assert float(fa[..., :9].abs().max()) == 0 and float(fb[..., 35:].abs().max()) == 0
    assert float((mean.cpu() - torch.mean(x[:, :, 0:2], dim=2)).abs().max()) < 1e-7
    # export / import / copy round trip
    src = rnd(n, 5, h, w, seed=12)
    buf = torch.zeros(n, h, w, 8, device=DEV)
    A.check(A.lib().demfi_import_nchw(src.to(DEV).data_ptr(), n, h, w, 5, buf.data_ptr(), 8, stream()), "import")
    dst = torch.zeros(n, h, w, 12, device=DEV)
    A.check(A.lib().demfi_copy_channels(buf.data_ptr(), 8, dst.data_ptr() + 4 * 3, 12, 5, n * h * w, A.ACT_NONE, stream()), "copy")
    out = torch.zeros(n, 5, h, w, device=DEV)
    A.check(A.lib().demfi_export_nchw(dst.data_ptr() + 4 * 3, 12, n, h, w, 5, A.ACT_NONE, out.data_ptr(), stream()), "export")
    sg = torch.zeros(n, 1, h, w, device=DEV)
    A.check(A.lib().demfi_export_nchw(dst.data_ptr() + 4 * 7, 12, n, h, w, 1, A.ACT_SIGMOID, sg.data_ptr(), stream()), "export")
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), src)
    assert float((sg.cpu() - torch.sigmoid(src[:, 4:5])).abs().max()) < 1e-6


def test_errors_are_loud():
    lib = A.lib()
    rc = lib.demfi_pack_input(None, 1, 15, 16, None, None, 0, None, 0, None, stream())
    assert rc != 0 and b"even" in lib.demfi_last_error()
    with pytest.raises(A.DemfiError):
        A.set_option("no_such_option", 1)


def test_upsample2x():
    n, h, w, C = 2, 6, 10, 128
    src = rnd(n, C, h, w, seed=13)
    sb, _ = nhwc(src, 132)
    dst = torch.zeros(n, 2 * h, 2 * w, 160, device=DEV)
    A.check(A.lib().demfi_upsample2x(sb.data_ptr(), 132, n, h, w, C, dst.data_ptr() + 4 * 16, 160, stream()), "upsample2x")
    torch.cuda.synchronize()
    want = src.repeat_interleave(2, 2).repeat_interleave(2, 3)
    assert torch.equal(dst[..., 16:16 + C].permute(0, 3, 1, 2).cpu(), want)
    assert float(dst[..., :16].abs().max()) == 0 and float(dst[..., 16 + C:].abs().max()) == 0


def test_gather_channels_assembles_a_row_in_one_pass():
    """ref_list / Agg3 assembly (DeMFInet.py:117-123, 151-155): five slices of four buffers into one 32-channel row"""
    import ctypes as C
    n, h, w = 2, 20, 28
    g = torch.Generator().manual_seed(3)
    srcs = [torch.randn(n, h, w, ld, generator=g).to(DEV) for ld in (4, 4, 8, 8)]
    dst = torch.full((n, h, w, 32), -7.0, device=DEV)
    spec = [(0, 0, 3, 0), (1, 0, 3, 3), (2, 0, 4, 21), (3, 0, 5, 25), (2, 4, 1, 30)]  # (buffer, src channel 0, nch, dst channel 0)
    arr = (A.Part * len(spec))()
    for i, (b, c0, nch, d0) in enumerate(spec):
        arr[i].src, arr[i].src_ld, arr[i].nch, arr[i].dst_c0 = srcs[b].data_ptr() + 4 * c0, srcs[b].shape[3], nch, d0
    A.check(A.lib().demfi_gather_channels(arr, len(spec), dst.data_ptr(), 32, n * h * w, stream()), "gather")
    torch.cuda.synchronize()
    want = torch.full((n, h, w, 32), -7.0, device=DEV)
    for b, c0, nch, d0 in spec:
        want[..., d0:d0 + nch] = srcs[b][..., c0:c0 + nch]
    assert torch.equal(dst, want)


def test_channel_absmean():
    """mean_c |a - b| per pixel (FGAC difference / visualisation maps, DeMFInet.py:456-491): channel slices of wider rows,
    with and without the second operand, a ragged pixel count"""
    n, h, w = 2, 7, 13
    a = rnd(n, 64, h, w, seed=31)
    b = rnd(n, 64, h, w, seed=32)
    ab, _ = nhwc(a, 132)
    bb, _ = nhwc(b, 68)
    for second in (True, False):
        out = torch.full((n * h * w + 5,), -3.0, device=DEV)
        A.check(A.lib().demfi_channel_absmean(ab.data_ptr(), 132, bb.data_ptr() if second else None, 68 if second else 0,
                                              n * h * w, 64, out.data_ptr(), stream()), "channel_absmean")
        torch.cuda.synchronize()
        want = ((a.double() - b.double()) if second else a.double()).abs().mean(1).reshape(-1)
        assert float((out[:n * h * w].cpu().double() - want).abs().max()) < 1e-6
        assert float((out[n * h * w:] + 3.0).abs().max()) == 0   # nothing written past npix
    A.lib().demfi_channel_absmean(ab.data_ptr(), 132, None, 0, 10, 6, out.data_ptr(), stream())  # C % 4 != 0
    assert b"channel_absmean" in A.lib().demfi_last_error()
