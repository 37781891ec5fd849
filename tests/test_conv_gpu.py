"""Parity of demfi_conv2d (CUDA-core and tcgen05 kernels) against torch conv2d in float64 on
identical inputs, for the conv shapes of DeMFInet.py: virtual concat, slices of
wider buffers, up-sampled sources, stride 2, pixel-shuffle store and every fused epilogue."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from demfi_b200 import _abi as A
from gpu_util import CONV_TC16_H3, DEV, from_nhwc, nhwc, run_conv, s16_decode, s16_encode

pytestmark = pytest.mark.gpu
KINDS = [pytest.param(A.CONV_FFMA, id="ffma"), pytest.param(CONV_TC16_H3, id="tc16-h3"),
         pytest.param(A.CONV_TC16, id="tc16-s3")]
TOL = 2e-5  # max-abs relative to max(1, max|ref|): fp32 conv noise level (ref self-noise 2-3e-5)


def rnd(*shape, seed=0, scale=1.0):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((g.standard_normal(shape) * scale).astype(np.float32))


def wb(co, ci, kh, kw, seed=1):
    std = math.sqrt(2.0 / ((ci + co) * kh * kw))
    return rnd(co, ci, kh, kw, seed=seed, scale=std), rnd(co, seed=seed + 1, scale=0.1)


def ref_conv(x, w, b, stride=1, pad=None):
    if pad is None:
        pad = (w.shape[2] // 2, w.shape[3] // 2)
    return F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=pad)


def check(got, want, what, tol=None):
    err = float((got.double() - want).abs().max())
    lim = (TOL if tol is None else tol) * max(1.0, float(want.abs().max()))
    print(f"{what}: max-abs err {err:.3e} (limit {lim:.3e}, max|ref| {float(want.abs().max()):.3f})")
    assert err <= lim, f"{what}: {err} > {lim}"


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n,h,w_,ci,co,k", [(2, 24, 40, 64, 64, (3, 3)), (1, 8, 16, 64, 64, (3, 3)), (1, 16, 48, 48, 96, (5, 5)),
                                             (1, 24, 24, 32, 32, (7, 7)), (3, 16, 32, 64, 64, (1, 1)), (1, 40, 56, 128, 64, (1, 5)),
                                             (1, 40, 56, 128, 64, (5, 1))])
def test_plain_relu(kind, n, h, w_, ci, co, k):
    x = rnd(n, ci, h, w_, seed=3)
    w, b = wb(co, ci, *k)
    xb, _ = nhwc(x)
    out = torch.zeros(n, h, w_, co, device=DEV)
    run_conv(w, b, [(xb, ci, 0)], (h, w_), kind, [dict(ch0=0, nch=co, dst=out, act=A.ACT_RELU)])
    check(from_nhwc(out, co), F.relu(ref_conv(x, w, b)), f"relu {ci}->{co} {k}")


@pytest.mark.parametrize("kind", KINDS)
def test_more_tiles_than_sms(kind):
    """persistent tile loop: 2*20*20 = 800 tiles > 148 CTAs, ragged right/bottom tiles (152 = 9.5*16, 156 = 19.5*8)"""
    n, h, w_ = 2, 156, 312
    x = rnd(n, 64, h, w_, seed=80)
    w, b = wb(64, 64, 3, 3)
    xb, _ = nhwc(x)
    res = rnd(n, 64, h, w_, seed=81)
    rb, _ = nhwc(res)
    out = torch.zeros(n, h, w_, 64, device=DEV)
    run_conv(w, b, [(xb, 64, 0)], (h, w_), kind, [dict(ch0=0, nch=64, dst=out, res=rb)])
    check(from_nhwc(out, 64), ref_conv(x, w, b) + res.double(), "many tiles + residual")


@pytest.mark.parametrize("kind", KINDS)
def test_wide_cout_two_n_blocks(kind):
    n, h, w_ = 1, 40, 48
    x = rnd(n, 96, h, w_, seed=82)
    w, b = wb(144, 96, 3, 3)
    xb, _ = nhwc(x)
    out = torch.zeros(n, h, w_, 144, device=DEV)
    run_conv(w, b, [(xb, 96, 0)], (h, w_), kind, [dict(ch0=0, nch=144, dst=out, act=A.ACT_TANH)])
    check(from_nhwc(out, 144), torch.tanh(ref_conv(x, w, b)), "cout 144 = 128 + 16")


@pytest.mark.parametrize("kind", KINDS)
def test_residual_into_slice(kind):
    # RDB-style: read channels [0,160) of a 224-wide buffer, write 32 channels into the same buffer
    n, h, w_ = 1, 16, 48
    x = rnd(n, 224, h, w_, seed=4)
    w, b = wb(32, 160, 3, 3)
    xb, _ = nhwc(x)
    run_conv(w, b, [(xb, 160, 0)], (h, w_), kind, [dict(ch0=0, nch=32, dst=xb, dst_c0=160, act=A.ACT_RELU)])
    got = from_nhwc(xb, 224)
    check(got[:, 160:192], F.relu(ref_conv(x[:, :160], w, b)), "rdb slice")
    assert torch.equal(got[:, :160], x[:, :160]) and torch.equal(got[:, 192:], x[:, 192:]), "neighbour channels touched"


@pytest.mark.parametrize("kind", KINDS)
def test_lff_dual_destination_residual(kind):
    n, h, w_ = 1, 16, 32
    x = rnd(n, 224, h, w_, seed=5)
    w, b = wb(96, 224, 1, 1)
    xb, _ = nhwc(x)
    o1 = torch.zeros(n, h, w_, 96, device=DEV)
    o2 = torch.zeros(n, h, w_, 1152, device=DEV)
    run_conv(w, b, [(xb, 224, 0)], (h, w_), kind,
             [dict(ch0=0, nch=96, dst=o1, res=xb), dict(ch0=0, nch=96, dst=o2, dst_c0=192, res=xb)])
    want = ref_conv(x, w, b) + x[:, :96].double()
    check(from_nhwc(o1, 96), want, "LFF dst1")
    check(from_nhwc(o2, 1152)[:, 192:288], want, "LFF dst2")


@pytest.mark.parametrize("kind", KINDS)
def test_pixel_shuffle_store(kind):
    n, h, w_ = 1, 16, 32
    x = rnd(n, 96, h, w_, seed=6)
    w, b = wb(256, 96, 3, 3)
    xb, _ = nhwc(x)
    out = torch.zeros(n, 2 * h, 2 * w_, 64, device=DEV)
    run_conv(w, b, [(xb, 96, 0)], (h, w_), kind, [dict(ch0=0, nch=256, dst=out, store=A.STORE_PIXEL_SHUFFLE2)],
             out_map=[(i % 64) * 4 + i // 64 for i in range(256)])
    check(from_nhwc(out, 64), F.pixel_shuffle(ref_conv(x, w, b), 2), "pixel shuffle")


@pytest.mark.parametrize("kind", KINDS)
def test_head_three_segments(kind):
    n, h, w_ = 1, 24, 32
    x = rnd(n, 64, h, w_, seed=7)
    w, b = wb(133, 64, 3, 3)
    xb, _ = nhwc(x)
    f01 = torch.zeros(2 * n, h, w_, 64, device=DEV)
    fo = torch.zeros(n, h, w_, 8, device=DEV)
    f1_view = f01[n:]
    run_conv(w, b, [(xb, 64, 0)], (h, w_), kind,
             [dict(ch0=0, nch=64, dst=f01, act=A.ACT_TANH), dict(ch0=64, nch=64, dst=f1_view, act=A.ACT_TANH),
              dict(ch0=128, nch=8, dst=fo)])
    r = ref_conv(x, w, b)
    check(from_nhwc(f01[:n], 64), torch.tanh(r[:, :64]), "head F0")
    check(from_nhwc(f01[n:], 64), torch.tanh(r[:, 64:128]), "head F1")
    check(from_nhwc(fo, 5), r[:, 128:133], "head flows/occ")


@pytest.mark.parametrize("kind", KINDS)
def test_three_sources_7x7_tanh(kind):
    n, h, w_ = 1, 24, 32
    xs = [rnd(n, 64, h, w_, seed=10 + i) for i in range(3)]
    w, b = wb(64, 192, 7, 7)
    bufs = [nhwc(x)[0] for x in xs]
    out = torch.zeros(n, h, w_, 64, device=DEV)
    run_conv(w, b, [(bb, 64, 0) for bb in bufs], (h, w_), kind, [dict(ch0=0, nch=64, dst=out, act=A.ACT_TANH)])
    check(from_nhwc(out, 64), torch.tanh(ref_conv(torch.cat(xs, 1), w, b)), "Ch_Reducer-like")


@pytest.mark.parametrize("kind", KINDS)
def test_gru_epilogues(kind):
    n, h, w_ = 1, 16, 32
    hh, xx = torch.tanh(rnd(n, 64, h, w_, seed=20)), F.relu(rnd(n, 64, h, w_, seed=21))
    wz, bz = wb(64, 128, 1, 5, seed=30)
    wr, br = wb(64, 128, 1, 5, seed=32)
    wq, bq = wb(64, 128, 1, 5, seed=34)
    hb, xb = nhwc(hh)[0], nhwc(xx)[0]
    Z = torch.zeros(n, h, w_, 64, device=DEV)
    RH = torch.zeros(n, h, w_, 64, device=DEV)
    H2 = torch.zeros(n, h, w_, 64, device=DEV)
    run_conv(torch.cat([wz, wr], 0), torch.cat([bz, br], 0), [(hb, 64, 0), (xb, 64, 0)], (h, w_), kind,
             [dict(ch0=0, nch=64, dst=Z, act=A.ACT_SIGMOID), dict(ch0=64, nch=64, dst=RH, act=A.ACT_SIGMOID_MUL, res=hb)])
    hx = torch.cat([hh, xx], 1)
    z = torch.sigmoid(ref_conv(hx, wz, bz))
    r = torch.sigmoid(ref_conv(hx, wr, br))
    check(from_nhwc(Z, 64), z, "gru z")
    check(from_nhwc(RH, 64), r * hh.double(), "gru r*h")
    run_conv(wq, bq, [(RH, 64, 0), (xb, 64, 0)], (h, w_), kind, [dict(ch0=0, nch=64, dst=H2, act=A.ACT_GRU, res=hb, res2=Z)])
    q = torch.tanh(ref_conv(torch.cat([(r * hh.double()).float(), xx], 1), wq, bq))
    check(from_nhwc(H2, 64), (1 - z) * hh.double() + z * q, "gru h'")


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("co,act", [(1, A.ACT_SIGMOID), (3, A.ACT_NONE), (9, A.ACT_NONE)])
def test_small_cout(kind, co, act):
    n, h, w_ = 2, 16, 16
    x = rnd(n, 64, h, w_, seed=40)
    w, b = wb(co, 64, 3, 3)
    xb, _ = nhwc(x)
    ld = (co + 3) // 4 * 4
    res = rnd(n, ld, h, w_, seed=41)
    rb, _ = nhwc(res)
    out = torch.zeros(n, h, w_, ld, device=DEV)
    run_conv(w, b, [(xb, 64, 0)], (h, w_), kind, [dict(ch0=0, nch=ld, dst=out, act=act, res=rb if act == A.ACT_NONE else None)])
    r = ref_conv(x, w, b)
    want = torch.sigmoid(r) if act == A.ACT_SIGMOID else r + res[:, :co].double()
    check(from_nhwc(out, co), want, f"cout={co}")


@pytest.mark.parametrize("kind", KINDS)
def test_two_sources_permuted_padded(kind):
    # D2's first conv: 36-wide assembled buffer (35 used, permuted) + 64-channel F_rec
    n, h, w_ = 1, 16, 32
    a3, fr = rnd(n, 36, h, w_, seed=50), rnd(n, 64, h, w_, seed=51)
    w, b = wb(64, 99, 3, 3)
    in_map = (list(range(9)) + [73, 74, 75, 76, 77, 80, 81, 78, 79, 82, 83, 84, 85, 86] + list(range(87, 99)) + [-1]
              + list(range(9, 73)))
    ab, fb = nhwc(a3)[0], nhwc(fr)[0]
    out = torch.zeros(n, h, w_, 64, device=DEV)
    run_conv(w, b, [(ab, 36, 0), (fb, 64, 0)], (h, w_), kind, [dict(ch0=0, nch=64, dst=out, act=A.ACT_RELU)], in_map=in_map)
    xin = torch.zeros(n, 99, h, w_)
    internal = torch.cat([a3, fr], 1)
    for kk, m in enumerate(in_map):
        if m >= 0:
            xin[:, m] = internal[:, kk]
    check(from_nhwc(out, 64), F.relu(ref_conv(xin, w, b)), "agg3 conv")


@pytest.mark.parametrize("kind", [pytest.param(A.CONV_FFMA, id="ffma"), pytest.param(A.CONV_TC16, id="tc16")])
@pytest.mark.parametrize("h,w_,ci,co", [(32, 48, 204, 64), (24, 40, 64, 128), (40, 72, 128, 256)])
def test_stride2_4x4(kind, h, w_, ci, co):
    """UNet encoders (DeMFInet.py:575-577): 4x4, stride 2, pad 1; ragged tiles (12 = 1.5*8, 20 = 1.25*16, 36 = 2.25*16)"""
    n = 1
    x = rnd(n, ci, h, w_, seed=60)
    w, b = wb(co, ci, 4, 4)
    xb, _ = nhwc(x)
    out = torch.zeros(n, h // 2, w_ // 2, co, device=DEV)
    run_conv(w, b, [(xb, ci, 0)], (h // 2, w_ // 2), kind, [dict(ch0=0, nch=co, dst=out, act=A.ACT_RELU)], stride=2, pad=(1, 1))
    check(from_nhwc(out, co), F.relu(ref_conv(x, w, b, 2, (1, 1))), f"enc 4x4 s2 {ci}->{co}")


def test_ffma_upsample_concat():
    n, h, w_ = 1, 16, 24
    lo, sk = rnd(n, 128, h // 2, w_ // 2, seed=61), rnd(n, 64, h, w_, seed=62)
    w, b = wb(64, 192, 3, 3)
    lb, sb = nhwc(lo)[0], nhwc(sk)[0]
    out = torch.zeros(n, h, w_, 64, device=DEV)
    run_conv(w, b, [(lb, 128, 1), (sb, 64, 0)], (h, w_), A.CONV_FFMA, [dict(ch0=0, nch=64, dst=out, act=A.ACT_RELU)])
    up = lo.repeat_interleave(2, 2).repeat_interleave(2, 3)
    check(from_nhwc(out, 64), F.relu(ref_conv(torch.cat([up, sk], 1), w, b)), "dec2 up+cat")


@pytest.mark.parametrize("kind", [pytest.param(A.CONV_FFMA, id="ffma"), pytest.param(CONV_TC16_H3, id="tc16-h3"),
                                  pytest.param(A.CONV_TC16, id="tc16-s3")])
def test_tiny_cin_7x7(kind):
    n, h, w_ = 1, 24, 24
    x = rnd(n, 5, h, w_, seed=63)
    w, b = wb(32, 5, 7, 7)
    xb, _ = nhwc(x, 8)
    out = torch.zeros(n, h, w_, 32, device=DEV)
    run_conv(w, b, [(xb, 8, 0)], (h, w_), kind, [dict(ch0=0, nch=32, dst=out, act=A.ACT_RELU)])
    check(from_nhwc(out, 32), F.relu(ref_conv(x, w, b)), "conv_delta1")


def test_tc16_operand_range():
    """3xFP16 split (conv_h3.cu): values far below the fp16 normal range (6e-5) and up to a few thousand keep
    fp32-level accuracy -- the low half is taken from the exact fp32 residual and scaled by 2048."""
    n, h, w_ = 1, 16, 32
    w, b = wb(64, 64, 3, 3)
    for scale in (1e-6, 1e-4, 1.0, 3e3):
        x = rnd(n, 64, h, w_, seed=72) * scale
        xb, _ = nhwc(x)
        out = torch.zeros(n, h, w_, 64, device=DEV)
        run_conv(w, torch.zeros(64), [(xb, 64, 0)], (h, w_), A.CONV_TC16, [dict(ch0=0, nch=64, dst=out)])
        want = ref_conv(x, w, torch.zeros(64))
        err = float((from_nhwc(out, 64).double() - want).abs().max())
        ref32 = float((F.conv2d(x, w, None, padding=1).double() - want).abs().max())
        print(f"scale {scale:g}: tc16 err {err:.3e}, torch fp32 conv err {ref32:.3e}, max|ref| {float(want.abs().max()):.3e}")
        assert err <= max(4.0 * ref32, 2e-5 * float(want.abs().max())), (scale, err, ref32)


def test_tc16_operand_and_weight_ranges():
    """3xFP16 over the whole representable range: activations uniform up to 6e4 (fp16 max 65504), weights scaled 1e-4 .. 1e2 --
    fp32-level accuracy relative to the result's magnitude; beyond the range the split saturates instead of producing inf / NaN"""
    n, h, w_ = 1, 16, 32
    g = np.random.Generator(np.random.PCG64(7))
    w0, _ = wb(64, 64, 3, 3)
    for xs, ws in ((6e4, 1.0), (6e4, 1e-4), (1.0, 1e2), (1e-3, 1e-4), (3e2, 1e2)):
        x = torch.from_numpy(g.uniform(-xs, xs, (n, 64, h, w_)).astype(np.float32))
        w = w0 * ws
        xb, _ = nhwc(x)
        for s16 in (False, True):   # fp32 source split by the converter warps / S16 source written by s16_encode
            out = torch.zeros(n, h, w_, 64, device=DEV)
            src = (s16_encode(xb), 64, 0, A.FMT_S16) if s16 else (xb, 64, 0)
            run_conv(w, torch.zeros(64), [src], (h, w_), A.CONV_TC16, [dict(ch0=0, nch=64, dst=out)])
            xv = s16_decode(src[0]).permute(0, 3, 1, 2).cpu() if s16 else x
            want = ref_conv(xv, w, torch.zeros(64))
            err = float((from_nhwc(out, 64).double() - want).abs().max())
            print(f"|x| <= {xs:g}, weights x {ws:g}, s16={s16}: err {err:.3e}, max|ref| {float(want.abs().max()):.3e}")
            assert err <= 2e-5 * float(want.abs().max()), (xs, ws, s16, err)
    # out of range: finite results (saturated operands), never inf / NaN
    x = torch.full((n, 64, h, w_), 1e6)
    x[:, ::2] = -3e5
    xb, _ = nhwc(x)
    out = torch.zeros(n, h, w_, 64, device=DEV)
    run_conv(w0, torch.zeros(64), [(xb, 64, 0)], (h, w_), A.CONV_TC16, [dict(ch0=0, nch=64, dst=out, fmt=A.SEG_DST_S16)])
    assert bool(torch.isfinite(s16_decode(out)).all())


def test_tc_accumulation_slope_is_what_the_compensation_assumes():
    """The tensor core accumulates with truncation: a GAIN error of about -0.27 * 2^-24 per chained MMA (profiles/r1_tc_numerics.md),
    which the epilogue compensates (tc_comp_milli = 270).  Re-measured here with the compensation off and one accumulation chain
    per tile: if a driver / firmware change moved the rounding behaviour by more than 20 %, ~100 chained layers would drift."""
    n, h, w_ = 1, 48, 64
    x = rnd(n, 64, h, w_, seed=3)
    w, _ = wb(64, 64, 3, 3, seed=1)
    xb, _ = nhwc(x)
    want = ref_conv(x, w, torch.zeros(64))
    slopes = {}
    try:
        for comp in (0, 270):
            A.set_option("tc_flush", 1000)     # one chain of 2 * 18 main MMAs per tile
            A.set_option("tc_comp_milli", comp)
            out = torch.zeros(n, h, w_, 64, device=DEV)
            run_conv(w, torch.zeros(64), [(xb, 64, 0)], (h, w_), A.CONV_TC16, [dict(ch0=0, nch=64, dst=out)])
            err = (from_nhwc(out, 64).double() - want).flatten()
            slopes[comp] = float((err * want.flatten()).sum() / (want.flatten() ** 2).sum())
    finally:
        A.set_option("tc_flush", 20)
        A.set_option("tc_comp_milli", 270)
    per_mma = -slopes[0] / 36 * 2 ** 24
    print(f"accumulation gain error per chained MMA: {per_mma:.3f} x 2^-24 (compensation assumes 0.27); residual slope with compensation {slopes[270]:.2e}")
    assert 0.8 * 0.27 <= per_mma <= 1.2 * 0.27, per_mma
    assert abs(slopes[270]) < 0.25 * abs(slopes[0])


# ---- CTA pairs (DEMFI_CONV_TC16P): cta_group::2 MMAs over two pixel tiles, each CTA holding half of the weight rows
PAIR_CASES = {
    # name: (n, h, w, srcC list, co, k, src S16, dst S16, residual, act)
    "resblock_s16_res": (2, 40, 56, [64], 64, (3, 3), True, True, True, A.ACT_NONE),
    "resblock_f32_relu": (1, 24, 40, [64], 64, (3, 3), False, False, False, A.ACT_RELU),
    "odd_tile_count": (1, 48, 24, [64], 64, (3, 3), True, True, False, A.ACT_RELU),      # 3 x 3 = 9 tiles: the last pair is half empty
    "single_tile": (1, 16, 8, [64], 64, (3, 3), True, False, False, A.ACT_RELU),
    "many_tiles": (2, 156, 312, [64], 64, (3, 3), True, True, True, A.ACT_RELU),         # 800 tiles > 74 pairs, ragged edges
    "n32_dense": (1, 40, 48, [96, 32], 32, (3, 3), True, True, False, A.ACT_RELU),
    "ring_7x7_three_sources": (1, 32, 40, [64, 64, 64], 64, (7, 7), False, True, False, A.ACT_TANH),
    "mixed_formats_1x5": (1, 24, 40, [64, 64], 64, (1, 5), None, True, False, A.ACT_RELU),   # first source S16, second fp32
    "ragged_cin_5": (1, 24, 40, [8], 32, (7, 7), False, True, False, A.ACT_RELU),
    "one_by_one": (3, 16, 32, [64], 64, (1, 1), True, False, False, A.ACT_NONE),
}


@pytest.mark.parametrize("case", list(PAIR_CASES))
def test_cta_pair_kernel(case):
    n, h, w_, srcC, co, k, s16, d16, res, act = PAIR_CASES[case]
    ci = sum(srcC)
    x = rnd(n, ci, h, w_, seed=11)
    w, b = wb(co, ci, *k, seed=13)
    srcs, vals, c0 = [], [], 0
    for i, c in enumerate(srcC):
        buf, _ = nhwc(x[:, c0:c0 + c])
        as16 = (s16 is True or (s16 is None and i == 0)) and c % 32 == 0
        if as16:
            buf = s16_encode(buf)
            vals.append(s16_decode(buf)[..., :c].permute(0, 3, 1, 2).cpu())
        else:
            vals.append(x[:, c0:c0 + c])
        srcs.append((buf, c, 0, A.FMT_S16 if as16 else A.FMT_F32))
        c0 += c
    if srcC == [8]:   # Mixer.conv_delta1: five real channels in an eight-channel slice
        w[:, 5:] = 0
    out = torch.zeros(n, h, w_, co, device=DEV)
    seg = dict(ch0=0, nch=co, dst=out, act=act, fmt=A.SEG_DST_S16 if d16 else 0)
    want = ref_conv(torch.cat(vals, 1), w, b)
    if res:
        r = rnd(n, co, h, w_, seed=17)
        rb, _ = nhwc(r)
        if d16:
            rb = s16_encode(rb)
            r = s16_decode(rb).permute(0, 3, 1, 2).cpu()
            seg["fmt"] |= A.SEG_RES_S16
        seg["res"] = rb
        want = want + r.double()
    want = {A.ACT_NONE: lambda v: v, A.ACT_RELU: F.relu, A.ACT_TANH: torch.tanh}[act](want)
    run_conv(w, b, srcs, (h, w_), A.CONV_TC16P, [seg])
    got = s16_decode(out) if d16 else out
    check(from_nhwc(got, co), want, f"cta pair: {case}")


def test_cta_pair_close_to_single_cta():
    """the pair kernel performs the same MMAs on the same operands and drains the same segments; only the two correction
    products are accumulated in columns of their own instead of on top of each other: last-bit differences"""
    n, h, w_ = 2, 72, 104
    x = rnd(n, 64, h, w_, seed=23)
    w, b = wb(64, 64, 3, 3, seed=29)
    xb = s16_encode(nhwc(x)[0])
    outs = []
    for kind in (A.CONV_TC16, A.CONV_TC16P):
        out = torch.zeros(n, h, w_, 64, device=DEV)
        run_conv(w, b, [(xb, 64, 0, A.FMT_S16)], (h, w_), kind, [dict(ch0=0, nch=64, dst=out, act=A.ACT_RELU, fmt=A.SEG_DST_S16)])
        outs.append(out.clone())
    a, b_ = s16_decode(outs[0]), s16_decode(outs[1])
    assert float((a - b_).abs().max()) <= 2.0 ** -20 * float(a.abs().max())


@pytest.mark.parametrize("grid", [2, 6, 148])
@pytest.mark.parametrize("h,w_,act", [(156, 312, A.ACT_NONE), (48, 24, A.ACT_RELU), (80, 40, A.ACT_NONE)])
def test_cta_pair_skip_operand_two_tiles_in_turn(grid, h, w_, act):
    """ResBlock conv2 on CTA pairs (S16 everywhere, one S16 skip operand): operand and result share one of TWO shared-memory
    tiles used in turn, the operand requested two tiles ahead (conv_s3: P.res_sep == 2).  Few CTAs = long tile sequences per
    CTA, odd tile counts = a half-empty last pair; the result must equal, bit for bit, the kernel with ONE operand tile
    (tc_diag & 32768), repeatedly, and the float64 convolution within the usual tolerance."""
    n = 2
    x, r = rnd(n, 64, h, w_, seed=61), rnd(n, 64, h, w_, seed=62)
    xb, rb = s16_encode(nhwc(x)[0]), s16_encode(nhwc(r)[0])
    xv = s16_decode(xb).permute(0, 3, 1, 2).cpu()
    rv = s16_decode(rb).permute(0, 3, 1, 2).cpu()
    w, b = wb(64, 64, 3, 3, seed=63)
    outs = {}
    try:
        A.set_option("tc_grid", grid)
        for diag in (32768, 0, 0, 0):
            A.set_option("tc_diag", diag)
            out = torch.zeros(n, h, w_, 64, device=DEV)
            run_conv(w, b, [(xb, 64, 0, A.FMT_S16)], (h, w_), A.CONV_TC16P,
                     [dict(ch0=0, nch=64, dst=out, res=rb, act=act, fmt=A.SEG_DST_S16 | A.SEG_RES_S16)])
            outs.setdefault(diag, []).append(out.clone())
    finally:
        A.set_option("tc_grid", 0)
        A.set_option("tc_diag", 0)
    want = ref_conv(xv, w, b) + rv.double()
    if act == A.ACT_RELU:
        want = F.relu(want)
    check(from_nhwc(s16_decode(outs[0][0]), 64), want, f"two operand tiles in turn, grid {grid}, {h}x{w_}")
    for o in outs[0]:
        assert torch.equal(o.view(torch.int32), outs[32768][0].view(torch.int32)), "differs from the one-operand-tile kernel"


@pytest.mark.parametrize("co,ci,k,src_c", [(64, 64, (3, 3), 64), (3, 64, (3, 3), 64), (96, 48, (5, 5), 48), (256, 96, (3, 3), 96),
                                           (32, 5, (7, 7), 8), (64, 128, (1, 5), 128), (133, 64, (3, 3), 64), (64, 204, (4, 4), 208)])
def test_device_weight_packing_equals_host_packing(co, ci, k, src_c):
    """demfi_pack_weights_device (the training step repacks every weight once per optimizer step, stream-ordered) writes the
    bytes demfi_pack_weights writes on the host: N blocking, hi / lo rows, 64-byte swizzle, zero padding of channels."""
    import ctypes as C
    lib = A.lib()
    w, _ = wb(co, ci, *k, seed=71)
    cout_pad = (co + 15) // 16 * 16
    sC = (A.i32 * 1)(src_c)
    n = lib.demfi_packed_weight_floats(A.CONV_TC16, k[0], k[1], sC, 1, cout_pad)
    host = np.empty(n, dtype=np.float32)
    wn = np.ascontiguousarray(w.numpy())
    A.check(lib.demfi_pack_weights(A.CONV_TC16, wn.ctypes.data, co, ci, k[0], k[1], (A.i32 * src_c)(*(list(range(ci)) + [-1] * (src_c - ci))),
                                   sC, 1, (A.i32 * cout_pad)(*(list(range(co)) + [-1] * (cout_pad - co))), cout_pad, host.ctypes.data), "pack")
    wd = w.to(DEV).contiguous()
    out = torch.full((n,), float("nan"), device=DEV)
    A.check(lib.demfi_pack_weights_device(A.CONV_TC16, wd.data_ptr(), co, ci, k[0], k[1], src_c, cout_pad, out.data_ptr(),
                                          torch.cuda.current_stream(DEV).cuda_stream), "pack_device")
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), host.view(np.uint32))


# ---- S16 ("split fp16") activation format: conv_s3 reads it without a conversion pass and writes it from its epilogue
def test_s16_roundtrip_host():
    x = rnd(2, 5, 7, 64, seed=5, scale=3.0)
    back = s16_decode(s16_encode(x))
    assert float((back - x).abs().max()) <= 2.0 ** -21 * float(x.abs().max())


@pytest.mark.parametrize("src_s16,dst_s16", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("k", [(3, 3), (1, 1), (1, 5)])
def test_s16_conv_relu(src_s16, dst_s16, k):
    n, h, w_, ci, co = 2, 40, 24, 64, 64
    x = rnd(n, ci, h, w_, seed=3)
    w, b = wb(co, ci, *k)
    xb, _ = nhwc(x)
    if src_s16:
        xb = s16_encode(xb)
    out = torch.zeros(n, h, w_, co, device=DEV)
    run_conv(w, b, [(xb, ci, 0, A.FMT_S16 if src_s16 else A.FMT_F32)], (h, w_), A.CONV_TC16,
             [dict(ch0=0, nch=co, dst=out, act=A.ACT_RELU, fmt=A.SEG_DST_S16 if dst_s16 else 0)])
    got = s16_decode(out) if dst_s16 else out
    check(from_nhwc(got, co), F.relu(ref_conv(x, w, b)), f"s16 src={src_s16} dst={dst_s16} {k}")


def test_s16_residual_two_sources_n32():
    """RDB-like: two S16 sources (96 + 32 channels of one trunk buffer), 32 output channels written S16 into the trunk, and an
    LFF-like 1x1 with an S16 residual and two S16 destinations"""
    n, h, w_ = 1, 32, 40
    x = rnd(n, 160, h, w_, seed=9)
    trunk, ld = nhwc(x, ld=224)
    trunk = s16_encode(trunk)
    w, b = wb(32, 128, 3, 3)
    out = trunk  # source: channels [0,128); destination: channels [128,160) of the same buffer
    run_conv(w, b, [(trunk, 128, 0, A.FMT_S16)], (h, w_), A.CONV_TC16,
             [dict(ch0=0, nch=32, dst=out, dst_c0=128, act=A.ACT_RELU, fmt=A.SEG_DST_S16)])
    got = s16_decode(trunk)[..., 128:160]
    check(from_nhwc(got, 32), F.relu(ref_conv(x[:, :128], w, b)), "s16 rdb conv into trunk slice")
    # LFF-like: 1x1 over the 160 channels, residual = channels [0,96), two destinations
    x2 = s16_decode(trunk)[..., :160].permute(0, 3, 1, 2).cpu()
    w1, b1 = wb(96, 160, 1, 1, seed=4)
    o1 = torch.zeros(n, h, w_, 96, device=DEV)
    o2 = torch.zeros(n, h, w_, 192, device=DEV)
    segs = [dict(ch0=0, nch=96, dst=o1, res=trunk, fmt=A.SEG_DST_S16 | A.SEG_RES_S16),
            dict(ch0=0, nch=96, dst=o2, dst_c0=96, res=trunk, fmt=A.SEG_DST_S16 | A.SEG_RES_S16)]
    run_conv(w1, b1, [(trunk, 160, 0, A.FMT_S16)], (h, w_), A.CONV_TC16, segs)
    want = ref_conv(x2, w1, b1) + x2[:, :96].double()
    check(from_nhwc(s16_decode(o1), 96), want, "s16 lff dst 1")
    check(from_nhwc(s16_decode(o2)[..., 96:192], 96), want, "s16 lff dst 2")


def test_gru_gates_on_cta_pairs_with_the_lean_activation_epilogue():
    """z, r and q as 64-channel convolutions on CTA pairs (the engine's form since round 2): sigmoid, sigmoid x h with an S16
    operand in place, and the GRU update with two S16 operand tiles -- LEAN == 2 kernels"""
    n, hh, w_ = 2, 40, 56
    hx = rnd(n, 128, hh, w_, seed=21, scale=0.7)
    hb, _ = nhwc(hx[:, :64]); xb, _ = nhwc(hx[:, 64:])
    hs, xs = s16_encode(hb), s16_encode(xb)
    h_val = s16_decode(hs).permute(0, 3, 1, 2).cpu().double()
    srcs = [(hs, 64, 0, A.FMT_S16), (xs, 64, 0, A.FMT_S16)]
    S = A.SEG_DST_S16
    wz, bz = wb(64, 128, 1, 5, seed=7)
    wr, br = wb(64, 128, 1, 5, seed=9)
    Z = torch.zeros(n, hh, w_, 64, device=DEV); RH = torch.zeros(n, hh, w_, 64, device=DEV)
    run_conv(wz, bz, srcs, (hh, w_), A.CONV_TC16P, [dict(ch0=0, nch=64, dst=Z, act=A.ACT_SIGMOID, fmt=S)])
    run_conv(wr, br, srcs, (hh, w_), A.CONV_TC16P, [dict(ch0=0, nch=64, dst=RH, act=A.ACT_SIGMOID_MUL, res=hs, fmt=S | A.SEG_RES_S16)])
    check(from_nhwc(s16_decode(Z), 64), torch.sigmoid(ref_conv(hx, wz, bz)), "pair lean2: z")
    check(from_nhwc(s16_decode(RH), 64), torch.sigmoid(ref_conv(hx, wr, br)) * h_val, "pair lean2: r*h")
    wq, bq = wb(64, 128, 5, 1, seed=8)
    H1 = torch.zeros(n, hh, w_, 64, device=DEV)
    run_conv(wq, bq, [(RH, 64, 0, A.FMT_S16), (xs, 64, 0, A.FMT_S16)], (hh, w_), A.CONV_TC16P,
             [dict(ch0=0, nch=64, dst=H1, act=A.ACT_GRU, res=hs, res2=Z, fmt=S | A.SEG_RES_S16 | A.SEG_RES2_S16)])
    rh_val = s16_decode(RH).permute(0, 3, 1, 2).cpu()
    z_val = s16_decode(Z).permute(0, 3, 1, 2).cpu().double()
    q = torch.tanh(ref_conv(torch.cat([rh_val, hx[:, 64:]], 1), wq, bq))
    check(from_nhwc(s16_decode(H1), 64), (1 - z_val) * h_val + z_val * q, "pair lean2: GRU update")
    # fp32 destination and fp32 operand in place (UNet dec3: tanh(conv + aligned feature)); S16 operand beside an fp32 destination
    # (the last FAC-FB ResBlock writes the fp32 SE buffer)
    xs64 = rnd(2, 64, 40, 56, seed=41)
    xb64 = s16_encode(nhwc(xs64)[0])
    xq64 = s16_decode(xb64).permute(0, 3, 1, 2).cpu()
    w3, b3 = wb(64, 64, 3, 3, seed=42)
    r32 = rnd(2, 64, 40, 56, seed=43)
    rb32, _ = nhwc(r32)
    o32 = torch.zeros(2, 40, 56, 64, device=DEV)
    run_conv(w3, b3, [(xb64, 64, 0, A.FMT_S16)], (40, 56), A.CONV_TC16P, [dict(ch0=0, nch=64, dst=o32, act=A.ACT_TANH, res=rb32)])
    check(from_nhwc(o32, 64), torch.tanh(ref_conv(xq64, w3, b3) + r32.double()), "pair lean2: fp32 dst, fp32 operand, tanh")
    rs16 = s16_encode(rb32)
    r16v = s16_decode(rs16).permute(0, 3, 1, 2).cpu().double()
    o32b = torch.zeros(2, 40, 56, 64, device=DEV)
    run_conv(w3, b3, [(xb64, 64, 0, A.FMT_S16)], (40, 56), A.CONV_TC16P, [dict(ch0=0, nch=64, dst=o32b, res=rs16, fmt=A.SEG_RES_S16)])
    check(from_nhwc(o32b, 64), ref_conv(xq64, w3, b3) + r16v, "pair lean2: fp32 dst, S16 operand in the second tile")
    o32c = torch.zeros(2, 40, 56, 64, device=DEV)
    run_conv(w3, b3, [(nhwc(xs64)[0], 64, 0)], (40, 56), A.CONV_TC16P, [dict(ch0=0, nch=64, dst=o32c, act=A.ACT_RELU)])
    check(from_nhwc(o32c, 64), F.relu(ref_conv(xs64, w3, b3)), "pair lean2: fp32 source and destination, relu")
    # tanh head with a long K loop (Ch_Reducer's form) on one CTA per tile and on pairs
    x3 = rnd(1, 192, 24, 40, seed=33)
    w7, b7 = wb(64, 192, 7, 7, seed=34)
    s3 = []
    for i in range(3):
        bf, _ = nhwc(x3[:, 64 * i:64 * i + 64])
        s3.append((s16_encode(bf), 64, 0, A.FMT_S16))
    xq = torch.cat([s16_decode(b_[0]).permute(0, 3, 1, 2).cpu() for b_ in s3], 1)
    for kind in (A.CONV_TC16, A.CONV_TC16P):
        o7 = torch.zeros(1, 24, 40, 64, device=DEV)
        run_conv(w7, b7, s3, (24, 40), kind, [dict(ch0=0, nch=64, dst=o7, act=A.ACT_TANH, fmt=S)])
        check(from_nhwc(s16_decode(o7), 64), torch.tanh(ref_conv(xq, w7, b7)), f"lean2 tanh 7x7 kind={kind}")


@pytest.mark.parametrize("zr_kind", [pytest.param(A.CONV_TC16, id="zr-2x64"), pytest.param(A.CONV_TC16W, id="zr-1x128")])
def test_s16_gru_epilogues(zr_kind):
    """zr conv: Z = sigmoid (S16 out), RH = sigmoid * h (S16 operand and out); q conv: (1 - z) h + z tanh(q) with S16 h and z.
    DEMFI_CONV_TC16W: z | r as ONE N block of 128 channels (N' = 256 MMAs) whose four 32-channel boxes are planned one by one."""
    n, hh, w_ = 1, 24, 40
    hx = rnd(n, 128, hh, w_, seed=21, scale=0.7)
    hb, _ = nhwc(hx[:, :64]); xb, _ = nhwc(hx[:, 64:])
    hs, xs = s16_encode(hb), s16_encode(xb)
    h_val = s16_decode(hs).permute(0, 3, 1, 2).cpu().double()
    wz, bz = wb(128, 128, 1, 5, seed=7)
    Z = torch.zeros(n, hh, w_, 64, device=DEV); RH = torch.zeros(n, hh, w_, 64, device=DEV)
    run_conv(wz, bz, [(hs, 64, 0, A.FMT_S16), (xs, 64, 0, A.FMT_S16)], (hh, w_), zr_kind,
             [dict(ch0=0, nch=64, dst=Z, act=A.ACT_SIGMOID, fmt=A.SEG_DST_S16),
              dict(ch0=64, nch=64, dst=RH, act=A.ACT_SIGMOID_MUL, res=hs, fmt=A.SEG_DST_S16 | A.SEG_RES_S16)])
    full = ref_conv(hx, wz, bz)
    z_ref, r_ref = torch.sigmoid(full[:, :64]), torch.sigmoid(full[:, 64:])
    check(from_nhwc(s16_decode(Z), 64), z_ref, "s16 gru z")
    check(from_nhwc(s16_decode(RH), 64), r_ref * h_val, "s16 gru r*h")
    wq, bq = wb(64, 128, 1, 5, seed=8)
    H1 = torch.zeros(n, hh, w_, 64, device=DEV)
    run_conv(wq, bq, [(RH, 64, 0, A.FMT_S16), (xs, 64, 0, A.FMT_S16)], (hh, w_), A.CONV_TC16,
             [dict(ch0=0, nch=64, dst=H1, act=A.ACT_GRU, res=hs, res2=Z, fmt=A.SEG_DST_S16 | A.SEG_RES_S16 | A.SEG_RES2_S16)])
    rh_val = s16_decode(RH).permute(0, 3, 1, 2).cpu()
    z_val = s16_decode(Z).permute(0, 3, 1, 2).cpu().double()
    q = torch.tanh(ref_conv(torch.cat([rh_val, hx[:, 64:]], 1), wq, bq))
    check(from_nhwc(s16_decode(H1), 64), (1 - z_val) * h_val + z_val * q, "s16 gru update")


@pytest.mark.parametrize("hw", [(32, 40), (37, 45)])
def test_dense_block_push_form_equals_pull_form(hw):
    """A residual dense block (RDB_Conv x 4, DeMFInet.py:256-287) the way the engine runs it (Engine._rdb_push_ops): every source
    convolved once with the weight slices of all later layers, fp32 partial sums P1..P3 accumulated in place through the epilogue
    operand path, the first 32-channel box of each launch finishing a layer (ReLU, S16 into the trunk) -- against the reference
    formulation (layer c convolves [x, g0..g_{c-1}]) in float64.  Exercises per-box epilogue plans: S16 + ReLU + fp32 operand in
    one box, fp32 read-modify-write in the others, of one N = 128 / 96 / 64 MMA tile."""
    n, (h, w_) = 1, hw
    x = rnd(n, 96, h, w_, seed=41)
    Ws, bs = zip(*[wb(32, 96 + 32 * c, 3, 3, seed=50 + c) for c in range(4)])
    trunk, _ = nhwc(torch.cat([x, torch.zeros(n, 128, h, w_)], 1), ld=224)
    trunk = s16_encode(trunk)
    xq = s16_decode(trunk)[..., :96].permute(0, 3, 1, 2).cpu()
    PS = torch.zeros(n, h, w_, 96, device=DEV)
    z32 = torch.zeros(32)
    g = lambda c: dict(dst=trunk, dst_c0=96 + 32 * c)
    cat = lambda layers, k0, k1: torch.cat([Ws[c][:, k0:k1] for c in layers], 0).contiguous()
    S = A.SEG_DST_S16
    run_conv(cat((0, 1, 2, 3), 0, 96), torch.cat([bs[0], z32, z32, z32]), [(trunk, 96, 0, A.FMT_S16)], (h, w_), A.CONV_TC16W,
             [dict(ch0=0, nch=32, act=A.ACT_RELU, fmt=S, **g(0)), dict(ch0=32, nch=96, dst=PS)])
    def src_of(c):
        v = trunk.view(-1)[96 + 32 * c:]
        class V:  # minimal buffer stand-in for run_conv: data_ptr() + the row stride of the trunk
            shape = trunk.shape
            def data_ptr(self_):
                return v.data_ptr()
        return [(V(), 32, 0, A.FMT_S16)]
    run_conv(cat((1, 2, 3), 96, 128), torch.cat([bs[1], z32, z32]), src_of(0), (h, w_), A.CONV_TC16,
             [dict(ch0=0, nch=32, act=A.ACT_RELU, fmt=S, res=PS, **g(1)), dict(ch0=32, nch=64, dst=PS, dst_c0=32, res=PS, res_c0=32)])
    run_conv(cat((2, 3), 128, 160), torch.cat([bs[2], z32]), src_of(1), (h, w_), A.CONV_TC16,
             [dict(ch0=0, nch=32, act=A.ACT_RELU, fmt=S, res=PS, res_c0=32, **g(2)), dict(ch0=32, nch=32, dst=PS, dst_c0=64, res=PS, res_c0=64)])
    run_conv(cat((3,), 160, 192), bs[3], src_of(2), (h, w_), A.CONV_TC16,
             [dict(ch0=0, nch=32, act=A.ACT_RELU, fmt=S, res=PS, res_c0=64, **g(3))])
    got = s16_decode(trunk).permute(0, 3, 1, 2).cpu()
    feats = [xq.double()]
    for c in range(4):
        feats.append(F.relu(ref_conv(torch.cat(feats, 1), Ws[c], bs[c])))
        check(got[:, 96 + 32 * c:128 + 32 * c], feats[-1], f"dense block growth slice {c} (push form)", tol=2e-5 * (c + 1))


@pytest.mark.parametrize("res_s16,dst_s16", [(True, False), (False, True)])
def test_s16_mixed_residual(res_s16, dst_s16):
    """skip connection stored in the other format than the result (FAC-FB: last ResBlock writes the fp32 SE buffer)"""
    n, h, w_ = 2, 24, 40
    x = rnd(n, 64, h, w_, seed=31)
    r = rnd(n, 64, h, w_, seed=32)
    w, b = wb(64, 64, 3, 3)
    xb, _ = nhwc(x)
    rb, _ = nhwc(r)
    if res_s16:
        rb = s16_encode(rb)
    r_val = (s16_decode(rb) if res_s16 else rb).permute(0, 3, 1, 2).cpu().double()
    out = torch.zeros(n, h, w_, 64, device=DEV)
    run_conv(w, b, [(xb, 64, 0)], (h, w_), A.CONV_TC16,
             [dict(ch0=0, nch=64, dst=out, res=rb, fmt=(A.SEG_DST_S16 if dst_s16 else 0) | (A.SEG_RES_S16 if res_s16 else 0))])
    got = s16_decode(out) if dst_s16 else out
    check(from_nhwc(got, 64), ref_conv(x, w, b) + r_val, f"mixed residual res_s16={res_s16} dst_s16={dst_s16}")


@pytest.mark.parametrize("case", ["resident_res_s16", "ring_7x7_f32", "gru_s16"])
def test_s3_repeated_launches_are_bitwise_identical(case):
    """race hunt: persistent tile loop (800 tiles on 148 CTAs), staging-tile reuse behind asynchronous TMA stores, in-place
    conversion, weight ring -- the kernel has a fixed accumulation order, so every launch must reproduce the first bit for bit"""
    import ctypes as C
    lib = A.lib()
    n, h, w_ = 2, 156, 312
    if case == "resident_res_s16":
        x, r = rnd(n, 64, h, w_, seed=50), rnd(n, 64, h, w_, seed=51)
        xb, rb = s16_encode(nhwc(x)[0]), s16_encode(nhwc(r)[0])
        w, b = wb(64, 64, 3, 3)
        out = torch.zeros(n, h, w_, 64, device=DEV)
        args = (w, b, [(xb, 64, 0, A.FMT_S16)], (h, w_), A.CONV_TC16,
                [dict(ch0=0, nch=64, dst=out, res=rb, fmt=A.SEG_DST_S16 | A.SEG_RES_S16)])
    elif case == "ring_7x7_f32":
        x = rnd(n, 96, h, w_, seed=52)
        xb = nhwc(x)[0]
        w, b = wb(64, 96, 7, 7)
        out = torch.zeros(n, h, w_, 64, device=DEV)
        args = (w, b, [(xb, 96, 0)], (h, w_), A.CONV_TC16, [dict(ch0=0, nch=64, dst=out, act=A.ACT_TANH)])
    else:
        hx = rnd(n, 128, h, w_, seed=53, scale=0.7)
        hs, xs = s16_encode(nhwc(hx[:, :64])[0]), s16_encode(nhwc(hx[:, 64:])[0])
        w, b = wb(128, 128, 1, 5)
        out = torch.zeros(n, h, w_, 128, device=DEV)
        args = (w, b, [(hs, 64, 0, A.FMT_S16), (xs, 64, 0, A.FMT_S16)], (h, w_), A.CONV_TC16,
                [dict(ch0=0, nch=64, dst=out, act=A.ACT_SIGMOID, fmt=A.SEG_DST_S16),
                 dict(ch0=64, nch=64, dst=out, dst_c0=64, act=A.ACT_SIGMOID_MUL, res=hs, fmt=A.SEG_DST_S16 | A.SEG_RES_S16)])
    run_conv(*args)
    ref = out.clone()
    for it in range(12):
        out.zero_()
        run_conv(*args)
        assert torch.equal(out.view(torch.int32), ref.view(torch.int32)), f"{case}: launch {it} differs from the first"
