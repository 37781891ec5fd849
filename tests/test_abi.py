"""CPU: the C-ABI library loads, exports every symbol include/demfi_b200.h declares, the ctypes
mirror has the C layout, and the host-side weight repacking is correct (no GPU calls)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from demfi_b200 import _abi as A

HEADER = os.path.join(ROOT, "include", "demfi_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(demfi_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 18
    lib = C.CDLL(A.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported by the library"
        assert n in A.SYMBOLS, f"{n} has no ctypes prototype in demfi_b200/_abi.py"
    assert set(A.SYMBOLS) == set(names)
    assert A.lib().demfi_version() == 1


def test_ctypes_struct_layout_matches_c(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "demfi_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(demfi_src_t),sizeof(demfi_seg_t),sizeof(demfi_conv_t),offsetof(demfi_conv_t,src),'
                    'offsetof(demfi_conv_t,seg),offsetof(demfi_conv_t,wpack),offsetof(demfi_seg_t,ch0),sizeof(demfi_part_t),'
                    'offsetof(demfi_part_t,dst_c0));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(A.Src), C.sizeof(A.Seg), C.sizeof(A.Conv), A.Conv.src.offset, A.Conv.seg.offset, A.Conv.wpack.offset,
            A.Seg.ch0.offset, C.sizeof(A.Part), A.Part.dst_c0.offset]
    assert got == want, (got, want)


def _pack(kind, w, src_C, in_map, out_map, cout_pad):
    lib = A.lib()
    Co, Ci, KH, KW = w.shape
    sC = (A.i32 * len(src_C))(*src_C)
    n = lib.demfi_packed_weight_floats(kind, KH, KW, sC, len(src_C), cout_pad)
    out = np.full(n, np.nan, dtype=np.float32)
    kt = sum(src_C)
    A.check(lib.demfi_pack_weights(kind, w.ctypes.data, Co, Ci, KH, KW, (A.i32 * kt)(*in_map), sC, len(src_C),
                                   (A.i32 * cout_pad)(*out_map), cout_pad, out.ctypes.data), "pack")
    return out


def test_pack_weights_ffma_layout():
    rng = np.random.default_rng(0)
    w = rng.standard_normal((5, 7, 3, 3)).astype(np.float32)
    in_map = [6, 5, 4, 3, 2, 1, 0, -1]
    out_map = [4, 3, 2, 1, 0] + [-1] * 11
    p = _pack(A.CONV_FFMA, w, [8], in_map, out_map, 16).reshape(9, 8, 16)
    for tap in range(9):
        for k in range(8):
            for n in range(16):
                want = 0.0 if in_map[k] < 0 or out_map[n] < 0 else w[out_map[n], in_map[k], tap // 3, tap % 3]
                assert p[tap, k, n] == want


def test_retired_first_generation_kind_is_rejected():
    """DEMFI_CONV_TC (3xTF32, round 1) is no longer in the library: packing or describing with it is an error, not a fallback"""
    w = np.zeros((16, 32, 3, 3), dtype=np.float32)
    sC = (A.i32 * 1)(32)
    out = np.zeros(16, dtype=np.float32)
    rc = A.lib().demfi_pack_weights(A.CONV_TC, w.ctypes.data, 16, 32, 3, 3, (A.i32 * 32)(*range(32)), sC, 1, (A.i32 * 16)(*range(16)), 16,
                                    out.ctypes.data)
    assert rc != 0 and b"retired" in A.lib().demfi_last_error()
    assert _describe(64, [(0, 64, A.ACT_RELU, False, 0)], kind=A.CONV_TC) is None


def test_pack_weights_tc16_layout_fp16_hi_lo_swizzle64():
    """the layout conv_h3 / conv_s3 stream: per (N block, 32-channel chunk, tap) a [2N x 32] fp16 K-major tile, 64-byte swizzle
    (16-byte group g of row r stored at g ^ ((r >> 1) & 3)); rows 0..N-1 = fp16(w), rows N..2N-1 = fp16((w - hi) * 2048)"""
    rng = np.random.default_rng(2)
    w = (rng.standard_normal((70, 44, 3, 1)) * 0.3).astype(np.float32)
    src_C = [36, 8]                      # chunks: (src 0: 0..31), (src 0: 32..35 + pad), (src 1: 0..7 + pad)
    in_map = list(range(36)) + list(range(36, 44))
    cout_pad = 112                       # > 96 -> N blocks of 64 and 48
    out_map = list(range(70)) + [-1] * 42
    p = _pack(A.CONV_TC16, w, src_C, in_map, out_map, cout_pad)
    halves = p.view(np.float16)
    chunks = [(0, 0), (0, 32), (1, 0)]
    kbase = [0, 36]
    taps = 3
    base = 0
    for nb, N in ((0, 64), (1, 48)):
        for ci, (s_, c0) in enumerate(chunks):
            for tap in range(taps):
                tile = halves[base + (ci * taps + tap) * 2 * N * 32: base + (ci * taps + tap + 1) * 2 * N * 32].reshape(2 * N, 32)
                for n in range(N):
                    co = out_map[nb * 64 + n]
                    for k in range(32):
                        c = c0 + k
                        v = np.float32(0.0)
                        if c < src_C[s_] and co >= 0:
                            v = w[co, in_map[kbase[s_] + c], tap, 0]
                        hi = np.float16(v)
                        lo = np.float16((np.float32(v) - np.float32(hi)) * np.float32(2048.0))
                        for row, want in ((n, hi), (N + n, lo)):
                            col = (((k >> 3) ^ ((row >> 1) & 3)) << 3) + (k & 7)
                            assert tile[row, col] == want, (nb, ci, tap, n, k)
                        # the pair reconstructs the weight to ~2^-22 relative
                        assert abs(float(hi) + float(lo) / 2048.0 - float(v)) <= 2.0 ** -21 * max(abs(float(v)), 2.0 ** -14)
        base += len(chunks) * taps * 2 * N * 32
    assert base == halves.size


def test_pack_weights_device_validates_before_it_launches():
    """demfi_pack_weights_device (device -> device, the training step's packer): argument errors are reported without a GPU"""
    lib = A.lib()
    assert lib.demfi_pack_weights_device(A.CONV_TC16, None, 64, 64, 3, 3, 64, 64, None, None) != 0              # null pointers
    assert b"null" in lib.demfi_last_error()
    buf = (C.c_float * 4)()
    assert lib.demfi_pack_weights_device(A.CONV_TC16P, C.addressof(buf), 64, 64, 3, 3, 64, 64, C.addressof(buf), None) != 0
    assert b"DEMFI_CONV_TC16 only" in lib.demfi_last_error()
    assert lib.demfi_pack_weights_device(A.CONV_TC16, C.addressof(buf), 64, 64, 3, 3, 32, 64, C.addressof(buf), None) != 0   # src_c < Ci
    assert lib.demfi_pack_weights_device(A.CONV_TC16, C.addressof(buf), 64, 64, 3, 3, 64, 48, C.addressof(buf), None) != 0   # cout_pad < Co


def test_pack_weights_rejects_bad_maps():
    w = np.zeros((4, 4, 1, 1), dtype=np.float32)
    with pytest.raises(A.DemfiError):
        _pack(A.CONV_FFMA, w, [4], [0, 1, 2, 9], [0, 1, 2, 3] + [-1] * 12, 16)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc = A.lib().demfi_copy_channels(None, 4, None, 4, 1, 1, 0, None)
    assert rc != 0 and b"no CPU path" in A.lib().demfi_last_error() or rc != 0


def _describe(cout_pad, segs, k=(3, 3), srcC=(64,), stride=1, kind=A.CONV_TC16, hw=(64, 96)):
    d = A.Conv()
    d.N, d.H, d.W = 1, hw[0], hw[1]
    d.Hi, d.Wi = hw[0] * stride, hw[1] * stride
    d.KH, d.KW, d.stride = k[0], k[1], stride
    d.pad_h, d.pad_w = (k[0] // 2, k[1] // 2) if stride == 1 else (1, 1)
    d.nsrc, d.nseg, d.cout_pad, d.kind = len(srcC), len(segs), cout_pad, kind
    for i, c in enumerate(srcC):
        d.src[i].ptr, d.src[i].C, d.src[i].ld = 0x1000, c, c
    for i, (ch0, nch, act, res, fmt) in enumerate(segs):
        sg = d.seg[i]
        sg.dst, sg.dst_ld, sg.ch0, sg.nch, sg.act, sg.fmt = 0x2000 + 0x100000 * i, 256, ch0, nch, act, fmt
        if res:
            sg.res, sg.res_ld = 0x9000000, 256
    info = (A.i32 * 16)()
    if A.lib().demfi_conv_describe(C.byref(d), info) != 0:
        return None
    return list(info)


def test_conv_describe_reports_kernel_and_epilogue_plan():
    """host-only planning (no GPU): which kernel, TMA or generic epilogue, resident weights"""
    plain = _describe(64, [(0, 64, A.ACT_RELU, False, 0)])
    assert plain[0] == 3 and plain[1] == 1 and plain[2] == 1 and plain[9] == 18 and plain[8] == 18  # 64->64 3x3: resident, one segment of 18 stages
    pair = _describe(64, [(0, 64, A.ACT_RELU, False, A.SEG_DST_S16)], kind=A.CONV_TC16P)
    assert pair[0] == 3 and pair[2] == 1 and pair[3] >= 4 and pair[12] == 1   # CTA pair: half the filter bank per CTA -> 4+ halo buffers
    assert _describe(96, [(0, 96, A.ACT_RELU, False, 0)], kind=A.CONV_TC16P, srcC=(48,)) is None   # pairs: 32 or 64 output channels only
    heads = _describe(144, [(0, 64, A.ACT_TANH, True, 0), (64, 64, A.ACT_TANH, True, 0), (128, 8, A.ACT_NONE, True, 0)])
    assert heads[0] == 3 and heads[1] == 1 and heads[6] == 3          # three heads = three N blocks, each alike -> TMA epilogue
    lff = _describe(96, [(0, 96, A.ACT_NONE, True, A.SEG_DST_S16 | A.SEG_RES_S16)] * 2, k=(1, 1), srcC=(224,))
    assert lff[1] == 1 and lff[6] == 1                                  # one result, two destinations
    mixed = _describe(64, [(0, 32, A.ACT_RELU, False, 0), (32, 32, A.ACT_TANH, False, 0)])
    assert mixed[0] == 3 and mixed[1] == 1 and mixed[10] == 2           # two different segments inside one N block: planned per 32-channel box
    ragged = _describe(64, [(0, 16, A.ACT_RELU, False, 0), (16, 48, A.ACT_TANH, False, 0)])
    assert ragged[0] == 3 and ragged[1] == 0                            # ... unless they do not fall on box boundaries: generic epilogue
    push = _describe(96, [(0, 32, A.ACT_RELU, True, A.SEG_DST_S16), (32, 64, A.ACT_NONE, True, 0)], srcC=(32,))
    assert push[1] == 1 and push[10] == 3 and push[2] == 1              # dense-block push conv: S16 head + fp32 partial sums, weights resident
    zr = _describe(128, [(0, 64, A.ACT_SIGMOID, False, A.SEG_DST_S16), (64, 64, A.ACT_SIGMOID_MUL, True, A.SEG_DST_S16 | A.SEG_RES_S16)],
                   k=(1, 5), srcC=(64, 64), kind=A.CONV_TC16W)
    assert zr[0] == 3 and zr[1] == 1 and zr[6] == 1 and zr[11] == 128 and zr[10] == 4   # GRU z | r: ONE N block of 128, four boxes
    zr64 = _describe(128, [(0, 64, A.ACT_SIGMOID, False, 0), (64, 64, A.ACT_SIGMOID_MUL, True, 0)], k=(1, 5), srcC=(64, 64))
    assert zr64[6] == 2 and zr64[11] == 64                              # DEMFI_CONV_TC16 keeps blocks of 64 (layout shared with conv_h3)
    big = _describe(64, [(0, 64, A.ACT_TANH, False, 0)], k=(7, 7), srcC=(64, 64, 64))
    assert big[2] == 0 and big[4] >= 2 and big[5] >= 2                  # 2.4 MB of weights: ring of multi-stage groups
    enc = _describe(64, [(0, 64, A.ACT_RELU, False, 0)], k=(4, 4), srcC=(204,), stride=2)
    assert enc[0] == 2                                                  # stride 2 -> conv_h3
    assert _describe(64, [(0, 64, A.ACT_RELU, False, 0)], kind=A.CONV_FFMA)[0] == 0
