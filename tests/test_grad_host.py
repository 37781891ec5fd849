"""Host-side logic of demfi_b200/grad.py that needs no GPU: weight packing for the forward and for the dx convolution (the
packing routine of the C ABI is host code), and the opt-in pack cache keyed on the parameters' version counters."""
import numpy as np
import torch

from demfi_b200 import grad

CPU = torch.device("cpu")


def test_forward_and_rotated_packs_equal_the_plain_expressions():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(40, 300, 3, 3, generator=g)          # Cin > 256: the dx convolution goes in two slices
    b = torch.randn(40, generator=g)
    wd, bd, cp = grad._pack_forward(w, b, 304, CPU)
    wd0, bd0, cp0 = grad._pack(w.numpy(), b.numpy(), 304, CPU)
    assert cp == cp0 == 48 and torch.equal(wd, wd0) and torch.equal(bd, bd0)
    wd1, bd1, _ = grad._pack_forward(w, None, 304, CPU)
    assert torch.equal(wd1, wd0) and float(bd1.abs().max()) == 0.0
    w_rot = w.flip(2, 3).transpose(0, 1).contiguous()     # [Cin, Cout, KH, KW]
    for c0, c1 in ((0, 256), (256, 300)):
        got = grad._pack_rotated(w, c0, c1, 40, CPU)
        want = grad._pack(w_rot[c0:c1].contiguous().numpy(), np.zeros(c1 - c0, dtype=np.float32), 40, CPU)
        assert got[2] == want[2] and torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    # the rotation is the one of a transposed convolution: w_t[ci, co, ky, kx] = W[co, ci, KH-1-ky, KW-1-kx]
    assert float(w_rot[7, 3, 0, 2]) == float(w[3, 7, 2, 0])


def test_pack_cache_follows_the_version_counter(monkeypatch):
    monkeypatch.setenv("DEMFI_GRAD_PACK_CACHE", "1")
    w = torch.nn.Parameter(torch.randn(16, 8, 3, 3))
    b = torch.nn.Parameter(torch.zeros(16))
    a1 = grad._pack_forward(w, b, 8, CPU)
    assert grad._pack_forward(w, b, 8, CPU)[0] is a1[0] and len(w._demfi_packed) == 1           # hit; kept ON the parameter
    r1 = grad._pack_rotated(w, 0, 8, 16, CPU)
    assert grad._pack_rotated(w, 0, 8, 16, CPU)[0] is r1[0]
    with torch.no_grad():
        w.mul_(2.0)                                                                              # any in-place update
    a2 = grad._pack_forward(w, b, 8, CPU)
    assert a2[0] is not a1[0] and not torch.equal(a2[0], a1[0])                                 # repacked with the new values
    with torch.no_grad():
        torch.autograd.graph.increment_version(w)                                                # what train.Adam does
    assert grad._pack_forward(w, b, 8, CPU)[0] is not a2[0]
    # a squeezed view of a Conv3d weight shares its base's counter
    w5 = torch.nn.Parameter(torch.randn(16, 8, 1, 3, 3))
    v = w5.squeeze(2)
    p1 = grad._pack_forward(v, None, 8, CPU)
    with torch.no_grad():
        w5.add_(1.0)
    assert grad._pack_forward(w5.squeeze(2), None, 8, CPU)[0] is not p1[0]
    assert grad._pack_forward(w5.squeeze(2), None, 8, CPU)[0] is grad._pack_forward(w5.squeeze(2), None, 8, CPU)[0]  # views hit through the base
    # another parameter at the same address (the first model of a process freed, the second allocated) never sees stale weights:
    # the packed tensors live on the parameter object and die with it
    w_new = torch.nn.Parameter(torch.randn(16, 8, 3, 3))
    assert not hasattr(w_new, "_demfi_packed")
    monkeypatch.setenv("DEMFI_GRAD_PACK_CACHE", "0")
    w_off = torch.nn.Parameter(torch.randn(16, 8, 3, 3))
    grad._pack_forward(w_off, b, 8, CPU)
    assert not hasattr(w_off, "_demfi_packed")                                                   # off: nothing is kept
