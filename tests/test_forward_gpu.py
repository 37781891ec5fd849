"""End-to-end parity of the B200 forward (through the DeMFInet module -> C ABI) against
(a) the committed golden vectors produced by the unmodified reference and (b) the oracle,
including named intermediates, at sizes the oracle finishes in seconds.

Tolerance: BASELINE.json north_star -- 5e-4 max-abs vs the reference fp32 forward, PSNR within
0.01 dB.  The reference's own 1-vs-8-thread noise on these cases is 1-4e-5 (tests/golden/meta.json)."""
import numpy as np
import pytest
import torch

from conftest import case_inputs, load_golden
from demfi_b200 import synth
from demfi_b200.DeMFInet import DeMFInet
from demfi_b200.engine import Engine
from oracle import demfi_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
TOL = 5e-4


def psnr255(a, b):
    """utils.psnr on np.around(denorm255_np(.)) (utils.py:652-660, main.py:763-770)"""
    A_ = np.around(np.clip((a.numpy().astype(np.float64) + 1.0) / 2.0, 0, 1) * 255.0)
    B_ = np.around(np.clip((b.numpy().astype(np.float64) + 1.0) / 2.0, 0, 1) * 255.0)
    mse = np.mean((A_ - B_) ** 2)
    return float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


@pytest.fixture(scope="module")
def net(state_dict):
    m = DeMFInet(synth.default_args()).to(DEV).eval()
    m.load_state_dict(state_dict, strict=True)
    return m


@pytest.mark.parametrize("case", ["c32x32_n1_noise", "c64x96_n3", "c48x40_n2_b2", "c256x256_n1"])
def test_forward_matches_reference_golden(net, golden_meta, case):
    cfg = golden_meta["cases"][case]["cfg"]
    x, t = case_inputs(cfg)
    with torch.no_grad():
        res = net(x.to(DEV), t.to(DEV), cfg["n"])
    torch.cuda.synchronize()
    got = {k: v.cpu() for k, v in O.flatten_outputs(res).items()}
    gold = load_golden(case)
    worst = 0.0
    for k, g in gold.items():
        if k not in got:
            continue
        d = float((got[k] - torch.from_numpy(g)).abs().max())
        worst = max(worst, d)
        frac = float(((got[k] - torch.from_numpy(g)).abs() > TOL).float().mean())
        line = f"{case}:{k}: max-abs {d:.3e} frac>5e-4 {frac:.2e}"
        if k.startswith("S"):
            line += f" PSNR(ours,ref) {psnr255(got[k], torch.from_numpy(g)):.1f} dB"
        print(line)
    assert worst < TOL, f"{case}: max-abs {worst}"


def test_intermediates_match_oracle(state_dict, golden_meta):
    cfg = golden_meta["cases"]["c64x96_n3"]["cfg"]
    x, t = case_inputs(cfg)
    inter = {}
    O.forward(state_dict, x, t, cfg["n"], intermediates=inter)
    B, H, W = 1, cfg["h"], cfg["w"]
    eng = Engine(state_dict, B, H, W, DEV, arena=False)  # every intermediate keeps memory of its own to be read back below
    eng.forward(x.to(DEV), t.to(DEV), cfg["n"])
    torch.cuda.synchronize()
    v = eng.views
    mine = {
        "F0": v["F01"].frames(0, B), "F1": v["F01"].frames(B, B),
        "flow_01": v["FO"].ch(4, 2), "flow_10": v["FO"].ch(6, 2), "occ_logit_ff": v["FO"].ch(0, 1),
        "flow_t0": v["AGG1"].ch(192, 2), "flow_t1": v["AGG1"].ch(194, 2), "Ft": v["AGG1"].ch(128, 64),
        "enc0": v["SE"].frames(0, B).ch(0, 64), "enc1": v["SE"].frames(B, B).ch(0, 64),
        "fgac_sampled0": v["SMP"].frames(0, B), "fgac_sampled1": v["SMP"].frames(B, B),
        "fgac_w0": v["WL"].frames(0, B).ch(0, 1), "fgac_w1": v["WL"].frames(B, B).ch(0, 1),
        "aF0": v["AGG1"].ch(0, 64), "aF1": v["AGG1"].ch(64, 64),
        "rF0": v["DECIN"].frames(0, B), "rF1": v["DECIN"].frames(B, B), "rFt": v["DECIN"].frames(2 * B, B),
        "F_rec3": v["FR0"],  # rotates by 2 per iteration: after 3 iterations the state is back in FR0
    }
    worst = {}
    for k, view in mine.items():
        d = float((view.to_nchw().cpu() - inter[k]).abs().max())
        worst[k] = d
        print(f"intermediate {k}: max-abs {d:.3e} (max|ref| {float(inter[k].abs().max()):.2f})")
    bad = {k: d for k, d in worst.items() if d > TOL}
    assert not bad, bad


def test_arena_workspace_changes_nothing(state_dict, golden_meta):
    """The liveness-planned arena (buffers with disjoint lifetimes share memory) against one allocation per buffer: same
    outputs for a full call, a reuse_prefix call at another t and a call with another N_tst (the CFR splat's fp32 atomics make
    two runs agree to rounding, not bit for bit: 2e-5, as below)."""
    cfg = golden_meta["cases"]["c64x96_n3"]["cfg"]
    x, _ = case_inputs(cfg)
    xd = x.to(DEV)
    outs = []
    for arena in (True, False):
        eng = Engine(state_dict, 1, cfg["h"], cfg["w"], DEV, arena=arena)
        assert (eng._arena is not None) == arena
        o = [eng.forward(xd, torch.tensor([[0.25]], device=DEV), 3),
             eng.forward(xd, torch.tensor([[0.75]], device=DEV), 3, reuse_prefix=True),
             eng.forward(xd, torch.tensor([[0.5]], device=DEV), 5),
             eng.forward(xd, torch.tensor([[0.5]], device=DEV), 2, reuse_prefix=True, final_only=True)]
        torch.cuda.synchronize()
        outs.append([O.flatten_outputs(r) if r[1][0] is not None else {"St": r[1][-1][2], "flow": r[2][-1]} for r in o])
        if arena:
            small = eng.workspace_bytes()
        else:
            assert small < 0.75 * eng.workspace_bytes()
    for a_, b_ in zip(*outs):
        for k in a_:
            d = float((a_[k] - b_[k]).abs().max())
            assert d <= 2e-5, (k, d)


def test_prefix_reuse_and_final_only_change_nothing(net, golden_meta):
    """Skipping the t-independent prefix / the D2 decodes of non-final iterations must not change what is
    returned.  The CFR splat adds with fp32 atomics (as the reference's CUDA put_ does), so two runs agree to
    rounding, not bit-for-bit: compare at 2e-5."""
    cfg = golden_meta["cases"]["c64x96_n3"]["cfg"]
    x, _ = case_inputs(cfg)
    xd = x.to(DEV)
    ts = [torch.tensor([[v]], device=DEV) for v in (0.125, 0.5, 0.875)]
    with torch.no_grad():
        base = [net(xd, tt, 3) for tt in ts]
        net(xd, ts[0], 3)
        reused = [net(xd, tt, 3, reuse_prefix=True) for tt in ts[1:]]
        net.final_only = True
        try:
            fo = net(xd, ts[1], 3)
        finally:
            net.final_only = False
    torch.cuda.synchronize()
    close = lambda a, b: float((a - b).abs().max()) < 2e-5
    for b_, r_ in zip(base[1:], reused):
        for k, a in O.flatten_outputs(b_).items():
            assert close(a, O.flatten_outputs(r_)[k]), f"prefix reuse changed {k}"
    assert fo[1][0] is None and fo[1][1] is None
    for j in range(3):
        assert close(fo[1][2][j], base[1][1][2][j]), "final_only changed the last Sharps_final"
    assert close(fo[2][3], base[1][2][3]) and close(fo[3][3], base[1][3][3])


def test_cuda_core_only_engine_matches_too(state_dict, golden_meta):
    """The CUDA-core conv kernel alone (DEMFI_CONV_KIND=ffma) is an independent implementation of every conv:
    both kernel families must agree with the reference."""
    case = "c32x32_n1_noise"
    cfg = golden_meta["cases"][case]["cfg"]
    x, t = case_inputs(cfg)
    gold = load_golden(case)
    # ffma: CUDA cores only; tc16f32: the default kernels with every buffer in fp32 (no S16 storage, converter path everywhere);
    # auto (default, S16) is what every other test runs
    for kind in ("ffma", "tc16f32"):
        eng = Engine(state_dict, 1, cfg["h"], cfg["w"], DEV, conv_kind=kind)
        res = eng.forward(x.to(DEV), t.to(DEV), cfg["n"])
        torch.cuda.synchronize()
        got = O.flatten_outputs(res)
        worst = max(float((got[k].cpu() - torch.from_numpy(g)).abs().max()) for k, g in gold.items())
        print(f"conv_kind={kind}: worst max-abs vs golden {worst:.3e}")
        assert worst < TOL


def test_return_structure(net):
    x = synth.make_frames(32, 32, seed=1).to(DEV)
    t = torch.tensor([[0.5]], device=DEV)
    with torch.no_grad():
        r = net(x, t)  # num_update=None -> 1 (DeMFInet.py:126-128)
        assert len(r) == 5 and len(r[0]) == 3 and len(r[1]) == 1 and len(r[1][0]) == 3
        assert len(r[2]) == 2 and r[2][0].shape == (1, 4, 32, 32) and r[3][1].shape == (1, 1, 32, 32)
        assert r[4].shape == (1, 3, 32, 32)
        r7 = net(x, t, 2, is_training=True)
        assert len(r7) == 7 and len(r7[5]) == 4 and r7[5][0].shape == (1, 1, 32, 32) and len(r7[6][0]) == 2
        # the two extra items of the training tuple (DeMFInet.py:170-172) against the oracle: FGAC difference maps
        # (min-max normalised, :456-462) and the refined flows of Stage I
        want = O.forward(net.state_dict_cpu() if hasattr(net, "state_dict_cpu") else {k: v.cpu() for k, v in net.state_dict().items()},
                         x.cpu(), t.cpu(), 2, is_training=True)
        for got_m, want_m in zip(r7[5], want[5]):
            assert float((got_m.cpu() - want_m).abs().max()) < 2e-3
        for got_f, want_f in zip(r7[6][0], want[6][0]):
            assert float((got_f.cpu() - want_f).abs().max()) < TOL
    r_grad = net(x, t, 1, is_training=True)   # under grad mode: the differentiable forward (tests/test_train_net_gpu.py)
    assert len(r_grad) == 7 and r_grad[1][-1][2].requires_grad
    with pytest.raises(ValueError):
        with torch.no_grad():
            net(synth.make_frames(20, 24), t, 1)


def test_visualisation_tuple_matches_reference_golden(state_dict):
    """eval + args.visualization_flag: the 7-tuple of DeMFInet.py:174-176 against the reference's own output
    (tests/golden/c48x64_vis_b2.npz).  Measured 1-5e-6 on every map; the bound is the north-star 5e-4."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c48x64_vis_b2.npz"))
    m = DeMFInet(synth.default_args(visualization_flag=True)).to(DEV).eval()
    m.load_state_dict(state_dict, strict=True)
    x = synth.make_frames(48, 64, seed=0, batch=2).to(DEV)
    t = torch.tensor([[0.25], [0.625]], device=DEV)
    with torch.no_grad():
        res = m(x, t, 1)
    assert len(res) == 7 and len(res[5]) == 5 and len(res[6]) == 4 and res[5][2] is res[5][0] and res[5][3] is res[5][1]
    assert res[6][2] is res[6][0] and len(res[5][0]) == 6 and len(res[5][4]) == 2
    worst = 0.0
    for i in range(2):
        for j in range(6):
            assert res[5][i][j].shape == (2, 1, 48, 64)
            d = float((res[5][i][j].cpu() - torch.from_numpy(g[f"bw{i}_{j}"])).abs().max())
            print(f"visualisation map {i}.{j}: max-abs {d:.2e}")
            assert d < TOL, (i, j, d)
            worst = max(worst, d)
        assert float((res[6][i].cpu() - torch.from_numpy(g[f"diff{i}"])).abs().max()) < TOL
    assert float((res[5][4][0].cpu() - torch.from_numpy(g["flow_01"])).abs().max()) < TOL
    assert float((res[5][4][1].cpu() - torch.from_numpy(g["flow_10"])).abs().max()) < TOL
    assert float((res[1][0][2].cpu() - torch.from_numpy(g["St_final0"])).abs().max()) < TOL
    # is_training wins over the flag (DeMFInet.py:170-173)
    with torch.no_grad():
        r7 = m(x, t, 1, is_training=True)
    assert len(r7) == 7 and len(r7[6]) == 1 and len(r7[6][0]) == 2


# ---- BASELINE.json's full size (1280x720 padded to 1280x736, N_tst=3): size-independent properties
FULL_H, FULL_W = 736, 1280


def test_full_size_two_implementations_agree_on_every_conv_stack(state_dict):
    """At the north-star size the oracle takes minutes, so parity is shown through a property: the tensor-core path (3xFP16,
    S16 buffers) and the independent CUDA-core fp32 path agree on the same input on everything upstream of the
    discontinuous operators (FF_RDB features and flows, FAC-FB encoder), and downstream differ only in isolated flips."""
    from demfi_b200.engine import Engine
    x = synth.make_frames(FULL_H, FULL_W, seed=11).to(DEV)
    t = torch.tensor([[0.625]], device=DEV)
    outs = {}
    for kind in ("auto", "ffma"):
        eng = Engine(state_dict, 1, FULL_H, FULL_W, DEV, conv_kind=kind)
        res = eng.forward(x, t, 3)
        torch.cuda.synchronize()
        outs[kind] = {"F01": eng.views["F01"].to_nchw(), "FO": eng.views["FO"].to_nchw(), "SE": eng.views["SE"].to_nchw()[:, :64],
                      "St": res[1][-1][2].clone(), "flow": res[2][-1].clone()}
        del eng, res
        torch.cuda.empty_cache()
    for k in ("F01", "FO", "SE"):
        err = float((outs["auto"][k] - outs["ffma"][k]).abs().max())
        print(f"full size {k}: tensor-core vs CUDA-core max-abs {err:.3e}")
        assert err < 1.5e-4, (k, err)
    for k in ("St", "flow"):
        e = (outs["auto"][k] - outs["ffma"][k]).abs()
        frac = float((e > 5e-4).float().mean())
        print(f"full size {k}: max-abs {float(e.max()):.3e}, fraction > 5e-4 {frac:.2e}")
        assert frac < 5e-3, (k, frac)


def test_reuse_prefix_needs_a_prefix(state_dict):
    """a fresh engine has no t-independent stage to reuse: asking for it is an error, not a forward on zero-filled buffers"""
    from demfi_b200.engine import Engine
    eng = Engine(state_dict, 1, 32, 32, DEV)
    x = synth.make_frames(32, 32, seed=3).to(DEV)
    t = torch.tensor([[0.5]], device=DEV)
    with pytest.raises(RuntimeError, match="reuse_prefix"):
        eng.forward(x, t, 1, reuse_prefix=True)
    eng.forward(x, t, 1)
    eng.forward(x, t, 1, reuse_prefix=True)


def test_full_size_forward_is_bitwise_repeatable_and_batch_invariant(state_dict):
    """fixed accumulation order per output pixel: the same input gives the same bits run after run, and a sample gives the same
    bits alone or as part of a batch (the splat uses fp32 atomics, so this is checked on the conv stacks upstream of it and on
    a whole forward up to isolated splat-order effects)"""
    from demfi_b200.engine import Engine
    h, w = 96, 160
    xa, xb = synth.make_frames(h, w, seed=21), synth.make_frames(h, w, seed=22)
    t2 = torch.tensor([[0.25], [0.75]], device=DEV)
    e2 = Engine(state_dict, 2, h, w, DEV)
    e2.forward(torch.cat([xa, xb]).to(DEV), t2, 2)
    f2 = e2.views["F01"].to_nchw().clone()   # [2B, 64, H, W] = F0 of both samples, then F1 of both samples
    e1 = Engine(state_dict, 1, h, w, DEV)
    for i, (xi, ti) in enumerate(((xa, 0.25), (xb, 0.75))):
        e1.forward(xi.to(DEV), torch.tensor([[ti]], device=DEV), 2)
        f1 = e1.views["F01"].to_nchw()
        assert torch.equal(f1[0], f2[i]) and torch.equal(f1[1], f2[2 + i]), f"sample {i}: batched conv stack differs from the single run"
    eng = Engine(state_dict, 1, FULL_H, FULL_W, DEV)
    x = synth.make_frames(FULL_H, FULL_W, seed=12).to(DEV)
    t = torch.tensor([[0.375]], device=DEV)
    eng.forward(x, t, 1)
    a = eng.views["SE"].to_nchw().clone()
    eng.forward(x, t, 1)
    assert torch.equal(a, eng.views["SE"].to_nchw()), "full-size conv stack is not bitwise repeatable"


def test_cuda_graph_replay_matches_eager_bitwise(state_dict):
    """graph=True replays the whole call (~240 launches) as one CUDA graph: same kernels, same order -> same bits up to the
    atomics of the splat; meant for the launch-bound regime of small frames"""
    import time
    from demfi_b200.engine import Engine
    h, w = 128, 160
    eng = Engine(state_dict, 1, h, w, DEV)
    xs = [synth.make_frames(h, w, seed=s).to(DEV) for s in (31, 32)]
    ts = [torch.tensor([[tv]], device=DEV) for tv in (0.25, 0.75)]
    for x in xs:
        eager = [eng.forward(x, ts[0], 2), eng.forward(x, ts[1], 2, reuse_prefix=True)]
        graph = [eng.forward(x, ts[0], 2, graph=True), eng.forward(x, ts[1], 2, reuse_prefix=True, graph=True)]
        for e, g in zip(eager, graph):
            fe, fg = O.flatten_outputs(e), O.flatten_outputs(g)
            assert fe.keys() == fg.keys()
            worst = max(float((fe[k] - fg[k]).abs().max()) for k in fe)
            assert worst <= 2e-4, worst  # the splat's fp32 atomics land in a different order (as between two eager runs)
            assert torch.equal(fe["two_blurry"], fg["two_blurry"])
    def run(graph, n=20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            eng.forward(xs[i % 2], ts[0], 2, graph=graph)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    print(f"{h}x{w} N=2: eager {run(False):.2f} ms, CUDA graph {run(True):.2f} ms per forward")
