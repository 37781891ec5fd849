"""Wiring of the differentiable training forward (demfi_b200/train_net.py) checked WITHOUT a GPU: the four kernel-backed
operator families are replaced by torch implementations (below, from the oracle's closed forms, themselves pinned to autograd
through the reference's functions in tests/test_oracle.py) and the whole graph -- 260 parameters, N_trn = 2, two samples with
different t -- is compared with `total_loss.backward()` through the UNMODIFIED reference module (tests/golden/train_grads.npz,
oracle/gen_golden_train.py).  On a B200 the same graph runs with `KernelOps`, whose operators are checked one by one in
tests/test_grad_gpu.py."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from demfi_b200 import synth, train_net
from demfi_b200.DeMFInet import DeMFInet
from oracle import demfi_oracle as O
from oracle import train_oracle as TO
from oracle.gen_golden_train import CFG, FULL, case_tensors, summarise

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_grads.npz"))
ACT = {"none": lambda v: v, "relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}


class TorchOps:
    """test-only stand-ins for the kernel-backed operators (never importable from the package)"""

    @staticmethod
    def conv2d(x, w, b, act="none", stride=1):
        pad = (w.shape[2] // 2, w.shape[3] // 2) if stride == 1 else (1, 1)
        return ACT[act](F.conv2d(x, w, b, stride=stride, padding=pad))

    @staticmethod
    def bwarp_blend(a, b, flow, occ_logit, t):
        return O.eq2_blend(a, flow[:, 0:2], b, flow[:, 2:4], occ_logit, t.view(-1, 1, 1, 1))

    @staticmethod
    def cfr(flow_01, flow_10, t):
        return O.cfr_flow_t_align(flow_01, flow_10, t.view(-1, 1, 1, 1))

    @staticmethod
    def fgac_sample(refk, flow):
        H, W = refk.shape[-2:]
        px = ((2 * flow[:, 0] / (W - 1) - 1) + 1.0) / 2.0 * (W - 1)
        py = ((2 * flow[:, 1] / (H - 1) - 1) + 1.0) / 2.0 * (H - 1)
        return O.bilinear_gather(refk, px, py)[0]


def test_training_forward_and_gradients_match_the_reference():
    model = DeMFInet(synth.default_args())
    model.load_state_dict(synth.make_state_dict(0))
    x, t, gts = case_tensors()
    res = train_net.forward_train(model, x, t, CFG["n"], ops=TorchOps)
    assert len(res) == 7 and len(res[1]) == CFG["n"] and len(res[2]) == CFG["n"] + 1 and len(res[5]) == 4 and len(res[6][0]) == 2
    assert float((res[1][-1][2].detach() - torch.from_numpy(GOLD["St_final_last"])).abs().max()) < 5e-5
    assert float((res[2][-1].detach() - torch.from_numpy(GOLD["flow_last"])).abs().max()) < 5e-5
    total, d1, d2 = TO.rec_losses(res[0], res[1], *gts)
    assert np.allclose([float(total.detach()), float(d1.detach()), float(d2.detach())], GOLD["losses"], rtol=2e-6)
    total.backward()
    names = [n for n, _ in model.named_parameters()]
    assert names == list(GOLD["names"])
    grads = [(n, p.grad if p.grad is not None else torch.zeros(1)) for n, p in model.named_parameters()]
    # the only parameters without a gradient: conv_source_k, dead for rr = sr = 0 (the reference gives them exact zeros)
    assert [n for n, p in model.named_parameters() if p.grad is None] == [
        "FAC_FB_Module.shared_FGAC.conv_source_k.weight", "FAC_FB_Module.shared_FGAC.conv_source_k.bias"]
    got, want = summarise(grads), GOLD["summary"]
    numel = np.asarray([p.numel() for p in model.parameters()], dtype=np.float64)
    # each statistic against its natural scale: the gradient's own L2 norm (norm, projection on a unit direction) and
    # sqrt(numel) * norm for the plain sum (|sum| <= ||g||_1 <= sqrt(n) ||g||_2; correlated 1e-5-level differences add up in it)
    scale = np.maximum(want[:, 0:1], 1e-6) * np.stack([np.ones_like(numel), np.sqrt(numel), np.ones_like(numel)], 1)
    err = np.abs(got - want) / scale
    worst = int(err.max(1).argmax())
    print(f"worst parameter {names[worst]}: relative error {err[worst].max():.2e}; median {np.median(err.max(1)):.2e}")
    assert err.max() < 2e-4, (names[worst], err[worst])
    for n in FULL:
        g = dict(grads)[n]
        w = torch.from_numpy(GOLD["full:" + n])
        assert float((g - w).abs().max()) < 2e-3 * float(w.abs().max()) + 1e-7, n


def test_kernel_ops_refuse_cpu_tensors():
    from demfi_b200._abi import DemfiError
    with pytest.raises(DemfiError):
        train_net.KernelOps.cfr(torch.zeros(1, 2, 8, 8), torch.zeros(1, 2, 8, 8), torch.tensor([0.5]))
    with pytest.raises(DemfiError):
        train_net.KernelOps.fgac_sample(torch.zeros(1, 64, 8, 8), torch.zeros(1, 2, 8, 8))
    with pytest.raises(DemfiError):
        train_net.KernelOps.conv2d(torch.zeros(1, 8, 8, 8), torch.zeros(8, 8, 4, 4), None, "relu", stride=2)


# ------------------------------------------------------------------------------- data parallel: two ranks over gloo, on CPU
def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    from demfi_b200.train import allreduce_gradients
    torch.set_num_threads(2)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        model = DeMFInet(synth.default_args())
        model.load_state_dict(synth.make_state_dict(0))
        x, t, gts = case_tensors()
        sl = slice(rank, rank + 1)                                     # one sample of the golden batch per rank
        res = train_net.forward_train(model, x[sl], t[sl], CFG["n"], ops=TorchOps)
        total, _, _ = TO.rec_losses(res[0], res[1], *[g[sl] for g in gts])
        total.backward()
        live = [p for p in model.parameters() if p.grad is not None]
        nbytes = allreduce_gradients(live)
        grads = [(n, p.grad if p.grad is not None else torch.zeros(1)) for n, p in model.named_parameters()]
        q.put((rank, nbytes, float(total.detach()), summarise(grads)))
    finally:
        dist.destroy_process_group()


def test_two_rank_data_parallel_step_equals_the_batched_reference_gradient():
    """SURVEY.md section 8(e), training: each rank runs the differentiable forward on its own sample, the gradients go through
    `train.allreduce_gradients` (one flat bucket, SUM / world) -- and every rank ends up with the gradient the unmodified
    reference computes for the batch of both samples (the L1 means make the batch loss the average of the per-sample losses)."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = GOLD["summary"]
    model = DeMFInet(synth.default_args())
    numel = np.asarray([p.numel() for p in model.parameters()], dtype=np.float64)
    scale = np.maximum(want[:, 0:1], 1e-6) * np.stack([np.ones_like(numel), np.sqrt(numel), np.ones_like(numel)], 1)
    assert abs(0.5 * (res[0][2] + res[1][2]) - GOLD["losses"][0]) < 2e-6 * GOLD["losses"][0]
    for rank, nbytes, _, got in res:
        assert nbytes == 4 * int(numel.sum() - 64 * 64 - 64)      # all parameters but the dead conv_source_k (4096 + 64 values)
        err = np.abs(got - want) / scale
        assert err.max() < 2e-4, (rank, err.max())
    assert np.array_equal(res[0][3], res[1][3])                    # both ranks hold the same averaged gradient


def test_two_iterations_of_the_training_loop_follow_the_reference():
    """`train.train_step` = the loop body of train() (main.py:386-448): batch split, forward, losses, backward, exchange, Adam.
    Two iterations with torch stand-ins for the kernels (and torch.optim.Adam, since train.Adam is GPU-only) against two
    iterations of the same loop on the unmodified reference module: the losses of both steps, and where the parameters end up."""
    from demfi_b200 import train
    from oracle.gen_golden_train import training_frames
    model = DeMFInet(synth.default_args())
    model.load_state_dict(synth.make_state_dict(0))
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.999), weight_decay=0)
    sched = train.MultiStepLR(opt, milestones=[1, 3], gamma=0.5)
    frames, t = training_frames()
    got = [train.train_step(model, frames, t, opt, CFG["n"], ops=TorchOps, loss_fn=TO.rec_losses) for _ in range(2)]
    want = GOLD["step_losses"]
    assert np.allclose(got[0], want[0], rtol=2e-6), (got[0], want[0])
    assert np.allclose(got[1], want[1], rtol=2e-4), (got[1], want[1])      # after one Adam step of every parameter
    after = summarise([(n, p) for n, p in model.named_parameters()])
    numel = np.asarray([p.numel() for p in model.parameters()], dtype=np.float64)
    moved = 2e-4 * np.sqrt(numel)                                           # two Adam steps of at most lr per element
    d_norm = np.abs(after[:, 0] - GOLD["params_after_2_steps"][:, 0]) / moved
    d_proj = np.abs(after[:, 2] - GOLD["params_after_2_steps"][:, 2]) / moved
    print(f"parameters after two steps: norm off by {d_norm.max():.3f}, projection off by {d_proj.max():.3f} of the distance moved")
    assert d_norm.max() < 0.05 and d_proj.max() < 0.05
    # the scheduler mirror (main.py:186, :511)
    sched.step()
    assert opt.param_groups[0]["lr"] == 0.5e-4 and sched.get_last_lr() == [0.5e-4]
    sched.step(); sched.step()
    assert opt.param_groups[0]["lr"] == 0.25e-4


def test_split_training_batch_follows_main_py():
    from demfi_b200.train import split_training_batch
    f = torch.arange(2 * 3 * 9 * 4 * 4, dtype=torch.float32).reshape(2, 3, 9, 4, 4)
    inp, s0, s1, ft = split_training_batch(f)
    assert inp.shape == (2, 3, 4, 4, 4) and torch.equal(inp, f[:, :, :4]) and torch.equal(ft, f[:, :, 4])
    assert torch.equal(s0, f[:, :, 5]) and torch.equal(s1, f[:, :, 6])


def test_optimizer_and_scheduler_checkpoints_are_interchangeable_with_torch():
    """main.py:270-271 saves `optimizer.state_dict()` / `scheduler.state_dict()`, :214-217 loads them: a checkpoint written by
    torch.optim.Adam + MultiStepLR loads into train.Adam + train.MultiStepLR and the other way round (host logic only)."""
    from demfi_b200 import train
    g = torch.Generator().manual_seed(2)
    ps = [torch.nn.Parameter(torch.randn(3, 2, generator=g)), torch.nn.Parameter(torch.randn(5, generator=g))]
    ref = torch.optim.Adam(ps, lr=1e-4, betas=(0.9, 0.999), weight_decay=0)
    rs = torch.optim.lr_scheduler.MultiStepLR(ref, milestones=[2, 5], gamma=0.5)
    for _ in range(3):
        for p in ps:
            p.grad = torch.randn(p.shape, generator=g)
        ref.step()
        rs.step()
    mine = train.Adam(ps, lr=123.0)
    ms = train.MultiStepLR(mine, milestones=[9], gamma=0.1)
    mine.load_state_dict(ref.state_dict())
    ms.load_state_dict(rs.state_dict())
    assert mine.param_groups[0]["lr"] == ref.param_groups[0]["lr"] == 0.5e-4 and mine.state[1]["step"] == 3
    assert torch.equal(mine.state[0]["exp_avg_sq"], ref.state[ps[0]]["exp_avg_sq"])
    assert ms.last_epoch == 3 and ms.milestones == [2, 5] and ms.gamma == 0.5 and ms.base_lr == 1e-4
    for _ in range(2):
        ms.step()
        rs.step()
    assert mine.param_groups[0]["lr"] == ref.param_groups[0]["lr"] == 0.25e-4
    # and back: torch accepts what train.Adam / MultiStepLR write
    ref2 = torch.optim.Adam(ps, lr=1.0)
    ref2.load_state_dict(mine.state_dict())
    assert ref2.param_groups[0]["lr"] == 0.25e-4 and float(ref2.state[ps[1]]["step"]) == 3.0
    assert torch.equal(ref2.state[ps[0]]["exp_avg"], ref.state[ps[0]]["exp_avg"])
    rs2 = torch.optim.lr_scheduler.MultiStepLR(ref2, milestones=[100], gamma=0.9)
    rs2.load_state_dict(ms.state_dict())
    assert rs2.last_epoch == 5 and rs2.gamma == 0.5 and sorted(rs2.milestones.elements()) == [2, 5]
    with pytest.raises(ValueError):
        train.Adam(ps[:1]).load_state_dict(ref.state_dict())
