"""CPU: the caller mirror (utils.py:1339-1477 semantics), clip enumeration and the world_size-2 sharding
(gloo) of the N>1 path.  The model is replaced by the oracle (torch CPU) so that no GPU is needed."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from conftest import ROOT
from demfi_b200 import synth
from demfi_b200.caller import interpolate, pad_to_multiple, patch_forward_DeFInet_itr
from demfi_b200.clip import pair_indices, pair_input, run_clip, shard_pairs, t_values
from oracle import demfi_oracle as O


class OracleModel:
    """stands in for DeMFInet on CPU (test infrastructure only)"""

    def __init__(self, sd):
        self.sd = sd
        self.calls = []

    def __call__(self, x, t, n, is_training=None, reuse_prefix=False):
        self.calls.append((tuple(x.shape), float(t.reshape(-1)[0]), reuse_prefix))
        return O.forward(self.sd, x, t, n)


def test_reflect_pad_to_multiple_of_32_right_bottom_only():
    x = torch.arange(1 * 3 * 4 * 40 * 44, dtype=torch.float32).reshape(1, 3, 4, 40, 44)
    y, oh, ow = pad_to_multiple(x, 32)
    assert (oh, ow) == (40, 44) and y.shape == (1, 3, 4, 64, 64)
    assert torch.equal(y[..., :40, :44], x)
    want = F.pad(x.view(1, 12, 40, 44), [0, 20, 0, 24], mode="reflect").view(1, 3, 4, 64, 64)
    assert torch.equal(y, want)
    z, _, _ = pad_to_multiple(torch.zeros(1, 3, 4, 64, 96), 32)
    assert z.shape == (1, 3, 4, 64, 96)


def test_patch_forward_mirror_returns_reference_structure(state_dict):
    m = OracleModel(state_dict)
    x = synth.make_frames(24, 40, seed=3)
    t = torch.tensor([[0.25]])
    two, Sp, Sf, gt, flows, occs = patch_forward_DeFInet_itr(m, x, None, t, 2, (1, 1), 32)
    assert m.calls[0][0] == (1, 3, 4, 32, 64)  # padded to x32
    assert two.shape == (3, 24, 40) and two.dtype == np.float64
    assert len(Sp) == 3 and len(Sf) == 3 and all(a.shape == (3, 24, 40) for a in Sp + Sf)
    assert gt == 0 and flows[0][0].shape == (2, 24, 40) and flows[1][1].shape == (2, 24, 40)
    assert occs[0].shape == (1, 24, 40) and occs[1].shape == (1, 24, 40)
    # same numbers as calling the model on the padded input and cropping
    xp, _, _ = pad_to_multiple(x, 32)
    ref = O.forward(state_dict, xp, t, 2)
    assert np.abs(Sf[2] - ref[1][-1][2][0, :, :24, :40].numpy()).max() < 1e-6
    with pytest.raises(TypeError):
        patch_forward_DeFInet_itr(m, x, None, t, 1, (2, 2), 32)


def test_clip_enumeration_matches_custom_test_loader():
    # utils.py:563-571: a clip of F frames has F-3 pairs; x8 -> 7 t values per pair; 64 frames -> 427 frames
    assert pair_indices(64) == list(range(1, 62)) and len(pair_indices(64)) * len(t_values(8)) == 427
    assert t_values(8) == [i / 8 for i in range(1, 8)] and t_values(2) == [0.5]
    fr = torch.arange(6 * 3 * 2 * 2, dtype=torch.float32).reshape(6, 3, 2, 2)
    x = pair_input(fr, 2)
    assert x.shape == (1, 3, 4, 2, 2)
    for slot, f in enumerate((2, 3, 1, 4)):  # B0, B1, B-1, B2
        assert torch.equal(x[0, :, slot], fr[f])
    parts = [shard_pairs(pair_indices(64), r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == pair_indices(64) and [len(p) for p in parts] == [8, 8, 8, 8, 8, 7, 7, 7]


def test_run_clip_reuses_prefix_only_within_a_pair(state_dict):
    m = OracleModel(state_dict)
    frames = synth.make_frames(32, 32, seed=5)[0].permute(1, 0, 2, 3)  # 4 frames
    frames = torch.cat([frames, frames[:1]], 0)  # 5 frames -> 2 pairs
    got = []
    n = run_clip(m, frames, 4, 1, sink=lambda idx, t, out: got.append((idx, t)))
    assert n == 6 and [g[0] for g in got] == [1, 1, 1, 2, 2, 2]
    assert [c[2] for c in m.calls] == [False, True, True, False, True, True]


def _worker(rank, world, port, sd_path, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    torch.set_num_threads(2)
    sd = torch.load(sd_path)
    m = OracleModel(sd)
    frames = synth.make_frames(32, 32, seed=7, batch=2)  # [2,3,4,32,32] -> 8 frames
    frames = frames.permute(0, 2, 1, 3, 4).reshape(8, 3, 32, 32)
    mine = {}
    n = run_clip(m, frames, 2, 1, rank=rank, world=world, sink=lambda idx, t, out: mine.__setitem__(idx, out[2].clone()))
    cnt = torch.tensor([n])
    dist.all_reduce(cnt)  # bookkeeping only: the data path has no collective
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: v.numpy() for k, v in mine.items()})
    if rank == 0:
        q.put((int(cnt), gathered))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo(state_dict, tmp_path):
    sd_path = str(tmp_path / "sd.pt")
    torch.save(state_dict, sd_path)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sd_path, q)) for r in range(2)]
    [p.start() for p in procs]
    total, gathered = q.get(timeout=300)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert total == 5  # 8 frames -> 5 pairs x 1 t
    assert sorted(gathered[0]) == [1, 3, 5] and sorted(gathered[1]) == [2, 4]
    # sharded results == single-process results
    frames = synth.make_frames(32, 32, seed=7, batch=2).permute(0, 2, 1, 3, 4).reshape(8, 3, 32, 32)
    single = {}
    run_clip(OracleModel(state_dict), frames, 2, 1, sink=lambda idx, t, out: single.__setitem__(idx, out[2].numpy()))
    for part in gathered:
        for idx, arr in part.items():
            assert np.abs(arr - single[idx]).max() < 1e-4  # worker uses 2 CPU threads, this process all of them


# ---- folder runner: mirrors Custom_Test / test_custom (utils.py:522-593, main.py:1109-1196)
def _write_scene(root, scene, n, h, w, seed):
    import cv2
    os.makedirs(os.path.join(root, scene), exist_ok=True)
    g = np.random.Generator(np.random.PCG64(seed))
    base = g.integers(0, 256, size=(h + 8, w + 8, 3), dtype=np.uint8)
    base = cv2.blur(base, (5, 5))
    for i in range(n):
        cv2.imwrite(os.path.join(root, scene, f"{i:05d}.png"), base[i % 4:i % 4 + h, (2 * i) % 7:(2 * i) % 7 + w])


def test_enumerate_custom_matches_reference_naming(tmp_path):
    from demfi_b200.clip import enumerate_custom
    _write_scene(str(tmp_path), "sceneA", 6, 24, 40, 1)
    _write_scene(str(tmp_path), "sceneB", 5, 24, 40, 2)
    work = enumerate_custom(str(tmp_path), 8)
    # F frames -> F-3 pairs (idx = 1 .. F-3), 7 time indices each, named <B0 stem>_<suffix:03d>.png
    assert [(w[0], w[1]) for w in work] == [("sceneA", 1), ("sceneA", 2), ("sceneA", 3), ("sceneB", 1), ("sceneB", 2)]
    scene, idx, paths, st, s0, s1 = work[1]
    assert [os.path.basename(p) for p in paths] == ["00002.png", "00003.png", "00001.png", "00004.png"]  # B0, B1, B-1, B2
    assert [n for _, n in st] == [f"00002_{m:03d}.png" for m in range(7)]
    assert np.allclose([t for t, _ in st], np.linspace(1 / 8, 7 / 8, 7)) and (s0, s1) == ("00002.png", "00003.png")


class BlendModel:
    """deterministic stand-in with the model's call signature and return structure: S0 = B0 * 0.9, S1 = B1 * 0.9 + 0.05,
    St = (1 - t) B0 + t B1 (bit-reproducible, unlike a multi-threaded conv stack)"""

    def __init__(self):
        self.calls = []

    def __call__(self, x, t, n, is_training=None, reuse_prefix=False):
        self.calls.append((tuple(x.shape), float(t.reshape(-1)[0]), reuse_prefix))
        b0, b1 = x[:, :, 0], x[:, :, 1]
        tt = t.reshape(-1, 1, 1, 1)
        fin = [b0 * 0.9, b1 * 0.9 + 0.05, (1 - tt) * b0 + tt * b1]
        return [b0, b1, b0], [fin] * n, [None] * (n + 1), [None] * (n + 1), (b0 + b1) / 2


def test_folder_runner_writes_what_the_reference_loop_writes(tmp_path):
    """End to end on CPU: frames decoded once, prefix reuse inside a pair, and byte-identical PNGs to a literal restatement of
    test_custom's save path (float64 denorm255_np, truncating uint8)."""
    import cv2
    from demfi_b200.clip import FolderRunner
    root = str(tmp_path)
    _write_scene(root, "clip", 5, 24, 40, 3)  # 2 pairs
    m = BlendModel()
    stats = FolderRunner(m, multiple=4, num_update=1, io_threads=2).run(root)
    # three deblurred files for two pairs: 00002.png is both S1 of pair 1 and S0 of pair 2 -- the reference's sequential loop leaves
    # pair 2's S0 on disk (main.py:1165-1172), so that is the one write made (no unordered double write)
    assert stats["pairs"] == 2 and stats["interpolated"] == 6 and stats["deblurred"] == 3 and len(stats["files"]) == 9
    assert len(set(stats["files"])) == len(stats["files"])
    assert [c[2] for c in m.calls] == [False, True, True, False, True, True]  # prefix recomputed once per pair
    assert all(c[0] == (1, 3, 4, 32, 64) for c in m.calls)  # reflect-padded to x32
    out_dir = os.path.join(root, "clip_sharply_interpolated_x4")
    assert sorted(os.listdir(out_dir)) == sorted(["00001.png", "00002.png", "00003.png"] +
                                                 [f"0000{i}_{k:03d}.png" for i in (1, 2) for k in range(3)])
    # literal reference path (utils.py:224-238, 1339-1477; main.py:1150-1178) for pair idx = 1
    frames = np.stack([cv2.imread(os.path.join(root, "clip", f"{i:05d}.png")) for i in (1, 2, 0, 3)], axis=0)
    x = torch.Tensor(frames.transpose(3, 0, 1, 2).astype(float)).mul_(1.0)
    x = ((x / 255.0 - 0.5) * 2).unsqueeze(0)  # RGBframes_np2Tensor
    xp, oh, ow = pad_to_multiple(x, 32)
    ref = BlendModel()
    for k, t in enumerate((0.25, 0.5, 0.75)):
        res = ref(xp, torch.tensor([[t]], dtype=torch.float32), 1)
        st = np.squeeze(res[1][-1][2].numpy()).astype(np.float64)[..., :oh, :ow]
        img = (((np.transpose(st, [1, 2, 0])[:, :, ::-1] + 1) / 2).clip(0, 1) * 255).astype(np.uint8)[:, :, ::-1]
        assert np.array_equal(cv2.imread(os.path.join(out_dir, f"00001_{k:03d}.png")), img), k
    # the deblurred pair is written at the FIRST time index of the pair only (main.py:1160-1169)
    res0 = ref(xp, torch.tensor([[0.25]]), 1)
    s0 = np.squeeze(res0[1][-1][0].numpy()).astype(np.float64)[..., :oh, :ow]
    img0 = np.transpose(((s0 + 1) / 2).clip(0, 1) * 255, [1, 2, 0]).astype(np.uint8)
    assert np.array_equal(cv2.imread(os.path.join(out_dir, "00001.png")), img0)
    # the shared file holds the S0 of the LATER pair (idx = 2: frames 2, 3, 1, 4), the last pair's S1 is written too
    frames2 = np.stack([cv2.imread(os.path.join(root, "clip", f"{i:05d}.png")) for i in (2, 3, 1, 4)], axis=0)
    x2 = ((torch.Tensor(frames2.transpose(3, 0, 1, 2).astype(float)) / 255.0 - 0.5) * 2).unsqueeze(0)
    x2p, _, _ = pad_to_multiple(x2, 32)
    r2 = ref(x2p, torch.tensor([[0.25]]), 1)
    for name, k in (("00002.png", 0), ("00003.png", 1)):
        v = np.squeeze(r2[1][-1][k].numpy()).astype(np.float64)[..., :oh, :ow]
        assert np.array_equal(cv2.imread(os.path.join(out_dir, name)), np.transpose(((v + 1) / 2).clip(0, 1) * 255, [1, 2, 0]).astype(np.uint8)), name


def test_folder_runner_shards_pairs_across_ranks(tmp_path):
    """world_size 2: the two ranks split the frame pairs (all t of a pair stay on one rank), together they write every file
    exactly as one rank does"""
    from demfi_b200.clip import FolderRunner
    roots = []
    for world in (1, 2):
        root = str(tmp_path / f"w{world}")
        _write_scene(root, "s", 7, 24, 40, 9)  # 4 pairs
        counts = [FolderRunner(BlendModel(), multiple=2, num_update=1, io_threads=2, rank=r, world=world).run(root) for r in range(world)]
        assert sum(c["pairs"] for c in counts) == 4 and sum(c["interpolated"] for c in counts) == 4
        if world == 2:
            assert [c["pairs"] for c in counts] == [2, 2]
        roots.append(os.path.join(root, "s_sharply_interpolated_x2"))
    import cv2
    assert sorted(os.listdir(roots[0])) == sorted(os.listdir(roots[1]))
    for name in os.listdir(roots[0]):  # every file has exactly one writer, whatever the sharding: identical on disk
        assert np.array_equal(cv2.imread(os.path.join(roots[0], name)), cv2.imread(os.path.join(roots[1], name))), name


def test_clip_schedule_balances_the_last_round():
    """BASELINE config 3: 64 frames = 61 pairs x 7 time indices on 8 ranks.  Whole pairs round-robin would leave 8,8,8,8,8,7,7,7
    pairs (the last rank idle for a pair's worth of time, efficiency 61/64); cutting the last round's five pairs into (pair, t)
    units gives 53 or 54 units to every rank, every unit exactly once"""
    from demfi_b200.clip import pair_indices, schedule_units
    pairs = pair_indices(64)
    assert len(pairs) == 61
    assert [len(schedule_units(pairs, 7, r, 8, balance_tail=False)) for r in range(8)] == [8, 8, 8, 8, 8, 7, 7, 7]
    for world in (1, 2, 3, 8):
        seen, loads = set(), []
        for r in range(world):
            units = schedule_units(pairs, 7, r, world)
            loads.append(sum(len(js) for _, js in units))
            for p, js in units:
                assert js == sorted(js)
                for j in js:
                    assert (p, j) not in seen
                    seen.add((p, j))
        assert len(seen) == 61 * 7 and max(loads) - min(loads) <= (1 if world == 8 else 7), (world, loads)


def _allreduce_worker(rank, world, port, q):
    import torch.distributed as dist
    from demfi_b200.train import allreduce_gradients
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        params = [torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2))]
        params[0].grad = torch.full((4, 3), float(rank + 1))
        params[1].grad = torch.arange(5, dtype=torch.float32) * (rank + 1)
        nbytes = allreduce_gradients(params)            # params[2] has no gradient: skipped
        q.put((rank, nbytes, params[0].grad.clone(), params[1].grad.clone(), params[2].grad))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    """the one exchange step of data-parallel training, world_size 2 over gloo on CPU: the flat bucket
    is summed over ranks and averaged, tensors without a gradient stay out of the bucket"""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, nbytes, g0, g1, g2 in res:
        assert nbytes == (12 + 5) * 4 and g2 is None
        assert torch.equal(g0, torch.full((4, 3), 1.5)) and torch.equal(g1, torch.arange(5, dtype=torch.float32) * 1.5)
