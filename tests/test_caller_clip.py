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
