"""Clip-level runner: the loop structure of `test_custom` (main.py:1109-1196) over a synthetic or decoded
clip, sharded over one-process-per-GPU ranks.

Units: a clip of F frames has F-3 frame pairs (`utils.py:556-571`: inputs are frames
idx, idx+1, idx-1, idx+2); x M interpolation produces M-1 interpolated frames per pair.  Pairs are
independent, so rank r of W takes pairs p with p % W == r and no collective touches the data path; all the
t values of one pair stay on one rank so the t-independent prefix (FF_RDB + FAC_FB) can be reused.
"""
from __future__ import annotations

from typing import Callable, Iterator, List, Sequence, Tuple

import torch

from .caller import interpolate


def pair_indices(num_frames: int) -> List[int]:
    """centre indices idx = 1 .. F-3 (utils.py:563-566)"""
    return list(range(1, num_frames - 2))


def shard_pairs(pairs: Sequence[int], rank: int, world: int) -> List[int]:
    return [p for i, p in enumerate(pairs) if i % world == rank]


def t_values(multiple: int) -> List[float]:
    return [i / multiple for i in range(1, multiple)]


def pair_input(frames: torch.Tensor, idx: int) -> torch.Tensor:
    """frames [F,3,H,W] -> x [1,3,4,H,W] in the reference's slot order (B0, B1, B-1, B2) = (idx, idx+1, idx-1, idx+2)"""
    sel = frames[[idx, idx + 1, idx - 1, idx + 2]]  # [4,3,H,W]
    return sel.permute(1, 0, 2, 3).unsqueeze(0).contiguous()


@torch.no_grad()
def run_clip(model_net, frames: torch.Tensor, multiple: int, num_update: int, rank: int = 0, world: int = 1,
             reuse_prefix: bool = True, patch_boundary: int = 32,
             sink: Callable[[int, float, Tuple[torch.Tensor, ...]], None] | None = None) -> int:
    """Process this rank's pairs of a clip already resident on the model's device.  `sink(pair_idx, t, (S0,S1,St))`
    receives the results.  Returns the number of interpolated frames produced on this rank."""
    dev = frames.device
    done = 0
    for idx in shard_pairs(pair_indices(frames.shape[0]), rank, world):
        x = pair_input(frames, idx)
        for j, t in enumerate(t_values(multiple)):
            tt = torch.tensor([[t]], dtype=torch.float32, device=dev)
            out = interpolate(model_net, x, tt, num_update, patch_boundary, reuse_prefix=reuse_prefix and j > 0)
            if sink is not None:
                sink(idx, t, out)
            done += 1
    return done
