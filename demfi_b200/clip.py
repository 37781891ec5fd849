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


def schedule_units(pairs: Sequence[int], n_t: int, rank: int, world: int, balance_tail: bool = True) -> List[Tuple[int, List[int]]]:
    """This rank's work on a clip as [(pair index, [time indices])].  Whole pairs go round-robin (`pair % world`, all time
    indices of a pair on one rank: the t-independent prefix of the network is computed once per pair).  With balance_tail the
    pairs of the last, incomplete round -- 61 pairs on 8 GPUs leave 5 pairs for 8 ranks -- are cut into their (pair, t)
    units and dealt out in contiguous runs, so that no rank idles for a whole pair at the end (a rank that receives part of
    a pair recomputes that pair's prefix: 12 ms against 33 ms per time index at 1280x720)."""
    pairs = list(pairs)
    full = len(pairs) // world * world if balance_tail else len(pairs)
    if balance_tail and len(pairs) - full == 0:
        full = len(pairs)
    mine = [(p, list(range(n_t))) for i, p in enumerate(pairs[:full]) if i % world == rank]
    units = [(p, j) for p in pairs[full:] for j in range(n_t)]
    lo, hi = len(units) * rank // world, len(units) * (rank + 1) // world
    for p, j in units[lo:hi]:
        if mine and mine[-1][0] == p:  # (a tail pair: pair indices are unique, so this never merges into a whole pair)
            mine[-1][1].append(j)
        else:
            mine.append((p, [j]))
    return mine


@torch.no_grad()
def run_clip(model_net, frames: torch.Tensor, multiple: int, num_update: int, rank: int = 0, world: int = 1,
             reuse_prefix: bool = True, patch_boundary: int = 32,
             sink: Callable[[int, float, Tuple[torch.Tensor, ...]], None] | None = None, balance_tail: bool = False) -> int:
    """Process this rank's share of a clip already resident on the model's device (see schedule_units).
    `sink(pair_idx, t, (S0,S1,St))` receives the results.  Returns the number of interpolated frames produced on this rank."""
    dev = frames.device
    done = 0
    ts = t_values(multiple)
    for idx, js in schedule_units(pair_indices(frames.shape[0]), len(ts), rank, world, balance_tail):
        x = pair_input(frames, idx)
        for k, j in enumerate(js):
            tt = torch.tensor([[ts[j]]], dtype=torch.float32, device=dev)
            out = interpolate(model_net, x, tt, num_update, patch_boundary, reuse_prefix=reuse_prefix and k > 0)
            if sink is not None:
                sink(idx, ts[j], out)
            done += 1
    return done


# ---------------------------------------------------------------------------------------------------------------
# Folder runner: the `--phase test_custom` path of the reference (Custom_Test + make_2D_dataset_Custom_Test,
# utils.py:522-593; test_custom, main.py:1109-1196) with the I/O taken off the critical path.
#
#   <custom_path>/<scene>/*.png   ->   <custom_path>/<scene>_sharply_interpolated_x<M>/
#        <frame idx>.png, <frame idx+1>.png             deblurred S0_final / S1_final of the pair (written at its first t)
#        <frame idx stem>_<suffix:03d>.png               St_final for t = (suffix + 1) / M
#
# The reference decodes four PNGs, runs the model and encodes up to three PNGs synchronously on the main thread for
# every (pair, t).  Here: every frame is decoded ONCE (thread pool, cv2 releases the GIL), normalised on the GPU,
# kept in a small device cache (a frame serves four pairs); the t-independent prefix of the network is reused across
# the M-1 time indices of a pair; results are converted to uint8 on the GPU exactly as the reference does on the host
# (float64 denorm255_np, truncating astype(uint8)), copied back into pinned buffers and encoded by the pool while the
# next forward runs.
import glob
import os
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def enumerate_custom(custom_path: str, multiple: int):
    """[(scene, idx, [B0, B1, Bm1, B2] paths, [(t, St name)], S0 name, S1 name)] in the reference's order
    (make_2D_dataset_Custom_Test, utils.py:558-593: pairs idx = 1 .. F-3, t = linspace(1/M, 1-1/M, M-1))."""
    out = []
    ts = np.linspace(1 / multiple, 1 - 1 / multiple, multiple - 1)
    for scene_folder in sorted(glob.glob(os.path.join(custom_path, "*", ""))):
        if "_sharply_interpolated_x" in os.path.basename(os.path.dirname(scene_folder)):
            continue  # an output folder of an earlier run (the reference would re-process it as a scene)
        frames = sorted(glob.glob(scene_folder + "*.png"))
        scene = scene_folder.split(os.path.join(custom_path, ""))[-1].split(os.sep)[0]
        for idx in range(1, len(frames)):
            if idx == len(frames) - 2:
                break
            stem = os.path.basename(frames[idx]).split(".")[0]
            st = [(float(ts[m]), f"{stem}_{str(m).zfill(3)}.png") for m in range(multiple - 1)]
            out.append((scene, idx, [frames[idx], frames[idx + 1], frames[idx - 1], frames[idx + 2]], st,
                        os.path.basename(frames[idx]), os.path.basename(frames[idx + 1])))
    return out


_NORM_LUT = ((torch.arange(256, dtype=torch.float32) / 255.0 - 0.5) * 2)  # CPU fp32 arithmetic, as the reference's loader does


def normalize_bgr_u8(img_u8: torch.Tensor) -> torch.Tensor:
    """[H,W,3] uint8 (cv2 BGR) -> [3,H,W] fp32 in [-1,1]: RGBframes_np2Tensor (utils.py:224-238).  The reference computes
    (x / 255.0 - 0.5) * 2 on a CPU fp32 tensor; a 256-entry table built with exactly that arithmetic gives the same bits on
    any device (CUDA's tensor / scalar is a multiplication by the reciprocal and differs in the last place)."""
    return _NORM_LUT.to(img_u8.device)[img_u8.permute(2, 0, 1).long()]


def denorm255_u8(x: torch.Tensor) -> torch.Tensor:
    """[.., 3, H, W] in [-1,1] -> [.., H, W, 3] uint8: denorm255_np (utils.py:718-721) on the float64 copy the caller makes
    (utils.py:1415), then the truncating astype(np.uint8) of main.py:1165-1178."""
    y = ((x.to(torch.float64) + 1) / 2).clamp(0, 1) * 255
    return y.to(torch.uint8).movedim(-3, -1).contiguous()


class FolderRunner:
    def __init__(self, model_net, multiple: int, num_update: int, patch_boundary: int = 32, io_threads: int = 8,
                 rank: int = 0, world: int = 1, cache_frames: int = 8):
        self.net, self.M, self.N, self.pb = model_net, multiple, num_update, patch_boundary
        self.rank, self.world = rank, world
        self.pool = ThreadPoolExecutor(max_workers=io_threads)
        self.cache: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self.cache_frames = cache_frames
        self.dev = next(model_net.parameters()).device if hasattr(model_net, "parameters") else torch.device("cpu")
        self.cuda = self.dev.type == "cuda"
        self.copy_stream = torch.cuda.Stream(self.dev) if self.cuda else None

    # -- input side
    def _decode(self, path):
        import cv2
        img = cv2.imread(path)
        if img is None:
            raise FileNotFoundError(path)
        t = torch.from_numpy(img)
        return t.pin_memory() if self.cuda else t

    def _prefetch(self, paths, pending):
        for p in paths:
            if p not in self.cache and p not in pending:
                pending[p] = self.pool.submit(self._decode, p)

    def _frame(self, path, pending):
        if path in self.cache:
            self.cache.move_to_end(path)
            return self.cache[path]
        host = pending.pop(path).result() if path in pending else self._decode(path)
        if self.cuda:
            with torch.cuda.stream(self.copy_stream):
                f = normalize_bgr_u8(host.to(self.dev, non_blocking=True))
            cur = torch.cuda.current_stream(self.dev)
            cur.wait_stream(self.copy_stream)
            f.record_stream(cur)  # allocated on the copy stream, consumed on the compute stream: no reuse before that is done
        else:
            f = normalize_bgr_u8(host)
        self.cache[path] = f
        while len(self.cache) > self.cache_frames:
            self.cache.popitem(last=False)
        return f

    # -- output side
    @staticmethod
    def _write(path, img_hwc_u8: torch.Tensor, ready_event):
        import cv2
        if ready_event is not None:
            ready_event.synchronize()
        cv2.imwrite(path, img_hwc_u8.numpy())
        return path

    def _save(self, path, img_chw: torch.Tensor, futures):
        u8 = denorm255_u8(img_chw)
        if self.cuda:
            host = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
            host.copy_(u8, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        else:
            host, ev = u8, None
        futures.append(self.pool.submit(self._write, path, host, ev))

    @torch.no_grad()
    def run(self, custom_path: str) -> dict:
        """Returns {"pairs", "interpolated", "deblurred", "files"} for this rank's share of the work."""
        work = enumerate_custom(custom_path, self.M)
        mine = [w for i, w in enumerate(work) if i % self.world == self.rank]
        # The deblurred S1 of pair i and the deblurred S0 of pair i + 1 carry the same file name (main.py:1165-1172); the
        # reference writes sequentially, so the S0 of the later pair is what stays on disk.  Here writes are asynchronous and
        # neighbouring pairs may sit on different ranks: every such file is written ONCE, by the pair whose S0 it is; an S1 is
        # written only where no later pair produces that name (the last pair of a scene).
        s0_files = {(w[0], w[4]) for w in work}
        pending, futures = {}, []
        stats = {"pairs": 0, "interpolated": 0, "deblurred": 0, "files": []}
        for k, (scene, idx, paths, st, s0_name, s1_name) in enumerate(mine):
            self._prefetch(paths, pending)
            if k + 1 < len(mine):
                self._prefetch(mine[k + 1][2], pending)  # decode the next pair's frames while this one computes
            x = torch.stack([self._frame(p, pending) for p in paths], dim=1).unsqueeze(0)  # [1,3,4,H,W]: B0, B1, B-1, B2
            out_dir = os.path.join(custom_path, f"{scene}_sharply_interpolated_x{self.M}")
            os.makedirs(out_dir, exist_ok=True)
            for j, (t, st_name) in enumerate(st):
                tt = torch.tensor([[t]], dtype=torch.float32, device=x.device)
                s0, s1, stf = interpolate(self.net, x, tt, self.N, self.pb, reuse_prefix=j > 0)
                if j == 0:  # main.py:1160-1169: the deblurred pair is written at the first time index only
                    self._save(os.path.join(out_dir, s0_name), s0[0], futures)
                    stats["deblurred"] += 1
                    if (scene, s1_name) not in s0_files:
                        self._save(os.path.join(out_dir, s1_name), s1[0], futures)
                        stats["deblurred"] += 1
                self._save(os.path.join(out_dir, st_name), stf[0], futures)
                stats["interpolated"] += 1
            stats["pairs"] += 1
        stats["files"] = [f.result() for f in futures]
        return stats
