#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its config: interpolated frames/s, 1280x720, x8 MFI, N_tst=3.

  python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (B200 kernels through the C ABI)
  python bench.py --impl reference [...]                       # reference arm: the CPU path on the host cores

A "step" is one pass of the hot path over one batch of synthetic input = one DeMFInet forward for one
(frame pair, t) = ONE interpolated frame (plus the two deblurred frames it returns as by-products).
The step does all the work the reference does for that call: the whole network including the
t-independent prefix and every boosting iteration's D2 decode (no caching, nothing skipped); the
prefix-cached / final-only variants are reported separately in `extra`.

One JSON line on stdout (rank 0).  Multi-GPU: launched under torchrun, one process per GPU, each rank
works on its own frame pairs (weak scaling, no data-path collective); the timed region is bracketed by a
barrier + synchronize and the maximum over ranks is used.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

H0, W0 = 720, 1280      # BASELINE.json config[1]; the caller reflect-pads to 736x1280 (utils.py:1351-1365)
N_TST, MFI = 3, 8
# reference arm / cpu_baseline: always the WHOLE padded frame (CPU conv throughput is not area-linear, so nothing is cropped or
# extrapolated); what is bounded is the number of forwards: as many of the requested K as fit in REF_BUDGET_S seconds of CPU
# work (one forward takes ~16 s on the B200 box's 16 host cores), and `steps` / `ms_per_step` report what was actually run.
REF_BUDGET_S = float(os.environ.get("DEMFI_REF_BUDGET_S", "150"))


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []  # the upper half of the samples = under load
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_sample(steps: int, warmup: int, budget_s: float = REF_BUDGET_S):
    """The reference's CPU implementation of the path on this box's host cores.  The reference is Python and is
    not present on the GPU box, so this is the oracle PORT (oracle/demfi_oracle.py, pinned to the reference by
    tests/test_oracle.py and, at this very size, by tests/golden/full_736x1280_n3.npz) with all host threads.  Each step = one
    forward on the whole reflect-padded frame.  Returns (frames/s, ms per step, cores, sample description, steps run)."""
    from demfi_b200 import synth
    from oracle import demfi_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp, wp = (H0 + 31) // 32 * 32, (W0 + 31) // 32 * 32
    sd = synth.make_state_dict(0)
    x = synth.make_frames(hp, wp, seed=0)
    ts = [torch.tensor([[t]]) for t in synth.mfi_t_values(MFI)]
    t_start = time.perf_counter()
    done_w = 0
    for i in range(warmup):  # at least one untimed forward (thread pool, allocator), more only while they fit in a third of the budget
        if i > 0 and (time.perf_counter() - t_start) * (i + 1) / i > budget_s / 3:
            break
        O.forward(sd, x, ts[i % len(ts)], N_TST)
        done_w += 1
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        if done > 0 and (time.perf_counter() - t_start) + (time.perf_counter() - t0) / done > budget_s:
            break
        O.forward(sd, x, ts[i % len(ts)], N_TST)
        done += 1
    dt = (time.perf_counter() - t0) / max(done, 1)
    sample = (f"oracle port (torch CPU fp32, {cores} threads), {done} timed forward(s) (+{done_w} warm-up) on the whole {hp}x{wp} "
              f"padded frame, N_tst={N_TST} ({dt:.2f} s per forward); {steps} requested, bounded by {budget_s:.0f} s of CPU work")
    return 1.0 / dt, dt * 1000.0, cores, sample, done


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    fps, ms, cores, sample, done = cpu_reference_sample(args.steps, args.warmup)
    hp = (H0 + 31) // 32 * 32
    line = {"impl": "reference", "metric": "interpolated_frames_per_sec", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": done, "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{W0}x{H0} x{MFI} MFI, N_tst={N_TST}: 1 interpolated frame (one DeMFInet forward on the "
                                   f"{W0}x{hp} reflect-padded pair) per step, full network every step"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def measure_gemm_peak(dev, dtype=torch.bfloat16):
    """cuBLAS dense GEMM burst rate on this GPU, same method as MEASURED_PEAKS.json (8192^3, best of 10, CUDA events);
    reported beside the driver's number so that the roofline denominator can be cross-checked in the same run."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev).to(dtype)
        b = torch.randn(n, n, device=dev).to(dtype)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        return 2 * n ** 3 / best / 1e9
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
        del a, b
        torch.cuda.empty_cache()


def conv_groups(prof, K):
    """ms per step of every conv launch, grouped by the reference module it belongs to"""
    groups = {}
    for fam in ("conv_tc", "conv_ffma"):
        for label, v in prof.get(fam, {}).get("by_label", {}).items():
            g = label.split(".")[0]
            if g.startswith("Dec") or g == "Ch_Reducer":
                g = "D1" if not label.split(".")[0].endswith("_2") and "res_2" not in label else "D2"
                if label.startswith("Ch_Reducer"):
                    g = "Ch_Reducer"
            d = groups.setdefault(g, {"ms": 0.0, "tflops": 0.0, "macs": 0, "ffma_ms": 0.0})
            d["ms"] += v["ms"] / K
            d["macs"] += v["macs"] / K
            if fam == "conv_ffma":
                d["ffma_ms"] += v["ms"] / K
    return {g: {"ms": round(d["ms"], 2), "TFLOP/s": round(2 * d["macs"] / (d["ms"] / 1e3) / 1e12, 1), "of_which_cuda_core_ms": round(d["ffma_ms"], 2)}
            for g, d in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])}


def run_ours(args):
    import torch.distributed as dist
    from demfi_b200 import _abi as A
    from demfi_b200 import synth
    from demfi_b200.DeMFInet import DeMFInet
    from demfi_b200.caller import interpolate, patch_forward_DeFInet_itr

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    assert torch.cuda.is_available(), "bench.py (our arm) needs a B200: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, Wm = args.steps, max(args.warmup, 3)
    for kv in filter(None, os.environ.get("DEMFI_OPTS", "").split(",")):  # diagnostics: library options for an A/B run, e.g. tc_flush=10
        k_, v_ = kv.split("=")
        A.set_option(k_, int(v_))

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    bf16_live = measure_gemm_peak(dev) if rank == 0 else None

    sd = synth.make_state_dict(0)
    net = DeMFInet(synth.default_args(gpu=local)).to(dev).eval()
    net.load_state_dict(sd, strict=True)
    # each rank owns different frame pairs (seeded by rank): weak scaling, no data-path collective
    pairs = [synth.make_frames(H0, W0, seed=100 * rank + i) for i in range(2)]
    x_dev = [p.to(dev) for p in pairs]
    x_pin = [p.pin_memory() for p in pairs]
    tvals = synth.mfi_t_values(MFI)
    t_dev = [torch.tensor([[t]], device=dev) for t in tvals]
    out_pin = torch.empty((1, 3, H0, W0), dtype=torch.float32).pin_memory()

    def step(i, reuse=False):
        return interpolate(net, x_dev[(i // len(tvals)) % 2], t_dev[i % len(tvals)], N_TST, 32, reuse_prefix=reuse)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(Wm):
        step(i)
    eng = next(iter(net._engines.values()))
    # ---- timed region: K full forwards, inputs resident in HBM (clocks sampled during it)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = A.launch_count()
    ms_total = timed(lambda i: step(i), K)
    launches = A.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = world * K / (ms_total / 1e3)
    # ---- the same K steps once more with a CUDA-event pair around every C-ABI call: per-kernel durations for the roofline
    # (kept out of the headline region: ~500 event records per step cost ~2 %)
    eng.profile = []
    ms_profiled = timed(lambda i: step(i), K)
    prof = eng.profile_summary()
    eng.profile = None

    # ---- end to end: pinned host frames -> H2D -> forward -> D2H of the interpolated frame, every step
    def e2e_step(i):
        xd = x_pin[(i // len(tvals)) % 2].to(dev, non_blocking=True)
        s0, s1, st = interpolate(net, xd, t_dev[i % len(tvals)], N_TST, 32)
        out_pin.copy_(st, non_blocking=False)

    e2e_step(0)
    ms_e2e = timed(e2e_step, K)
    e2e_value = world * K / (ms_e2e / 1e3)

    # ---- the same through the reference's own inference boundary (utils.py:1339-1477, mirrored by
    # caller.patch_forward_DeFInet_itr): pinned host frames -> H2D -> forward -> thirteen float64 host arrays
    ref_bytes = [0]

    def e2e_ref_step(i):
        xd = x_pin[(i // len(tvals)) % 2].to(dev, non_blocking=True)
        out = patch_forward_DeFInet_itr(net, xd, None, t_dev[i % len(tvals)], N_TST, (1, 1), 32)
        if i == 0:
            flat = [out[0]] + list(out[1]) + list(out[2]) + [a for pr in out[4] for a in pr] + list(out[5])
            ref_bytes[0] = int(sum(a.size for a in flat) * 4)  # device -> host moves fp32; the float64 widening happens on the host

    e2e_ref_step(0)
    Kr = max(2, min(K, 7))
    ms_e2e_ref = timed(e2e_ref_step, Kr)
    e2e_ref_value = world * Kr / (ms_e2e_ref / 1e3)

    # ---- the optimised variants that keep the returned frames identical (reported, not the headline)
    def cached_step(i):
        step(i, reuse=(i % len(tvals)) != 0)
    ms_cached = timed(cached_step, len(tvals))
    net.final_only = True
    ms_cached_fo = timed(cached_step, len(tvals))
    net.final_only = False

    lt = torch.tensor([float(launches)], device=dev)
    if world > 1:
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (tcgen05 conv), from the live CUDA-event durations above.
    # Denominator: the convs run kind::f16 MMAs, three products per MAC (3xFP16 split, fp32 parity), inside a long step
    # => measured cuBLAS bf16 SUSTAINED rate / 3 (MEASURED_PEAKS.json; B200_PROFILING.md's fallback when the file is absent).
    tc = prof.get("conv_tc", {"ms": 0.0, "macs": 0, "launches": 0, "by_label": {}})
    tc_share = tc["ms"] / sum(d["ms"] for d in prof.values())
    achieved = 2 * tc["macs"] / (tc["ms"] / 1e3) / 1e12 if tc["ms"] > 0 else 0.0
    if peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops"):
        dense = float(peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops"))
        basis = (f"of measured: MEASURED_PEAKS.json cuBLAS bf16 {'sustained' if peaks.get('bf16_tflops_sustained') else 'burst'} "
                 f"{dense:.0f} TFLOP/s / 3 products per MAC (3xFP16)")
    else:
        dense = 1400.0
        basis = ("of fallback: MEASURED_PEAKS.json absent, B200_PROFILING.md fallback 1.4 PFLOP/s sustained bf16 (1.59 burst) "
                 "/ 3 products per MAC (3xFP16)")
    basis += f"; cuBLAS bf16 burst measured live in this run: {bf16_live:.0f} TFLOP/s"
    peak = dense / 3.0
    top = sorted(tc["by_label"].items(), key=lambda kv: -kv[1]["ms"])[:5]
    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", "conv_s3_ncu_summary.json")) as f:
            ncu = json.load(f)
    except OSError:
        pass
    roofline = {
        "bound": "tensor", "achieved": round(achieved, 2), "peak": round(peak, 1), "unit": "TFLOP/s",
        "frac": round(achieved / peak, 4), "traffic": ncu.get("dram_bytes_per_launch"),
        "kernel": "demfi::conv_s3_kernel<NMAX> (tcgen05 kind::f16 SS-form, 3xFP16 fp32-parity split; the three stride-2 UNet "
                  "encoders run demfi::conv_h3_kernel), all tensor-core conv launches in the timed region",
        "launches_per_step": tc["launches"] // K, "share_of_step_time": round(tc_share, 3),
        "profiled_pass_ms_per_step": round(ms_profiled / K, 3),
        "algorithmic_flops_per_step": 2 * tc["macs"] // K,
        "peak_basis": basis,
        "traffic_basis": ncu.get("kernel"),
        "traffic_capture": {"commit": ncu.get("commit"), "file": "profiles/conv_s3_ncu_summary.json", "duration_us": ncu.get("duration_us"),
                            "algorithmic_bytes_per_launch": ncu.get("algorithmic_bytes_per_launch"),
                            "tensor_pipe_active_pct": ncu.get("tensor_pipe_active_pct")},
        "top_shapes": [{"conv": k, "launches_per_step": v["launches"] // K, "ms_per_launch": round(v["ms"] / v["launches"], 3),
                        "TFLOP/s": round(2 * v["macs"] / (v["ms"] / 1e3) / 1e12, 1)} for k, v in top],
        "other_kernels_ms_per_step": {k: round(v["ms"] / K, 3) for k, v in prof.items() if k != "conv_tc"},
        "conv_ms_per_step_by_layer_group": conv_groups(prof, K),
        "hbm_bound_kernels": {k: {"launches_per_step": v["launches"] // K, "ms_per_launch": round(v["ms"] / v["launches"], 4),
                                  "GB/s": round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1),
                                  "frac_of_hbm_peak": round(v["bytes"] / (v["ms"] / 1e3) / 1e9 / (peaks.get("hbm_gbs") or 6650.0), 3)}
                              for k, v in prof.items() if v.get("bytes", 0) > 0 and v["ms"] > 0},
    }
    cpu_fps, _, cores, sample, _ = cpu_reference_sample(1, 1)
    hp = (H0 + 31) // 32 * 32
    line = {
        "metric": "interpolated_frames_per_sec", "value": round(value, 4), "unit": "frames/s", "n_gpus": world, "steps": K,
        "warmup": Wm, "ms_per_step": round(ms_total / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{W0}x{H0} x{MFI} MFI, N_tst={N_TST}: 1 interpolated frame (one DeMFInet forward on the "
                               f"{W0}x{hp} reflect-padded pair) per step, full network every step",
                   "frames_per_rank": "2 synthetic frame pairs x 7 t values, cycled", "parallelism": f"pair-sharded x{world}",
                   "l2": "per-step activations (13.7 GB workspace) >> 126 MB L2; inputs 44 MB/pair", "conv_precision": "3xFP16 split on kind::f16 tensor cores, fp32 accumulate (fp32 parity: <=5e-4 max-abs end to end)"},
        "e2e": {"value": round(e2e_value, 4), "unit": "frames/s", "h2d_bytes_per_step": int(x_pin[0].numel() * 4 + 4),
                "d2h_bytes_per_step": int(out_pin.numel() * 4), "steps": K,
                "api": "demfi_b200.caller.interpolate(DeMFInet, pinned host frames, t) -> host St_final"},
        "e2e_reference_boundary": {"value": round(e2e_ref_value, 4), "unit": "frames/s", "h2d_bytes_per_step": int(x_pin[0].numel() * 4 + 4),
                                   "d2h_bytes_per_step": ref_bytes[0], "steps": Kr,
                                   "api": "demfi_b200.caller.patch_forward_DeFInet_itr (mirror of utils.py:1339-1477): pinned host frames "
                                          "-> 13 float64 numpy arrays (two_blurry, 3 + 3 sharp frames, 4 flows, 2 occlusion maps), "
                                          "the host-side float64 conversion inside the timed region"},
        "gpu_launches": int(lt.item()),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "extra": {"frames_per_sec_prefix_cached": round(world * len(tvals) / (ms_cached / 1e3), 4),
                  "frames_per_sec_prefix_cached_final_only": round(world * len(tvals) / (ms_cached_fo / 1e3), 4),
                  "note": "cached variants reuse the t-independent FF_RDB+FAC_FB stage across the 7 t of a pair and/or decode "
                          "D2 only for the last boosting iteration; outputs read by the inference caller are unchanged",
                  "bf16_dense_tflops_live": round(bf16_live, 1), "hbm_gbs_measured": peaks.get("hbm_gbs"),
                  "hbm_gbs_basis": "MEASURED_PEAKS.json" if peaks.get("hbm_gbs") else "fallback 6650 GB/s (B200_PROFILING.md)"},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_clip_workload(args, H0=H0, W0=W0, MFI=MFI, N_TST=N_TST, label="BASELINE config 3"):
    """BASELINE config 5 (--workload 4k) is the same loop on a 3840x2160 clip, x16 MFI, N_tst = 5: the whole reflect-padded frame
    (3840x2176) is one engine call per interpolated frame -- the reference tiles 4K frames for memory (utils.py:1757-1798); here the
    liveness-planned workspace of a whole 4K frame takes ~65 GB of the 180 GB, so the sharding is by frame pair only.
    BASELINE config 3 as written: ONE 64-frame 1280x720 clip (61 frame pairs x 7 time indices = 427 interpolated frames)
    sharded over the ranks -- strong scaling.  The clip sits in pinned host memory on every rank; a rank copies the frames of
    its pairs to the device as it goes, reuses the t-independent prefix across the time indices of a pair and copies every
    interpolated frame back to pinned host memory.  Timed from a barrier to the last rank done (device events, max over ranks).
    --balance 1 (default) cuts the pairs of the last incomplete round into (pair, t) units (clip.schedule_units)."""
    import torch.distributed as dist
    from demfi_b200 import synth
    from demfi_b200.DeMFInet import DeMFInet
    from demfi_b200.caller import interpolate
    from demfi_b200.clip import pair_indices, pair_input, schedule_units, t_values
    world, rank, local = env_int("WORLD_SIZE", 1), env_int("RANK", 0), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = DeMFInet(synth.default_args(gpu=local)).to(dev).eval()
    net.load_state_dict(synth.make_state_dict(0), strict=True)
    F_ = args.frames
    base = synth.make_frames(H0, W0, seed=7)[0]                       # [3,4,H,W]: four distinct frames, cycled with a drift
    clip = torch.stack([torch.roll(base[:, i % 4], shifts=i // 4, dims=2) for i in range(F_)]).contiguous().pin_memory()  # [F,3,H,W]
    ts = t_values(MFI)
    pairs = pair_indices(F_)
    units = schedule_units(pairs, len(ts), rank, world, balance_tail=bool(args.balance))
    # results go back to pinned host memory asynchronously through a small ring of buffers (a slot is reused only after its
    # copy has completed): the host never waits for the GPU inside the loop, so the ~200 launches of the next forward are
    # queued while this one runs
    NOUT = 4
    out_pin = [torch.empty((1, 3, H0, W0), dtype=torch.float32).pin_memory() for _ in range(NOUT)]
    out_evt = [None] * NOUT
    t_dev = [torch.tensor([[t]], device=dev) for t in ts]
    stage = [torch.empty((4, 3, H0, W0), dtype=torch.float32, device=dev) for _ in range(2)]

    def load_pair(idx, k):
        # four asynchronous copies straight from the pinned clip (no host-side gather), slot order B0, B1, B-1, B2
        for i, f in enumerate((idx, idx + 1, idx - 1, idx + 2)):
            stage[k % 2][i].copy_(clip[f], non_blocking=True)
        return stage[k % 2].permute(1, 0, 2, 3).unsqueeze(0).contiguous()

    def one_pass():
        n = 0
        for k, (idx, js) in enumerate(units):
            x = load_pair(idx, k)
            for q, j in enumerate(js):
                s0, s1, st = interpolate(net, x, t_dev[j], N_TST, 32, reuse_prefix=q > 0)
                slot = n % NOUT
                if out_evt[slot] is not None:
                    out_evt[slot].synchronize()
                out_pin[slot].copy_(st, non_blocking=True)
                out_evt[slot] = torch.cuda.Event()
                out_evt[slot].record()
                n += 1
        return n

    x = pair_input(clip, pairs[0]).to(dev)
    for w_ in range(3):  # warm-up: engine construction, weight packing, function attributes
        interpolate(net, x, t_dev[w_], N_TST, 32, reuse_prefix=w_ > 0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = one_pass()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    tot = torch.tensor([float(done)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"metric": "interpolated_frames_per_sec", "value": round(float(tot) / (float(ms) / 1e3), 3), "unit": "frames/s",
                          "n_gpus": world, "steps": 1, "warmup": 3, "ms_per_step": round(float(ms), 1), "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"{label}: one {F_}-frame {W0}x{H0} clip, x{MFI} MFI, N_tst={N_TST}: {len(pairs)} pairs x "
                                                 f"{len(ts)} t = {int(tot)} interpolated frames, pairs sharded over {world} rank(s), prefix reused "
                                                 f"within a pair, host frames in / host frames out", "balance_tail": bool(args.balance),
                                     "workspace_GB": round(max(e.workspace_bytes() for e in net._engines.values()) / 1e9, 2),
                                     "units_per_rank_max": max(sum(len(js) for _, js in schedule_units(pairs, len(ts), r, world, bool(args.balance)))
                                                               for r in range(world))}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train_workload(args):
    """BASELINE config 4: one training step on 2 samples of 256x256 per GPU (batch 16 on 8 GPUs), N_trn = 5: differentiable forward
    -> L1 losses with fused gradients -> backward -> gradient all-reduce (NCCL, one flat 29.6 MB bucket) -> Adam."""
    sys.argv = [sys.argv[0], "--steps", str(max(args.steps, 2))]
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_train
    bench_train.main()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=14)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mfi", choices=["mfi", "clip", "train", "4k"],
                    help="mfi: BASELINE's headline (default); clip: config 3 (one clip, strong scaling); train: config 4 (training step); "
                         "4k: config 5 (3840x2160 clip, x16 MFI, N_tst = 5, whole frames, pairs sharded over the ranks)")
    ap.add_argument("--frames", type=int, default=0, help="--workload clip / 4k: frames in the clip (default 64 / 11)")
    ap.add_argument("--balance", type=int, default=1, help="--workload clip: cut the last round's pairs into (pair, t) units")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "clip":
        args.frames = args.frames or 64
        run_clip_workload(args)
    elif args.workload == "4k":
        args.frames = args.frames or 11
        run_clip_workload(args, H0=2160, W0=3840, MFI=16, N_TST=5, label="BASELINE config 5 (4K, whole frames, no tiler)")
    elif args.workload == "train":
        run_train_workload(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
