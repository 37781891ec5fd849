"""BASELINE config 4 per GPU (2 samples of 256x256, N_trn = 5): one training step = differentiable forward
(train_net.forward_train) -> L1 losses with fused gradients -> backward -> gradient all-reduce (when launched under torchrun)
-> Adam.  CUDA-event time per step and its split.  NOT yet run on a GPU (written at the end of round 1, after the GPU budget
was spent): first thing to run in round 2, together with DEMFI_TRAIN_E2E=1 pytest tests/test_train_net_gpu.py.

    python tools/bench_train.py [--size 256] [--batch 2] [--n-trn 5] [--steps 3]
    DEMFI_GRAD_PACK_CACHE=1 python tools/bench_train.py        # weights packed once per optimizer step instead of per call
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_train.py
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from demfi_b200 import synth, train, train_net
from demfi_b200.DeMFInet import DeMFInet


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--n-trn", type=int, default=5)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = DeMFInet(synth.default_args(gpu=local)).to(dev)
    model.load_state_dict(synth.make_state_dict(0))
    opt = train.Adam(model.parameters(), lr=1e-4)
    x = synth.make_frames(args.size, args.size, seed=rank, batch=args.batch).to(dev)
    gt = synth.make_frames(args.size, args.size, seed=100 + rank, batch=args.batch).to(dev)
    gts = [gt[:, :, i].contiguous() for i in range(3)]
    t = torch.rand(args.batch, 1, generator=torch.Generator().manual_seed(rank)).clamp(0.125, 0.875).to(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    rows = []
    WARM = 2                                               # steps 0-1 = warm-up (allocator growth, first packing of every weight)
    for step in range(args.steps + WARM):
        e = [ev() for _ in range(5)]
        t_host = time.perf_counter()
        opt.zero_grad()
        e[0].record()
        res = train_net.forward_train(model, x, t, args.n_trn)
        e[1].record()
        total, d1, d2, g_prime, g_final = train.rec_losses(res[0], res[1], *gts, with_grads=True)
        torch.autograd.backward(list(res[0]) + [s for tri in res[1] for s in tri], list(g_prime) + [g for tri in g_final for g in tri])
        e[2].record()
        nbytes = train.allreduce_gradients(list(model.parameters())) if world > 1 else 0
        e[3].record()
        opt.step()
        e[4].record()
        torch.cuda.synchronize(dev)
        if step >= WARM:
            rows.append([e[i].elapsed_time(e[i + 1]) for i in range(4)] + [total])
            if os.environ.get("DEMFI_TRAIN_VERBOSE") == "1" and rank == 0:
                print("step", step, [round(v, 1) for v in rows[-1][:4]], "host s", round(time.perf_counter() - t_host, 3), flush=True)
    if world > 1:
        worst = torch.tensor([sum(r[:4]) for r in rows], device=dev).mean().reshape(1)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        step_ms = float(worst)
    else:
        step_ms = sum(sum(r[:4]) for r in rows) / len(rows)
    if rank == 0:
        mean = lambda i: sum(r[i] for r in rows) / len(rows)
        print(json.dumps({"config": f"{args.batch} x {args.size}x{args.size} per GPU, N_trn={args.n_trn}, {world} GPU(s)",
                          "ms_per_step": round(step_ms, 2), "samples_per_s": round(world * args.batch / step_ms * 1e3, 3),
                          "forward_ms": round(mean(0), 2), "loss_backward_ms": round(mean(1), 2), "allreduce_ms": round(mean(2), 3),
                          "adam_ms": round(mean(3), 3), "allreduce_bytes": nbytes, "last_loss": rows[-1][4]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
