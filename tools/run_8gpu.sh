#!/bin/bash
# 8-GPU pass: BASELINE config 3 (one clip, strong scaling; with and without the balanced tail) and config 4 (training step with NCCL all-reduce)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --workload clip --gpus 8 > gpurun_out/r2_clip_8gpu.json 2> gpurun_out/r2_clip_8gpu.err; tail -1 gpurun_out/r2_clip_8gpu.json
$TR bench.py --workload clip --gpus 8 --balance 0 > gpurun_out/r2_clip_8gpu_unbalanced.json 2> gpurun_out/r2_clip_8gpu_unbalanced.err; tail -1 gpurun_out/r2_clip_8gpu_unbalanced.json
$TR bench.py --workload train --gpus 8 --steps 3 > gpurun_out/r2_train_8gpu.json 2> gpurun_out/r2_train_8gpu.err; tail -1 gpurun_out/r2_train_8gpu.json
python -m pytest tests/test_grad_gpu.py tests/test_train_net_gpu.py -m gpu -x -q 2>&1 | tail -1
