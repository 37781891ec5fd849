#!/bin/bash
# BASELINE config 5 on the 8 GPUs of one box: 8 frame pairs of a 3840x2160 clip x 15 time indices, one pair per rank
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --workload 4k > gpurun_out/r2_4k_8gpu.json 2> gpurun_out/r2_4k_8gpu.err; echo "4k x8 rc=$?"
cat gpurun_out/r2_4k_8gpu.json; tail -3 gpurun_out/r2_4k_8gpu.err
