#!/bin/bash
# config 4 on 8 GPUs after the device-side weight packing and the mma wgrad
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --workload train --steps 8 > gpurun_out/r2w_train_8gpu.json 2> gpurun_out/r2w_train_8gpu.err; echo "train x8 rc=$?"
tail -2 gpurun_out/r2w_train_8gpu.json; tail -3 gpurun_out/r2w_train_8gpu.err
