#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_grad_gpu.py tests/test_train_net_gpu.py -m gpu -q -x -s > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"; grep "dW" gpurun_out/r2u_tests.log | head -14; tail -3 gpurun_out/r2u_tests.log
timeout 200 python tools/bench_grad.py 2>&1 | tail -12
timeout 300 python bench.py --workload train --steps 3 2>&1 | tail -1 | tee gpurun_out/r2u_train_1gpu.json
