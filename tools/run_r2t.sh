#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_grad_gpu.py tests/test_train_net_gpu.py -m gpu -q -x -k "device_weight or grad or train" > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2t_tests.log
DEMFI_GRAD_HOST_PACK=1 timeout 300 python bench.py --workload train --steps 3 2>&1 | tail -1
timeout 300 python bench.py --workload train --steps 3 2>&1 | tail -1 | tee gpurun_out/r2t_train_1gpu.json
