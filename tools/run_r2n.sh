#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -q -x -k "pair or repeated or s16" > gpurun_out/r2n_tests.log 2>&1; echo "conv tests rc=$?"; tail -5 gpurun_out/r2n_tests.log
for d in 32768 0 32768 0; do
  timeout 120 python tools/bench_conv.py --kinds tc16p --s16 --only "conv2 64->64" --opts tc_diag=$d 2>&1 | tail -1
done
timeout 200 python tools/role_timers_net.py "conv2|conv1" > gpurun_out/r2n_role_timers.jsonl 2>&1; echo "timers rc=$?"
timeout 300 python tools/time_forward.py --kinds auto --iters 8 2>&1 | tail -2
timeout 300 python tools/opt_ab.py 2>&1 | tail -8
