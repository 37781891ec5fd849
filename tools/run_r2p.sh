#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2p_tests.log
for d in 65536 0; do
  timeout 200 python tools/bench_conv.py --kinds tc16p --s16 --only "sigmoid" --opts tc_diag=$d 2>&1 | tail -2
done
timeout 300 python tools/opt_ab.py tc_diag 65536 0 2>&1 | tail -7
timeout 300 python tools/layer_table.py > gpurun_out/r2p_layer_table.txt 2>&1; head -12 gpurun_out/r2p_layer_table.txt
