#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2r_prefetch.log
for pf in 0 1 2 4; do
  echo "== tc_prefetch=$pf" >> gpurun_out/r2r_prefetch.log
  timeout 200 python tools/bench_conv.py --kinds tc16p --s16 --opts tc_prefetch=$pf --only "x3 frames" 2>&1 | grep conv >> gpurun_out/r2r_prefetch.log
  timeout 200 python tools/bench_conv.py --kinds tc16p --s16 --opts tc_prefetch=$pf --only "sigmoid" 2>&1 | grep conv >> gpurun_out/r2r_prefetch.log
  timeout 200 python tools/bench_conv.py --kinds tc16 --s16 --opts tc_prefetch=$pf --only "LFF" 2>&1 | grep conv >> gpurun_out/r2r_prefetch.log
done
cat gpurun_out/r2r_prefetch.log | cut -c1-150
timeout 300 python tools/opt_ab.py tc_prefetch 0 2 2>&1 | tail -7
timeout 300 python tools/opt_ab.py tc_prefetch 0 1 2>&1 | tail -6
