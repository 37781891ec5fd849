#!/bin/bash
# round-2 GPU pass: full-size golden parity, L2-prefetch A/B on the conv shapes, role timers
mkdir -p gpurun_out
python -m pytest tests/test_full_size_parity.py -m gpu -x -q -s > gpurun_out/r2_parity_full.log 2>&1; echo "parity rc=$?"
for pf in 0 2 4; do
  echo "== tc_prefetch=$pf" >> gpurun_out/r2_bench_conv_pf.log
  python tools/bench_conv.py --kinds tc16 --s16 --opts tc_prefetch=$pf >> gpurun_out/r2_bench_conv_pf.log 2>&1
done
python tools/diag_timers.py > gpurun_out/r2_diag_timers_pf.log 2>&1
tail -5 gpurun_out/r2_parity_full.log
cat gpurun_out/r2_bench_conv_pf.log
