mkdir -p gpurun_out
: > gpurun_out/r2_flush_sweep.log
for f in 10 18 36 1000; do
  echo "== tc_flush=$f" >> gpurun_out/r2_flush_sweep.log
  python tools/bench_conv.py --kinds tc16 --s16 --opts tc_flush=$f >> gpurun_out/r2_flush_sweep.log 2>&1
done
cat gpurun_out/r2_flush_sweep.log
