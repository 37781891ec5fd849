mkdir -p gpurun_out
timeout 120 tools/mma_probe_pair > gpurun_out/r2_probe_pair.log 2>&1
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r2_probe_pair_clocks.csv &
SMI=$!
timeout 120 tools/mma_probe_pair sustained > gpurun_out/r2_probe_pair_sustained.log 2>&1
kill $SMI
cat gpurun_out/r2_probe_pair.log
