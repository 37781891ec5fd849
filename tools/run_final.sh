#!/bin/bash
# final pass of the round: tests, bench, per-layer table, ncu launch list of one bench step, ncu --set full of the dominant kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_final_tests.log
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"
python tools/layer_table.py > gpurun_out/r2_final_layer_table.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_final_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_s3 -s 4 -c 1 -o gpurun_out/r2_final_pair_conv1 -f python tools/bench_conv.py --kinds tc16p --s16 --only "x3 frames, D1" > gpurun_out/r2_final_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_s3 -s 4 -c 1 -o gpurun_out/r2_final_pair_conv2 -f python tools/bench_conv.py --kinds tc16p --s16 --only "conv2 64->64" > gpurun_out/r2_final_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"pwb_kernel|bwarp_blend|fgac|cfr_" -c 12 -o gpurun_out/r2_final_ops -f python tools/layer_table.py > gpurun_out/r2_final_ncu3.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_final_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e_reference_boundary"]["value"], d["clocks"], d["gpu_launches"])
PY
ls -la gpurun_out/r2_final* | awk '{print $5, $9}'
