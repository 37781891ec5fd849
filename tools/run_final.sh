#!/bin/bash
# final pass of the round: tests, bench, per-layer table, ncu launch list of one bench step, ncu --set full of the dominant kernels
T=${1:-r2s}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${T}_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python tools/layer_table.py > gpurun_out/${T}_layer_table.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_s3 -s 4 -c 1 -o gpurun_out/${T}_pair_conv1 -f python tools/bench_conv.py --kinds tc16p --s16 --only "x3 frames, D1" > gpurun_out/${T}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_s3 -s 4 -c 1 -o gpurun_out/${T}_pair_conv2 -f python tools/bench_conv.py --kinds tc16p --s16 --only "conv2 64->64" > gpurun_out/${T}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_s3 -s 4 -c 1 -o gpurun_out/${T}_gru_z -f python tools/bench_conv.py --kinds tc16p --s16 --only "GRU z 128->64 1x5 sigmoid" > gpurun_out/${T}_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"pwb_kernel|bwarp_blend|fgac|cfr_" -c 12 -o gpurun_out/${T}_ops -f python tools/layer_table.py > gpurun_out/${T}_ncu3.log 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e_reference_boundary"]["value"], d["clocks"], d["gpu_launches"])
PY
ls -la gpurun_out/${T}* | awk '{print $5, $9}'
