#!/bin/bash
# last check of the round: the whole GPU suite and the bench line on the final tree; ncu of the mma wgrad kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2x_tests.log
timeout 600 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_mma -s 2 -c 1 -o gpurun_out/r2x_wgrad_mma -f python tools/bench_grad.py > gpurun_out/r2x_ncu_wgrad.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2x_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e_reference_boundary"]["value"], d["clocks"], d["gpu_launches"])
PY
