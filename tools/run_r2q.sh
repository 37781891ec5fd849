#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2q_tests.log
timeout 600 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2q_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"], d["gpu_launches"])
print(d["roofline"]["conv_ms_per_step_by_layer_group"])
PY
timeout 600 python bench.py --workload 4k > gpurun_out/r2q_4k_1gpu.json 2> gpurun_out/r2q_4k_1gpu.err; echo "4k rc=$?"; cat gpurun_out/r2q_4k_1gpu.json; tail -3 gpurun_out/r2q_4k_1gpu.err
