#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2j_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2j_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2j_bench.json"))
    print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"])
    print(json.dumps(d["roofline"]["conv_ms_per_step_by_layer_group"]))
    print(json.dumps(d["roofline"]["other_kernels_ms_per_step"]))
except Exception as e:
    print("bench parse failed", e)
PY
python tools/layer_table.py > gpurun_out/r2j_layer_table.txt 2>&1; head -40 gpurun_out/r2j_layer_table.txt
