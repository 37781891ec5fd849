#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_forward_gpu.py tests/test_full_size_parity.py tests/test_clip_gpu.py -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2d_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2d_bench.json"))
    print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e_reference_boundary"]["value"], d["cpu_baseline"]["value"])
    print(json.dumps(d["roofline"]["hbm_bound_kernels"]))
except Exception as e:
    print("bench parse failed", e)
PY
