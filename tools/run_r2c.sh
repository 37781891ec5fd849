#!/bin/bash
# round-2 GPU pass: per-box epilogue plans / wide N / dense-block push form
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -x -q > gpurun_out/r2c_conv.log 2>&1; echo "conv rc=$?"; tail -5 gpurun_out/r2c_conv.log
timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_full_size_parity.py -m gpu -x -q > gpurun_out/r2c_fwd.log 2>&1; echo "fwd rc=$?"; tail -5 gpurun_out/r2c_fwd.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c_bench.json"))
    print(d["value"], d["ms_per_step"], d["roofline"]["frac"], json.dumps(d["roofline"]["conv_ms_per_step_by_layer_group"]))
except Exception as e:
    print("bench parse failed", e)
PY
