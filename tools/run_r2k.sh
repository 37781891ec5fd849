#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_forward_gpu.py tests/test_full_size_parity.py -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2k_tests.log
python tools/bench_conv.py --kinds tc16,tc16p --s16 --only resblock
python tools/bench_conv.py --kinds tc16p --s16 --only resblock --opts tc_diag=8192
run() { name=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$name.json"))
g = d["roofline"]["conv_ms_per_step_by_layer_group"]
print("$name", d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], {k: v["ms"] for k, v in g.items()})
PY
}
run sep DEMFI_X=1
run nosep DEMFI_OPTS=tc_diag=8192
