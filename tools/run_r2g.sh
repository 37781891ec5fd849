#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_forward_gpu.py -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2g_tests.log
python tools/bench_conv.py --kinds tc16,tc16p --s16 --only resblock
python tools/diag_pair.py 2>&1 | cut -c1-30,150-1100
run() { name=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$name.json"))
g = d["roofline"]["conv_ms_per_step_by_layer_group"]
print("$name", d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], {k: v["ms"] for k, v in g.items()})
PY
}
run pair1 DEMFI_PAIR=1
run pair2 DEMFI_PAIR=2
run pair1_nooff DEMFI_PAIR=1 DEMFI_OPTS=tc_diag=2048
