#!/bin/bash
mkdir -p gpurun_out
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
run pair1_nolean DEMFI_PAIR=1 DEMFI_OPTS=tc_diag=4096
