#!/bin/bash
# A/B of engine / library switches inside ONE box (clocks differ between boxes): ms per step and per layer group
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$name.json"))
g = d["roofline"]["conv_ms_per_step_by_layer_group"]
print("$name", d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], {k: v["ms"] for k, v in g.items()})
PY
}
run base DEMFI_X=0
run flush10 DEMFI_OPTS=tc_flush=10
run nopair DEMFI_PAIR=0
run nopair_flush10 DEMFI_PAIR=0 DEMFI_OPTS=tc_flush=10
