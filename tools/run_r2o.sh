#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/bench_conv.py --kinds tc16p --s16 --only "GRU" 2>&1 | grep -v "^$" | tail -6
timeout 100 python tools/bench_conv.py --kinds tc16p --s16 --only "one source" 2>&1 | tail -1
timeout 100 python tools/bench_conv.py --kinds tc16p --s16 --only "5x5" 2>&1 | tail -1
timeout 100 python tools/bench_conv.py --kinds tc16p --s16 --only "x1 frame" 2>&1 | tail -2
timeout 300 python tools/role_timers_net.py "GB\.conv|Ch_Reducer|UPNet.2.F0|dec3.rF0|conv_ref_k" > gpurun_out/r2o_role_timers.jsonl 2>&1; echo "timers rc=$?"
