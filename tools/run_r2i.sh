#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_forward_gpu.py -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2i_tests.log
python tools/bench_conv.py --kinds tc16,tc16p --s16 --only resblock
python tools/bench_conv.py --kinds tc16,tc16p --s16 --only resblock --opts tc_diag=4096
python tools/diag_pair.py 2>&1 | cut -c1-30,150-1100
