#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_full_size_parity.py -m gpu -q -x > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2z_tests.log
timeout 300 python tools/hoist_ab.py 2>&1 | tail -8
