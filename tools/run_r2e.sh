#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -x -q -k "cta_pair" > gpurun_out/r2e_pair.log 2>&1; echo "pair rc=$?"; grep -v "timed out" gpurun_out/r2e_pair.log | tail -25
