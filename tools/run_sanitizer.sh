#!/bin/bash
# compute-sanitizer on the kernels added in the last session of round 2
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 400 $S --tool memcheck python -m pytest tests/test_conv_gpu.py -m gpu -q -k "two_tiles_in_turn and 48-24 or device_weight_packing or gru_gates_on_cta_pairs" > gpurun_out/r2_sanitizer_memcheck_conv.log 2>&1; echo "memcheck conv rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_memcheck_conv.log | tail -3
timeout 300 $S --tool memcheck python -m pytest tests/test_grad_gpu.py -m gpu -q -k "layer_gradients and wgrad-mma and (64-64 or 5-32 or 96-133)" > gpurun_out/r2_sanitizer_memcheck_wgrad.log 2>&1; echo "memcheck wgrad rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_memcheck_wgrad.log | tail -3
timeout 300 $S --tool racecheck python -m pytest tests/test_grad_gpu.py -m gpu -q -k "layer_gradients and wgrad-mma and 64-64" > gpurun_out/r2_sanitizer_racecheck_wgrad.log 2>&1; echo "racecheck wgrad rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_racecheck_wgrad.log | tail -3
