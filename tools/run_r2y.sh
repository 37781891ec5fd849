#!/bin/bash
timeout 300 python -m pytest tests/test_grad_gpu.py tests/test_train_net_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 200 python tools/bench_grad.py 2>&1 | grep layer | cut -c1-260
timeout 300 python bench.py --workload train --steps 8 2>&1 | tail -1
