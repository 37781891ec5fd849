#!/bin/bash
nproc; uptime
for i in 1 2; do DEMFI_TRAIN_VERBOSE=1 timeout 300 python bench.py --workload train --steps 8 2>&1 | tail -10; done
