#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/graph_ab.py > gpurun_out/r2m_graph_ab.log 2>&1; echo "graph rc=$?"; cat gpurun_out/r2m_graph_ab.log | tail -6
timeout 400 python tools/role_timers_net.py > gpurun_out/r2m_role_timers_net.jsonl 2> gpurun_out/r2m_role_timers_net.err; echo "timers rc=$?"; tail -3 gpurun_out/r2m_role_timers_net.err; wc -l gpurun_out/r2m_role_timers_net.jsonl
