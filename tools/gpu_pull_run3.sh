#!/bin/bash
# GPU box helper: the run decomposition of the phase recurrence (NVSP_PULL_PHASE=runs) on the device
cd "$(dirname "$0")/.."
O=gpurun_out/pull3
mkdir -p $O
NVSP_PULL_PHASE=runs timeout 25 python -m pytest tests/test_gpu_pull.py -q -m gpu > $O/pytest_pull_runs.log 2>&1; echo "pull tests (runs) rc=$?"; tail -2 $O/pytest_pull_runs.log
NVSP_PULL_PHASE=runs NVSP_PULL_DEBUG=1 NVSP_PROBE_ONLY=stream timeout 15 python tools/latency_probe.py > $O/latency_runs.txt 2> $O/phases_runs.txt; cat $O/latency_runs.txt; grep -m3 "n=8192" $O/phases_runs.txt | tail -1; grep "n=2048" $O/phases_runs.txt | sed -n 5,5p
