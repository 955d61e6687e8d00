#!/bin/bash
# GPU box helper: the optimised pull kernel -- tests, per-phase cycle counts, latency with and without zero-copy
cd "$(dirname "$0")/.."
O=gpurun_out/pull2
mkdir -p $O
timeout 100 python -m pytest tests/test_gpu_pull.py -q -m gpu > $O/pytest_pull.log 2>&1; echo "pull tests rc=$?"; tail -3 $O/pytest_pull.log
NVSP_PROBE_ONLY=stream timeout 60 python tools/latency_probe.py > $O/latency_zerocopy.txt 2> $O/latency.err; echo "probe rc=$?"; cat $O/latency_zerocopy.txt
NVSP_PULL_ZEROCOPY=0 NVSP_PROBE_ONLY=stream timeout 60 python tools/latency_probe.py > $O/latency_copies.txt 2>> $O/latency.err; echo "probe rc=$?"; cat $O/latency_copies.txt
NVSP_PULL_DEBUG=1 NVSP_PROBE_ONLY=stream timeout 60 python tools/latency_probe.py > /dev/null 2> $O/phases.txt; grep -m3 "n=8192" $O/phases.txt | tail -2; grep "n=2048" $O/phases.txt | sed -n 5,6p; grep "n=512" $O/phases.txt | sed -n 5,5p
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
