#!/bin/bash
# GPU box helper, last call of round 1: whole GPU suite + smoke on the final build, pull-path latency (with the reference
# on a host core beside it), the pull bench line, per-phase cycles and one full ncu capture of klatt_pull_kernel
cd "$(dirname "$0")/.."
O=gpurun_out/final4
mkdir -p $O
timeout 100 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -2 $O/pytest_gpu.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 40 python tools/latency_probe.py > $O/latency.txt 2> $O/latency.err; echo "probe rc=$?"; cat $O/latency.txt
timeout 60 python bench.py --workload pull > $O/bench_pull.json 2> $O/bench_pull.err; echo "bench pull rc=$?"; cut -c1-300 $O/bench_pull.json; tail -2 $O/bench_pull.err
NVSP_PULL_DEBUG=1 NVSP_PROBE_ONLY=stream timeout 30 python tools/latency_probe.py > /dev/null 2> $O/phases.txt; grep -m4 "n=8192" $O/phases.txt | tail -2; grep "n=2048" $O/phases.txt | sed -n 5,5p; grep "n=512" $O/phases.txt | sed -n 5,5p
NVSP_PROBE_ONLY=stream timeout 60 ncu --set full --clock-control none --import-source on -k regex:klatt_pull_kernel -s 3 -c 2 -f -o $O/prof_pull python tools/latency_probe.py > $O/ncu.log 2>&1; echo "ncu rc=$?"; ls -la $O/prof_pull.ncu-rep
