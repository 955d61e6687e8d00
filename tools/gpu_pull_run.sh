#!/bin/bash
# GPU box helper: checks of the low-latency pull path, then the whole GPU suite, a short bench and the pull kernel's launch times
cd "$(dirname "$0")/.."
O=gpurun_out/pull
mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_pull.py -q -m gpu > $O/pytest_pull.log 2>&1; echo "pull tests rc=$?"; tail -3 $O/pytest_pull.log
timeout 60 python tools/latency_probe.py > $O/latency.txt 2> $O/latency.err; echo "probe rc=$?"; cat $O/latency.txt
timeout 120 python -m pytest tests -q -m gpu > $O/pytest_all.log 2>&1; echo "all tests rc=$?"; tail -2 $O/pytest_all.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
timeout 120 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > $O/bench_config3.json 2> $O/bench_config3.err; echo "bench rc=$?"; cut -c1-400 $O/bench_config3.json
NVSP_PROBE_ONLY=stream timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:klatt_pull_kernel -c 60 --csv --log-file $O/pull_launches.csv python tools/latency_probe.py > $O/ncu_probe.txt 2>&1; echo "ncu rc=$?"; tail -5 $O/pull_launches.csv
