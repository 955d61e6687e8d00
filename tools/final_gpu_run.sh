#!/bin/bash
# GPU box: the measurements that are copied into profiles/ (run under gpurun from the repo root)
mkdir -p gpurun_out/final
cd "$(dirname "$0")/.."
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/final/pytest_gpu.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -6) > gpurun_out/final/smoke.log
timeout 600 python bench.py > gpurun_out/final/bench_config3.json 2> gpurun_out/final/bench_config3.err
timeout 600 python bench.py --impl reference > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
timeout 600 python bench.py --workload long --no-cpu-baseline > gpurun_out/final/bench_long.json 2> gpurun_out/final/bench_long.err
timeout 900 python bench.py --workload midi --no-cpu-baseline --steps 3 > gpurun_out/final/bench_midi.json 2> gpurun_out/final/bench_midi.err
timeout 900 python bench.py --workload vowel --voices 256 --no-cpu-baseline --steps 3 > gpurun_out/final/bench_vowel.json 2> gpurun_out/final/bench_vowel.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final/launches_bench.log 2>&1
cat gpurun_out/final/pytest_gpu.log gpurun_out/final/smoke.log
for f in config3 reference long midi vowel; do echo "== $f"; cut -c1-700 gpurun_out/final/bench_$f.json; tail -2 gpurun_out/final/bench_$f.err; done
