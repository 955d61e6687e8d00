#!/bin/bash
mkdir -p gpurun_out/final2
cd "$(dirname "$0")/.."
timeout 600 python bench.py > gpurun_out/final2/bench_config3.json 2> gpurun_out/final2/bench_config3.err
timeout 600 python bench.py --workload midi --no-cpu-baseline --steps 3 > gpurun_out/final2/bench_midi.json 2> gpurun_out/final2/bench_midi.err
timeout 600 python bench.py --workload vowel --voices 256 --no-cpu-baseline --steps 20 > gpurun_out/final2/bench_vowel.json 2> gpurun_out/final2/bench_vowel.err
NVSP_SCHED=rounds timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/final2/bench_config3_rounds.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final2/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final2/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:klatt_f32_sched_kernel -s 3 -c 1 -o gpurun_out/final2/prof_sched python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final2/prof_sched.log 2>&1
for f in config3 midi vowel config3_rounds; do echo "== $f"; cut -c1-300 gpurun_out/final2/bench_$f.json; done
tail -2 gpurun_out/final2/prof_sched.log | cut -c1-200
