#!/bin/bash
mkdir -p gpurun_out/final3
cd "$(dirname "$0")/.."
timeout 600 python bench.py > gpurun_out/final3/bench_config3.json 2> gpurun_out/final3/bench_config3.err
timeout 600 python bench.py --workload midi --no-cpu-baseline --steps 3 > gpurun_out/final3/bench_midi.json 2> gpurun_out/final3/bench_midi.err
timeout 600 python bench.py --workload vowel --voices 256 --no-cpu-baseline --steps 20 > gpurun_out/final3/bench_vowel.json 2> gpurun_out/final3/bench_vowel.err
for f in config3 midi vowel; do echo "== $f"; cut -c1-260 gpurun_out/final3/bench_$f.json; done
