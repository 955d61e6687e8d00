#!/bin/bash
# usage: tools/gpu_try.sh "<env assignments>" [bench args...]   -> one compact line per run (GPU box helper)
envs="$1"; shift
out=$(env $envs timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 "$@" 2>gpurun_out/try_err.txt)
echo "$envs :: $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('ms/step %.1f  audio-s/s %.0f  frac %.4f  launches %d' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['gpu_launches']))
except Exception as e:
    print('FAILED', e)
")"
grep -h "scheduler\|rror" gpurun_out/try_err.txt | head -3
