#!/bin/bash
# GPU box helper: tools/gpu_run.sh <stage> [args].  One script for every gpurun call of the round; outputs under gpurun_out/<stage>/.
#   try "<env>" [bench args]   one compact line for a short bench run (A/B of builds: NVSP_LIB=tools/_variants/libX.so)
cd "$(dirname "$0")/.."
stage="$1"; shift
O=gpurun_out/$stage
mkdir -p $O
try() {  # try "<env assignments>" [bench args...]
	local envs="$1"; shift
	local out
	out=$(env $envs timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 "$@" 2>$O/try_err.txt)
	echo "$envs $* :: $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('ms/step %.2f  audio-s/s %.0f  frac %.4f  launches %d' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['gpu_launches']))
except Exception as e:
    print('FAILED', e)
")"
	grep -h "sched profile\|block profile\|rror" $O/try_err.txt | head -4
}
case "$stage" in
c1)  # baseline data of round 2: GPU suite, A/B of cheap scheduler variants, ncu source-level capture, kernel (b) captures, D2H probe
	timeout 400 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/pytest_gpu.log
	try "NVSP_X=base"
	try "NVSP_LIB=$PWD/tools/_variants/libnoflip.so"
	try "NVSP_SCHED_HOLD_SMS=44"
	try "NVSP_SCHED_HOLD_SMS=52"
	try "NVSP_SCHED_HOLD_SMS=60"
	try "NVSP_LIB=$PWD/tools/_variants/libprof.so"
	timeout 300 ncu --set full --clock-control none --import-source on -k regex:klatt_f32_sched_kernel -s 3 -c 1 -f -o $O/prof_sched \
		python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_sched.log 2>&1; echo "ncu sched rc=$?"; ls -la $O/prof_sched.ncu-rep
	timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/long_launches.csv \
		python bench.py --workload long --steps 1 --warmup 1 --no-cpu-baseline > $O/long_launches.log 2>&1; echo "ncu long list rc=$?"
	timeout 300 ncu --set full --clock-control none --import-source on -k regex:klatt_long_stage -s 40 -c 3 -f -o $O/prof_long \
		python bench.py --workload long --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_long.log 2>&1; echo "ncu long rc=$?"; ls -la $O/prof_long.ncu-rep
	timeout 120 tools/d2h_probe 1 1024 6 > $O/d2h_1gpu.json 2>&1; cat $O/d2h_1gpu.json
	;;
c2)  # parity round: whole GPU suite with the new tests, smoke, the long bench line with its full-hour oracle check
	timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -15 $O/pytest_gpu.log
	timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -8 $O/smoke.log
	timeout 400 python bench.py --workload long --steps 3 --warmup 3 > $O/bench_long.json 2> $O/bench_long.err; echo "bench long rc=$?"; cut -c1-1500 $O/bench_long.json; tail -3 $O/bench_long.err
	;;
san)  # compute-sanitizer on the long path (short stream)
	timeout 300 compute-sanitizer --tool memcheck --print-limit 8 python -c "
import numpy as np
from nvspeechplayer_b200 import player, workloads
fr, m, f, nul, ux = workloads.random_stream(9, 0.5, 22050)
out = player.synthesize_long(22050, fr, m, f, nul, seed=4, stream_id=9, chunk_ticks=256)
print('ok', len(out[0]))
" > $O/san.log 2>&1; echo "sanitizer rc=$?"; grep -v "^=========     at\|^=========         in" $O/san.log | head -60
	;;
c3)  # block scheduler first light: suite, smoke, A/B against the ring scheduler, cell-length sweep, in-kernel cycle profile
	timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -12 $O/pytest_gpu.log
	timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -9 $O/smoke.log
	try "NVSP_SCHED=block"
	try "NVSP_SCHED=rings"
	try "NVSP_BLOCK_HOLD_TICKS=64"
	try "NVSP_BLOCK_HOLD_TICKS=256"
	try "NVSP_LIB=$PWD/tools/_variants/libbprof.so"
	timeout 400 python bench.py --workload long --steps 3 --warmup 3 > $O/bench_long.json 2> $O/bench_long.err; echo "bench long rc=$?"; cut -c1-1800 $O/bench_long.json; tail -3 $O/bench_long.err
	;;
c4)  # block scheduler with staged records: bitwise test, A/B, cycle profile, ncu capture; kernel (b) after the timeline / tile fixes
	timeout 600 python -m pytest tests/test_gpu_parity_f32.py tests/test_gpu_long.py -q -m gpu > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -5 $O/pytest_gpu.log
	try "NVSP_SCHED=block"
	try "NVSP_SCHED=rings"
	try "NVSP_BLOCK_HOLD_TICKS=64"
	try "NVSP_BLOCK_HOLD_TICKS=256"
	try "NVSP_LIB=$PWD/tools/_variants/libbprof.so"
	timeout 300 ncu --set full --clock-control none --import-source on -k regex:klatt_f32_block_kernel -s 3 -c 1 -f -o $O/prof_block \
		python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_block.log 2>&1; echo "ncu block rc=$?"; ls -la $O/prof_block.ncu-rep
	timeout 400 python bench.py --workload long --steps 3 --warmup 3 > $O/bench_long.json 2> $O/bench_long.err; echo "bench long rc=$?"; cut -c1-400 $O/bench_long.json; tail -3 $O/bench_long.err
	timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/long_launches.csv \
		python bench.py --workload long --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/long_launches.log 2>&1; echo "ncu long list rc=$?"
	;;
c5)  # block scheduler: how many classes (loop pairs) may be in flight per SM
	timeout 300 python -m pytest tests/test_gpu_parity_f32.py -q -m gpu -k "rounds_equal or batch_random" > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/pytest_gpu.log
	try "NVSP_BLOCK_MAX_CLASSES=2"
	try "NVSP_BLOCK_MAX_CLASSES=1"
	try "NVSP_BLOCK_MAX_CLASSES=3"
	try "NVSP_BLOCK_MAX_CLASSES=2 NVSP_LIB=$PWD/tools/_variants/libbprof.so"
	try "NVSP_BLOCK_MAX_CLASSES=1 NVSP_LIB=$PWD/tools/_variants/libbprof.so"
	timeout 300 ncu --set full --clock-control none --import-source on -k regex:klatt_f32_block_kernel -s 3 -c 1 -f -o $O/prof_block \
		python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_block.log 2>&1; echo "ncu block rc=$?"; ls -la $O/prof_block.ncu-rep
	;;
c6)  # completeness round: whole suite (multi-batch, sinks, batched pull, kernel b), kernel (b) timings + captures, pull bench
	timeout 1200 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -8 $O/pytest_gpu.log
	timeout 400 python bench.py --workload long --steps 3 --warmup 3 > $O/bench_long.json 2> $O/bench_long.err; echo "bench long rc=$?"; cut -c1-300 $O/bench_long.json; tail -3 $O/bench_long.err
	timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/long_launches.csv \
		python bench.py --workload long --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/long_launches.log 2>&1; echo "ncu long list rc=$?"
	timeout 300 ncu --set full --clock-control none --import-source on -k regex:klatt_long_stage -s 60 -c 3 -f -o $O/prof_long \
		python bench.py --workload long --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/ncu_long.log 2>&1; echo "ncu long rc=$?"
	timeout 300 python bench.py --workload pull > $O/bench_pull.json 2> $O/bench_pull.err; echo "bench pull rc=$?"; cut -c1-2000 $O/bench_pull.json; tail -3 $O/bench_pull.err
	;;
c7)  # validation after the timeline rewrite and the multi-batch test fix
	timeout 1200 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -5 $O/pytest_gpu.log
	timeout 400 python bench.py --workload long --steps 3 --warmup 3 > $O/bench_long.json 2> $O/bench_long.err; echo "bench long rc=$?"; cut -c1-300 $O/bench_long.json; tail -3 $O/bench_long.err
	timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/long_launches.csv \
		python bench.py --workload long --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/long_launches.log 2>&1; echo "ncu long list rc=$?"
	timeout 200 python tools/multibatch_bench.py --gpus 1 --streams-per-gpu 16384 > $O/multibatch_1gpu.json 2> $O/multibatch.err; echo "multibatch rc=$?"; cat $O/multibatch_1gpu.json; tail -2 $O/multibatch.err
	;;
mg)  # N GPUs of one box (gpurun --gpus N): concurrent D2H ceiling, the in-library multi-GPU line, the torchrun bench
	N=${1:-8}
	for g in 1 2 4 8; do [ $g -le $N ] && { timeout 300 tools/d2h_probe $g 1024 6 > $O/d2h_${g}gpu.json 2>&1; cat $O/d2h_${g}gpu.json; }; done
	timeout 600 python tools/multibatch_bench.py --gpus $N --streams-per-gpu 16384 > $O/multibatch_${N}gpu.json 2> $O/multibatch.err; echo "multibatch rc=$?"; cat $O/multibatch_${N}gpu.json; tail -2 $O/multibatch.err
	timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline \
		> $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "bench x$N rc=$?"; cut -c1-2500 $O/bench_${N}gpu.json; tail -3 $O/bench_${N}gpu.err
	;;
final)  # the round's evidence on one GPU: suite, smoke, the three bench lines, launch lists, full captures
	timeout 1200 python -m pytest tests -q -m gpu -rA > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/pytest_gpu.log
	timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -9 $O/smoke.log
	timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-3000 $O/bench.json; tail -3 $O/bench.err
	timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench ref rc=$?"; cut -c1-600 $O/bench_reference.json
	timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
		python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/launches.log 2>&1; echo "ncu list rc=$?"
	timeout 300 ncu --set full --clock-control none --import-source on -k regex:klatt_f32_sched_kernel -s 3 -c 1 -f -o $O/prof_sched \
		python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_sched.log 2>&1; echo "ncu sched rc=$?"
	timeout 400 python bench.py --workload long --steps 5 --warmup 3 > $O/bench_long.json 2> $O/bench_long.err; echo "bench long rc=$?"; cut -c1-300 $O/bench_long.json; tail -3 $O/bench_long.err
	timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/long_launches.csv \
		python bench.py --workload long --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/long_launches.log 2>&1; echo "ncu long list rc=$?"
	timeout 300 ncu --set full --clock-control none --import-source on -k regex:klatt_long_stage -s 60 -c 3 -f -o $O/prof_long \
		python bench.py --workload long --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/ncu_long.log 2>&1; echo "ncu long rc=$?"
	timeout 300 python bench.py --workload pull > $O/bench_pull.json 2> $O/bench_pull.err; echo "bench pull rc=$?"; cut -c1-300 $O/bench_pull.json; tail -3 $O/bench_pull.err
	timeout 200 python bench.py --workload midi --no-cpu-baseline --no-e2e > $O/bench_midi.json 2> $O/bench_midi.err; echo "bench midi rc=$?"; cut -c1-300 $O/bench_midi.json
	timeout 200 python bench.py --workload vowel --no-cpu-baseline --no-e2e > $O/bench_vowel.json 2> $O/bench_vowel.err; echo "bench vowel rc=$?"; cut -c1-300 $O/bench_vowel.json
	;;
c9)  # kernel (b) after the timeline fix; staging size of the host pipeline
	timeout 600 python -m pytest tests/test_gpu_long.py tests/test_gpu_parity_f32.py tests/test_gpu_configs.py -q -m gpu > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/pytest_gpu.log
	timeout 400 python bench.py --workload long --steps 5 --warmup 3 > $O/bench_long.json 2> $O/bench_long.err; echo "bench long rc=$?"; cut -c1-300 $O/bench_long.json; tail -3 $O/bench_long.err
	timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/long_launches.csv \
		python bench.py --workload long --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/long_launches.log 2>&1; echo "ncu long list rc=$?"
	timeout 300 ncu --set full --clock-control none --import-source on -k regex:klatt_long_stage -s 60 -c 3 -f -o $O/prof_long \
		python bench.py --workload long --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/ncu_long.log 2>&1; echo "ncu long rc=$?"
	for mb in 1024 2048 4096; do
		NVSP_STAGE_MB=$mb timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > $O/bench_stage$mb.json 2> $O/bench_stage$mb.err
		python -c "
import json; d=json.load(open('$O/bench_stage$mb.json')); e=d['e2e']; print('NVSP_STAGE_MB=$mb: device %.1f ms, e2e %.1f ms/step, %.1f GB/s achieved, host ceiling %.1f GB/s' % (d['ms_per_step'], e['ms_per_step'], e['d2h_GBps_achieved'], e['host_d2h_GBps_all_ranks_concurrent']))"
	done
	;;
lg)  # kernel (b) only: tests, bench line, launch list, stage-kernel capture
	timeout 600 python -m pytest tests/test_gpu_long.py -q -m gpu > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/pytest_gpu.log
	timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
	timeout 400 python bench.py --workload long --steps 5 --warmup 3 > $O/bench_long.json 2> $O/bench_long.err; echo "bench long rc=$?"; cut -c1-300 $O/bench_long.json; tail -3 $O/bench_long.err
	timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/long_launches.csv \
		python bench.py --workload long --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/long_launches.log 2>&1; echo "ncu long list rc=$?"
	timeout 300 ncu --set full --clock-control none --import-source on -k regex:klatt_long_stage -s 60 -c 3 -f -o $O/prof_long \
		python bench.py --workload long --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/ncu_long.log 2>&1; echo "ncu long rc=$?"
	;;
rl)  # ring scheduler with compact staged records (KLATT_SCHED_LITE) vs round 1's AoS state with -dlcm=cg
	timeout 600 python -m pytest tests/test_gpu_parity_f32.py tests/test_gpu_configs.py tests/test_gpu_multibatch.py tests/test_gpu_sinks.py -q -m gpu > $O/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/pytest_gpu.log
	timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
	try "NVSP_X=lite"
	try "NVSP_LIB=$PWD/tools/_variants/libnolite.so"
	try "NVSP_X=lite"
	try "NVSP_LIB=$PWD/tools/_variants/libnolite.so"
	try "NVSP_LIB=$PWD/tools/_variants/libprof.so"
	try "NVSP_SCHED_GEN_TICKS=512"
	try "NVSP_SCHED_GEN_TICKS=384 NVSP_SCHED_HOLD_TICKS=384"
	try "NVSP_SCHED_HOLD_TICKS=256 NVSP_SCHED_GEN_TICKS=512"
	timeout 300 ncu --set full --clock-control none --import-source on -k regex:klatt_f32_sched_kernel -s 3 -c 1 -f -o $O/prof_sched \
		python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_sched.log 2>&1; echo "ncu sched rc=$?"
	;;
rl2)  # ring scheduler (compact staged records): chunk-length sweep, streaming output stores
	timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
	for hg in "256 512" "256 384" "256 256" "192 384" "384 512" "128 256" "256 640" "320 448" "192 512" "128 384"; do
		set -- $hg
		try "NVSP_SCHED_HOLD_TICKS=$1 NVSP_SCHED_GEN_TICKS=$2"
	done
	try "NVSP_LIB=$PWD/tools/_variants/libstream.so NVSP_SCHED_HOLD_TICKS=256 NVSP_SCHED_GEN_TICKS=512"
	NVSP_LIB=$PWD/tools/_variants/libstream.so NVSP_SCHED_HOLD_TICKS=256 NVSP_SCHED_GEN_TICKS=512 timeout 300 ncu --set full --clock-control none -k regex:klatt_f32_sched_kernel -s 3 -c 1 -f -o $O/prof_sched_stream \
		python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_sched.log 2>&1; echo "ncu sched rc=$?"
	;;
rl3)  # ring scheduler: new defaults (256/384, streaming output stores, hold write-back trim), L2 persistence window on/off
	timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
	timeout 600 python -m pytest tests -q -m gpu -x -k "parity_f32 or batch_handles or sinks or multibatch" > $O/pytest_sub.log 2>&1; echo "gpu subset rc=$?"; tail -3 $O/pytest_sub.log
	try "NVSP_L2_PERSIST=1"
	try "NVSP_L2_PERSIST=0"
	try "NVSP_L2_PERSIST=1 NVSP_SCHED_HOLD_TICKS=192 NVSP_SCHED_GEN_TICKS=320"
	try "NVSP_L2_PERSIST=1 NVSP_SCHED_HOLD_TICKS=256 NVSP_SCHED_GEN_TICKS=256"
	try "NVSP_L2_PERSIST=1 NVSP_SCHED_HOLD_TICKS=128 NVSP_SCHED_GEN_TICKS=256"
	for pz in 1 0; do
		NVSP_L2_PERSIST=$pz timeout 300 ncu --set full --clock-control none --import-source on -k regex:klatt_f32_sched_kernel -s 3 -c 1 -f -o $O/prof_sched_p$pz \
			python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_sched_p$pz.log 2>&1; echo "ncu sched persist=$pz rc=$?"
	done
	;;
rl4)  # ring scheduler: evict-last prefetch of the fade plans on/off (time, DRAM bytes)
	timeout 300 python -m pytest tests -q -m gpu -x -k "parity_f32 or batch_handles" > $O/pytest_sub.log 2>&1; echo "gpu subset rc=$?"; tail -3 $O/pytest_sub.log
	try "NVSP_X=default"
	try "NVSP_LIB=$PWD/tools/_variants/libnopf.so"
	try "NVSP_X=default"
	try "NVSP_LIB=$PWD/tools/_variants/libnopf.so"
	for v in default nopf; do
		lib=$PWD/nvspeechplayer_b200/libspeechPlayer.so; [ $v = nopf ] && lib=$PWD/tools/_variants/libnopf.so
		NVSP_LIB=$lib timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:klatt_f32_sched_kernel -s 3 -c 1 \
			python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_$v.log 2>&1; echo "ncu $v rc=$?"; grep -A8 "klatt_f32_sched_kernel" $O/ncu_$v.log | grep "dram\|duration\|lts" 
	done
	;;
rl5)  # ring scheduler: hold chunks stretched to the shortest quiet span of their 32 streams
	timeout 300 python -m pytest tests -q -m gpu -x -k "parity_f32 or batch_handles or configs" > $O/pytest_sub.log 2>&1; echo "gpu subset rc=$?"; tail -3 $O/pytest_sub.log
	for hm in 256 512 1024 2048 4096; do
		try "NVSP_SCHED_HOLD_MAX=$hm"
		try "NVSP_SCHED_HOLD_MAX=$hm" --workload vowel
		try "NVSP_SCHED_HOLD_MAX=$hm" --workload midi
	done
	;;
rl6)  # stretched hold chunks: the fixed test, a few thresholds around the default
	timeout 300 python -m pytest tests -q -m gpu -x -k "parity_f32" > $O/pytest_sub.log 2>&1; echo "gpu subset rc=$?"; tail -3 $O/pytest_sub.log
	try "NVSP_SCHED_HOLD_TICKS=192 NVSP_SCHED_GEN_TICKS=384"
	try "NVSP_SCHED_HOLD_TICKS=128 NVSP_SCHED_GEN_TICKS=384"
	try "NVSP_SCHED_HOLD_TICKS=256 NVSP_SCHED_GEN_TICKS=320"
	try "NVSP_SCHED_HOLD_TICKS=128 NVSP_SCHED_GEN_TICKS=256"
	try "NVSP_SCHED_HOLD_TICKS=192 NVSP_SCHED_GEN_TICKS=384" --workload vowel
	try "NVSP_SCHED_HOLD_TICKS=128 NVSP_SCHED_GEN_TICKS=384" --workload midi
	;;
san2)  # compute-sanitizer on the ring scheduler with staged records (memcheck, racecheck), the batched pull and the device sinks
	cat > $O/san_case.py <<'PY'
import os, sys
import numpy as np
os.environ.update({"NVSP_SCHED": "rings", "NVSP_ROUNDS_MIN_STREAMS": "1", "NVSP_SCHED_BLOCKS": "2", "NVSP_SCHED_HOLD_TICKS": "128", "NVSP_SCHED_GEN_TICKS": "128"})
from nvspeechplayer_b200 import player, workloads
sr = 16000
vc = workloads.vowel_chart(3, sr)
pick = list(range(0, len(vc.stream_ids), 211))[:24]
streams = [vc.stream(s) for s in pick] + [workloads.random_stream(900 + s, 0.4, sr) for s in range(43)]
ids = np.concatenate([vc.stream_ids[pick], np.arange(900, 943, dtype=np.uint64)])
fb = workloads._concat(sr, streams, ids)
b = player.Batch(sr, len(ids), precision=player.PRECISION_FP32, seed=5, stream_ids=fb.stream_ids)
b.set_frames_host(fb)
tot = 0
for c in (3001, 2000):
    o, w = b.synthesize_host(c)
    tot += int(w.sum())
print("ring scheduler ok", tot, b.launch_stats())
b.close()
if "pull" in sys.argv:
    ps = [player.SpeechPlayer(sr, precision=player.PRECISION_STREAM, noise=player.NOISE_PHILOX, seed=3, streamId=100 + i) for i in range(5)]
    for i, p in enumerate(ps):
        fr, m, f, nul, ux = workloads.random_stream(100 + i, 0.3, sr)
        p.queue_frames(fr, m, f, ux, nul)
    out = player.synthesize_batch(ps, 1024) if hasattr(player, "synthesize_batch") else None
    print("pull batch ok", None if out is None else np.asarray(out[0]).shape)
    for p in ps:
        p.close()
PY
	timeout 500 env PYTHONPATH=$PWD compute-sanitizer --tool memcheck --print-limit 8 python $O/san_case.py pull > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -v "^=========     at\|^=========         in" $O/memcheck.log | tail -12
	timeout 700 env PYTHONPATH=$PWD compute-sanitizer --tool racecheck --print-limit 8 python $O/san_case.py > $O/racecheck.log 2>&1; echo "racecheck rc=$?"; grep -v "^=========     at\|^=========         in" $O/racecheck.log | tail -12
	;;
rl7)  # ring scheduler: L2 evict-last policy on the plan loads of the general loop, on / off (time, DRAM bytes)
	timeout 300 python -m pytest tests -q -m gpu -x -k "parity_f32 or batch_handles" > $O/pytest_sub.log 2>&1; echo "gpu subset rc=$?"; tail -3 $O/pytest_sub.log
	try "NVSP_X=default"
	try "NVSP_LIB=$PWD/tools/_variants/libnoel.so"
	try "NVSP_X=default"
	try "NVSP_LIB=$PWD/tools/_variants/libnoel.so"
	try "NVSP_X=default" --workload midi
	try "NVSP_LIB=$PWD/tools/_variants/libnoel.so" --workload midi
	for v in default noel; do
		lib=$PWD/nvspeechplayer_b200/libspeechPlayer.so; [ $v = noel ] && lib=$PWD/tools/_variants/libnoel.so
		NVSP_LIB=$lib timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:klatt_f32_sched_kernel -s 3 -c 1 \
			python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_$v.log 2>&1; echo "ncu $v rc=$?"; grep "dram__\|gpu__time\|lts__" $O/ncu_$v.log
	done
	;;
rl8)  # evict-last plan loads with a larger persisting set-aside (records + plan rows)
	timeout 300 python -m pytest tests -q -m gpu -x -k "parity_f32 or batch_handles" > $O/pytest_sub.log 2>&1; echo "gpu subset rc=$?"; tail -3 $O/pytest_sub.log
	for mb in 0 16 32 64 max; do
		e="NVSP_L2_SETASIDE_MB=$mb"; [ $mb = max ] && e="NVSP_X=max"
		try "$e"
		env $e NVSP_VERBOSE=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:klatt_f32_sched_kernel -s 3 -c 1 \
			python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_$mb.log 2>&1; echo "ncu $mb rc=$?"; grep "dram__\|gpu__time\|lts__" $O/ncu_$mb.log; grep -m1 "L2: persisting" $O/ncu_$mb.log
	done
	;;
rl9)  # persisting window on the records: warm multi-step time, on / off, alternating
	for i in 1 2 3; do
		try "NVSP_L2_PERSIST=1" --steps 5
		try "NVSP_L2_PERSIST=0" --steps 5
	done
	try "NVSP_L2_PERSIST=1" --workload vowel
	try "NVSP_L2_PERSIST=0" --workload vowel
	try "NVSP_L2_PERSIST=1" --workload midi
	try "NVSP_L2_PERSIST=0" --workload midi
	;;
rl10)  # streaming (st.global.cs) vs plain output stores on the three workload families, persistence off
	for w in batch vowel midi; do
		try "NVSP_X=cs" --workload $w
		try "NVSP_LIB=$PWD/tools/_variants/libplain.so" --workload $w
	done
	try "NVSP_X=cs"
	try "NVSP_LIB=$PWD/tools/_variants/libplain.so"
	;;
mg2)  # the torchrun bench line only, N GPUs of one box (gpurun --gpus N)
	N=${1:-8}
	timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline \
		> $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "bench x$N rc=$?"; cut -c1-2500 $O/bench_${N}gpu.json; tail -3 $O/bench_${N}gpu.err
	;;
pl)  # batched pull with helper threads: tests, bench with and without helpers
	timeout 400 python -m pytest tests/test_gpu_pull.py tests/test_gpu_batch_handles.py -q -m gpu > $O/pytest_pull.log 2>&1; echo "pull tests rc=$?"; tail -3 $O/pytest_pull.log
	timeout 300 python bench.py --workload pull > $O/bench_pull.json 2> $O/bench_pull.err; echo "bench pull rc=$?"; python -c "
import json; d=json.load(open('$O/bench_pull.json')); print(d['ms_per_step'], json.dumps(d.get('batch_of_players') or d['config'].get('batch_of_players') or {k:v for k,v in d.items() if 'batch' in k})[:600])"
	NVSP_HOST_THREADS=0 timeout 300 python bench.py --workload pull > $O/bench_pull_nohelpers.json 2> $O/bench_pull_nohelpers.err; echo "bench pull (no helpers) rc=$?"; python -c "
import json; d=json.load(open('$O/bench_pull_nohelpers.json')); print(d['ms_per_step'], json.dumps(d.get('batch_of_players') or d['config'].get('batch_of_players') or {k:v for k,v in d.items() if 'batch' in k})[:600])"
	;;
pl2)  # batched pull: host phases (NVSP_PULL_TIMING) with and without helpers, bench
	NVSP_PULL_TIMING=1 timeout 300 python bench.py --workload pull > $O/bench_pull.json 2> $O/bench_pull.err; echo "bench pull rc=$?"; grep "pull batch" $O/bench_pull.err | tail -4; python -c "
import json; d=json.load(open('$O/bench_pull.json')); print(d['ms_per_step'], json.dumps(d.get('batch_of_players'))[:600])"
	NVSP_PULL_TIMING=1 NVSP_HOST_THREADS=0 timeout 300 python bench.py --workload pull > $O/bench_pull_nohelpers.json 2> $O/bench_pull_nohelpers.err; echo "bench pull (no helpers) rc=$?"; grep "pull batch" $O/bench_pull_nohelpers.err | tail -4; python -c "
import json; d=json.load(open('$O/bench_pull_nohelpers.json')); print(d['ms_per_step'], json.dumps(d.get('batch_of_players'))[:600])"
	NVSP_PULL_TIMING=1 NVSP_HOST_THREADS=3 timeout 300 python bench.py --workload pull > $O/bench_pull_3.json 2> $O/bench_pull_3.err; echo "bench pull (3 helpers) rc=$?"; grep "pull batch" $O/bench_pull_3.err | tail -3
	;;
fd)  # ring scheduler with the fade class: bitwise tests, then fade thresholds / SM roles on config 3
	timeout 400 python -m pytest tests/test_gpu_parity_f32.py -q -m gpu -x > $O/pytest_sub.log 2>&1; echo "gpu subset rc=$?"; tail -3 $O/pytest_sub.log
	try "NVSP_SCHED_FADE_TICKS=0"
	try "NVSP_SCHED_FADE_TICKS=128"
	try "NVSP_SCHED_FADE_TICKS=128 NVSP_SCHED_FADE_SMS=50"
	try "NVSP_SCHED_FADE_TICKS=128 NVSP_SCHED_FADE_SMS=74"
	try "NVSP_SCHED_FADE_TICKS=192 NVSP_SCHED_FADE_SMS=60 NVSP_SCHED_GEN_TICKS=256"
	try "NVSP_SCHED_FADE_TICKS=64 NVSP_SCHED_FADE_SMS=74 NVSP_SCHED_GEN_TICKS=256"
	try "NVSP_SCHED_FADE_TICKS=128 NVSP_SCHED_FADE_SMS=60 NVSP_LIB=$PWD/tools/_variants/libprof.so"
	;;
fd2)  # fade class with SM roles (a primary class per SM)
	timeout 400 python -m pytest tests/test_gpu_parity_f32.py -q -m gpu -x -k "rounds_equal or long_hold" > $O/pytest_sub.log 2>&1; echo "gpu subset rc=$?"; tail -3 $O/pytest_sub.log
	try "NVSP_SCHED_FADE_TICKS=128 NVSP_SCHED_HOLD_SMS=27 NVSP_SCHED_FADE_SMS=30"
	try "NVSP_SCHED_FADE_TICKS=128 NVSP_SCHED_HOLD_SMS=27 NVSP_SCHED_FADE_SMS=40"
	try "NVSP_SCHED_FADE_TICKS=128 NVSP_SCHED_HOLD_SMS=20 NVSP_SCHED_FADE_SMS=30"
	try "NVSP_SCHED_FADE_TICKS=192 NVSP_SCHED_FADE_MAX=1024 NVSP_SCHED_HOLD_SMS=27 NVSP_SCHED_FADE_SMS=30"
	try "NVSP_SCHED_FADE_TICKS=128 NVSP_SCHED_HOLD_SMS=27 NVSP_SCHED_FADE_SMS=30 NVSP_LIB=$PWD/tools/_variants/libprof.so"
	try "NVSP_SCHED_FADE_TICKS=0 NVSP_SCHED_HOLD_SMS=28"
	;;
pl3)  # exact hold glide in the pull path: pull tests (oracle parity, batch vs solo), smoke, pull bench
	timeout 400 python -m pytest tests/test_gpu_pull.py tests/test_gpu_batch_handles.py tests/test_gpu_reference_wrapper.py -q -m gpu > $O/pytest_pull.log 2>&1; echo "pull tests rc=$?"; tail -3 $O/pytest_pull.log
	timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
	timeout 300 python bench.py --workload pull > $O/bench_pull.json 2> $O/bench_pull.err; echo "bench pull rc=$?"; python -c "
import json; d=json.load(open('$O/bench_pull.json')); print(d['ms_per_step'], json.dumps(d.get('batch_of_players'))[:400], json.dumps(d.get('latency_ms'))[:300])"
	;;
si)  # stage-in copy with 6 loads in flight per thread: bitwise tests, tries, the bench line
	timeout 300 python -m pytest tests/test_gpu_parity_f32.py tests/test_gpu_batch_handles.py -q -m gpu -x > $O/pytest_sub.log 2>&1; echo "gpu subset rc=$?"; tail -2 $O/pytest_sub.log
	try "NVSP_X=stagein6"
	try "NVSP_X=stagein6" --workload vowel
	try "NVSP_X=stagein6" --workload midi
	timeout 300 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-330 $O/bench.json
	;;
*) echo "unknown stage $stage"; exit 2;;
esac
