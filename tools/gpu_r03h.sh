#!/bin/bash
set -u
OUT=gpurun_out/${1:-r03h}
mkdir -p $OUT
python -m agent0_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_hotloop.py tests/test_gpu_replay.py -m gpu -x -q --timeout=240 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
Q="--no-extra --no-cpu-baseline --no-learner --min-seconds 0.1"
run() { n=$1; shift
  timeout 300 python bench.py $Q "$@" > $OUT/$n.json 2> $OUT/$n.err; echo "$n rc=$? $(python - <<P
import json
try:
    d=json.loads(open('$OUT/$n.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['run'].get('gather_waves_batches'), d['run'].get('gather_window_draws'))
except Exception as e:
    print('no line', e)
P
)"
}
run b32
A0_K3_DEEP_MAX=128 run b32_deep
run b32_pdl --pdl-at-joins
A0_K3_DEEP_MAX=128 run b32_deep_pdl --pdl-at-joins
A0_K3_DEEP_MAX=128 run b32_deep_w96 --gather-window 96
A0_K3_DEEP_MAX=128 run b32_deep_w64 --gather-window 64
A0_K3_DEEP_MAX=1024 run b512_deep_w400 --workload c51_b512
A0_K3_DEEP_MAX=1024 run b512_deep_w300 --workload c51_b512 --gather-window 300
A0_K3_DEEP_MAX=1024 run b512_deep_w220 --workload c51_b512 --gather-window 220
run b512 --workload c51_b512
export A0_LIB=agent0_b200/libagent0_b200_trace.so
timeout 200 python tools/trace_step.py 32 20 > $OUT/trace_b32.txt 2>&1
A0_K3_DEEP_MAX=128 timeout 200 python tools/trace_step.py 32 20 > $OUT/trace_b32_deep.txt 2>&1
tail -30 $OUT/trace_b32.txt | head -12; tail -30 $OUT/trace_b32_deep.txt | head -12
