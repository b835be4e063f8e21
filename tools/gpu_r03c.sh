#!/bin/bash
# gather waves + ordered window + K4 priority stream: tests, A/B, timelines
set -u
TAG=${1:-r03c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_hotloop.py tests/test_gpu_edges.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
Q="--no-extra --no-cpu-baseline --no-learner --min-seconds 0.1"
run() { # name, args
  n=$1; shift
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
run b32_auto
run b32_auto_noprio --no-k4-priority
run b32_w64 --gather-window 64
run b32_w128 --gather-window 128
run b32_w192 --gather-window 192
run b32_w0 --gather-window 0
run b32_1_1_2_16 --gather-waves 1,1,2,16
run b32_2_2_4_12_w128 --gather-waves 2,2,4,12 --gather-window 128
run b32_auto_pdl --pdl-at-joins
A0_K2B_SPARSE=0 run b32_auto_nosparse
run b512_auto --workload c51_b512
run b512_auto_noprio --workload c51_b512 --no-k4-priority
run qr_auto --workload qr_b512
run dqn_uniform --workload dqn_b32_uniform
export A0_LIB=agent0_b200/libagent0_b200_trace.so
timeout 200 python tools/trace_step.py 32 20 > $OUT/trace_b32_auto.txt 2>&1
timeout 200 python tools/trace_step.py 512 20 > $OUT/trace_b512_auto.txt 2>&1
tail -32 $OUT/trace_b32_auto.txt; tail -28 $OUT/trace_b512_auto.txt
