#!/bin/bash
set -u
OUT=gpurun_out/${1:-r03g}
mkdir -p $OUT
python -m agent0_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_hotloop.py tests/test_gpu_replay.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -m gpu -x -q --timeout=240 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
Q="--no-extra --no-cpu-baseline --no-learner --min-seconds 0.1"
run() { n=$1; shift
  timeout 300 python bench.py $Q "$@" > $OUT/$n.json 2> $OUT/$n.err; echo "$n rc=$? $(python - <<P
import json
try:
    d=json.loads(open('$OUT/$n.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['run'].get('gather_waves_batches'), d['run'].get('gather_window_draws'), d['run'].get('update_groups'))
except Exception as e:
    print('no line', e)
P
)"
}
run b512 --workload c51_b512
A0_K2A_ROUNDS=0 run b512_rounds0 --workload c51_b512
run b512_nogroups --workload c51_b512 --update-groups 0
run b512_g2 --workload c51_b512 --update-groups 2
run qr --workload qr_b512
run qr_nogroups --workload qr_b512 --update-groups 0
run iqn --workload iqn_b512
run b32
export A0_LIB=agent0_b200/libagent0_b200_trace.so
timeout 200 python tools/trace_step.py 512 20 > $OUT/trace_b512.txt 2>&1
tail -28 $OUT/trace_b512.txt
