#!/bin/bash
set -u
OUT=gpurun_out/${1:-r03d}
mkdir -p $OUT
python -m agent0_b200.build > $OUT/build.log 2>&1
for w in none auto; do for c in 1 0; do timeout 200 python tools/diag_e2e.py c51_b512 $w $c 2>&1 | tail -1; done; done | tee $OUT/diag_e2e.txt
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
A0_K3_L2=3 run b512_l2hints --workload c51_b512
run b512_eager_prio --workload c51_b512 --no-graph
run b512_eager_noprio --workload c51_b512 --no-graph --no-k4-priority
run b512_w700 --workload c51_b512 --gather-window 700
run b512_w400 --workload c51_b512 --gather-window 400
A0_K3_L2=3 run b512_w700_l2 --workload c51_b512 --gather-window 700
run b32_w160 --gather-window 160
run b32_w128_fine --gather-window 128 --gather-waves 1,1,2,4,4,8
