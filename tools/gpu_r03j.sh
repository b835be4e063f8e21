#!/bin/bash
set -u
OUT=gpurun_out/${1:-r03j}
mkdir -p $OUT
python -m agent0_b200.build > $OUT/build.log 2>&1
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
)"; tail -2 $OUT/$n.err | cut -c1-300
}
export A0_NODE_PRIORITY=1
run np_b512 --workload c51_b512
run np_b512_w600 --workload c51_b512 --gather-window 600
run np_b512_w800 --workload c51_b512 --gather-window 800
run np_b512_w0 --workload c51_b512 --gather-window 0
run np_qr --workload qr_b512
run np_qr_w600 --workload qr_b512 --gather-window 600
run np_b32
run np_b32_w0 --gather-window 0
unset A0_NODE_PRIORITY
( time timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2>&1 | grep real
