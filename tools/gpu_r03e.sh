#!/bin/bash
set -u
OUT=gpurun_out/${1:-r03e}
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
)"
}
for w in 128 192 256 320 400 480 560; do run b512_w$w --workload c51_b512 --gather-window $w; done
run b512_w320_pdl --workload c51_b512 --gather-window 320 --pdl-at-joins
run b512_w320_2 --workload c51_b512 --gather-window 320 --gather-waves 2,2,2,2,2,2,2,2,2,2
for w in 192 320 400; do run qr_w$w --workload qr_b512 --gather-window $w; done
run iqn_w320 --workload iqn_b512 --gather-window 320
export A0_LIB=agent0_b200/libagent0_b200_trace.so
A0_GATHER_WINDOW=320 timeout 200 python tools/trace_step.py 512 20 > $OUT/trace_b512_w320.txt 2>&1
tail -28 $OUT/trace_b512_w320.txt
