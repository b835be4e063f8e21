#!/bin/bash
# Round-2 session-3 A/B: gather waves (side-stream gather, K4 joins wave by wave) and the sparse K2b climb.
set -u
TAG=${1:-r03a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_hotloop.py tests/test_gpu_edges.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
Q="--no-extra --no-cpu-baseline --no-learner --min-seconds 0.1"
run() { # name, args
  n=$1; shift
  timeout 300 python bench.py $Q "$@" > $OUT/$n.json 2> $OUT/$n.err; echo "$n rc=$? $(python - <<P
import json
try:
    d=json.loads(open('$OUT/$n.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['run'].get('gather_waves_batches'))
except Exception as e:
    print('no line', e)
P
)"
}
run b32_none --gather-waves none
run b32_auto
run b32_auto_pdl --pdl-at-joins
run b32_w1_1_2_16 --gather-waves 1,1,2,16
run b32_w1_19 --gather-waves 1,19
run b32_w2_2_4_12 --gather-waves 2,2,4,12
A0_K2B_SPARSE=0 run b32_auto_nosparse
run b512_none --workload c51_b512 --gather-waves none
run b512_auto --workload c51_b512
run b512_auto_pdl --workload c51_b512 --pdl-at-joins
run b512_w2 --workload c51_b512 --gather-waves 2,2,2,2,2,2,2,2,2,2
run qr_none --workload qr_b512 --gather-waves none
run qr_auto --workload qr_b512
