#!/bin/bash
set -u
OUT=gpurun_out/${1:-r03i}
mkdir -p $OUT
python -m agent0_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --timeout=240 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
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
A0_K2A_TOP=0 run b32_notop
run qr --workload qr_b512
A0_QH_SPEC=0 run qr_nospec --workload qr_b512
run b512 --workload c51_b512
A0_K2A_TOP=0 run b512_notop --workload c51_b512
timeout 200 python tools/bench_qr.py > $OUT/bench_qr.json 2> $OUT/bench_qr.err; cat $OUT/bench_qr.json | head -c 600; echo
A0_QH_SPEC=0 timeout 200 python tools/bench_qr.py > $OUT/bench_qr_nospec.json 2> $OUT/bench_qr_nospec.err; cat $OUT/bench_qr_nospec.json | head -c 600; echo
