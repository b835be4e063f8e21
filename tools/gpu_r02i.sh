#!/bin/bash
set -u
OUT=gpurun_out/r02i
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_hotloop.py tests/test_gpu_replay.py -m gpu -q -x --timeout=120 > $OUT/pytest_a.log 2>&1; echo "rc=$?" >> $OUT/pytest_a.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-learner > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err
tail -4 $OUT/pytest_a.log; tail -3 $OUT/bench.err; python -c "
import json; d=json.load(open('$OUT/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('extend_api_value'), d['e2e'].get('extend_api_extend_ms')); print(d['extra'].get('stale_by_one_schedule')); print(d['extra']['compat_extend']['ms'], d['extra']['compat_extend']['phase_us_last_call'])
for c in d['configs']: print(c)"
