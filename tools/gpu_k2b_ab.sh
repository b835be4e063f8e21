#!/bin/bash
# K2b chunk schedule with the match list: tests, then the step with the chunk schedule taking 10 240 indices or not
set -u
OUT=gpurun_out/${1:-k2b_ab}
mkdir -p $OUT
python -m agent0_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_hotloop.py tests/test_gpu_fullsize.py tests/test_gpu_replay.py -m gpu -x -q --timeout=300 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
A0_K2B_CHUNK_MAX=16384 timeout 600 python -m pytest tests/test_gpu_edges.py tests/test_gpu_fullsize.py "tests/test_gpu_hotloop.py::test_update_report_hands_the_result_to_mapped_host_memory" -m gpu -x -q --timeout=300 > $OUT/pytest_16384.log 2>&1; echo "pytest(16384) rc=$?" >> $OUT/pytest_16384.log
tail -3 $OUT/pytest_16384.log
Q="--no-cpu-baseline --no-learner --min-seconds 0.1"
run() { n=$1; shift
  timeout 400 python bench.py $Q "$@" > $OUT/$n.json 2> $OUT/$n.err; echo "$n rc=$? $(python - <<P
import json
try:
    d=json.loads(open('$OUT/$n.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
    for k,v in d.get('extra',{}).get('workloads',{}).items(): print('   ',k, v['transitions_per_s'], v['ms_per_step'], 'k2b_us', v['k2b_us'])
except Exception as e:
    print('no line', e)
P
)"
}
run default
A0_K2B_CHUNK_MAX=16384 run chunkmax16384
