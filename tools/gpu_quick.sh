#!/bin/bash
# Quick correctness + timing pass: GPU tests, smoke, bench (no CPU baseline), host profile of e2e.
set -u
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
timeout 900 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
timeout 300 python tools/profile_e2e.py c51_b32 > $OUT/profile_e2e.log 2>&1
tail -15 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; tail -3 $OUT/bench.err; cat $OUT/bench.json; head -45 $OUT/profile_e2e.log
