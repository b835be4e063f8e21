#!/bin/bash
# Checkpoint: all GPU tests, smoke, the default bench line (both arms optional).
set -u
OUT=gpurun_out/${1:-check}
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout=240 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" >> $OUT/bench_default.err
tail -6 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; grep -E "bench rc" $OUT/bench_default.err; head -c 1500 $OUT/bench_default.json; echo
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; head -c 300 $OUT/bench_reference.json; echo
