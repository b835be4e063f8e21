#!/bin/bash
# Iteration pass: selected GPU tests + default bench without the CPU baseline / extras.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_iter.sh <tag> "<pytest -k expr>" "<bench args>"'
set -u
TAG=${1:-iter}
KEXPR=${2:-}
BARGS=${3:---no-cpu-baseline --no-extra}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
else
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
fi
timeout 900 python bench.py $BARGS > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
tail -15 $OUT/pytest_gpu.log; tail -3 $OUT/bench.err; cat $OUT/bench.json
