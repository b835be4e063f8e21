#!/bin/bash
# One gpurun call: GPU parity tests, smoke, default bench, ncu launch list, ncu --set full of K3.
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" >> $OUT/bench_default.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:a0_ -c 600 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-graph \
    > $OUT/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:a0_k3_gather -s 3 -c 2 \
    -o $OUT/k3_full python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-graph \
    > $OUT/ncu_k3.log 2>&1
python -m agent0_b200.build --trace > /dev/null 2>&1
timeout 300 python tools/trace_step.py 32 20 1 > $OUT/timeline_c51_b32.txt 2>&1
timeout 300 python tools/trace_step.py 32 20 0 > $OUT/timeline_c51_b32_no_overlap.txt 2>&1
timeout 300 python tools/trace_step.py 512 20 1 > $OUT/timeline_c51_b512.txt 2>&1
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; head -c 1500 $OUT/bench_default.json; echo; cat $OUT/bench_reference.json
