#!/bin/bash
set -u
OUT=gpurun_out/r02b
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:a0_ -c 200 --csv --log-file $OUT/extend_launches.csv python tools/bench_extend.py > $OUT/ncu_extend.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:a0_k6 -s 6 -c 3 -o $OUT/k6_full python tools/bench_extend.py > $OUT/ncu_k6.log 2>&1
tail -15 $OUT/pytest_gpu.log; grep -c a0_ $OUT/extend_launches.csv
