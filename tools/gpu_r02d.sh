#!/bin/bash
set -u
OUT=gpurun_out/r02d
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_losses.py -m gpu -q > $OUT/pytest_losses.log 2>&1; echo "rc=$?" >> $OUT/pytest_losses.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:a0_k4_quantile_sorted -s 2 -c 2 -o $OUT/qr_sorted python tools/bench_qr.py > $OUT/ncu_qr.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" >> $OUT/bench_default.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?" >> $OUT/bench_reference.err
tail -8 $OUT/pytest_losses.log; tail -5 $OUT/bench_default.err; head -c 3000 $OUT/bench_default.json; echo; tail -3 $OUT/bench_reference.err; head -c 1500 $OUT/bench_reference.json
