#!/bin/bash
set -u
OUT=gpurun_out/r02c
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_extend.py tests/test_gpu_losses.py tests/test_gpu_edges.py -m gpu -q > $OUT/pytest_sel.log 2>&1; echo "rc=$?" >> $OUT/pytest_sel.log
timeout 300 python tools/bench_extend.py > $OUT/extend_atari.json 2> $OUT/extend_atari.err
timeout 300 python tools/bench_qr.py > $OUT/bench_qr.json 2> $OUT/bench_qr.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:a0_k6 -c 30 --csv --log-file $OUT/extend_launches.csv python tools/bench_extend.py > $OUT/ncu_extend.log 2>&1
tail -25 $OUT/pytest_sel.log; cat $OUT/extend_atari.json; cat $OUT/bench_qr.json; tail -3 $OUT/bench_qr.err; grep lz4 $OUT/extend_launches.csv | tail -3; cat gpurun_out/parity_r02.json | head -80
