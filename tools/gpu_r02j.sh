#!/bin/bash
set -u
OUT=gpurun_out/r02j
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_losses.py tests/test_gpu_edges.py -m gpu -q --timeout=120 > $OUT/pytest_a.log 2>&1; echo "rc=$?" >> $OUT/pytest_a.log
timeout 200 python tools/bench_qr.py > $OUT/bench_qr.json 2> $OUT/bench_qr.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:a0_k4_quantile_warp -s 2 -c 1 -o $OUT/qr_warp python tools/bench_qr.py > $OUT/ncu_qr.log 2>&1
tail -12 $OUT/pytest_a.log; cat $OUT/bench_qr.json; tail -2 $OUT/bench_qr.err
