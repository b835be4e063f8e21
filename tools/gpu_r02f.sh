#!/bin/bash
set -u
OUT=gpurun_out/r02f
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 300 python tools/bench_qr.py > $OUT/bench_qr.json 2> $OUT/bench_qr.err
timeout 600 python -m pytest tests/test_gpu_extend.py tests/test_gpu_losses.py tests/test_gpu_hotloop.py -m gpu -q -x --timeout=120 > $OUT/pytest_a.log 2>&1; echo "rc=$?" >> $OUT/pytest_a.log
timeout 200 python tools/bench_extend.py > $OUT/extend_atari.json 2> $OUT/extend_atari.err
A0_K6_SPLIT=0 timeout 200 python tools/bench_extend.py > $OUT/extend_atari_onewarp.json 2> $OUT/extend_atari_onewarp.err
timeout 900 python -m pytest tests -m gpu -q --timeout=180 --deselect tests/test_gpu_extend.py --deselect tests/test_gpu_losses.py --deselect tests/test_gpu_hotloop.py > $OUT/pytest_b.log 2>&1; echo "rc=$?" >> $OUT/pytest_b.log
cat $OUT/bench_qr.json; tail -15 $OUT/pytest_a.log; python -c "
import json
for f in ('extend_atari','extend_atari_onewarp'):
    d=json.load(open('$OUT/'+f+'.json')); print(f, [(c['call_ms'], c['device_decode_label_us']) for c in d['calls']], d['last_entries_bit_exact'])"; tail -25 $OUT/pytest_b.log
