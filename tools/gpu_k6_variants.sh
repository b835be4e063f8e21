#!/bin/bash
set -u
OUT=gpurun_out/r02k
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_extend.py -m gpu -q -x --timeout=120 > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
for g in 0 1; do A0_K6_GLOBAL=$g timeout 200 python tools/bench_extend.py > $OUT/extend_g$g.json 2> $OUT/extend_g$g.err; done
tail -5 $OUT/pytest.log
python -c "
import json
for g in (0,1):
    d=json.load(open('$OUT/extend_g%d.json'%g)); print(g, [(c['call_ms'], c['device_decode_label_us']) for c in d['calls'][2:]], d['last_entries_bit_exact'])"
