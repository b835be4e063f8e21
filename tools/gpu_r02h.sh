#!/bin/bash
set -u
OUT=gpurun_out/r02h
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
for v in "0 0" "1 0" "1 20" "1 200"; do
  set -- $v
  A0_K6_SPLIT=$1 A0_K6_SLEEP=$2 timeout 200 python tools/bench_extend.py > $OUT/extend_$1_$2.json 2> $OUT/extend_$1_$2.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:a0_k4_quantile_sorted -s 2 -c 1 -o $OUT/qr_sorted python tools/bench_qr.py > $OUT/ncu_qr.log 2>&1
python -c "
import json,glob
for f in sorted(glob.glob('$OUT/extend_*.json')):
    d=json.load(open(f)); print(f, [(c['call_ms'], c['device_decode_label_us']) for c in d['calls'][2:]], d['last_entries_bit_exact'])"
