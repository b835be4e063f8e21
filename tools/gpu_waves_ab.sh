#!/bin/bash
# A/B of the step's gather schedule (hotloop.ReplayTargetLoop: waves on a side stream, ordered-fetch window, PDL at the joins,
# sampler top levels, QR all-action fetch) -- the sweeps behind profiles/r02s3_gather_waves_ab.json, with the current flags.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_waves_ab.sh waves_ab'
set -u
OUT=gpurun_out/${1:-waves_ab}
mkdir -p $OUT
python -m agent0_b200.build > $OUT/build.log 2>&1
Q="--no-extra --no-cpu-baseline --no-learner --min-seconds 0.1"
run() { n=$1; shift
  timeout 300 python bench.py $Q "$@" > $OUT/$n.json 2> $OUT/$n.err; echo "$n rc=$? $(python - <<P
import json
try:
    d=json.loads(open('$OUT/$n.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['run'].get('gather_waves_batches'), d['run'].get('gather_window_draws'))
except Exception as e:
    print('no line', e)
P
)"
}
run b32_none --gather-waves none
run b32_auto
for w in 0 64 96 128 192; do run b32_w$w --gather-window $w; done
run b32_no_pdl_at_joins --no-pdl-at-joins
A0_K2A_TOP=0 run b32_sampler_all_levels_from_l2
A0_K2B_SPARSE=0 run b32_k2b_full_chunks
run b512_none --workload c51_b512 --gather-waves none
for w in 0 256 320 400 480 560; do run b512_w$w --workload c51_b512 --gather-window $w; done
run qr_none --workload qr_b512 --gather-waves none
run qr_auto --workload qr_b512
A0_QH_SPEC=0 run qr_dependent_fetch --workload qr_b512
run iqn_auto --workload iqn_b512
