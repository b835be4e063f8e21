set -u
OUT=gpurun_out/opts; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
for cfg in "3 1" "4 1" "3 3" "4 3" "3 7"; do
  set -- $cfg
  A0_K2B_LEVELS=$1 A0_PDL=$2 timeout 600 python bench.py --no-cpu-baseline --steps 200 > $OUT/bench_l$1_p$2.json 2> $OUT/bench_l$1_p$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/opts/bench_*.json')):
    d=json.loads(open(f).read())
    w=d['extra']['workloads']
    print(f.split('/')[-1], 'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], 'b32 k2b',w['c51_b32']['k2b_us'],'k2a',w['c51_b32']['k2a_us'],'| b512 step',w['c51_b512']['ms_per_step'],'k2b',w['c51_b512']['k2b_us'],'k2a',w['c51_b512']['k2a_us'],'k3',w['c51_b512']['k3_us'])
PY
