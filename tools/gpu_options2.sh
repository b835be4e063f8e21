set -u
OUT=gpurun_out/opts2; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
for bm in 16385 8192 2048 512; do
  A0_K2B_BULK_MIN=$bm timeout 600 python bench.py --no-cpu-baseline --steps 100 > $OUT/bench_bm$bm.json 2> $OUT/bench_bm$bm.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/opts2/bench_*.json')):
    d=json.loads(open(f).read()); w=d['extra']['workloads']
    print(f.split('/')[-1], 'b32 step',d['ms_per_step'],'k2b',w['c51_b32']['k2b_us'],'| b512 step',w['c51_b512']['ms_per_step'],'k2b',w['c51_b512']['k2b_us'])
PY
