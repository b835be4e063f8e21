python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/trace_step.py 32 20 1 2>&1 | grep "^K2b \|K4\[19\]"
for ch in 1 0; do
A0_K2B_CHUNKS=$ch python bench.py --no-cpu-baseline --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunks=$ch b32 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
A0_K2B_CHUNKS=$ch python bench.py --no-cpu-baseline --no-extra --workload c51_b512 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunks=$ch b512 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
done
