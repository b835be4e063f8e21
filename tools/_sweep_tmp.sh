python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/trace_step.py 32 20 1 2>&1 | grep "K2b\|K4 (C51)" -A1 | grep -v "^--"
for sm in 0 1; do
A0_K2B_SMALL=$sm python bench.py --no-cpu-baseline --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('small=$sm value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])"
done
