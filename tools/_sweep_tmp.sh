for m in 0 2 3; do
A0_K3_L2=$m python bench.py --no-cpu-baseline --no-extra --workload c51_b512 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('b512 L2=$m value',d['value'],'ms',d['ms_per_step'],'k3 frac',d['roofline']['frac'])"
done
for m in 0 3; do
A0_K3_L2=$m python bench.py --no-cpu-baseline --no-extra --workload qr_b512 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('qr512 L2=$m value',d['value'],'ms',d['ms_per_step'],'k3 frac',d['roofline']['frac'])"
done
