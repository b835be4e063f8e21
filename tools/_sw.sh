for w in 1 2 4; do
A0_C51_WARPS=$w python bench.py --no-cpu-baseline --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('warps=$w b32',d['value'],d['ms_per_step'])"
A0_C51_WARPS=$w python bench.py --no-cpu-baseline --no-extra --workload c51_b512 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('warps=$w b512',d['value'],d['ms_per_step'])"
done
