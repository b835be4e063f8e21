set -u
OUT=gpurun_out/ncu_q; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'a0_k4_quantile' -s 30 -c 2 -o $OUT/k4q python bench.py --workload qr_b512 --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-graph --ring 200000 > $OUT/ncu.log 2>&1
ncu -i $OUT/k4q.ncu-rep --page raw --csv > $OUT/k4q_raw.csv 2>/dev/null
ncu -i $OUT/k4q.ncu-rep --page details --csv > $OUT/k4q_details.csv 2>/dev/null
tail -3 $OUT/ncu.log
