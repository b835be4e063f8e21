#!/bin/bash
# compute-sanitizer (memcheck + racecheck + synccheck) over the small GPU tests, and ncu --set full of
# the K2a / K2b / K4 kernels at the headline and the batch-512 launch shapes.
set -u
TAG=${1:-evidence}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
SEL="tests/test_gpu_replay.py tests/test_gpu_losses.py tests/test_gpu_edges.py tests/test_gpu_actor.py tests/test_gpu_hotloop.py"
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 86 --log-file $OUT/sanitizer_$tool.log \
      python -m pytest $SEL -m gpu -x -q -k "not 70000 and not 3000000 and not learner and not real_network" > $OUT/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?" >> $OUT/sanitizer_summary.txt
  tail -3 $OUT/sanitizer_$tool.log >> $OUT/sanitizer_summary.txt
  tail -1 $OUT/sanitizer_${tool}_pytest.log >> $OUT/sanitizer_summary.txt
done
for wl in c51_b32 c51_b512 qr_b512; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'a0_k4_|a0_k2a|a0_k2b' -s 30 -c 8 \
      -o $OUT/k_$wl python bench.py --workload $wl --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-graph --ring 200000 \
      > $OUT/ncu_$wl.log 2>&1
done
for wl in c51_b32 c51_b512 qr_b512; do
  python tools/summarize_ncu.py full $OUT/k_$wl.ncu-rep > $OUT/k_$wl.json 2>/dev/null
done
cat $OUT/sanitizer_summary.txt
