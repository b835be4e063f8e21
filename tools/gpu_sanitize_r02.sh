#!/bin/bash
# compute-sanitizer memcheck + racecheck over the round-2 kernels' tests (K6, sorted QR, early K2b, bounded mailbox)
set -u
OUT=gpurun_out/r02s
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
SEL="tests/test_gpu_extend.py tests/test_gpu_losses.py tests/test_gpu_hotloop.py tests/test_gpu_edges.py"
rm -f $OUT/sanitizer_summary.txt
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 86 --log-file $OUT/sanitizer_$tool.log \
      python -m pytest $SEL -m gpu -x -q --timeout=600 -k "not 70000 and not 3000000 and not learner and not real_network and not takes_the_decisions" > $OUT/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?" >> $OUT/sanitizer_summary.txt
  tail -3 $OUT/sanitizer_$tool.log >> $OUT/sanitizer_summary.txt
  tail -1 $OUT/sanitizer_${tool}_pytest.log >> $OUT/sanitizer_summary.txt
done
cat $OUT/sanitizer_summary.txt
