#!/bin/bash
# compute-sanitizer over the hot-loop tests (overlapped sampler/gather, result report, reporting graphs) and the
# actor tests the main evidence run stops before; each tool under its own timeout.
set -u
TAG=${1:-sanitize_hotloop}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 86 --log-file $OUT/sanitizer_$tool.log \
      python -m pytest tests/test_gpu_hotloop.py tests/test_gpu_actor.py -m gpu -q -k "not whole_loop and not real_network" > $OUT/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?" >> $OUT/sanitizer_summary.txt
  tail -3 $OUT/sanitizer_$tool.log >> $OUT/sanitizer_summary.txt
  tail -1 $OUT/sanitizer_${tool}_pytest.log >> $OUT/sanitizer_summary.txt
done
cat $OUT/sanitizer_summary.txt
