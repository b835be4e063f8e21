#!/bin/bash
# Final pass of a session: every GPU test, smoke, both bench arms, sanitizer over the tests that touch the kernels changed last.
set -u
OUT=gpurun_out/${1:-final}
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout=240 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" >> $OUT/bench_default.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?" >> $OUT/bench_reference.err
SEL="tests/test_gpu_hotloop.py tests/test_gpu_edges.py tests/test_gpu_losses.py"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 86 --log-file $OUT/sanitizer_$tool.log \
      python -m pytest $SEL -m gpu -x -q --timeout=600 -k "not 70000 and not 3000000 and not learner and not real_network" > $OUT/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?" >> $OUT/sanitizer_summary.txt
  tail -3 $OUT/sanitizer_$tool.log >> $OUT/sanitizer_summary.txt
  tail -1 $OUT/sanitizer_${tool}_pytest.log >> $OUT/sanitizer_summary.txt
done
tail -4 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; tail -1 $OUT/bench_default.err; tail -1 $OUT/bench_reference.err; head -c 400 $OUT/bench_default.json; echo; cat $OUT/sanitizer_summary.txt
