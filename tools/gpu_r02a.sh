#!/bin/bash
# Round-2 first light: the new K6 tests, the whole GPU suite, extend timing.
set -u
OUT=gpurun_out/r02a
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_extend.py tests/test_gpu_hotloop.py::test_uniform_replay_stays_uniform_under_the_loop tests/test_gpu_edges.py::test_unpaired_mailbox_gather_times_out_instead_of_hanging -m gpu -x -q > $OUT/pytest_new.log 2>&1; echo "rc=$?" >> $OUT/pytest_new.log
timeout 300 python tools/bench_extend.py > $OUT/extend_atari.json 2> $OUT/extend_atari.err
timeout 300 python tools/bench_extend.py noise > $OUT/extend_noise.json 2> $OUT/extend_noise.err
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
tail -30 $OUT/pytest_new.log; cat $OUT/extend_atari.json; tail -3 $OUT/extend_atari.err; cat $OUT/extend_noise.json; tail -5 $OUT/pytest_gpu.log
