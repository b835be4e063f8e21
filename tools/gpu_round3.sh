#!/bin/bash
# Round-2 (session 3: gather waves) evidence in one gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list, ncu --set full of
# K3 (640 and 10 240 transitions), K6 and the sorted QR K4, device timelines, sanitizer passes over the new kernels.
# Usage (from the repo root): gpurun --timeout 2400 -- 'bash tools/gpu_round3.sh r02s3'
set -u
TAG=${1:-r02s3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout=240 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" >> $OUT/bench_default.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
NCU="--no-extra --no-cpu-baseline --no-graph --eager-waves --no-learner --min-seconds 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:a0_ -c 600 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 $NCU > $OUT/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:a0_k3_gather -s 3 -c 2 \
    -o $OUT/k3_full_b32 python bench.py --steps 3 --warmup 3 --gather-waves none $NCU > $OUT/ncu_k3_b32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:a0_k3_gather -s 3 -c 2 \
    -o $OUT/k3_full_b512 python bench.py --workload c51_b512 --steps 3 --warmup 3 --gather-waves none $NCU > $OUT/ncu_k3_b512.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:a0_ -c 900 --csv \
    --log-file $OUT/launches_b512.csv python bench.py --workload c51_b512 --steps 3 --warmup 3 $NCU > $OUT/ncu_launches_b512.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:a0_k4_quantile_sorted -s 2 -c 1 -o $OUT/qr_sorted python tools/bench_qr.py > $OUT/ncu_qr.log 2>&1
timeout 200 python tools/bench_extend.py > $OUT/extend_atari.json 2> $OUT/extend_atari.err
timeout 200 python tools/bench_extend.py noise > $OUT/extend_noise.json 2> $OUT/extend_noise.err
timeout 200 python tools/bench_qr.py > $OUT/bench_qr.json 2> $OUT/bench_qr.err
python -m agent0_b200.build --trace > /dev/null 2>&1
export A0_LIB=agent0_b200/libagent0_b200_trace.so
timeout 300 python tools/trace_step.py 32 20 1 > $OUT/timeline_c51_b32.txt 2>&1
timeout 300 python tools/trace_step.py 512 20 1 > $OUT/timeline_c51_b512.txt 2>&1
A0_GATHER_WAVES=none timeout 300 python tools/trace_step.py 32 20 1 > $OUT/timeline_c51_b32_single_gather.txt 2>&1
A0_GATHER_WAVES=none timeout 300 python tools/trace_step.py 512 20 1 > $OUT/timeline_c51_b512_single_gather.txt 2>&1
unset A0_LIB
SEL="tests/test_gpu_extend.py tests/test_gpu_losses.py tests/test_gpu_hotloop.py tests/test_gpu_edges.py"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 86 --log-file $OUT/sanitizer_$tool.log \
      python -m pytest $SEL -m gpu -x -q --timeout=600 -k "not 70000 and not 3000000 and not learner and not real_network and not takes_the_decisions" > $OUT/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?" >> $OUT/sanitizer_summary.txt
  tail -3 $OUT/sanitizer_$tool.log >> $OUT/sanitizer_summary.txt
  tail -1 $OUT/sanitizer_${tool}_pytest.log >> $OUT/sanitizer_summary.txt
done
tail -4 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; head -c 600 $OUT/bench_default.json; echo; head -c 400 $OUT/bench_reference.json; echo; cat $OUT/sanitizer_summary.txt
