#!/bin/bash
# N-GPU box: (N = 2) the NCCL tests; (any N) bench.py under torch.distributed.run as the driver launches it -- the
# headline workload with learner_scaling, and configs[4] (batch 512, 8 M transitions sharded over the GPUs)
set -u
N=${1:-2}
OUT=gpurun_out/r02s3m_n$N
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
if [ "$N" = "2" ]; then
  timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -v --timeout=200 > $OUT/pytest_multi.log 2>&1; echo "rc=$?" >> $OUT/pytest_multi.log
  tail -8 $OUT/pytest_multi.log
fi
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_c51_b32.json 2> $OUT/bench_c51_b32.err; echo "rc=$?" >> $OUT/bench_c51_b32.err
timeout 600 $RUN --master-port 29712 bench.py --gpus $N --steps 20 --warmup 5 --workload c51_b512 --total-ring 8000000 --no-learner > $OUT/bench_c51_b512_8M.json 2> $OUT/bench_c51_b512_8M.err; echo "rc=$?" >> $OUT/bench_c51_b512_8M.err
python - <<PY
import json
for f in ("bench_c51_b32","bench_c51_b512_8M"):
    try:
        d=json.load(open("$OUT/"+f+".json")); print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]); print(json.dumps(d.get("learner_scaling"))); print(d["extra"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 $OUT/bench_c51_b32.err
