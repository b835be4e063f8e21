#!/bin/bash
# N GPUs (default 2): the NCCL tests, then the learner-step variants
set -u
N=${1:-2}
OUT=gpurun_out/r02g_n$N
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=240 > $OUT/pytest_multi.log 2>&1; echo "rc=$?" >> $OUT/pytest_multi.log
fi
P=29600
for v in "2 1" "1 1" "2 0"; do
  set -- $v
  P=$((P+1))
  A0_GRAD_BUCKETS=$1 A0_NCCL_HIGH_PRIORITY=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/bench_learner_dp.py 2>> $OUT/dp.err | grep '^{' >> $OUT/dp.jsonl
done
tail -6 $OUT/pytest_multi.log 2>/dev/null; cat $OUT/dp.jsonl; tail -3 $OUT/dp.err
