#!/bin/bash
# N-GPU bench (ours + reference arm), launched the way the driver does.
set -u
N=${1:-2}
TAG=${2:-multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
PORT=29531
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 200 --warmup 10 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?" >> $OUT/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
    bench.py --gpus $N --steps 100 --warmup 10 --workload c51_b512 --total-ring 8000000 > $OUT/bench_n${N}_b512_8M.json 2> $OUT/bench_n${N}_b512_8M.err
tail -3 $OUT/bench_n$N.err; cat $OUT/bench_n$N.json | cut -c1-1800; echo; cat $OUT/bench_n${N}_b512_8M.json | cut -c1-1500
