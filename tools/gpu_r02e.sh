#!/bin/bash
# 2-GPU box: the whole GPU suite (incl. the NCCL tests), QR timing, learner-scaling bench at N=1 and N=2
set -u
OUT=gpurun_out/r02e
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.csv 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $OUT/pytest_multi.log 2>&1; echo "rc=$?" >> $OUT/pytest_multi.log
timeout 300 python tools/bench_qr.py > $OUT/bench_qr.json 2> $OUT/bench_qr.err
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?" >> $OUT/bench_n2.err
tail -15 $OUT/pytest_multi.log; cat $OUT/bench_qr.json; tail -12 $OUT/pytest_gpu.log; tail -5 $OUT/bench_n2.err; python -c "
import json; d=json.load(open('$OUT/bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['learner_scaling'])); print(d['extra'])"
