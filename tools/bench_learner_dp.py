"""Data-parallel learner step under torchrun: ms per update with the gradient all-reduce inside the timed region,
for the exchange schedules (A0_GRAD_BUCKETS = 1 | 2) and NCCL stream priorities.  One variant per process launch:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_learner_dp.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from agent0_b200.dist import init_nccl  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
init_nccl(local, high_priority=os.environ.get("A0_NCCL_HIGH_PRIORITY", "1") != "0")


def reduce_max(x):
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class A:
    no_graph = os.environ.get("A0_NO_GRAPH", "0") == "1"


out = bench.learner_scaling(A, torch, dist, world, rank, dist.barrier, reduce_max, 4)
out["variant"] = {"A0_GRAD_BUCKETS": os.environ.get("A0_GRAD_BUCKETS", "2"), "A0_NCCL_HIGH_PRIORITY": os.environ.get("A0_NCCL_HIGH_PRIORITY", "1"),
                  "A0_NO_GRAPH": os.environ.get("A0_NO_GRAPH", "0")}
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
