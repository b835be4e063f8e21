"""Kernel-level timings under the library's runtime options (PDL on/off, K2b levels per phase).
Run on the GPU box: python tools/bench_kernels.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from agent0_b200 import _lib  # noqa: E402
from agent0_b200.config import make_config  # noqa: E402
from agent0_b200.replay import ReplayDataset  # noqa: E402

lib = _lib.load()
cfg = make_config("c51", per=True, n_step=3, batch_size=32, replay_size=1_000_000, double_q=True, dueling=True,
                  num_envs=16, action_dim=4)
rp = ReplayDataset(cfg, native_nstep=True)
bench.fill_shard(rp, 1_000_000, 16, 1234, torch)
out = []
for name in ("c51_b32", "c51_b512"):
    wl = bench.WORKLOADS[name]
    for pdl in (1,):
        for levels in (3,):
            lib.a0_set_option(1, pdl)
            lib.a0_set_option(2, levels)
            hp = bench.HotPath(rp, wl, 20, 4, torch)
            secs = bench.time_graphed(hp, 100, 5, torch, True, lambda: None)
            hp.draw_pool()
            row = {"workload": name, "pdl": pdl, "k2b_levels": levels, "step_us": round(secs / 100 * 1e6, 2),
                   "k4_us": round(bench.time_kernel(lambda i: hp.loss_k(i % 20), 100, torch) * 1e6, 2),
                   "k3_us": round(bench.time_kernel(lambda i: hp.gather(pool=i), 60, torch) * 1e6, 2),
                   "k3_split_us": round(bench.time_kernel(lambda i: hp.gather(pool=i, variant=3), 60, torch) * 1e6, 2),
                   "k3_full_us": round(bench.time_kernel(lambda i: hp.gather(pool=i, variant=2), 60, torch) * 1e6, 2),
                   "k2a_us": round(bench.time_kernel(lambda i: hp.sample(), 60, torch) * 1e6, 2),
                   "k2b_us": round(bench.time_kernel(lambda i: hp.update(), 60, torch) * 1e6, 2)}
            print(json.dumps(row), flush=True)
            out.append(row)
            del hp
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_kernels.json"), "w"), indent=1)
