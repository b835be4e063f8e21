"""Host-side profile of the e2e loop (bench.e2e_loop): where the Python/ctypes time goes.
Run on the GPU box:  python tools/profile_e2e.py [workload]"""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from agent0_b200.config import make_config  # noqa: E402
from agent0_b200.replay import ReplayDataset  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c51_b32"
wl = bench.WORKLOADS[name]
cfg = make_config(wl["algo"], per=wl["per"], n_step=wl["n"], batch_size=wl["B"], replay_size=200_000,
                  double_q=wl["double"], dueling=True, num_envs=16, action_dim=4)
rp = ReplayDataset(cfg, native_nstep=True)
bench.fill_shard(rp, 200_000, 16, 1234, torch)
v, h2d, d2h = bench.e2e_loop(rp, wl, 20, 4, 100, 5, torch)
print(f"e2e {name}: {v:.0f} transitions/s  ({20 * wl['B'] / v * 1e6:.1f} us per step)")
pr = cProfile.Profile()
pr.enable()
bench.e2e_loop(rp, wl, 20, 4, 100, 0, torch)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
