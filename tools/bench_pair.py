"""Sample + gather pair timing: separate launches vs a0_rb_sample_gather (overlapped), eager-in-graph.
Usage: python tools/bench_pair.py [B] [L]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as BN
from agent0_b200.config import make_config
from agent0_b200.replay import ReplayDataset

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = int(sys.argv[2]) if len(sys.argv) > 2 else 20
wl = dict(BN.WORKLOADS["c51_b32"]); wl["B"] = B
cfg = make_config("c51", per=True, n_step=3, batch_size=B, replay_size=1_000_000, double_q=True, dueling=True, num_envs=16, action_dim=4)
rp = ReplayDataset(cfg, native_nstep=True)
BN.fill_shard(rp, 1_000_000, 16, 1234, torch)
hp = BN.HotPath(rp, wl, L, 4, torch)
rp.push_dynamic()
out = {}
def pair_sep(i):
    hp.sample(); hp.gather()
def pair_ovl(i):
    hp.sample_gather()
for name, fn in (("separate", pair_sep), ("overlapped", pair_ovl), ("sample_only", lambda i: hp.sample()), ("gather_only", lambda i: hp.gather())):
    out[name + "_us"] = round(BN.time_kernel(fn, 200, torch) * 1e6, 2)
for ovl in (False, True):
    hp.overlap_sg = ovl
    secs = BN.time_graphed(hp, 200, 10, torch, True, lambda: None)
    out["step_us_overlap_%s" % ovl] = round(secs / 200 * 1e6, 2)
print(json.dumps(out))
