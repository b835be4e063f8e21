"""Where an e2e step's time goes (bench.e2e_cabi): device time of the ingest and of the graph replay (CUDA events on the
caller's stream) and host time of the two calls, per step.  Usage: python tools/diag_e2e.py <workload> <waves auto|none> [copy_stream 0/1]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench as BN

name = sys.argv[1] if len(sys.argv) > 1 else "c51_b512"
waves = sys.argv[2] if len(sys.argv) > 2 else "auto"
copy_stream = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
wl = BN.WORKLOADS[name]
rp, _ = BN.make_shard(wl, 1_000_000, 4, 0, torch)
hp = BN.HotPath(rp, wl, 20, 4, torch, waves=waves)
rec = {"graph": [], "ingest": []}


def wrap(obj, attr, key):
    orig = getattr(obj, attr)

    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        r = orig(*a, **k)
        t1 = time.perf_counter()
        e1.record()
        rec[key].append((e0, e1, t1 - t0))
        return r
    setattr(obj, attr, f)


wrap(hp, "run", "graph")
wrap(rp, "append_steps", "ingest")
v, h2d, d2h = BN.e2e_cabi(hp, 30, 5, torch, graph=True, depth=2, copy_stream=copy_stream, min_seconds=0.0)
torch.cuda.synchronize()
out = {"workload": name, "waves": [w[1] for w in hp.waves] if hp.waves else None, "copy_stream": copy_stream, "e2e_tr_s": round(v, 1),
       "us_per_step": round(hp.total / v * 1e6, 1)}
for k, lst in rec.items():
    lst = lst[-30:]
    out[k + "_device_us"] = round(float(np.median([a.elapsed_time(b) for a, b, _ in lst])) * 1e3, 1)
    out[k + "_host_us"] = round(float(np.median([h for _, _, h in lst])) * 1e6, 1)
# gaps: end of step s's graph -> start of step s+1's ingest on the caller's stream
g, i = rec["graph"][-30:], rec["ingest"][-30:]
out["graph_end_to_next_ingest_start_us"] = round(float(np.median([g[j][1].elapsed_time(i[j + 1][0]) for j in range(len(g) - 1)])) * 1e3, 1)
print(out)
