"""ReplayDataset.extend(list of reference tuples) per Trainer.step-sized call (1280 entries = 80 steps x 16 envs,
agent0/deepq/config.py:111-112): wall clock per call and the phase breakdown a0_ex_last_timing reports.
Usage: python tools/bench_extend.py [noise]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agent0_b200.config import make_config  # noqa: E402
from agent0_b200.replay import ReplayDataset  # noqa: E402
from agent0_b200.synth import record_stream  # noqa: E402
from oracle import cpu_path as CP, reference_replay as OR_  # noqa: E402


def main(noise=False, calls=6, m=1280, E=16):
    s_ = record_stream(E, calls * m // E + 4, seed=77, noise=noise)
    fr_, a_, r_, d_ = OR_.pack_nstep(s_["obs"], s_["action"], s_["reward"], s_["done"], 3, 0.99)
    z = CP.lz4()
    tup = [(z.compress(fr_[i].tobytes()), a_[i], r_[i], d_[i]) for i in range(calls * m)]
    rq = ReplayDataset(make_config("c51", per=True, n_step=3, batch_size=32, replay_size=100_000, num_envs=E))
    out = []
    for c in range(calls):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rq.extend(tup[c * m:(c + 1) * m])
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        t = rq.extend_timing()
        out.append({"call_ms": round((t1 - t0) * 1e3, 3), "call_plus_drain_ms": round((t2 - t0) * 1e3, 3),
                    **{k: round(v, 1) for k, v in t.items()}})
    b = rq.gather(torch.arange(rq.top - 64, rq.top, device="cuda"))
    want = fr_[calls * m - 64:calls * m]
    ok = bool(np.array_equal(b.frames.cpu().numpy(), want))
    blob = float(np.mean([len(t[0]) for t in tup]))
    print(json.dumps({"frames": "noise" if noise else "atari-like synthetic", "entries_per_call": m, "mean_blob_bytes": round(blob, 1),
                      "stored_frames_per_entry": round(rq.index.head_fs / (calls * m), 3), "last_entries_bit_exact": ok, "calls": out}))


if __name__ == "__main__":
    main(noise="noise" in sys.argv[1:])
