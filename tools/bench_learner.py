"""Learner updates per second WITH the Nature-CNN (not the BASELINE metric, which excludes it):
the eager loop (one Python-driven update at a time, as the reference runs it) against the
CUDA-graphed loop (trainer.GraphedUpdates), without and with K3 writing the CNN's normalised f32
inputs directly (a0_rb_gather_f32), and the opt-in mixed-precision form (bf16 inputs from a0_rb_gather_bf16, the
networks under bf16 autocast; Trainer(amp=True)).  Run on the GPU box: python tools/bench_learner.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from agent0_b200.config import make_config  # noqa: E402
from agent0_b200.synth import fill_shard_synthetic  # noqa: E402
from agent0_b200.trainer import Trainer  # noqa: E402

rows = []
CASES = (("c51", 32), ("c51", 512), ("qr", 512), ("iqn", 512), ("dqn", 32))
if len(sys.argv) > 1:                      # e.g. "c51:512,qr:512"
    CASES = tuple((a, int(b)) for a, b in (x.split(":") for x in sys.argv[1].split(",")))
for algo, B in CASES:
    for graph, fused, amp, cl in ((False, False, False, False), (True, False, False, False), (True, True, False, False),
                                 (True, True, True, False), (True, True, True, True), (True, True, False, True)):
        cfg = make_config(algo, per=True, n_step=3, batch_size=B, double_q=True, dueling=True, replay_size=200_000,
                          num_envs=16)
        cfg.learner.learner_steps = 20
        cfg.learner.target_update_freq = 500
        tr = Trainer(cfg, native_nstep=True, graph=graph, fused_input=fused, amp=amp, channels_last=cl)
        fill_shard_synthetic(tr.replay, 200_000, 16, 1)
        for _ in range(3):
            tr.learn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        steps = 10
        for _ in range(steps):
            tr.learn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        row = {"algo": algo, "batch": B, "graphed": graph, "fused_f32_gather": fused, "bf16_autocast_and_gather": amp, "channels_last": cl, "updates_per_s": round(steps * 20 / dt, 1),
               "transitions_per_s": round(steps * 20 * B / dt, 1), "ms_per_update": round(dt / (steps * 20) * 1e3, 3)}
        print(json.dumps(row), flush=True)
        rows.append(row)
        del tr
        torch.cuda.empty_cache()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_learner.json"), "w"), indent=1)
