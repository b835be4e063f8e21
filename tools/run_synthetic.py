"""The whole deepq loop of the reference's Trainer.run (agent0/deepq/trainer.py:171-184) on the B200 path,
with a synthetic vector env in place of ALE: ShardActor (batched on-GPU action selection, 1-step
transitions appended straight into the HBM shard) -> Trainer.learn (prioritized draw + fused gather ->
CNN -> fused target/loss -> Adam -> priority write-back, L updates replayed as one CUDA graph).
Prints env steps/s and learner updates/s.  Not the BASELINE metric (that one excludes the CNN and the
env); an integration run of every piece together.   python tools/run_synthetic.py [algo] [batch]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from agent0_b200.actor import ActorPolicy, ShardActor  # noqa: E402
from agent0_b200.config import make_config  # noqa: E402
from agent0_b200.synth import SyntheticStreams  # noqa: E402
from agent0_b200.trainer import Trainer  # noqa: E402

algo = sys.argv[1] if len(sys.argv) > 1 else "c51"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 12
E = 16
cfg = make_config(algo, per=True, n_step=3, batch_size=B, double_q=True, dueling=True, replay_size=200_000, num_envs=E)
cfg.actor.sample_steps = 80               # config.py:111-112: 80 steps x 16 envs per Trainer.step
cfg.learner.learner_steps = 20
cfg.learner.target_update_freq = 500
tr = Trainer(cfg, native_nstep=True, graph=True, fused_input=True, sampler_seed=7)
envs = SyntheticStreams(E, seed=3, noise=True)
actor = ShardActor(cfg, envs, ActorPolicy(cfg, tr.learner.model), tr.replay)
rows = []
for it in range(iters):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n, rs, qs = actor.sample(max(0.05, 1.0 - it / 10))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    out = tr.learn() if len(tr.replay) > 2000 else []
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    loss = float(torch.stack([q.mean() for q, _ in out]).mean()) if out else None
    rows.append(dict(iter=it, transitions=n, act_s=round(t1 - t0, 4), learn_s=round(t2 - t1, 4), loss=loss, shard=len(tr.replay)))
    print(json.dumps(rows[-1]), flush=True)
steady = rows[4:]
act = sum(r["transitions"] for r in steady) / sum(r["act_s"] for r in steady)
upd = 20 * len(steady) / sum(r["learn_s"] for r in steady)
print(json.dumps({"algo": algo, "batch": B, "env_steps_per_s_actor_side": round(act, 1), "learner_updates_per_s": round(upd, 1),
                  "agent_steps_per_s_whole_loop": round(sum(r["transitions"] for r in steady) / sum(r["act_s"] + r["learn_s"] for r in steady), 1),
                  "note": "synthetic numpy env (84x84 noise frames) stepping on one host core is part of act_s"}))
