"""The hot loop of the reference's ``Trainer.step`` (agent0/deepq/trainer.py:74-119) on the
HBM-resident shard: extend -> L x {sample, IS weights, learner.train, update_priority}.

What changes mechanically: the DataLoaderX/DataPrefetcher pump (trainer.py:63-72) disappears --
one K2a launch draws all ``learner_steps`` batches and one K3 launch gathers them; IS weights come
out of the sampler's epilogue instead of ``priority.sum().item()`` over a million CPU floats
(trainer.py:91-94); nothing is copied to the host inside the loop.  Drawing the L batches up front
matches what the reference effectively does: its 2 DataLoader workers and 3-deep prefetch queue
sample several batches ahead with a fork-time snapshot of the priorities.
"""
from __future__ import annotations

from .learner import make_learner
from .replay import ReplayDataset, split_batches


class Trainer:
    def __init__(self, cfg, process_group=None, native_nstep=False, **replay_kw):
        self.cfg = cfg
        self.replay = ReplayDataset(cfg, native_nstep=native_nstep, **replay_kw)
        self.learner = make_learner(cfg, process_group=process_group, device=self.replay.device,
                                    max_p=None)
        self.num_transitions = cfg.actor.sample_steps * cfg.actor.num_envs
        self.frame_count = 0
        self.Ls, self.FLs, self.Rs, self.Qs = [], [], [], []

    def learn(self, learner_steps=None):
        """The inner loop (trainer.py:82-109).  Returns per-update (q_loss, fraction_loss) device
        tensors; nothing is synchronised."""
        cfg = self.cfg
        L = int(learner_steps or cfg.learner.learner_steps)
        B = cfg.learner.batch_size
        out = []
        for b in split_batches(self.replay.sample(B, k_batches=L), B):
            data = (b.frames, b.actions, b.rewards_f32, b.terminals_f32, b.weights, b.indices)
            result = self.learner.train(data)
            self.replay.update_priority(result["indices"], result["q_loss"])
            out.append((result["q_loss"], result["fraction_loss"]))
        return out

    def step(self, transitions, returns=(), qmax=()):
        """trainer.py:74-119 with the same argument list: new transitions from the actor (reference
        tuples), episode returns and mean q-values for logging."""
        self.Qs.extend(qmax)
        self.Rs.extend(returns)
        self.replay.extend(transitions)
        self.frame_count += self.num_transitions
        if len(self.replay) > self.cfg.trainer.training_start_steps:
            for q_loss, fraction_loss in self.learn():
                if q_loss is not None:
                    self.Ls.append(q_loss.mean())
                if fraction_loss is not None:
                    self.FLs.append(fraction_loss.mean())
        mean = lambda xs, k: float(sum(float(x) for x in xs[-k:]) / len(xs[-k:])) if xs else None
        return dict(frames=self.frame_count, fraction_loss=mean(self.FLs, 20), loss=mean(self.Ls, 20),
                    return_train=mean(self.Rs, 20), return_train_max=max(self.Rs) if self.Rs else None,
                    qmax=mean(self.Qs, 100))
