"""numpy restatement of the reference's host replay path.  Test infrastructure only.

Follows, line by line:
  * Actor.sample's done rule and n-step tracker   agent0/deepq/agent.py:57-81
  * ReplayDataset                                  agent0/deepq/replay.py:14-59
  * LinearSchedule                                 agent0/common/utils.py:12-28
  * the .float() + IS-weight block of Trainer.step agent0/deepq/trainer.py:88-96
  * Actor.act's epsilon-greedy selection            agent0/deepq/agent.py:25-39
Pinned against tests/golden/{replay_n1,replay_n3,per_state,trainer_step,act}.npz, which are
outputs of the unmodified reference.
"""
from collections import deque

import numpy as np


def done_rule(terminal, life_loss, truncated):
    """agent.py:57-62: done = (terminal | life_loss) & ~truncated."""
    return np.logical_and(np.logical_or(terminal, life_loss), np.logical_not(truncated))


def pack_nstep_iter(obs, action, reward, done, n_step, discount):
    """Restates the tracker loop of Actor.sample (agent.py:64-81), one outer step at a time.

    obs[k] u8[E,4,H,W] is the stack before outer step k (len steps+1); action/reward/done are per
    step [steps,E].  Yields, per step k, (st u8[E,4,H,W], at i64[E], Rn f64[E], Dn bool[E],
    st_next u8[E,4,H,W]) - the five things the reference zips into entries (agent.py:78-81).
    """
    tracker = deque(maxlen=n_step)
    for k in range(action.shape[0]):
        tracker.append((obs[k], action[k], reward[k], done[k]))
        r_n = np.zeros_like(reward[k])
        d_n = np.zeros_like(reward[k], dtype=np.bool_)
        for _, _, rt, dt in reversed(tracker):
            d_n = np.logical_or(d_n, dt)
            # float64, three separately rounded ops, newest -> oldest (agent.py:69)
            r_n = r_n * discount * (1 - dt) + rt
        yield tracker[0][0], tracker[0][1], r_n, d_n, obs[k + 1]


def pack_nstep(obs, action, reward, done, n_step, discount):
    """All entries of pack_nstep_iter in the reference's ``data`` list order, entry i = k*E + e
    (step-major, env-minor): frames u8[M, 8*H*W] = concat(st, st_next), action i64[M],
    reward f64[M] (n-step return), done bool[M] (OR over the window)."""
    frames, acts, rews, dones = [], [], [], []
    for st, at, r_n, d_n, st_next in pack_nstep_iter(obs, action, reward, done, n_step, discount):
        for e in range(action.shape[1]):
            frames.append(np.concatenate((st[e], st_next[e]), axis=0).reshape(-1))
            acts.append(at[e]); rews.append(r_n[e]); dones.append(d_n[e])
    return (np.stack(frames), np.array(acts, dtype=np.int64),
            np.array(rews, dtype=np.float64), np.array(dones, dtype=np.bool_))


class LinearSchedule:
    """utils.py:12-28: returns the current value, then advances by inc*steps, clamped."""

    def __init__(self, start, end, steps):
        self.inc = (end - start) / float(steps)
        self.current, self.end = start, end
        self.bound = min if end > start else max

    def __call__(self, steps=1):
        val = self.current
        self.current = self.bound(self.current + self.inc * steps, self.end)
        return val


class RefReplay:
    """ReplayDataset restated (replay.py:14-59), including quirks Q1-Q3 of SURVEY.md:
    the no-op ``roll`` (new priorities land on the tail of the vector) and a priority
    vector that keeps 1.0 in never-written slots."""

    def __init__(self, size, prioritize, alpha=0.5, eps=0.01, beta0=0.4, total_steps=int(1e7)):
        self.size, self.prioritize = size, prioritize
        self.alpha, self.eps = alpha, eps
        self.data = deque(maxlen=size)
        self.priority = np.ones(size, dtype=np.float32)
        self.top = 0
        if prioritize:
            self.beta_schedule = LinearSchedule(beta0, 1.0, total_steps)
            self.beta = beta0
            self.max_p = 1.0

    def __len__(self):
        return self.top

    def extend(self, transitions):
        self.data.extend(transitions)
        k = len(transitions)
        self.top = min(self.top + k, self.size)
        if self.prioritize:
            # replay.py:51 `self.priority.roll(...)` is not in-place: a no-op (Q1)
            self.priority[-k:] = np.float32(self.max_p ** self.alpha)
            self.beta = self.beta_schedule(k)

    def getitem(self, idx):
        idx = idx % self.top
        frames, at, rt, dt = self.data[idx]
        return np.array(frames), at, rt, dt, self.priority[idx], idx

    def update_priority(self, ids, losses):
        losses = np.asarray(losses, dtype=np.float32)
        # torch's CPU pow special-cases exponent 0.5 as sqrt
        p = np.sqrt(losses + np.float32(self.eps)) if self.alpha == 0.5 else \
            np.power(losses + np.float32(self.eps), np.float32(self.alpha))
        for i, v in zip(np.asarray(ids), p):          # sequential: last writer wins
            self.priority[int(i)] = v
        self.max_p = max(float(losses.max()), self.max_p)


def is_weights(batch_priority, priority_sum_all, top, beta):
    """trainer.py:91-94 in fp32: w=(top*p/sum_all)^-beta; w/=(max w + 1e-8)."""
    p = np.asarray(batch_priority, dtype=np.float32)
    probs = p / np.float32(priority_sum_all)
    w = np.power(np.float32(top) * probs, np.float32(-beta)).astype(np.float32)
    return (w / (w.max() + np.float32(1e-8))).astype(np.float32)


def act_rule(q, epsilon, num_actions, rng=np.random):
    """Actor.act (agent.py:29-39) after the network: draws randint first, then rand (float64), from
    numpy's generator; action = where(rand > epsilon, argmax_a q, random); returns the mean of the
    per-env maximum in float32 like ``qt_max.mean().item()``."""
    q = np.asarray(q, dtype=np.float32)
    E = q.shape[0]
    action_random = rng.randint(0, num_actions, E)
    greedy = q.argmax(axis=-1)
    action = np.where(rng.rand(E) > epsilon, greedy, action_random)
    return action.astype(np.int64), float(q.max(axis=-1).mean(dtype=np.float32))


def obs_to_float(obs):
    """agent.py:27 / agent.py:129-131 on the CPU: uint8 -> float32, true division by 255."""
    return np.asarray(obs, dtype=np.uint8).astype(np.float32) / np.float32(255.0)
