"""Synthetic per-env Atari-like streams for tests, golden fixtures and bench.py.

The generator reproduces the *frame-stream discontinuities* the reference's env
wrappers create (agent0/common/atari_wrappers.py:20-56, FrameStack(4) at :62):

* ordinary step: the 4-frame stack shifts by one new frame;
* life loss (EpisodicLifeEnv + FIRE): the env is stepped 3 extra times inside
  one outer step, so the next stack is four new frames;
* terminal or truncation: vector auto-reset replaces the whole stack.

It is deterministic for a seed (numpy legacy ``RandomState``, whose stream is
frozen across numpy versions), so fixtures can be regenerated from the seed.
Frames are "background + a few rectangles" so that they compress like Atari
frames do (the reference stores lz4 blobs, agent0/deepq/agent.py:80).
"""
from __future__ import annotations

import numpy as np


class SyntheticStreams:
    """``num_envs`` independent streams with the gymnasium vector-env step contract
    that ``Actor.sample`` consumes (agent0/deepq/agent.py:55-62)."""

    def __init__(self, num_envs, seed=1234, frame_hw=(84, 84), action_dim=4,
                 p_terminal=1.0 / 200, p_life_loss=1.0 / 120, p_truncated=1.0 / 500,
                 noise=False):
        self.E = int(num_envs)
        self.hw = tuple(frame_hw)
        self.action_dim = int(action_dim)
        self.p_terminal, self.p_life_loss, self.p_truncated = p_terminal, p_life_loss, p_truncated
        self.noise = noise
        self.rng = np.random.RandomState(seed)
        self.stack = None

    # -- frame content -------------------------------------------------------
    def _frame(self):
        h, w = self.hw
        if self.noise:
            return self.rng.randint(0, 256, size=(h, w)).astype(np.uint8)
        f = np.full((h, w), self.rng.randint(0, 64), dtype=np.uint8)
        for _ in range(self.rng.randint(1, 9)):
            y0, x0 = self.rng.randint(0, h), self.rng.randint(0, w)
            y1, x1 = min(h, y0 + self.rng.randint(1, 16)), min(w, x0 + self.rng.randint(1, 16))
            f[y0:y1, x0:x1] = self.rng.randint(64, 256)
        return f

    def _fresh_stack(self):
        return np.stack([self._frame() for _ in range(4)])

    # -- vector-env contract ---------------------------------------------------
    def reset(self):
        self.stack = np.stack([self._fresh_stack() for _ in range(self.E)])
        return self.stack.copy(), {}

    def step(self, action=None):
        E = self.E
        reward = self.rng.choice(np.array([-1.0, 0.0, 1.0]), size=E, p=[0.05, 0.9, 0.05])
        terminal = self.rng.rand(E) < self.p_terminal
        truncated = (self.rng.rand(E) < self.p_truncated) & ~terminal
        life_loss = (self.rng.rand(E) < self.p_life_loss) & ~terminal & ~truncated
        nxt = np.empty_like(self.stack)
        for e in range(E):
            if terminal[e] or truncated[e] or life_loss[e]:
                nxt[e] = self._fresh_stack()
            else:
                nxt[e, :3] = self.stack[e, 1:]
                nxt[e, 3] = self._frame()
        self.stack = nxt
        info = {"life_loss": life_loss}
        return nxt.copy(), reward.astype(np.float64), terminal, truncated, info

    def sample_actions(self):
        return self.rng.randint(0, self.action_dim, size=self.E).astype(np.int64)

    def close(self):
        pass


def record_stream(num_envs, steps, seed, **kw):
    """Roll a ``SyntheticStreams`` for ``steps`` outer steps and return the whole
    stream as arrays: obs[k] is the stack *before* step k (so obs has steps+1 rows).

    ``done`` follows the reference's rule (agent0/deepq/agent.py:57-62):
    done = (terminal | life_loss) & ~truncated.
    """
    env = SyntheticStreams(num_envs, seed=seed, **kw)
    obs0, _ = env.reset()
    obs, act, rew, term, trunc, life = [obs0], [], [], [], [], []
    for _ in range(steps):
        a = env.sample_actions()
        o, r, t, tr, info = env.step(a)
        obs.append(o); act.append(a); rew.append(r); term.append(t); trunc.append(tr)
        life.append(info["life_loss"])
    term, trunc, life = np.array(term), np.array(trunc), np.array(life)
    done = np.logical_and(np.logical_or(term, life), np.logical_not(trunc))
    return dict(obs=np.stack(obs), action=np.stack(act), reward=np.stack(rew),
                terminal=term, truncated=trunc, life_loss=life, done=done)


def fill_shard_synthetic(rp, transitions, num_envs, seed):
    """Fill a ReplayDataset shard with device-generated uint8 noise frames through the native
    ingest (K1): ``num_envs`` streams, one new frame per step, a whole new stack with p = 1/100
    (life loss / reset), done with p = 1/200, rewards in {-1, 0, 1}.  Used by bench.py and the
    full-size GPU tests; content-independent kernels make noise as good as Atari frames here."""
    import torch
    F = rp.F
    E = int(num_envs)
    rng = np.random.RandomState(seed)
    dev = rp.device
    g = torch.Generator(device=dev).manual_seed(seed)
    rp.reset_streams(np.arange(E), torch.randint(0, 256, (E * 4, F), dtype=torch.uint8, device=dev, generator=g))
    steps_total = (transitions + E - 1) // E
    chunk_steps = max(1, min(rp.index.max_chunk // E, 4096))
    done_steps = 0
    while done_steps < steps_total:
        T = min(chunk_steps, steps_total - done_steps)
        m = T * E
        streams = np.tile(np.arange(E, dtype=np.int64), T)
        n_new = np.where(rng.rand(m) < 0.01, 4, 1).astype(np.int64)
        frames = torch.randint(0, 256, (int(n_new.sum()), F), dtype=torch.uint8, device=dev, generator=g)
        action = rng.randint(0, 4, m)
        reward = rng.choice([-1.0, 0.0, 1.0], m, p=[0.05, 0.9, 0.05])
        done = rng.rand(m) < (1.0 / 200)
        rp.append_steps(streams, n_new, frames, action, reward, done)
        done_steps += T
    torch.cuda.synchronize(dev)
