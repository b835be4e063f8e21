"""The reference's host replay-and-target pipeline, restated for timing.  Test/bench infrastructure only.

This is the CPU baseline bench.py reports (``cpu_baseline.kind = "port"``; /root/reference is a
Python package that cannot travel to the GPU box, so the unmodified reference cannot be timed
there).  It keeps the reference's *structure*, because that is what costs the time:

  deque of lz4-compressed 8-frame blobs            agent0/deepq/replay.py:18, agent.py:78-81
  __getitem__: deque[idx] -> lz4 decompress -> np.array copy     replay.py:32-37
  DataLoader(batch_size, shuffle=True, num_workers=W) + default_collate   trainer.py:63-72
  .float() of all six fields, IS weights with priority.sum() over all slots   trainer.py:88-96
  train_step target/loss in eager torch on CPU tensors (network outputs pre-generated, the CNN
  is excluded on both sides)                         agent.py:172-388
  update_priority on CPU tensors                     replay.py:55-59

lz4 goes through the system liblz4 (python-lz4's block framing: 4-byte LE size + raw block) so
the baseline pays the real decompression cost.  The loss functions are checked against the
golden outputs of the unmodified reference in tests/test_cpu_path.py.
"""
import ctypes as C
import time
from collections import deque

import numpy as np
import torch
import torch.nn.functional as F_
from torch.utils.data import DataLoader, Dataset


# ----------------------------------------------------------------------------------------- lz4
class _LZ4:
    def __init__(self):
        lib = C.CDLL("liblz4.so.1")
        lib.LZ4_compressBound.restype = C.c_int
        lib.LZ4_compressBound.argtypes = [C.c_int]
        lib.LZ4_compress_default.restype = C.c_int
        lib.LZ4_compress_default.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        lib.LZ4_decompress_safe.restype = C.c_int
        lib.LZ4_decompress_safe.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        self.lib = lib

    def compress(self, data: bytes) -> bytes:
        n = len(data)
        cap = self.lib.LZ4_compressBound(n)
        out = C.create_string_buffer(cap)
        k = self.lib.LZ4_compress_default(data, out, n, cap)
        return n.to_bytes(4, "little") + out.raw[:k]

    def decompress(self, blob: bytes) -> bytes:
        n = int.from_bytes(blob[:4], "little")
        out = C.create_string_buffer(n)
        k = self.lib.LZ4_decompress_safe(blob[4:], out, len(blob) - 4, n)
        assert k == n
        return out.raw


_lz4 = None


def lz4():
    global _lz4
    if _lz4 is None:
        _lz4 = _LZ4()
    return _lz4


# ----------------------------------------------------------------------------------------- replay
class CpuReplay(Dataset):
    """ReplayDataset's data path (replay.py:14-59) with real lz4."""

    def __init__(self, size, prioritize, alpha=0.5, eps=0.01, beta=0.4):
        self.size, self.prioritize, self.alpha, self.eps, self.beta = size, prioritize, alpha, eps, beta
        self.data = deque(maxlen=size)
        self.priority = torch.ones(size)
        self.top = 0
        self.max_p = 1.0

    def __len__(self):
        return self.top

    def __getitem__(self, idx):
        idx = idx % self.top
        frames, at, rt, dt = self.data[idx]
        frames = np.frombuffer(lz4().decompress(frames), dtype=np.uint8)
        return np.array(frames), at, rt, dt, self.priority[idx], idx

    def extend(self, transitions):
        self.data.extend(transitions)
        k = len(transitions)
        self.top = min(self.top + k, self.size)
        if self.prioritize:
            self.priority[-k:] = self.max_p ** self.alpha

    def update_priority(self, ids, priorities):
        self.priority[ids] = (priorities + self.eps).pow(self.alpha)
        self.max_p = max(priorities.max().item(), self.max_p)


def fill_replay(replay, streams, n_step, discount=0.99):
    """Pack a synthetic stream the way Actor.sample does (agent.py:64-81: concat + lz4.compress
    per env) and extend the deque once per outer step batch."""
    from .reference_replay import pack_nstep_iter
    z = lz4()
    count = 0
    data = []
    for st, at, r_n, d_n, st_next in pack_nstep_iter(streams["obs"], streams["action"], streams["reward"],
                                                     streams["done"], n_step, discount):
        for e in range(len(at)):
            data.append((z.compress(np.concatenate((st[e], st_next[e]), axis=0).tobytes()), at[e], r_n[e], d_n[e]))
        count += len(at)
        if len(data) >= 1024:
            replay.extend(data)
            data = []
    if data:
        replay.extend(data)
    return count


# ----------------------------------------------------------------------------------------- losses
def _td(r, d, gamma_n, boot):
    return r + gamma_n * (1 - d) * boot


def huber_qr(q, q_target, taus):
    huber = F_.smooth_l1_loss(q, q_target, reduction="none")
    loss = huber * (taus - q_target.lt(q).detach().float()).abs()
    return loss.sum(-1).mean(-1).view(-1)


def log_softmax_stable(logits, tau):
    logits = logits - logits.max(dim=-1, keepdim=True)[0]
    return logits - tau * torch.logsumexp(logits / tau, dim=-1, keepdim=True)


def loss_dqn(o, a, r, d, gamma_n):
    bi = torch.arange(a.shape[0])
    qn = o["tgt_next"]
    a_next = (o["qsel"] if o.get("qsel") is not None else qn).argmax(-1)
    tgt = _td(r, d, gamma_n, qn[bi, a_next])
    return F_.smooth_l1_loss(o["online"][bi, a], tgt, reduction="none").view(-1)


def loss_mdqn(o, a, r, d, gamma_n, tau=0.03, lo=-1.0):
    bi = torch.arange(a.shape[0])
    ln = o["tgt_next"]
    qn = ln - log_softmax_stable(ln, tau)
    qn = ln.softmax(-1).mul(qn).sum(-1)
    add = log_softmax_stable(o["tgt_cur"], tau)[bi, a].clamp(lo, 0)
    tgt = r + tau * add + gamma_n * (1 - d) * qn
    return F_.smooth_l1_loss(o["online"][bi, a], tgt, reduction="none").view(-1)


def loss_c51(o, a, r, d, gamma_n, atoms, vmin=-10.0, vmax=10.0):
    B = a.shape[0]
    bi = torch.arange(B)
    M = atoms.numel()
    delta = (vmax - vmin) / (M - 1)
    prob_next = o["tgt_next"].softmax(-1)
    if o.get("qsel") is not None:
        a_next = o["qsel"].argmax(-1)
    else:
        a_next = prob_next.mul(atoms.view(1, 1, -1)).sum(-1).argmax(-1)
    prob_next = prob_next[bi, a_next, :]
    atoms_next = r.view(-1, 1) + gamma_n * (1 - d.view(-1, 1)) * atoms.view(1, -1)
    atoms_next.clamp_(vmin, vmax)
    base = (atoms_next - vmin) / delta
    lo, up = base.floor().long(), base.ceil().long()
    lo[(up > 0) * (lo == up)] -= 1
    up[(lo < (M - 1)) * (lo == up)] += 1
    target = torch.zeros_like(prob_next)
    offset = (torch.arange(B) * M).view(-1, 1).expand(B, M)
    target.view(-1).index_add_(0, (lo + offset).view(-1), (prob_next * (up.float() - base)).view(-1))
    target.view(-1).index_add_(0, (up + offset).view(-1), (prob_next * (base - lo.float())).view(-1))
    logp = o["online"].log_softmax(-1)[bi, a, :]
    return target.mul(logp).sum(-1).neg().view(-1)


def loss_qr(o, a, r, d, gamma_n):
    bi = torch.arange(a.shape[0])
    qn = o["tgt_next"]
    a_next = (o["qsel"] if o.get("qsel") is not None else qn.mean(-1)).argmax(-1)
    tgt = r.view(-1, 1) + gamma_n * (1 - d.view(-1, 1)) * qn[bi, a_next, :]
    q = o["online"][bi, a, :]
    N = q.shape[-1]
    taus = ((2 * torch.arange(N) + 1) / (2.0 * N)).view(1, 1, -1)
    return huber_qr(q.unsqueeze(1), tgt.unsqueeze(2), taus)


def loss_iqn(o, a, r, d, gamma_n):
    bi = torch.arange(a.shape[0])
    a_next = o["qsel"].argmax(-1)
    tgt = r.view(-1, 1) + gamma_n * (1 - d.view(-1, 1)) * o["tgt_next"][bi, :, a_next]
    q = o["online"][bi, :, a]
    return huber_qr(q.unsqueeze(1), tgt.unsqueeze(2), o["taus"].view(q.shape[0], 1, -1))


def loss_fqf(o, a, r, d, gamma_n):
    loss = loss_iqn(dict(online=o["online"], tgt_next=o["tgt_next"], qsel=o["qsel"], taus=o["taus_hat"]),
                    a, r, d, gamma_n)
    bi = torch.arange(a.shape[0])
    q_hat = o["online"][bi, :, a].detach()
    q = o["q_bar"][bi, :, a]
    v1 = q - q_hat[:, :-1]
    s1 = q.gt(torch.cat((q_hat[:, :1], q[:, :-1]), dim=1))
    v2 = q - q_hat[:, 1:]
    s2 = q.lt(torch.cat((q[:, 1:], q_hat[:, -1:]), dim=1))
    grads = torch.where(s1, v1, -v1) + torch.where(s2, v2, -v2)
    frac = (grads * o["taus"][:, 1:-1]).sum(dim=1).view(-1)
    return loss, frac


LOSSES = dict(dqn=loss_dqn, mdqn=loss_mdqn, c51=loss_c51, qr=loss_qr, iqn=loss_iqn, fqf=loss_fqf)


# ----------------------------------------------------------------------------------------- the timed step
def trainer_step(replay, fetch, net_outputs, algo, learner_steps, gamma_n, extra=None, with_grad=True):
    """The inner loop of Trainer.step (trainer.py:82-104) with the CNN replaced by pre-generated
    outputs: fetch -> .float() -> IS weights -> loss (+ backward to the network output) ->
    update_priority.  Returns the number of transitions processed."""
    done = 0
    extra = extra or {}
    for it in range(learner_steps):
        data = fetch()
        frames, actions, rewards, terminals, priorities, indices = map(lambda x: x.float(), data)
        if replay.prioritize:
            probs = priorities / replay.priority.sum().item()
            weights = (replay.top * probs).pow(-replay.beta)
            weights = weights / weights.max().add(1e-8)
        else:
            weights = priorities
        frames = frames.reshape(-1, 8, 84, 84).div(255.0)          # agent.py:132-135
        obs, next_obs = torch.split(frames, 4, 1)
        o = net_outputs(it)
        if with_grad:
            o["online"].requires_grad_(True)
            o["online"].grad = None
        out = LOSSES[algo](o, actions.long(), rewards, terminals, gamma_n, **extra)
        q_loss = out[0] if isinstance(out, tuple) else out
        if with_grad:
            q_loss.mul(weights).sum().backward()
        q_loss = q_loss.detach()
        if replay.prioritize:
            replay.update_priority(indices.long(), q_loss)
        done += actions.shape[0]
    return done


def make_fetcher(replay, batch_size, workers, seed=0):
    """workers == 0: the single-thread pipeline; otherwise the reference's DataLoader pump
    (trainer.py:63-72) with ``workers`` processes (the reference uses 2)."""
    if workers == 0:
        rng = np.random.RandomState(seed)
        from torch.utils.data import default_collate

        def fetch():
            idx = rng.randint(0, replay.top, size=batch_size)
            return default_collate([replay[int(i)] for i in idx])
        return fetch, None
    loader = DataLoader(replay, batch_size=batch_size, shuffle=True, num_workers=workers, pin_memory=False,
                        persistent_workers=True, prefetch_factor=3)
    state = {"it": iter(loader)}

    def fetch():
        try:
            return next(state["it"])
        except StopIteration:
            state["it"] = iter(loader)
            return next(state["it"])
    return fetch, loader


def time_steps(step_fn, steps, warmup):
    for _ in range(warmup):
        step_fn()
    t0 = time.perf_counter()
    n = 0
    for _ in range(steps):
        n += step_fn()
    dt = time.perf_counter() - t0
    return n, dt
