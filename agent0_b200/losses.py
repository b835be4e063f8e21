"""Python entry points of the fused target/loss kernels (K4) -- thin wrappers over the C ABI.

Each function takes the network outputs the reference's ``train_step`` works from and returns
``LossOut(loss, grad, prio, ...)`` where ``loss`` is the per-sample loss the reference returns
(agent0/deepq/agent.py:172-388), ``grad`` is d[(loss*weights).sum()]/d[online output] -- so that
``online_out.backward(grad)`` is the reference's ``q_loss.mul(weights).sum().backward()``
(agent.py:154) -- and ``prio`` = (loss+eps)^alpha (agent0/deepq/replay.py:56-58).

The wrappers are on the per-update path (20 calls per Trainer.step), so they avoid everything
that costs microseconds without doing work: no device context switch when the tensors already
live on the current device, the raw stream handle, no re-validation of tensors this module
allocated itself.
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple

import torch

from . import _lib

LossOut = namedtuple("LossOut", ["loss", "grad", "prio", "target_prob", "fraction_loss", "grad_taus"],
                     defaults=[None, None, None])

_f32_t, _i64_t = torch.float32, torch.int64
_byref = C.byref


def _f32(t):
    """float32, contiguous, detached CUDA tensor (no copy when it already is one)."""
    if t is None:
        return None
    if t.dtype is not _f32_t or not t.is_contiguous():
        t = t.detach().to(_f32_t).contiguous()
    elif t.requires_grad:
        t = t.detach()
    if not t.is_cuda:
        raise RuntimeError("agent0_b200 kernels need CUDA tensors (there is no CPU path)")
    return t


def _p(t):
    return None if t is None else t.data_ptr()


class _on:
    """``with torch.cuda.device(d)`` only when d is not already the current device (the context
    manager costs several microseconds per K4 launch otherwise)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if device.index is None or device.index == torch.cuda.current_device() \
            else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _common(B, A, action, reward, done, weight, gamma_n, alpha, eps, max_p):
    if action.dtype is not _i64_t or not action.is_contiguous():
        action = action.to(_i64_t).contiguous()
    reward, done, weight = _f32(reward), _f32(done), _f32(weight)
    if action.numel() != B or reward.numel() != B or done.numel() != B or weight.numel() != B:
        raise ValueError("action/reward/done/weight must hold one value per sample")
    out = torch.empty((2, B), dtype=_f32_t, device=reward.device)     # loss, prio
    loss, prio = out[0], out[1]
    if max_p is not None and max_p.dtype is not _f32_t:
        raise TypeError("max_p must be a float32 device scalar")
    c = _lib.LossCommon(B, A, action.data_ptr(), reward.data_ptr(), done.data_ptr(), weight.data_ptr(),
                        gamma_n, alpha, eps, loss.data_ptr(), prio.data_ptr(), _p(max_p))
    return c, loss, prio, (action, reward, done, weight)


def dqn_loss(q, qt_next, action, reward, done, weight, gamma_n, qsel=None, alpha=0.5, eps=0.01, max_p=None):
    """DQNLearner.train_step (agent.py:173-190).  q/qt_next/qsel f32[B,A]."""
    lib = _lib.load()
    q, qt_next, qsel = _f32(q), _f32(qt_next), _f32(qsel)
    B, A = q.shape
    c, loss, prio, keep = _common(B, A, action, reward, done, weight, gamma_n, alpha, eps, max_p)
    grad = torch.empty_like(q)
    dev = q.device
    with _on(dev):
        _lib.check(lib.a0_loss_dqn(_byref(c), q.data_ptr(), qt_next.data_ptr(), _p(qsel), grad.data_ptr(),
                                   _lib.stream_ptr(dev)), "a0_loss_dqn")
    return LossOut(loss, grad, prio)


def mdqn_loss(q, qt_next, qt_cur, action, reward, done, weight, gamma_n, tau=0.03, lo=-1.0,
              alpha=0.5, eps=0.01, max_p=None):
    """MDQNLearner.train_step (agent.py:194-215)."""
    lib = _lib.load()
    q, qt_next, qt_cur = _f32(q), _f32(qt_next), _f32(qt_cur)
    B, A = q.shape
    c, loss, prio, keep = _common(B, A, action, reward, done, weight, gamma_n, alpha, eps, max_p)
    grad = torch.empty_like(q)
    dev = q.device
    with _on(dev):
        _lib.check(lib.a0_loss_mdqn(_byref(c), q.data_ptr(), qt_next.data_ptr(), qt_cur.data_ptr(), tau, lo,
                                    grad.data_ptr(), _lib.stream_ptr(dev)), "a0_loss_mdqn")
    return LossOut(loss, grad, prio)


def c51_loss(logits, tgt_logits, atoms, action, reward, done, weight, gamma_n, vmin, vmax, qsel=None,
             alpha=0.5, eps=0.01, max_p=None, want_target_prob=False):
    """C51Learner.train_step (agent.py:219-269).  logits/tgt_logits f32[B,A,M]; atoms f32[M]."""
    lib = _lib.load()
    logits, tgt_logits, qsel = _f32(logits), _f32(tgt_logits), _f32(qsel)
    atoms = _f32(atoms)
    B, A, M = logits.shape
    if atoms.numel() != M:
        raise ValueError("atoms must hold num_atoms values")
    c, loss, prio, keep = _common(B, A, action, reward, done, weight, gamma_n, alpha, eps, max_p)
    grad = torch.empty_like(logits)
    dev = logits.device
    tp = torch.empty((B, M), dtype=_f32_t, device=dev) if want_target_prob else None
    with _on(dev):
        _lib.check(lib.a0_loss_c51(_byref(c), logits.data_ptr(), tgt_logits.data_ptr(), _p(qsel), atoms.data_ptr(),
                                   M, vmin, vmax, grad.data_ptr(), _p(tp), _lib.stream_ptr(dev)), "a0_loss_c51")
    return LossOut(loss, grad, prio, tp)


def _quantile(layout, q, qt, taus, qsel, Ni, Nj, A, action, reward, done, weight, gamma_n, alpha, eps, max_p,
              q_bar=None, taus_full=None):
    lib = _lib.load()
    B = q.shape[0]
    c, loss, prio, keep = _common(B, A, action, reward, done, weight, gamma_n, alpha, eps, max_p)
    grad = torch.empty_like(q)
    dev = q.device
    frac = gt = None
    if q_bar is not None:
        frac = torch.empty(B, dtype=_f32_t, device=dev)
        gt = torch.empty((B, Nj + 1), dtype=_f32_t, device=dev)
    with _on(dev):
        _lib.check(lib.a0_loss_quantile(_byref(c), layout, q.data_ptr(), qt.data_ptr(), _p(taus), _p(qsel), Ni, Nj,
                                        grad.data_ptr(), _p(q_bar), _p(taus_full), _p(frac), _p(gt),
                                        _lib.stream_ptr(dev)), "a0_loss_quantile")
    return LossOut(loss, grad, prio, None, frac, gt)


def qr_loss(q, qt_next, action, reward, done, weight, gamma_n, qsel=None, alpha=0.5, eps=0.01, max_p=None):
    """QRLearner.train_step (agent.py:273-293).  q/qt_next f32[B,A,N]; tau_j=(2j+1)/2N."""
    q, qt_next, qsel = _f32(q), _f32(qt_next), _f32(qsel)
    B, A, N = q.shape
    return _quantile(0, q, qt_next, None, qsel, qt_next.shape[2], N, A, action, reward, done, weight, gamma_n,
                     alpha, eps, max_p)


def iqn_loss(q, taus, qt_next, qsel, action, reward, done, weight, gamma_n, alpha=0.5, eps=0.01, max_p=None):
    """IQNLearner.train_step (agent.py:297-327).  q f32[B,N,A] at taus f32[B,N(,1)]; qt_next f32[B,N',A];
    qsel = head.qval(...) f32[B,A]."""
    q, qt_next, qsel = _f32(q), _f32(qt_next), _f32(qsel)
    B, N, A = q.shape
    taus = _f32(taus.reshape(B, N))
    return _quantile(1, q, qt_next, taus, qsel, qt_next.shape[1], N, A, action, reward, done, weight, gamma_n,
                     alpha, eps, max_p)


def fqf_loss(q_hat, taus, taus_hat, qt_next, q_bar, qsel, action, reward, done, weight, gamma_n,
             alpha=0.5, eps=0.01, max_p=None):
    """FQFLearner.train_step (agent.py:340-388).  q_hat/qt_next f32[B,F,A] at taus_hat f32[B,F];
    q_bar f32[B,F-1,A] at the interior fractions; taus f32[B,F+1].  Also returns the fraction loss
    and d[(fraction_loss*weights).sum()]/d taus."""
    q_hat, qt_next, q_bar, qsel = _f32(q_hat), _f32(qt_next), _f32(q_bar), _f32(qsel)
    B, F_, A = q_hat.shape
    taus = _f32(taus.reshape(B, F_ + 1))
    taus_hat = _f32(taus_hat.reshape(B, F_))
    return _quantile(1, q_hat, qt_next, taus_hat, qsel, F_, F_, A, action, reward, done, weight, gamma_n,
                     alpha, eps, max_p, q_bar=q_bar, taus_full=taus)
