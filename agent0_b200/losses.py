"""Python entry points of the fused target/loss kernels (K4) -- thin wrappers over the C ABI.

Each function takes the network outputs the reference's ``train_step`` works from and returns
``LossOut(loss, grad, prio, ...)`` where ``loss`` is the per-sample loss the reference returns
(agent0/deepq/agent.py:172-388), ``grad`` is d[(loss*weights).sum()]/d[online output] -- so that
``online_out.backward(grad)`` is the reference's ``q_loss.mul(weights).sum().backward()``
(agent.py:154) -- and ``prio`` = (loss+eps)^alpha (agent0/deepq/replay.py:56-58).
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple

import torch

from . import _lib

LossOut = namedtuple("LossOut", ["loss", "grad", "prio", "target_prob", "fraction_loss", "grad_taus"],
                     defaults=[None, None, None])


def _f32(t):
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.detach().to(torch.float32).contiguous()
    return t.detach()


def _common(B, A, action, reward, done, weight, gamma_n, alpha, eps, max_p):
    dev = reward.device
    action = action.to(torch.int64).contiguous()
    reward, done, weight = _f32(reward), _f32(done), _f32(weight)
    loss = torch.empty(B, dtype=torch.float32, device=dev)
    prio = torch.empty(B, dtype=torch.float32, device=dev)
    c = _lib.LossCommon(B=B, A=A, action=_lib.ptr(action, torch.int64), reward=_lib.ptr(reward),
                        done=_lib.ptr(done), weight=_lib.ptr(weight), gamma_n=float(gamma_n),
                        alpha=float(alpha), eps=float(eps), loss=_lib.ptr(loss), prio=_lib.ptr(prio),
                        max_p=None if max_p is None else _lib.ptr(max_p, torch.float32))
    keep = (action, reward, done, weight)
    return c, loss, prio, keep


def dqn_loss(q, qt_next, action, reward, done, weight, gamma_n, qsel=None, alpha=0.5, eps=0.01, max_p=None):
    """DQNLearner.train_step (agent.py:173-190).  q/qt_next/qsel f32[B,A]."""
    lib = _lib.load()
    q, qt_next, qsel = _f32(q), _f32(qt_next), _f32(qsel)
    B, A = q.shape
    c, loss, prio, keep = _common(B, A, action, reward, done, weight, gamma_n, alpha, eps, max_p)
    grad = torch.empty_like(q)
    with torch.cuda.device(q.device):
        _lib.check(lib.a0_loss_dqn(C.byref(c), _lib.ptr(q), _lib.ptr(qt_next), _lib.ptr(qsel), _lib.ptr(grad),
                                   _lib.stream_ptr(q.device)), "a0_loss_dqn")
    return LossOut(loss, grad, prio)


def mdqn_loss(q, qt_next, qt_cur, action, reward, done, weight, gamma_n, tau=0.03, lo=-1.0,
              alpha=0.5, eps=0.01, max_p=None):
    """MDQNLearner.train_step (agent.py:194-215)."""
    lib = _lib.load()
    q, qt_next, qt_cur = _f32(q), _f32(qt_next), _f32(qt_cur)
    B, A = q.shape
    c, loss, prio, keep = _common(B, A, action, reward, done, weight, gamma_n, alpha, eps, max_p)
    grad = torch.empty_like(q)
    with torch.cuda.device(q.device):
        _lib.check(lib.a0_loss_mdqn(C.byref(c), _lib.ptr(q), _lib.ptr(qt_next), _lib.ptr(qt_cur), float(tau),
                                    float(lo), _lib.ptr(grad), _lib.stream_ptr(q.device)), "a0_loss_mdqn")
    return LossOut(loss, grad, prio)


def c51_loss(logits, tgt_logits, atoms, action, reward, done, weight, gamma_n, vmin, vmax, qsel=None,
             alpha=0.5, eps=0.01, max_p=None, want_target_prob=False):
    """C51Learner.train_step (agent.py:219-269).  logits/tgt_logits f32[B,A,M]; atoms f32[M]."""
    lib = _lib.load()
    logits, tgt_logits, qsel = _f32(logits), _f32(tgt_logits), _f32(qsel)
    atoms = _f32(atoms.reshape(-1))
    B, A, M = logits.shape
    c, loss, prio, keep = _common(B, A, action, reward, done, weight, gamma_n, alpha, eps, max_p)
    grad = torch.empty_like(logits)
    tp = torch.empty((B, M), dtype=torch.float32, device=logits.device) if want_target_prob else None
    with torch.cuda.device(logits.device):
        _lib.check(lib.a0_loss_c51(C.byref(c), _lib.ptr(logits), _lib.ptr(tgt_logits), _lib.ptr(qsel),
                                   _lib.ptr(atoms), M, float(vmin), float(vmax), _lib.ptr(grad), _lib.ptr(tp),
                                   _lib.stream_ptr(logits.device)), "a0_loss_c51")
    return LossOut(loss, grad, prio, tp)


def _quantile(layout, q, qt, taus, qsel, Ni, Nj, A, action, reward, done, weight, gamma_n, alpha, eps, max_p,
              q_bar=None, taus_full=None):
    lib = _lib.load()
    B = q.shape[0]
    c, loss, prio, keep = _common(B, A, action, reward, done, weight, gamma_n, alpha, eps, max_p)
    grad = torch.empty_like(q)
    frac = gt = None
    if q_bar is not None:
        frac = torch.empty(B, dtype=torch.float32, device=q.device)
        gt = torch.empty((B, Nj + 1), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(lib.a0_loss_quantile(C.byref(c), layout, _lib.ptr(q), _lib.ptr(qt), _lib.ptr(taus),
                                        _lib.ptr(qsel), Ni, Nj, _lib.ptr(grad), _lib.ptr(q_bar),
                                        _lib.ptr(taus_full), _lib.ptr(frac), _lib.ptr(gt),
                                        _lib.stream_ptr(q.device)), "a0_loss_quantile")
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
