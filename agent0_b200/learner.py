"""The six deepq learners behind the reference's interface, with the target/loss rules on the fused
sm_100a kernels (K4) and the Nature-CNN forward/backward left to PyTorch.

Drop-in surface (agent0/deepq/agent.py:96-388): classes ``DQNLearner, MDQNLearner, C51Learner,
QRLearner, IQNLearner, FQFLearner`` looked up by ``f"{algo.name.upper()}Learner"``
(agent0/deepq/trainer.py:31-34); ``__init__(cfg)``; attributes ``model``, ``model_target``,
``optimizer`` (``fqf_optimizer``), ``update_steps``; ``train(data)`` with
``data = (frames, actions, rewards, terminals, weights, indices)`` returning
``{"q_loss", "fraction_loss", "indices"}``.

How a step differs from the reference mechanically, not numerically: the reference builds the loss
with eager ATen ops and calls ``q_loss.mul(weights).sum().backward()`` (agent.py:154); here one K4
launch produces the per-sample loss *and* d[(loss*weights).sum()]/d[network output], and the CNN's
backward is started from that gradient (``out.backward(grad)``).  Results stay on the device (the
reference's ``.cpu()`` calls, agent.py:163-169, are a host sync per update); everything the
reference's ``Trainer.step`` does with them (``.mean().item()``, ``update_priority``) still works.

Data-parallel use (north_star): pass ``process_group``; gradients are SUM-all-reduced
over NCCL in one flat bucket and Adam's eps becomes 1e-2/(world*batch), so that G ranks at batch B
take the step one learner would take at batch G*B (the reference's loss is SUM-reduced).
"""
from __future__ import annotations

import copy

import numpy as np
import torch
import torch.nn as nn

from . import losses as L
from .dist import FlatGradBucket, OverlappedGradBucket, adam_eps
from .model import DeepQNet


def _algo_name(cfg):
    return getattr(cfg.learner.algo, "name", cfg.learner.algo)


class BaseLearner:
    def __init__(self, cfg, process_group=None, device=None, max_p=None, capturable=False, amp_dtype=None, channels_last=False):
        """capturable=True makes ``update`` CUDA-graph safe (trainer.GraphedUpdates): device-side
        optimizer step counters, IQN/FQF taus from the device generator instead of the CPU one
        (SURVEY Q13 is then not reproduced draw for draw) and no NaN guard (its host sync cannot be
        captured).
        amp_dtype=torch.bfloat16 (opt-in): the networks run under ``torch.autocast`` in bf16 -- on
        obs / next_obs that K3 already wrote as bf16 (a0_rb_gather_bf16, ``Trainer(amp=True)``) -- while parameters,
        Adam, the K4 rules (fed ``output.float()``) and every priority stay fp32.  NOT the reference's arithmetic:
        losses agree with the fp32 learner to bf16 rounding of the network outputs, not to 1e-5, so it is off by
        default and outside the parity contract."""
        self.cfg = cfg
        self.capturable = bool(capturable)
        assert amp_dtype in (None, torch.bfloat16), "amp_dtype: None (the reference's fp32) or torch.bfloat16"
        self.amp_dtype = amp_dtype
        self.channels_last = bool(channels_last)    # measured option: NHWC convolutions (obs converted by torch, one extra pass)
        dv = device if device is not None else getattr(cfg.device, "value", cfg.device)
        self.device = torch.device(dv)
        if self.device.type != "cuda":
            raise RuntimeError("agent0_b200 learners run their target/loss rules in CUDA kernels; "
                               "there is no CPU path (use the reference learner on CPU)")
        self.model = DeepQNet(cfg).to(self.device)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        if self.world > 1:
            # data-parallel replicas start from rank 0's weights (whatever each process's RNG drew), then stay
            # identical because every step applies the same all-reduced gradient
            from .actor import broadcast_model
            broadcast_model(self.model, src=0, process_group=process_group)
        if self.channels_last:
            self.model = self.model.to(memory_format=torch.channels_last)
        self.model_target = copy.deepcopy(self.model)       # agent.py:100 (no RNG consumed)
        self.optimizer = torch.optim.Adam(list(self.model.params()), cfg.learner.learning_rate,
                                          eps=adam_eps(cfg.learner.batch_size, self.world), capturable=self.capturable)
        if self.capturable:
            for net in (self.model, self.model_target):
                if hasattr(net.head, "device_taus"):
                    net.head.device_taus = True
        self.update_steps = 0
        self.gamma_n = float(np.float32(cfg.learner.discount ** cfg.learner.n_step_q))
        self.alpha, self.eps = float(cfg.replay.alpha), float(cfg.replay.eps)
        self.max_p = max_p                      # device scalar shared with the replay shard (optional)
        self.nan_guard = not self.capturable    # agent.py:152-158 (costs one device->host sync per update)
        # one flat gradient buffer, zeroed with one memset; with a process group it is all-reduced in two pieces
        # from backward hooks, so the NCCL calls run under the convolution backward (dist.OverlappedGradBucket)
        import os
        nb = int(os.environ.get("A0_GRAD_BUCKETS", "2"))        # 1: one all-reduce after the backward pass (nothing overlapped)
        self.bucket = OverlappedGradBucket(list(self.model.params()), process_group, n_buckets=nb) if self.world > 1 \
            else FlatGradBucket(list(self.model.params()), process_group)

    # ---- helpers ----------------------------------------------------------------------------------
    def _kw(self):
        return dict(alpha=self.alpha, eps=self.eps, max_p=self.max_p)

    def split_frames(self, frames):
        """agent.py:129-135: [B, 8*H*W] -> obs, next_obs in [0,1] (uint8 straight from K3, or the
        reference Trainer's float tensors)."""
        c = self.cfg.obs_shape[0]
        x = frames.reshape(-1, 2 * c, *self.cfg.obs_shape[1:]).float().div(255.0)
        return torch.split(x, c, 1)

    def train_step(self, obs, actions, rewards, terminals, next_obs, weights):
        raise NotImplementedError

    @staticmethod
    def _f32(x):
        """A network output as the fp32 tensor K4 reads (and whose gradient K4 writes); under autocast a cast node
        of the graph, so that ``net_out.backward(grad)`` flows back into the bf16 network."""
        return x if x is None or x.dtype == torch.float32 else x.float()

    def train(self, data):
        """agent.py:124-169: one update, then the periodic target sync."""
        result = self.update(data)
        self.sync_target()
        return result

    def sync_target(self, force=False):
        if force or self.update_steps % self.cfg.learner.target_update_freq == 0:
            self.model_target.load_state_dict(self.model.state_dict())   # deepcopy(model), agent.py:160-161

    def update(self, data):
        """The update without the target sync (everything here runs on the device; with
        capturable=True it can be captured into a CUDA graph)."""
        if self.cfg.learner.noisy_net:
            self.model.reset_noise()
            self.model_target.reset_noise()
        frames, actions, rewards, terminals, weights, indices = data
        dev = self.device
        if isinstance(frames, (tuple, list)):
            obs, next_obs = frames          # f32 [B,4,H,W] pair from the fused gather (ReplayDataset.sample(normalized=))
        else:
            obs, next_obs = self.split_frames(frames.to(dev, non_blocking=True))
        if self.channels_last:
            obs, next_obs = obs.contiguous(memory_format=torch.channels_last), next_obs.contiguous(memory_format=torch.channels_last)
        actions = actions.to(dev, non_blocking=True).long()
        rewards = rewards.to(dev, non_blocking=True).float()
        terminals = terminals.to(dev, non_blocking=True).float()
        weights = weights.to(dev, non_blocking=True).float()

        self.bucket.zero_()
        with torch.autocast("cuda", dtype=self.amp_dtype or torch.bfloat16, enabled=self.amp_dtype is not None):
            out, net_out = self.train_step(obs, actions, rewards, terminals, next_obs, weights)
        q_loss, fraction_loss = out.loss, out.fraction_loss
        skip = bool(torch.isnan(q_loss).any()) if self.nan_guard else False
        if not skip:
            net_out.backward(out.grad)                      # == q_loss.mul(weights).sum().backward()
            self.bucket.all_reduce()
            self.optimizer.step()
            self.update_steps += 1
        else:
            q_loss = None
        return {"q_loss": q_loss, "fraction_loss": fraction_loss,
                "indices": indices.to(dev, non_blocking=True).long()}


class DQNLearner(BaseLearner):
    """agent.py:172-190."""

    def train_step(self, obs, actions, rewards, terminals, next_obs, weights):
        with torch.no_grad():
            qt_next = self._f32(self.model_target(next_obs))
            qsel = self._f32(self.model.qval(next_obs)) if self.cfg.learner.double_q else None
        q = self._f32(self.model(obs))
        return L.dqn_loss(q, qt_next, actions, rewards, terminals, weights, self.gamma_n, qsel=qsel, **self._kw()), q


class MDQNLearner(BaseLearner):
    """agent.py:193-215 (both target evaluations come from the target net)."""

    def train_step(self, obs, actions, rewards, terminals, next_obs, weights):
        c = self.cfg.learner.mdqn
        with torch.no_grad():
            qt_next = self._f32(self.model_target(next_obs))
            qt_cur = self._f32(self.model_target(obs))
        q = self._f32(self.model(obs))
        return L.mdqn_loss(q, qt_next, qt_cur, actions, rewards, terminals, weights, self.gamma_n,
                           tau=c.tau, lo=c.lo, **self._kw()), q


class C51Learner(BaseLearner):
    """agent.py:218-269."""

    def train_step(self, obs, actions, rewards, terminals, next_obs, weights):
        c = self.cfg.learner.c51
        with torch.no_grad():
            tgt = self._f32(self.model_target(next_obs))
            qsel = self._f32(self.model.qval(next_obs)) if self.cfg.learner.double_q else None
        logits = self._f32(self.model(obs))
        return L.c51_loss(logits, tgt, self.model.head.atoms, actions, rewards, terminals, weights, self.gamma_n,
                          c.vmin, c.vmax, qsel=qsel, **self._kw()), logits


class QRLearner(BaseLearner):
    """agent.py:272-293."""

    def train_step(self, obs, actions, rewards, terminals, next_obs, weights):
        with torch.no_grad():
            qt_next = self._f32(self.model_target(next_obs))
            qsel = self._f32(self.model.qval(next_obs)) if self.cfg.learner.double_q else None
        q = self._f32(self.model(obs))
        return L.qr_loss(q, qt_next, actions, rewards, terminals, weights, self.gamma_n, qsel=qsel, **self._kw()), q


class IQNLearner(BaseLearner):
    """agent.py:296-327.  The tau draws keep the reference's order on torch's CPU generator:
    K (action selection) -> N_dash (target) -> N (online)."""

    def train_step(self, obs, actions, rewards, terminals, next_obs, weights):
        c = self.cfg.learner.iqn
        with torch.no_grad():
            next_convs = self.model_target.encoder(next_obs)
            if self.cfg.learner.double_q:
                qsel = self.model.head.qval(self.model.encoder(next_obs), n=c.K)
            else:
                qsel = self.model_target.head.qval(next_convs, n=c.K)
            qt_next, _ = self.model_target.head(next_convs, n=c.N_dash)
            qsel, qt_next = self._f32(qsel), self._f32(qt_next)
        q, taus = self.model.head(self.model.encoder(obs), n=c.N)
        q, taus = self._f32(q), self._f32(taus)
        return L.iqn_loss(q, taus, qt_next, qsel, actions, rewards, terminals, weights, self.gamma_n, **self._kw()), q


class FQFLearner(BaseLearner):
    """agent.py:330-388: quantile loss at tau_hat plus the fraction-proposal loss, which has its own
    RMSprop step taken *before* the q-loss backward (agent.py:139-148)."""

    def __init__(self, cfg, **kw):
        super().__init__(cfg, **kw)
        fp = list(self.model.head.fraction_net.parameters())
        self.fqf_optimizer = torch.optim.RMSprop(fp, lr=cfg.learner.learning_rate / 2e4, alpha=0.95, eps=0.00001,
                                                 capturable=self.capturable)
        self.frac_bucket = FlatGradBucket(fp, self.pg)

    def train_step(self, obs, actions, rewards, terminals, next_obs, weights):
        head = self.model.head
        convs = self.model.encoder(obs)
        taus, taus_hat, _ = head.prop_taus(convs.detach())
        q_hat, _ = head(convs, taus=taus_hat)
        with torch.no_grad():
            next_convs = self.model_target.encoder(next_obs)
            if self.cfg.learner.double_q:
                qsel = head.qval(self.model.encoder(next_obs))
            else:
                qsel = self.model_target.head.qval(next_convs)
            qt_next, _ = self.model_target.head(next_convs, taus=taus_hat)
            q_bar, _ = head(convs, taus=taus[:, 1:-1])
            qsel, qt_next, q_bar = self._f32(qsel), self._f32(qt_next), self._f32(q_bar)
        q_hat, taus, taus_hat = self._f32(q_hat), self._f32(taus), self._f32(taus_hat)
        out = L.fqf_loss(q_hat, taus, taus_hat, qt_next, q_bar, qsel, actions, rewards, terminals, weights,
                         self.gamma_n, **self._kw())
        # fraction step first (agent.py:139-148): only fraction_net receives this gradient
        self.frac_bucket.zero_()
        taus.squeeze(-1).backward(out.grad_taus)
        self.frac_bucket.all_reduce()
        if self.cfg.learner.max_grad_norm > 0:
            nn.utils.clip_grad_norm_(head.fraction_net.parameters(), self.cfg.learner.max_grad_norm)
        self.fqf_optimizer.step()
        return out, q_hat


LEARNERS = {c.__name__: c for c in (DQNLearner, MDQNLearner, C51Learner, QRLearner, IQNLearner, FQFLearner)}


def make_learner(cfg, **kw):
    """trainer.py:31-34: getattr(agents, f"{algo.name.upper()}Learner")(cfg)."""
    name = f"{_algo_name(cfg).upper()}Learner"
    if name not in LEARNERS:
        raise NotImplementedError(f"No such learner for {_algo_name(cfg)}")
    return LEARNERS[name](cfg, **kw)
