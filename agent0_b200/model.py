"""The Nature-CNN Q-network family, in plain PyTorch (north_star: "the Nature-CNN forward/backward
stays on PyTorch").  Out of the hot path's scope, but the learners need a network on hosts where the
reference package is not importable, and a drop-in must exchange ``state_dict``s with the reference's
actors (agent0/deepq/launch.py:33-36).  So parameter/buffer names, shapes, forward semantics and the
order in which the constructors consume torch's RNG follow agent0/deepq/model.py:28-338: with the
same ``torch.manual_seed`` both packages build bit-identical weights (tests/test_learner_golden).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

RELU_GAIN = math.sqrt(2.0)


def _ortho(layer, gain):
    nn.init.orthogonal_(layer.weight.data, gain)
    nn.init.zeros_(layer.bias.data)
    return layer


class NoisyLinear(nn.Module):
    """Factorised-noise linear layer (model.py:28-87)."""

    def __init__(self, in_features, out_features, std_init=0.4, noisy_layer_std=0.1):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.std_init, self.noisy_layer_std = std_init, noisy_layer_std
        z = torch.zeros
        self.weight_mu = nn.Parameter(z(out_features, in_features))
        self.weight_sigma = nn.Parameter(z(out_features, in_features))
        self.register_buffer("weight_epsilon", z(out_features, in_features))
        self.bias_mu = nn.Parameter(z(out_features))
        self.bias_sigma = nn.Parameter(z(out_features))
        self.register_buffer("bias_epsilon", z(out_features))
        self.register_buffer("noise_in", z(in_features))
        self.register_buffer("noise_out_weight", z(out_features))
        self.register_buffer("noise_out_bias", z(out_features))
        bound = 1.0 / math.sqrt(in_features)
        self.weight_mu.data.uniform_(-bound, bound)
        self.weight_sigma.data.fill_(std_init / math.sqrt(in_features))
        self.bias_mu.data.uniform_(-bound, bound)
        self.bias_sigma.data.fill_(std_init / math.sqrt(out_features))
        self.reset_noise()

    @staticmethod
    def _f(x):
        return x.sign() * x.abs().sqrt()

    def reset_noise(self):
        for buf in (self.noise_in, self.noise_out_weight, self.noise_out_bias):
            buf.normal_(std=self.noisy_layer_std)
        self.weight_epsilon.copy_(torch.outer(self._f(self.noise_out_weight), self._f(self.noise_in)))
        self.bias_epsilon.copy_(self._f(self.noise_out_bias))

    def forward(self, x):
        if self.training:
            return F.linear(x, self.weight_mu + self.weight_sigma * self.weight_epsilon,
                            self.bias_mu + self.bias_sigma * self.bias_epsilon)
        return F.linear(x, self.weight_mu, self.bias_mu)


def _dense(noisy, n_in, n_out, gain):
    """A head layer; orthogonal init only touches nn.Linear (model.py:15-18), as in the reference."""
    if noisy:
        return NoisyLinear(n_in, n_out)
    return _ortho(nn.Linear(n_in, n_out), gain)


class ConvEncoder(nn.Module):
    """32x8x8/4 -> 64x4x4/2 -> 64x3x3/1 -> flatten (model.py:90-105); 84x84 input -> 3136 features."""

    def __init__(self, chan_dim):
        super().__init__()
        self.convs = nn.Sequential(
            nn.Conv2d(chan_dim, 32, 8, stride=4), nn.ReLU(),
            nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(),
            nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(), nn.Flatten())
        for m in self.convs:
            if isinstance(m, nn.Conv2d):
                _ortho(m, RELU_GAIN)

    def forward(self, x):
        return self.convs(x)


class DQNHead(nn.Module):
    """model.py:108-134 (also the M-DQN head)."""

    def __init__(self, act_dim, feat_dim, dueling, noisy, cfg=None):
        super().__init__()
        self.first_dense = _dense(noisy, feat_dim, 512, RELU_GAIN)
        self.q_head = _dense(noisy, 512, act_dim, 0.01)
        self.value_head = _dense(noisy, 512, 1, 1.0) if dueling else None

    def forward(self, x):
        x = F.relu(self.first_dense(x))
        q = self.q_head(x)
        if self.value_head is not None:
            q = self.value_head(x) + (q - q.mean(dim=-1, keepdim=True))
        return q

    qval = forward


class C51Head(nn.Module):
    """model.py:137-177: [B, A, atoms] logits; atoms = linspace(vmin, vmax, M) as a [1,1,M] buffer."""

    def __init__(self, act_dim, feat_dim, dueling, noisy, cfg):
        super().__init__()
        self.action_dim, self.num_atoms = act_dim, cfg.num_atoms
        self.first_dense = _dense(noisy, feat_dim, 512, RELU_GAIN)
        self.q_head = _dense(noisy, 512, act_dim * cfg.num_atoms, 0.01)
        if cfg.vmax is not None:
            self.register_buffer("atoms", torch.linspace(cfg.vmin, cfg.vmax, cfg.num_atoms).view(1, 1, -1))
            self.delta = (cfg.vmax - cfg.vmin) / (cfg.num_atoms - 1)
        self.value_head = _dense(noisy, 512, cfg.num_atoms, 1.0) if dueling else None

    def forward(self, x):
        x = F.relu(self.first_dense(x))
        q = self.q_head(x).view(-1, self.action_dim, self.num_atoms)
        if self.value_head is not None:
            q = self.value_head(x).unsqueeze(1) + (q - q.mean(dim=1, keepdim=True))
        return q

    def qval(self, x):
        return self.forward(x).softmax(dim=-1).mul(self.atoms).sum(dim=-1)


class QRHead(C51Head):
    """model.py:180-192: quantile midpoints tau_j = (2j+1)/2N; Q = mean over quantiles."""

    def __init__(self, act_dim, feat_dim, dueling, noisy, cfg):
        super().__init__(act_dim, feat_dim, dueling, noisy, cfg)
        self.register_buffer("cumulative_density", (2 * torch.arange(cfg.num_atoms) + 1) / (2.0 * cfg.num_atoms))

    def qval(self, x):
        return self.forward(x).mean(dim=-1)


class IQNHead(nn.Module):
    """model.py:195-257: cosine tau embedding multiplied into the state features; output [B, n, A]."""

    def __init__(self, act_dim, feat_dim, dueling, noisy, cfg):
        super().__init__()
        self.cfg = cfg
        self.first_dense = _dense(noisy, feat_dim, 512, RELU_GAIN)
        self.q_head = _dense(noisy, 512, act_dim, 0.01)
        self.value_head = _dense(noisy, 512, 1, 1.0) if dueling else None
        self.cosine_emb = nn.Sequential(_ortho(nn.Linear(cfg.num_cosines, feat_dim), RELU_GAIN), nn.ReLU())
        # pi * (1..num_cosines) as a non-persistent buffer (not in the state_dict, so checkpoints stay
        # interchangeable with the reference): the reference rebuilds it on the CPU and copies it to the
        # device in every forward (model.py:240), which cannot be captured in a CUDA graph
        self.register_buffer("ipi", (math.pi * torch.arange(1, cfg.num_cosines + 1)).float(), persistent=False)
        self.device_taus = False       # True: draw taus with the device generator (graph-capturable)

    def feature_emb(self, x, n, taus):
        B = x.size(0)
        if taus is None:
            if self.device_taus:
                taus = torch.rand(B, n, 1, device=x.device, dtype=x.dtype)
            else:
                taus = torch.rand(B, n, 1).to(x)   # CPU default generator, then moved
        else:
            n = taus.size(1)
        ipi = self.ipi.to(x)
        cosine = (ipi.view(1, 1, -1) * taus).cos().view(B * n, -1)
        tau_embed = self.cosine_emb(cosine).view(B, n, -1)
        return (tau_embed * x.unsqueeze(1)).view(B * n, -1), taus, n

    def forward(self, x, n=None, taus=None):
        feats, taus, n = self.feature_emb(x, n, taus)
        feats = F.relu(self.first_dense(feats))
        q = self.q_head(feats)
        if self.value_head is not None:
            q = self.value_head(feats) + (q - q.mean(dim=-1, keepdim=True))
        return q.view(-1, n, q.size(-1)), taus

    def qval(self, x, n=None):
        return self.forward(x, self.cfg.K if n is None else n)[0].mean(dim=1)


class FQFHead(IQNHead):
    """model.py:260-284: fraction proposal net; Q = sum_i (tau_{i+1}-tau_i) * q(tau_hat_i)."""

    def __init__(self, act_dim, feat_dim, dueling, noisy, cfg):
        super().__init__(act_dim, feat_dim, dueling, noisy, cfg)
        self.fraction_net = nn.Linear(feat_dim, cfg.F)
        nn.init.xavier_uniform_(self.fraction_net.weight, gain=0.01)
        nn.init.constant_(self.fraction_net.bias, 0)

    def prop_taus(self, x):
        log_probs = self.fraction_net(x).log_softmax(dim=-1)
        probs = log_probs.exp()
        taus = torch.cat((x.new_zeros(x.size(0), 1), torch.cumsum(probs, dim=-1)), dim=-1)
        taus_hat = (taus[:, :-1] + taus[:, 1:]).detach() / 2.0
        entropies = probs.mul(log_probs).neg().sum(dim=-1, keepdim=True)
        return taus.unsqueeze(-1), taus_hat.unsqueeze(-1), entropies

    def qval(self, x):
        taus, taus_hat, _ = self.prop_taus(x.detach())
        q_hat, _ = self.forward(x, taus=taus_hat)
        return ((taus[:, 1:, :] - taus[:, :-1, :]) * q_hat).sum(dim=1)


_HEADS = {"dqn": (DQNHead, None), "mdqn": (DQNHead, None), "c51": (C51Head, "c51"), "qr": (QRHead, "qr"),
          "iqn": (IQNHead, "iqn"), "fqf": (FQFHead, "iqn")}


class DeepQNet(nn.Module):
    """model.py:287-338.  ``cfg.learner.algo`` selects the head by enum *name*."""

    def __init__(self, cfg):
        super().__init__()
        self.encoder = ConvEncoder(cfg.obs_shape[0])
        feat_dim = self.encoder(torch.rand(1, *cfg.obs_shape)).shape[-1]   # also consumes RNG like the reference
        algo = getattr(cfg.learner.algo, "name", cfg.learner.algo)
        head_cls, sub = _HEADS[algo]
        self.head = head_cls(cfg.action_dim, feat_dim, cfg.learner.dueling_head, cfg.learner.noisy_net,
                             getattr(cfg.learner, sub) if sub else None)

    def forward(self, x):
        return self.head(self.encoder(x))

    def qval(self, x):
        return self.head.qval(self.encoder(x))

    def params(self):
        return (v for k, v in self.named_parameters() if "fraction" not in k)

    def reset_noise(self):
        for m in self.modules():
            if isinstance(m, NoisyLinear):
                m.reset_noise()
