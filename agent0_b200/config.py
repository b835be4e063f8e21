"""Configuration objects with the attribute names of the reference's ``agent0.deepq.config``.

The replay shard and the learners only *read* attributes (``cfg.replay.size``,
``cfg.learner.n_step_q``, ...), so the reference's own ``ExpConfig`` instance can be passed
unchanged; this module exists for hosts where the reference package is not importable (the GPU
box, the tests, bench.py).  Defaults follow agent0/deepq/config.py:42-145.  Enum members are
compared by ``.name`` everywhere, so either package's enums work.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Any, Tuple

AlgoEnum = enum.Enum("AlgoEnum", "dqn c51 qr iqn fqf mdqn", start=0)          # config.py:6-12
ReplayEnum = enum.Enum("ReplayEnum", "uniform prioritize", start=0)           # config.py:21-23


class DeviceEnum(enum.Enum):                                                   # config.py:37-39
    cuda = "cuda"
    cpu = "cpu"


def _algo(x):
    return x if isinstance(x, AlgoEnum) else AlgoEnum[getattr(x, "name", x)]


@dataclass
class C51Config:
    num_atoms: int = 51
    vmin: float = -10
    vmax: float = 10


@dataclass
class QRConfig:
    num_atoms: int = 200
    vmin: Any = None
    vmax: Any = None


@dataclass
class IQNConfig:
    K: int = 32            # tau samples for action selection
    N: int = 64            # online tau samples
    N_dash: int = 64       # target tau samples
    num_cosines: int = 64
    F: int = 32            # FQF fractions


@dataclass
class MDQNConfig:
    tau: float = 0.03
    alpha: float = 0.9     # unused by the reference's loss; kept for parity of the surface
    lo: float = -1


@dataclass
class LearnerConfig:
    algo: Any = AlgoEnum.dqn
    discount: float = 0.99
    batch_size: int = 512
    learning_rate: float = 5e-4
    fraction_lr: float = 2.5e-8
    max_grad_norm: float = -1.0
    target_update_freq: int = 500
    learner_steps: int = 20
    double_q: bool = False
    dueling_head: bool = False
    n_step_q: int = 1
    noisy_net: bool = False
    reset_noise_freq: int = 4
    c51: C51Config = field(default_factory=C51Config)
    qr: QRConfig = field(default_factory=QRConfig)
    iqn: IQNConfig = field(default_factory=IQNConfig)
    mdqn: MDQNConfig = field(default_factory=MDQNConfig)


@dataclass
class TrainerConfig:
    total_steps: int = 10_000_000
    training_start_steps: int = 100_000
    exploration_steps: int = 1_000_000
    log_freq: int = 10
    test_freq: int = 500
    test_episodes: int = 20


@dataclass
class ActorConfig:
    num_envs: int = 16
    sample_steps: int = 80
    test_steps: int = 800
    min_eps: float = 0.01
    test_eps: float = 0.001


@dataclass
class ReplayConfig:
    size: int = 1_000_000
    policy: Any = ReplayEnum.uniform
    beta0: float = 0.4
    alpha: float = 0.5
    eps: float = 0.01


@dataclass
class ExpConfig:
    env_id: str = "Breakout"
    obs_shape: Tuple[int, ...] = (4, 84, 84)
    action_dim: int = 4
    num_actors: int = 3
    seed: int = 42
    device: Any = DeviceEnum.cuda
    name: str = "agent0"
    logdir: str = "logs"
    wandb: bool = False
    tb: bool = False
    learner: LearnerConfig = field(default_factory=LearnerConfig)
    trainer: TrainerConfig = field(default_factory=TrainerConfig)
    actor: ActorConfig = field(default_factory=ActorConfig)
    replay: ReplayConfig = field(default_factory=ReplayConfig)


def make_config(algo="dqn", *, per=False, n_step=1, batch_size=512, replay_size=1_000_000, double_q=False,
                dueling=False, action_dim=4, num_envs=16, device="cuda", **learner_kw) -> ExpConfig:
    """Shorthand used by tests and bench.py, e.g. BASELINE config 2:
    make_config("c51", per=True, n_step=3, double_q=True, dueling=True, batch_size=32)."""
    cfg = ExpConfig()
    cfg.action_dim = action_dim
    cfg.device = DeviceEnum(device) if isinstance(device, str) and device in ("cuda", "cpu") else device
    cfg.learner.algo = _algo(algo)
    cfg.learner.n_step_q = n_step
    cfg.learner.batch_size = batch_size
    cfg.learner.double_q = double_q
    cfg.learner.dueling_head = dueling
    for k, v in learner_kw.items():
        setattr(cfg.learner, k, v)
    cfg.replay.size = replay_size
    cfg.replay.policy = ReplayEnum.prioritize if per else ReplayEnum.uniform
    cfg.actor.num_envs = num_envs
    return cfg
