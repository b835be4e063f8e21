"""``bench.py --impl reference``: the UNMODIFIED reference (oracle/_ref/, copied by oracle/make_ref.py) timed on
the host cores.  Test/bench infrastructure only -- nothing under agent0_b200/ imports this.

What runs is the reference's own code, through its own call sites:
  agent0.deepq.trainer.Trainer.step            (trainer.py:74-119: the hot loop being replaced)
  agent0.deepq.replay.ReplayDataset            (replay.py:14-59: deque of lz4 blobs, __getitem__, update_priority)
  agent0.common.utils.DataLoaderX/DataPrefetcher  (utils.py:31-61, via Trainer.get_data_fetcher: shuffle=True,
                                                 num_workers=2, pin_memory=True, 3-deep background prefetch)
  agent0.deepq.agent.{DQN,MDQN,C51,QR}Learner.train / train_step   (agent.py:124-293)
with the Nature-CNN replaced by table-lookup networks (instance attributes on the learner, exactly as
tests/golden/make_golden.py does) because the BASELINE metric excludes the CNN on both arms.

Third-party modules the reference imports and this image lacks are supplied as sys.modules stubs:
  lz4.block      -> the system liblz4 with python-lz4's framing (4-byte size prefix), so the real decompression
                    cost is paid (lz4>=4.3.3, pyproject.toml)
  prefetch_generator.BackgroundGenerator -> the published behaviour restated: a daemon thread that runs the
                    wrapped generator ahead into a Queue(max_prefetch)
  agent0.common.atari_wrappers.make_atari -> a dummy vector env with the Atari shapes (Trainer.__init__ only
                    reads observation_space / action_space from it, trainer.py:26-29)
"""
import os
import queue
import sys
import tempfile
import threading
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "agent0", "deepq", "trainer.py"))


class BackgroundGenerator(threading.Thread):
    """prefetch_generator.BackgroundGenerator, restated (utils.py:59-61 wraps the DataLoader iterator in it)."""

    def __init__(self, generator, max_prefetch=1):
        super().__init__(daemon=True)
        self.queue = queue.Queue(max_prefetch)
        self.generator = generator
        self.start()

    def run(self):
        for item in self.generator:
            self.queue.put(item)
        self.queue.put(None)

    def __next__(self):
        item = self.queue.get()
        if item is None:
            raise StopIteration
        return item

    next = __next__

    def __iter__(self):
        return self


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_NS = None


def load(action_dim=4):
    """Import the copied reference with the three stubs; returns a namespace of its modules."""
    global _NS
    if _NS is not None:
        return _NS
    if not available():
        raise RuntimeError("oracle/_ref/ is missing: run `python oracle/make_ref.py` where /root/reference exists")
    from oracle import cpu_path as CP
    z = CP.lz4()
    lz4 = _stub("lz4")
    lz4.block = _stub("lz4.block", compress=z.compress, decompress=z.decompress)
    _stub("prefetch_generator", BackgroundGenerator=BackgroundGenerator)

    class _DummyVecEnv:
        def __init__(self, n):
            self.observation_space = types.SimpleNamespace(shape=(n, 4, 84, 84))
            self.action_space = [types.SimpleNamespace(n=action_dim)] * n

        def close(self):
            pass

    if REF not in sys.path:
        sys.path.insert(0, REF)
    import agent0.common  # noqa: F401
    _stub("agent0.common.atari_wrappers", make_atari=lambda env_id, n, **k: _DummyVecEnv(n))
    import agent0.deepq.agent as agents
    import agent0.deepq.trainer as trainer
    from agent0.deepq import config as cfgmod
    from agent0.deepq.replay import ReplayDataset
    assert os.path.realpath(trainer.__file__).startswith(os.path.realpath(REF)), trainer.__file__
    _NS = types.SimpleNamespace(agents=agents, trainer=trainer, config=cfgmod, ReplayDataset=ReplayDataset)
    return _NS


class TableNet:
    """Stands in for DeepQNet on a learner INSTANCE (the CNN is excluded on both arms): returns pre-generated
    outputs in the order the reference's train_step asks for them (agent.py:172-293)."""

    def __init__(self, outs, qval, head):
        self.outs, self._qval, self.head, self.k = list(outs), qval, head, 0

    def __call__(self, x):
        o = self.outs[self.k % len(self.outs)]
        self.k += 1
        return o

    def qval(self, x):
        return self._qval

    def reset_noise(self):
        pass


SUPPORTED = ("dqn", "mdqn", "c51", "qr")


def make_trainer(algo, per, n_step, double, B, L, A, outputs, replay_size=1_000_000):
    """The reference Trainer for one bench workload, table networks installed.  outputs: CPU tensors as
    bench.net_outputs makes them (online, tgt_next, tgt_cur, qsel)."""
    ns = load(A)
    C = ns.config
    cfg = C.ExpConfig()
    cfg.device = C.DeviceEnum.cpu
    cfg.obs_shape = (4, 84, 84)
    cfg.action_dim = A
    cfg.wandb = False
    cfg.tb = False
    cfg.logdir = tempfile.mkdtemp(prefix="a0_ref_")
    cfg.actor.num_envs = 16
    cfg.learner.algo = C.AlgoEnum[algo]
    cfg.learner.n_step_q = n_step
    cfg.learner.batch_size = B
    cfg.learner.double_q = bool(double)
    cfg.learner.dueling_head = True
    cfg.learner.learner_steps = L
    cfg.learner.target_update_freq = 1 << 60          # deepcopy(model) of a table would drop the target's rows
    cfg.replay.size = replay_size
    cfg.replay.policy = C.ReplayEnum.prioritize if per else C.ReplayEnum.uniform
    cfg.trainer.training_start_steps = 0
    tr = ns.trainer.Trainer(cfg, use_lp=True)
    lr = tr.learner
    head = lr.model.head
    online = outputs["online"][:B].clone().requires_grad_(True)
    tgt = [outputs["tgt_next"][:B]] + ([outputs["tgt_cur"][:B]] if algo == "mdqn" else [])
    lr.model = TableNet([online], outputs["qsel"][:B] if outputs.get("qsel") is not None else None, head)
    lr.model_target = TableNet(tgt, None, head)
    return tr


def fill(tr, entries, total):
    """``total`` deque entries made of the ``entries`` distinct reference tuples, repeated (the deque holds
    references, so a 1 M-entry deque -- the size replay.py:18 is configured for -- costs its index, not 1 GB)."""
    rep = -(-total // len(entries))
    chunk = (entries * rep)[:total]
    step = 65536
    for lo in range(0, total, step):
        tr.replay.extend(chunk[lo:lo + step])
    return len(tr.replay)


def fetcher_cpu(tr):
    """Trainer.get_data_fetcher needs a CUDA stream (utils.py:34); where there is no GPU (the build container)
    the same DataLoaderX is iterated directly -- what DataPrefetcher.next() returns, minus the device copy."""
    if torch.cuda.is_available():
        return None                                   # the reference builds its own (trainer.py:65-72)
    U = sys.modules["agent0.common.utils"]
    dl = U.DataLoaderX(tr.replay, batch_size=tr.cfg.learner.batch_size, shuffle=True, num_workers=2, pin_memory=False)
    it = iter(dl)
    return types.SimpleNamespace(next=lambda: next(it))


def step(tr):
    """One Trainer.step (trainer.py:74-119) with no new transitions: learner_steps x {fetch, IS weights,
    learner.train, update_priority}."""
    if tr.data_fetcher is None:
        f = fetcher_cpu(tr)
        if f is not None:
            tr.data_fetcher = f
    return tr.step([], [], [])
