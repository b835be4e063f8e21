"""K5 on the GPU through the C ABI: Actor.act's epsilon-greedy selection (agent0/deepq/agent.py:25-39)
against outputs of the unmodified reference (tests/golden/act.npz) and the oracle restatement.
Bit-exact: actions, normalised observations; <=1e-6 relative: the mean of the per-env max."""
import numpy as np
import pytest
import torch

from agent0_b200.config import make_config
from oracle import reference_replay as OR

pytestmark = pytest.mark.gpu


class _TableNet(torch.nn.Module):
    """Stands in for DeepQNet where the kernel's input must be exactly the reference's q-values."""

    def __init__(self, q):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1, device="cuda"))
        self.q = torch.as_tensor(q, device="cuda")
        self.seen = None

    def qval(self, st):
        self.seen = st.clone()
        return self.q


@pytest.mark.parametrize("algo", ["dqn", "c51"])
def test_actor_policy_matches_reference_act(golden, algo):
    from agent0_b200.actor import ActorPolicy
    from agent0_b200.replay import NORM_DIV
    g = golden("act")
    cfg = make_config(algo, action_dim=4, num_envs=16)
    obs = g[f"{algo}_obs_first"]
    for k, (eps, seed) in enumerate(zip(g["epsilon"], g["np_seed"])):
        net = _TableNet(g[f"{algo}_q"][k])
        pol = ActorPolicy(cfg, net, norm_mode=NORM_DIV)
        np.random.seed(int(seed))                       # numpy's global generator, as the reference
        a, m = pol.act(obs, float(eps))
        assert a.dtype == np.int64 and np.array_equal(a, g[f"{algo}_action"][k])
        assert abs(m - g[f"{algo}_qmax_mean"][k]) <= 1e-6 * max(1.0, abs(g[f"{algo}_qmax_mean"][k]))
        # the network saw exactly the reference's input (agent.py:27 on the CPU: true division)
        assert np.array_equal(net.seen.cpu().numpy().view(np.int32), g[f"{algo}_st_first"].view(np.int32))
        # the same generator state is consumed: the next draw agrees with the reference's stream
        np.random.seed(int(seed))
        OR.act_rule(g[f"{algo}_q"][k], float(eps), 4)
        nxt = np.random.rand()
        np.random.seed(int(seed))
        pol.act(obs, float(eps))
        assert np.random.rand() == nxt


def test_actor_policy_with_the_real_network_and_private_rng():
    """Real DeepQNet, many envs, a private RandomState: equals the oracle rule applied to the q-values
    torch computes from torch's own cast/divide of the same observations on this device."""
    from agent0_b200.actor import ActorPolicy
    from agent0_b200.model import DeepQNet
    cfg = make_config("qr", action_dim=18, num_envs=70)
    torch.manual_seed(3)
    net = DeepQNet(cfg).cuda()
    rs = np.random.RandomState(9)
    obs = rs.randint(0, 256, (70, 4, 84, 84)).astype(np.uint8)
    pol = ActorPolicy(cfg, net, rng=np.random.RandomState(42))
    for eps in (0.0, 0.5, 1.0):
        a, m = pol.act(obs, eps)
        with torch.no_grad():
            q = net.qval(torch.from_numpy(obs).cuda().float().div(255.0)).cpu().numpy()
        if eps == 0.0:
            assert np.array_equal(a, q.argmax(-1))
        assert a.min() >= 0 and a.max() < 18
        assert abs(m - float(q.max(-1).mean())) <= 1e-5 * abs(float(q.max(-1).mean())) + 1e-6
    ref_rng = np.random.RandomState(42)
    pol2 = ActorPolicy(cfg, net, rng=np.random.RandomState(42))
    for eps in (0.0, 0.5, 1.0):
        with torch.no_grad():
            q = net.qval(torch.from_numpy(obs).cuda().float().div(255.0)).cpu().numpy()
        want, _ = OR.act_rule(q, eps, 18, rng=ref_rng)
        got, _ = pol2.act(obs, eps)
        assert np.array_equal(got, want)


def test_u8_to_f32_all_modes_and_errors():
    from agent0_b200 import _lib
    lib = _lib.load()
    x = torch.arange(256 * 16, dtype=torch.int64).remainder(256).to(torch.uint8).cuda()
    out = torch.empty(x.numel(), dtype=torch.float32, device="cuda")
    xf = x.cpu().numpy().astype(np.float32)
    want = {0: xf / np.float32(255.0), 1: xf * (np.float32(1.0) / np.float32(255.0)), 2: xf}
    for mode in (0, 1, 2):
        _lib.check(lib.a0_u8_to_f32(x.data_ptr(), out.data_ptr(), x.numel(), mode, _lib.stream_ptr()), "a0_u8_to_f32")
        assert np.array_equal(out.cpu().numpy().view(np.int32), want[mode].view(np.int32))
    assert torch.equal(out.new_tensor(want[1]), x.float().div(255.0))          # mode 1 is torch's CUDA .div(255)
    assert lib.a0_u8_to_f32(x.data_ptr(), out.data_ptr(), 17, 0, _lib.stream_ptr()) == -1
    assert lib.a0_u8_to_f32(x.data_ptr(), out.data_ptr(), 16, 5, _lib.stream_ptr()) == -1
    assert lib.a0_act_epsilon_greedy(out.data_ptr(), 4, 64, 0.1, None, None, None, None, None, None, _lib.stream_ptr()) == -1
    assert lib.a0_u8_to_f32(None, None, 0, 0, _lib.stream_ptr()) == 0


class _ScriptedVecEnv:
    """The recorded stream of tests/golden/replay_n*.npz behind the vector-env contract (the same stand-in
    tests/golden/make_golden.py puts behind make_atari for the reference Actor)."""

    def __init__(self, g):
        self.g, self.k = g, 0

    def reset(self):
        return self.g["stream_obs"][0].copy(), {}

    def step(self, action):
        k = self.k
        assert np.array_equal(action, self.g["stream_action"][k])
        self.k += 1
        info = {"life_loss": self.g["stream_life_loss"][k]}
        return (self.g["stream_obs"][k + 1].copy(), self.g["stream_reward"][k].copy(), self.g["stream_terminal"][k].copy(),
                self.g["stream_truncated"][k].copy(), info)

    def close(self):
        pass


@pytest.mark.parametrize("n", [1, 3])
def test_shard_actor_fills_the_ring_like_the_reference_actor_fills_its_deque(golden, n):
    """ShardActor.sample over the recorded env stream (scripted actions, as the fixture's reference run):
    gathering the shard must give the reference Actor.sample's n-step entries bit for bit."""
    from agent0_b200.actor import ShardActor
    from agent0_b200.replay import ReplayDataset
    g = golden(f"replay_n{n}")
    E, T = int(g["num_envs"]), int(g["steps"])
    cfg = make_config("dqn", per=True, n_step=n, batch_size=8, replay_size=256, num_envs=E)
    cfg.actor.sample_steps = T
    rp = ReplayDataset(cfg, native_nstep=True)

    class _Script:
        model = None

        def __init__(self):
            self.it = iter(g["stream_action"])

        def act(self, obs, eps):
            return next(self.it), 0.5

    actor = ShardActor(cfg, _ScriptedVecEnv(g), _Script(), rp)
    count, rs, qs = actor.sample(1.0)
    assert count == T * E and len(qs) == T and rs == []
    assert rp.top == (T - n + 1) * E
    # the reference emitted T*E entries: its first (n-1)*E are warm-up entries (tracker shorter than n,
    # agent.py:64-73), which the ring rejects and counts
    assert actor.warmup_entries_skipped == (n - 1) * E == len(g["entry_action"]) - rp.top
    ks, es = np.meshgrid(np.arange(n - 1, T), np.arange(E), indexing="ij")
    ref_i = (ks * E + es).reshape(-1)
    pos = torch.as_tensor(((ks - n + 1) * E + es).reshape(-1), device="cuda")
    b = rp.gather(pos)
    assert np.array_equal(b.frames.cpu().numpy(), g["entry_frames"][ref_i])
    assert np.array_equal(b.actions.cpu().numpy(), g["entry_action"][ref_i])
    assert np.array_equal(b.rewards.cpu().numpy().view(np.int64), g["entry_reward"][ref_i].view(np.int64))
    assert np.array_equal(b.terminals.cpu().numpy(), g["entry_done"][ref_i])


def test_actor_and_trainer_run_the_whole_loop_together():
    """Trainer.run's loop (trainer.py:171-184) with every piece on the new path: ActorPolicy + ShardActor
    feed the shard one step at a time, Trainer(graph, fused input, the sampler's own generator) learns
    from it; counts, priorities and the tree stay consistent."""
    from agent0_b200.actor import ActorPolicy, ShardActor
    from agent0_b200.synth import SyntheticStreams
    from agent0_b200.trainer import Trainer
    E = 4
    cfg = make_config("c51", per=True, n_step=3, batch_size=8, double_q=True, dueling=True, replay_size=512, num_envs=E)
    cfg.actor.sample_steps = 20
    cfg.learner.learner_steps = 4
    cfg.learner.target_update_freq = 8
    tr = Trainer(cfg, native_nstep=True, graph=True, fused_input=True, sampler_seed=3)
    actor = ShardActor(cfg, SyntheticStreams(E, seed=5, noise=True), ActorPolicy(cfg, tr.learner.model, rng=np.random.RandomState(0)),
                       tr.replay)
    losses = []
    for it in range(8):                    # 8 x 20 x 4 = 640 transitions through a 512-record ring: it wraps
        n, rs, qs = actor.sample(0.5)
        assert n == 20 * E and len(qs) == 20 and all(np.isfinite(q) for q in qs)
        out = tr.learn()
        torch.cuda.synchronize()
        losses.append(float(torch.stack([q.mean() for q, _ in out]).mean()))
    assert all(np.isfinite(l) for l in losses) and tr.learner.update_steps == 32
    rp = tr.replay
    assert 0 < rp.top <= 512 and rp.index.tail_q > 0
    leaves = rp.priority.leaves().cpu().numpy()
    assert (leaves > 0).sum() == rp.top and (leaves != 1.0).sum() > 20       # priorities were rewritten
    tree = rp.tree.cpu().numpy()
    for node in (1, 2, 3, rp.P // 2, rp.P - 1):
        assert tree[node] == np.float32(tree[2 * node] + tree[2 * node + 1])
