"""The six learners (CUDA K4 + PyTorch CNN) against three real reference ``train()`` updates made
on CPU from the same seed (tests/golden/learner_*.npz): same initial weights (bit-identical
constructors, tests/test_model_parity.py), same batches, same IS weights.  The first loss depends
only on the forward pass and the K4 rule; the next two also on the backward from the kernel's
gradient, Adam/RMSprop and the target sync.  cuDNN-vs-CPU convolution rounding bounds the match."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = {"dqn": (True, False), "mdqn": (False, True), "c51": (True, True), "qr": (False, False),
         "iqn": (True, True), "fqf": (True, False)}


@pytest.mark.parametrize("algo", sorted(CASES))
def test_learner_train_matches_reference_updates(golden, algo):
    from agent0_b200.config import make_config
    from agent0_b200.learner import LEARNERS, make_learner
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = golden(f"learner_{algo}")
    entries = golden("replay_n3")
    double, dueling = CASES[algo]
    B = int(g["batch"])
    cfg = make_config(algo, per=True, n_step=3, batch_size=B, double_q=double, dueling=dueling, replay_size=256)
    cfg.learner.target_update_freq = 2
    torch.manual_seed(int(g["seed"]))
    learner = make_learner(cfg)
    assert type(learner) is LEARNERS[f"{algo.upper()}Learner"]
    for it in range(3):
        idx = g[f"idx{it}"]
        data = (torch.from_numpy(entries["entry_frames"][idx]).cuda(),          # uint8 straight from K3
                torch.from_numpy(entries["entry_action"][idx]).cuda(),
                torch.from_numpy(entries["entry_reward"][idx]).cuda(),          # float64, cast inside
                torch.from_numpy(entries["entry_done"][idx]).cuda(),
                torch.from_numpy(g[f"w{it}"]).cuda(), torch.from_numpy(idx).cuda())
        res = learner.train(data)
        rtol = 2e-3 if it == 0 else 5e-2
        np.testing.assert_allclose(res["q_loss"].cpu().numpy(), g[f"q_loss{it}"], rtol=rtol, atol=2e-7)
        if algo == "fqf":
            np.testing.assert_allclose(res["fraction_loss"].cpu().numpy(), g[f"fraction_loss{it}"], rtol=rtol, atol=2e-6)
        else:
            assert res["fraction_loss"] is None
        assert np.array_equal(res["indices"].cpu().numpy(), idx)
    assert learner.update_steps == int(g["update_steps"]) == 3
    ours = np.array([p.detach().abs().sum().item() for p in learner.model.parameters()])
    np.testing.assert_allclose(ours, g["param_abs_sum"], rtol=2e-4)
    tgt = np.array([p.detach().abs().sum().item() for p in learner.model_target.parameters()])
    np.testing.assert_allclose(tgt, g["target_abs_sum"], rtol=2e-4)     # synced at update 2, not at 3


def test_trainer_step_end_to_end(golden):
    """Reference tuples -> Trainer.step: extend, K2a/K3 sample of L batches, L learner updates,
    priority updates; the tree stays consistent and nothing leaves the device."""
    from agent0_b200.config import make_config
    from agent0_b200.trainer import Trainer
    g = golden("replay_n3")
    cfg = make_config("c51", per=True, n_step=3, batch_size=8, double_q=True, dueling=True, replay_size=256, num_envs=3)
    cfg.trainer.training_start_steps = 50
    cfg.learner.learner_steps = 4
    tr = Trainer(cfg)
    M = len(g["entry_action"])
    tup = [(g["entry_frames"][i].tobytes(), g["entry_action"][i], g["entry_reward"][i], g["entry_done"][i]) for i in range(M)]
    r0 = tr.step(tup[:45])
    assert r0["loss"] is None and tr.learner.update_steps == 0           # below training_start_steps
    r1 = tr.step(tup[45:])
    assert tr.learner.update_steps == 4 and r1["loss"] is not None and np.isfinite(r1["loss"])
    leaves = tr.replay.priority.leaves().cpu().numpy()
    assert (leaves[:M] > 0).all() and (leaves[M:] == 0).all()
    assert (leaves[:M] != 1.0).sum() >= 8                                # priorities were rewritten
    tree = tr.replay.tree.cpu().numpy()
    P = tr.replay.P
    for node in (1, 2, 3, P // 2, P - 1):
        assert tree[node] == np.float32(tree[2 * node] + tree[2 * node + 1])
    assert tr.replay.max_p >= 1.0


@pytest.mark.parametrize("algo", ["c51", "dqn", "qr", "iqn"])
def test_graphed_updates_match_the_eager_loop(golden, algo):
    """Trainer(graph=True): the L updates of a step replayed as one CUDA graph must follow the eager
    loop (same draws, same batches): first step eager+capture, later steps replayed."""
    from agent0_b200.config import make_config
    from agent0_b200.trainer import Trainer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = golden("replay_n3")
    M = len(g["entry_action"])
    tup = [(g["entry_frames"][i].tobytes(), g["entry_action"][i], g["entry_reward"][i], g["entry_done"][i]) for i in range(M)]
    runs = {}
    for graph in (False, True):
        cfg = make_config(algo, per=True, n_step=3, batch_size=8, double_q=True, dueling=True, replay_size=256, num_envs=3)
        cfg.trainer.training_start_steps = 10
        cfg.learner.learner_steps = 4
        cfg.learner.target_update_freq = 8
        torch.manual_seed(11)
        tr = Trainer(cfg, graph=graph)
        if algo == "iqn":                                   # same tau source on both arms
            tr.learner.model.head.device_taus = tr.learner.model_target.head.device_taus = True
        tr.replay.extend(tup)
        torch.cuda.manual_seed(5)
        losses = []
        for step in range(4):
            out = tr.learn()
            losses.append(torch.stack([q for q, _ in out]).clone())
        runs[graph] = (torch.stack(losses).cpu().numpy(), [p.detach().clone() for p in tr.learner.model.parameters()],
                       [p.detach().clone() for p in tr.learner.model_target.parameters()], tr.learner.update_steps,
                       tr.replay.tree.clone())
    assert runs[False][3] == runs[True][3] == 16
    tol = dict(rtol=5e-3, atol=1e-5) if algo != "iqn" else dict(rtol=5e-2, atol=1e-4)
    np.testing.assert_allclose(runs[True][0], runs[False][0], **tol)
    # Adam normalises every coordinate's step to ~lr, so a last-bit difference between the two
    # optimizer implementations (python-float vs device step counters) on a near-zero gradient can
    # move a weight by up to lr per update: bound the drift by updates*lr and require it to be rare
    lr, updates = cfg.learner.learning_rate, 16
    for arm in (1, 2):                                      # online net, target net (synced at 8 and 16)
        for a, b in zip(runs[True][arm], runs[False][arm]):
            diff = (a - b).abs()
            assert float(diff.max()) <= updates * lr * 1.01
            assert float(diff.mean()) < 0.05 * lr and float((diff > lr).float().mean()) < 1e-3
    torch.testing.assert_close(runs[True][4], runs[False][4], rtol=5e-3, atol=1e-5)


@pytest.mark.parametrize("graph", [False, True])
def test_fused_input_trainer_equals_unfused(golden, graph):
    """Trainer(fused_input=True): K3 writes the CNN's normalised f32 inputs directly (a0_rb_gather_f32,
    x * fl(1/255) like torch's CUDA .div(255)); losses, weights and tree must equal the u8-gather +
    torch cast/divide/split path bit for bit."""
    from agent0_b200.config import make_config
    from agent0_b200.trainer import Trainer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = golden("replay_n3")
    M = len(g["entry_action"])
    tup = [(g["entry_frames"][i].tobytes(), g["entry_action"][i], g["entry_reward"][i], g["entry_done"][i]) for i in range(M)]
    runs = {}
    for fused in (False, True):
        cfg = make_config("c51", per=True, n_step=3, batch_size=8, double_q=True, dueling=True, replay_size=256, num_envs=3)
        cfg.trainer.training_start_steps = 10
        cfg.learner.learner_steps = 4
        cfg.learner.target_update_freq = 8
        torch.manual_seed(11)
        tr = Trainer(cfg, graph=graph, fused_input=fused)
        tr.replay.extend(tup)
        torch.cuda.manual_seed(5)
        losses = [torch.stack([q for q, _ in tr.learn()]).clone() for _ in range(3)]
        runs[fused] = (torch.stack(losses), [p.detach().clone() for p in tr.learner.model.parameters()], tr.replay.tree.clone())
    # non-contiguous split views vs contiguous inputs can pick different cuDNN kernels: allow fp32 noise
    torch.testing.assert_close(runs[True][0], runs[False][0], rtol=2e-4, atol=1e-6)
    torch.testing.assert_close(runs[True][2], runs[False][2], rtol=2e-4, atol=1e-6)
    for a, b in zip(runs[True][1], runs[False][1]):
        assert float((a - b).abs().max()) <= 12 * cfg.learner.learning_rate * 1.01


def test_trainer_with_the_samplers_own_generator_draws_fresh_batches_per_replay(golden):
    """Trainer(graph=True, sampler_seed=...): uniforms come from the sampler's Philox stream, whose call
    counter lives on the device, so every replay of the captured updates draws new batches."""
    from agent0_b200.config import make_config
    from agent0_b200.trainer import Trainer
    g = golden("replay_n3")
    M = len(g["entry_action"])
    tup = [(g["entry_frames"][i].tobytes(), g["entry_action"][i], g["entry_reward"][i], g["entry_done"][i]) for i in range(M)]
    cfg = make_config("c51", per=True, n_step=3, batch_size=8, double_q=True, dueling=True, replay_size=256, num_envs=3)
    cfg.trainer.training_start_steps = 10
    cfg.learner.learner_steps = 4
    cfg.learner.target_update_freq = 8
    tr = Trainer(cfg, graph=True, fused_input=True, sampler_seed=99)
    tr.replay.extend(tup)
    seen = []
    for _ in range(4):
        out = tr.learn()
        torch.cuda.synchronize()
        assert all(torch.isfinite(q).all() for q, _ in out)
        seen.append(tr._graphed.static.indices.clone())
    assert tr.learner.update_steps == 16
    assert not torch.equal(seen[1], seen[2]) and not torch.equal(seen[2], seen[3])     # replays 2 and 3 drew different batches
    tree = tr.replay.tree.cpu().numpy()
    P = tr.replay.P
    for node in (1, 2, 3, P // 2, P - 1):
        assert tree[node] == np.float32(tree[2 * node] + tree[2 * node + 1])


@pytest.mark.parametrize("algo", sorted(CASES))
def test_bf16_autocast_learner_tracks_the_fp32_learner(golden, algo):
    """Opt-in mixed precision: the networks under bf16 autocast, K4 / Adam / priorities in fp32.
    NOT the reference's arithmetic -- so the check is the one a mixed-precision path can meet: from the same initial
    weights and the same batch the first per-sample losses agree with the fp32 learner to bf16 rounding of the network
    outputs, gradients flow (parameters move, fp32 master weights), everything stays finite over three updates."""
    from agent0_b200.config import make_config
    from agent0_b200.learner import make_learner
    g = golden(f"learner_{algo}")
    entries = golden("replay_n3")
    double, dueling = CASES[algo]
    B = int(g["batch"])
    cfg = make_config(algo, per=True, n_step=3, batch_size=B, double_q=double, dueling=dueling, replay_size=256)
    cfg.learner.target_update_freq = 2
    torch.manual_seed(int(g["seed"]))
    ref = make_learner(cfg)
    torch.manual_seed(int(g["seed"]))
    amp = make_learner(cfg, amp_dtype=torch.bfloat16)
    assert all(torch.equal(a, b) for a, b in zip(ref.model.parameters(), amp.model.parameters()))
    p0 = [p.detach().clone() for p in amp.model.parameters()]
    for it in range(3):
        idx = g[f"idx{it}"]
        data = (torch.from_numpy(entries["entry_frames"][idx]).cuda(), torch.from_numpy(entries["entry_action"][idx]).cuda(),
                torch.from_numpy(entries["entry_reward"][idx]).cuda(), torch.from_numpy(entries["entry_done"][idx]).cuda(),
                torch.from_numpy(g[f"w{it}"]).cuda(), torch.from_numpy(idx).cuda())
        if algo in ("iqn", "fqf"):
            torch.manual_seed(100 + it)             # the tau draws come from the CPU generator
        r32 = ref.train(data)
        if algo in ("iqn", "fqf"):
            torch.manual_seed(100 + it)
        r16 = amp.train(data)
        a, b = r16["q_loss"].float().cpu().numpy(), r32["q_loss"].cpu().numpy()
        assert r16["q_loss"].dtype == torch.float32 and np.isfinite(a).all()
        if it == 0:                                 # later updates start from slightly different weights
            assert abs(a.mean() - b.mean()) <= 3e-2 * abs(b.mean()) + 1e-3, (a.mean(), b.mean())
            assert np.abs(a - b).max() <= 0.1 * np.abs(b).max() + 1e-2
    moved = sum(float((p.detach() - q).abs().sum()) for p, q in zip(amp.model.parameters(), p0))
    assert moved > 0 and all(p.dtype == torch.float32 and torch.isfinite(p).all() for p in amp.model.parameters())
    assert amp.update_steps == 3


def test_amp_trainer_consumes_the_bf16_gather_eagerly_and_graphed():
    """Trainer(amp=True): K3 writes obs / next_obs as bf16 (a0_rb_gather_bf16) and the learner consumes them under
    autocast; as a CUDA graph the L updates replay with finite losses and priorities that keep changing."""
    from agent0_b200.config import make_config
    from agent0_b200.synth import fill_shard_synthetic
    from agent0_b200.trainer import Trainer
    for graph in (False, True):
        cfg = make_config("c51", per=True, n_step=3, batch_size=16, double_q=True, dueling=True, replay_size=4096, num_envs=8)
        cfg.learner.learner_steps = 4
        cfg.learner.target_update_freq = 8
        tr = Trainer(cfg, native_nstep=True, graph=graph, amp=True, sampler_seed=3)
        fill_shard_synthetic(tr.replay, 4096, 8, 1)
        root0 = float(tr.replay.tree[1])
        for _ in range(3):
            outs = tr.learn()
        torch.cuda.synchronize()
        assert len(outs) == 4 and all(torch.isfinite(q).all() and q.dtype == torch.float32 for q, _ in outs)
        assert tr.learner.update_steps == 12 and float(tr.replay.tree[1]) != root0
        if graph:
            assert tr._graphed.static.obs.dtype == torch.bfloat16
        tree = tr.replay.tree.cpu().numpy()
        for node in (1, 2, 3, tr.replay.P // 2, tr.replay.P - 1):
            assert tree[node] == np.float32(tree[2 * node] + tree[2 * node + 1])
