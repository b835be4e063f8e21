"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/*.npz,
made by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import losses as OL
from oracle import reference_replay as OR
from oracle.sumtree import SumTree, new_priority

RTOL = 1e-5   # fp32 losses/grads/priorities (north_star tolerance)


def close(a, b, rtol=RTOL, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64),
                               rtol=rtol, atol=atol)


# ------------------------------------------------------------------ n-step packer / entries
@pytest.mark.parametrize("n", [1, 3])
def test_pack_nstep_matches_reference_entries(golden, n):
    g = golden(f"replay_n{n}")
    done = OR.done_rule(g["stream_terminal"], g["stream_life_loss"], g["stream_truncated"])
    frames, a, r, d = OR.pack_nstep(g["stream_obs"], g["stream_action"], g["stream_reward"], done,
                                    int(g["n_step"]), float(g["discount"]))
    assert np.array_equal(frames, g["entry_frames"])            # bit-exact uint8 stacks
    assert np.array_equal(a, g["entry_action"])
    assert np.array_equal(r.view(np.int64), g["entry_reward"].view(np.int64))   # bit-exact f64
    assert np.array_equal(d, g["entry_done"])
    assert np.array_equal(g["entry_idx"], np.arange(len(a)))


def test_synth_stream_regenerates_from_seed(golden):
    from agent0_b200.synth import record_stream
    g = golden("replay_n3")
    s = record_stream(int(g["num_envs"]), int(g["steps"]), seed=int(g["seed"]), p_terminal=0.07,
                      p_life_loss=0.08, p_truncated=0.05)
    assert np.array_equal(s["obs"], g["stream_obs"])
    assert np.array_equal(s["reward"], g["stream_reward"])


# ------------------------------------------------------------------ PER bookkeeping
def test_per_bookkeeping_matches_reference(golden):
    g = golden("per_state")
    rp = OR.RefReplay(int(g["size"]), True, float(g["alpha"]), float(g["eps"]), float(g["beta0"]),
                      int(g["total_steps"]))
    dummy = (b"x", 0, 0.0, False)
    for i, op in enumerate(g["op"]):
        if op == 0:
            rp.extend([dummy] * int(g["count"][i]))
        else:
            rp.update_priority(g["ids"][i], g["loss"][i])
        close(rp.priority, g["priority"][i])
        assert rp.top == int(g["top"][i])
        assert rp.beta == float(g["beta"][i])
        close(rp.max_p, g["max_p"][i], rtol=1e-7)


def test_trainer_step_is_weights_and_priorities(golden):
    g = golden("trainer_step")
    size, top = int(g["size"]), int(g["top"])
    prio = np.ones(size, dtype=np.float32)        # state after the single extend (max_p = 1)
    max_p = 1.0
    for it in range(len(g["batches"])):
        idx = g["batches"][it] % top
        w = OR.is_weights(prio[idx], prio.sum(), top, float(g["beta_used"]) if it == 0 else float(g["beta_after"]))
        # beta: extend() stores the schedule value *before* advancing (replay.py:53)
        w0 = OR.is_weights(prio[idx], prio.sum(), top, float(g["beta_used"]))
        close(w0, g["weights"][it])
        assert np.array_equal(g["indices"][it], idx)
        prio[idx] = new_priority(g["q_loss"][it], float(g["eps"]), float(g["alpha"]))
        max_p = max(max_p, float(g["q_loss"][it].max()))
        close(prio, g["prio_after"][it])
        close(max_p, g["max_p_after"][it], rtol=1e-7)
        del w


# ------------------------------------------------------------------ static fns
def test_huber_qr_and_log_softmax_stable(golden):
    g = golden("static_fns")
    close(OL.huber_qr_loss(g["q"], g["q_target"], g["taus"]), g["huber_qr"])
    close(OL.log_softmax_stable(g["logits"], 0.03), g["lss_tau003"])
    close(OL.log_softmax_stable(g["logits"]), g["lss_default"])


# ------------------------------------------------------------------ six losses
def _common(g):
    return (g["actions"], g["rewards"], g["terminals"], g["weights"], float(g["discount"]), int(g["n_step"]))


@pytest.mark.parametrize("dq", ["single", "double"])
def test_dqn(golden, dq):
    g = golden(f"loss_dqn_{dq}")
    a, r, d, w, disc, n = _common(g)
    loss, grad = OL.dqn(g["online_cur"], g["tgt_next"], g["qval_next"] if dq == "double" else None,
                        a, r, d, w, disc, n)
    close(loss, g["loss"]); close(grad, g["grad"])


def test_mdqn(golden):
    g = golden("loss_mdqn_single")
    a, r, d, w, disc, n = _common(g)
    loss, grad = OL.mdqn(g["online_cur"], g["tgt_next"], g["tgt_cur"], a, r, d, w, disc, n,
                         float(g["tau"]), float(g["lo"]))
    close(loss, g["loss"]); close(grad, g["grad"])


@pytest.mark.parametrize("dq", ["single", "double"])
def test_c51(golden, dq):
    import torch
    g = golden(f"loss_c51_{dq}")
    a, r, d, w, disc, n = _common(g)
    M = int(g["num_atoms"])
    atoms = g["atoms"]      # model.head.atoms = torch.linspace(vmin, vmax, M) (model.py:149-152)
    assert np.array_equal(atoms, torch.linspace(float(g["vmin"]), float(g["vmax"]), M).numpy())
    loss, grad, m = OL.c51(g["online_cur"], g["tgt_next"], g["qval_next"] if dq == "double" else None,
                           a, r, d, w, disc, n, atoms, float(g["vmin"]), float(g["vmax"]))
    close(loss, g["loss"]); close(grad, g["grad"])
    close(m.sum(-1), np.ones(len(a)), rtol=1e-6)


@pytest.mark.parametrize("dq", ["single", "double"])
def test_qr(golden, dq):
    g = golden(f"loss_qr_{dq}")
    a, r, d, w, disc, n = _common(g)
    loss, grad = OL.qr(g["online_cur"], g["tgt_next"], g["qval_next"] if dq == "double" else None,
                       a, r, d, w, disc, n)
    close(loss, g["loss"]); close(grad, g["grad"], atol=1e-5)


@pytest.mark.parametrize("dq", ["single", "double"])
def test_iqn(golden, dq):
    g = golden(f"loss_iqn_{dq}")
    a, r, d, w, disc, n = _common(g)
    loss, grad = OL.iqn(g["q_cur"], g["taus_cur"][..., 0], g["q_next"], g["qval_next"], a, r, d, w, disc, n)
    close(loss, g["loss"]); close(grad, g["grad"])


@pytest.mark.parametrize("dq", ["single", "double"])
def test_fqf(golden, dq):
    g = golden(f"loss_fqf_{dq}")
    a, r, d, w, disc, n = _common(g)
    loss, grad, frac, gt = OL.fqf(g["q_hat"], g["taus"], g["taus_hat"], g["q_next"], g["q_bar"],
                                  g["qval_next"], a, r, d, w, disc, n)
    close(loss, g["loss"]); close(grad, g["grad"])
    close(frac, g["fraction_loss"]); close(gt, g["grad_taus"])


# ------------------------------------------------------------------ sum-tree (new piece)
def test_sumtree_invariants_and_law():
    rng = np.random.RandomState(3)
    N = 1000
    t = SumTree(N)
    pr = (np.abs(rng.randn(N)) + 0.01).astype(np.float32) ** 0.5
    pr[rng.rand(N) < 0.2] = 0.0
    t.set(np.arange(N), pr)
    assert np.array_equal(t.leaves(), pr)
    for lvl_node in (1, 2, 3, 77, 500):
        assert t.nodes[lvl_node] == np.float32(t.nodes[2 * lvl_node] + t.nodes[2 * lvl_node + 1])
    # duplicate index: last writer wins
    t.set(np.array([5, 5]), np.array([2.0, 3.0], dtype=np.float32))
    assert t.leaves()[5] == np.float32(3.0)
    # never lands on a zero leaf, and follows the reference's categorical law P(i)=p_i/sum
    counts = np.zeros(N)
    for rep in range(200):
        idx, p = t.sample_stratified(rng.rand(256).astype(np.float32))
        assert (p > 0).all()
        np.add.at(counts, idx, 1)
    law = t.leaves().astype(np.float64) / t.leaves().astype(np.float64).sum()
    emp = counts / counts.sum()
    assert np.abs(emp - law).max() < 4e-3
    assert abs(np.corrcoef(emp, law)[0, 1]) > 0.97


# ------------------------------------------------------------------ Actor.act
@pytest.mark.parametrize("algo", ["dqn", "c51"])
def test_act_rule_matches_reference_actor(golden, algo):
    g = golden("act")
    for k, (eps, seed) in enumerate(zip(g["epsilon"], g["np_seed"])):
        np.random.seed(int(seed))
        a, m = OR.act_rule(g[f"{algo}_q"][k], float(eps), 4)
        assert np.array_equal(a, g[f"{algo}_action"][k])
        close(m, g[f"{algo}_qmax_mean"][k], rtol=1e-6)
    assert np.array_equal(OR.obs_to_float(g[f"{algo}_obs_first"]).view(np.int32), g[f"{algo}_st_first"].view(np.int32))


# ------------------------------------------------------------------ the sampler's own generator
def test_philox4x32_10_known_answers():
    """Random123's published known-answer vectors for Philox4x32-10 (kat_vectors): the oracle of
    a0_pt_sample_rng's uniforms is pinned to the published algorithm."""
    from oracle.sumtree import philox4x32_10, philox_uniform
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = philox4x32_10(np.array([c], dtype=np.uint32), np.array([k], dtype=np.uint32))[0]
        assert tuple(int(x) for x in got) == want
    u = philox_uniform(1234, 5, 100000)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    assert not np.array_equal(u[:100], philox_uniform(1234, 6, 100)) and np.array_equal(u[:100], philox_uniform(1234, 5, 100))


def test_sorted_target_form_of_the_quantile_huber_sums_matches_the_pairwise_form():
    """oracle.losses.huber_qr_sorted (the O(N log N) specification of the planned QR-200 K4)
    against the pairwise restatement of agent.py:110-114 that is pinned to the reference's own huber_qr_loss:
    loss and gradient within 1e-5 relative, on QR-, IQN- and FQF-shaped inputs, with exact ties (T == q,
    |q - T| == 1), repeated targets and far-apart values."""
    from oracle import losses as OL
    rng = np.random.RandomState(7)
    for B, Ni, Nj, scale in ((6, 200, 200, 3.0), (5, 64, 64, 1.0), (4, 32, 32, 0.3), (3, 7, 5, 10.0)):
        T = (rng.randn(B, Ni) * scale).astype(np.float32)
        q = (rng.randn(B, Nj) * scale).astype(np.float32)
        q[0, 0] = T[0, 0]                      # u == 0
        q[0, 1] = T[0, 1] + np.float32(1.0)    # u == 1
        if Nj > 2:
            q[0, 2] = T[0, 2] - np.float32(1.0)
        T[1, : Ni // 2] = T[1, 0]              # repeated targets
        tau = rng.rand(B, Nj).astype(np.float32)
        w = (rng.rand(B) + 0.1).astype(np.float32)
        loss_ref = OL.huber_qr_loss(q[:, None, :], T[:, :, None], tau[:, None, :])
        grad_ref = OL._huber_qr_grad(q, T, tau, w)
        loss, grad = OL.huber_qr_sorted(q, T, tau, w)
        np.testing.assert_allclose(loss, loss_ref, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(grad, grad_ref, rtol=1e-5, atol=2e-6)
