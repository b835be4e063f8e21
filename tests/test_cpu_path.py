"""The timed CPU baseline (oracle/cpu_path.py) must compute what the reference computes: its loss
functions against the golden outputs of the unmodified reference, its lz4 framing round trip, and
its replay __getitem__ against the packed entries."""
import numpy as np
import pytest
import torch

from oracle import cpu_path as CP


def _t(x):
    return torch.as_tensor(np.asarray(x))


def _close(a, b):
    np.testing.assert_allclose(a.detach().numpy(), b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("algo,dq", [("dqn", "single"), ("dqn", "double"), ("mdqn", "single"), ("c51", "single"),
                                     ("c51", "double"), ("qr", "single"), ("qr", "double")])
def test_simple_losses(golden, algo, dq):
    g = golden(f"loss_{algo}_{dq}")
    online = _t(g["online_cur"]).requires_grad_(True)
    o = dict(online=online, tgt_next=_t(g["tgt_next"]), tgt_cur=_t(g["tgt_cur"]),
             qsel=_t(g["qval_next"]) if dq == "double" else None)
    gam = float(g["discount"]) ** int(g["n_step"])
    extra = dict(atoms=_t(g["atoms"])) if algo == "c51" else {}
    loss = CP.LOSSES[algo](o, _t(g["actions"]), _t(g["rewards"]), _t(g["terminals"]), gam, **extra)
    loss.mul(_t(g["weights"])).sum().backward()
    _close(loss, g["loss"]); _close(online.grad, g["grad"])


def test_iqn_fqf_losses(golden):
    g = golden("loss_iqn_double")
    gam = float(g["discount"]) ** int(g["n_step"])
    o = dict(online=_t(g["q_cur"]), tgt_next=_t(g["q_next"]), qsel=_t(g["qval_next"]), taus=_t(g["taus_cur"]))
    _close(CP.loss_iqn(o, _t(g["actions"]), _t(g["rewards"]), _t(g["terminals"]), gam), g["loss"])
    g = golden("loss_fqf_double")
    o = dict(online=_t(g["q_hat"]), tgt_next=_t(g["q_next"]), qsel=_t(g["qval_next"]), taus_hat=_t(g["taus_hat"]),
             taus=_t(g["taus"]), q_bar=_t(g["q_bar"]))
    loss, frac = CP.loss_fqf(o, _t(g["actions"]), _t(g["rewards"]), _t(g["terminals"]), gam)
    _close(loss, g["loss"]); _close(frac, g["fraction_loss"])


def test_lz4_and_getitem(golden):
    g = golden("replay_n3")
    z = CP.lz4()
    blob = z.compress(g["entry_frames"][5].tobytes())
    assert len(blob) < 8 * 84 * 84 // 4
    assert z.decompress(blob) == g["entry_frames"][5].tobytes()
    rp = CP.CpuReplay(256, True)
    s = dict(obs=g["stream_obs"], action=g["stream_action"], reward=g["stream_reward"],
             done=(g["stream_terminal"] | g["stream_life_loss"]) & ~g["stream_truncated"])
    assert CP.fill_replay(rp, s, 3) == len(g["entry_action"])
    fr, a, r, d, p, i = rp[7]
    assert np.array_equal(fr, g["entry_frames"][7]) and a == g["entry_action"][7] and r == g["entry_reward"][7]


def test_trainer_step_runs_single_thread(golden):
    g = golden("replay_n3")
    rp = CP.CpuReplay(256, True)
    s = dict(obs=g["stream_obs"], action=g["stream_action"], reward=g["stream_reward"],
             done=(g["stream_terminal"] | g["stream_life_loss"]) & ~g["stream_truncated"])
    CP.fill_replay(rp, s, 3)
    fetch, _ = CP.make_fetcher(rp, 8, 0)
    atoms = torch.linspace(-10, 10, 51)
    outs = lambda it: dict(online=torch.randn(8, 4, 51), tgt_next=torch.randn(8, 4, 51), qsel=torch.randn(8, 4))
    n = CP.trainer_step(rp, fetch, outs, "c51", 3, 0.99 ** 3, extra=dict(atoms=atoms))
    assert n == 24 and rp.max_p > 1.0
