"""agent0_b200.model vs the reference's agent0.deepq.model on CPU: same seed -> bit-identical
state_dict (names, shapes, values), same RNG consumption, same outputs.  Needs /root/reference
(build container only); skipped where it does not exist (the GPU box)."""
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


@pytest.mark.parametrize("algo", ["dqn", "c51", "qr", "iqn", "fqf", "mdqn"])
@pytest.mark.parametrize("dueling,noisy", [(False, False), (True, False), (True, True)])
def test_same_seed_same_network(algo, dueling, noisy):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from agent0.deepq.config import AlgoEnum, ExpConfig
    from agent0.deepq.model import DeepQNet as RefNet
    from agent0_b200.config import make_config
    from agent0_b200.model import DeepQNet
    cfg = make_config(algo, dueling=dueling, noisy_net=noisy, action_dim=6)
    rc = ExpConfig()
    rc.obs_shape, rc.action_dim = (4, 84, 84), 6
    rc.learner.algo, rc.learner.dueling_head, rc.learner.noisy_net = AlgoEnum[algo], dueling, noisy
    torch.manual_seed(3); ours = DeepQNet(cfg); r_ours = torch.rand(2)
    torch.manual_seed(3); ref = RefNet(rc); r_ref = torch.rand(2)
    sa, sb = ours.state_dict(), ref.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]), k
    assert torch.equal(r_ours, r_ref)                       # constructors consumed the RNG identically
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
    assert sum(p.numel() for p in ours.params()) == sum(p.numel() for p in ref.params())
    x = torch.rand(3, 4, 84, 84)
    torch.manual_seed(1); qa = ours.qval(x)
    torch.manual_seed(1); qb = ref.qval(x)
    assert torch.equal(qa, qb)
    if algo in ("iqn", "fqf"):
        enc = ours.encoder(x)
        torch.manual_seed(2); ya, ta = ours.head(enc, n=8)
        torch.manual_seed(2); yb, tb = ref.head(enc, n=8)
        assert torch.equal(ya, yb) and torch.equal(ta, tb)
    else:
        assert torch.equal(ours(x), ref(x))
    if algo == "fqf":
        for a, b in zip(ours.head.prop_taus(enc), ref.head.prop_taus(enc)):
            assert torch.equal(a, b)
    # the reference's actors can load our weights and vice versa (launch.py:33-36)
    ref.load_state_dict(ours.state_dict())
    ours.load_state_dict(ref.state_dict())
