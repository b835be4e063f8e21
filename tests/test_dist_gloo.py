"""World-size-2 gloo tests (CPU) of the N>1 host logic (agent0_b200/dist.py): stream sharding with
no data-path collective, the flat gradient bucket's SUM all-reduce with Adam eps = 1e-2/(G*B)
(so that G ranks at batch B equal one learner at batch G*B; the reference's loss is SUM-reduced,
agent0/deepq/agent.py:102-106,154), and the global priority statistics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from agent0_b200.dist import (FlatGradBucket, OverlappedGradBucket, adam_eps, global_batch_is_weights,
                              global_priority_stats, local_streams, shard_capacity, shard_of_stream)

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _net(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(12, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4))


def _batch(seed, B):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 12, generator=g)
    a = torch.randint(0, 4, (B,), generator=g)
    tgt = torch.randn(B, generator=g) * 2
    w = torch.rand(B, generator=g) + 0.1
    return x, a, tgt, w


def _update(net, bucket, opt, x, a, tgt, w):
    """One learner update with the K4 contract: per-sample Huber loss, SUM of loss*weight."""
    bucket.zero_()
    q = net(x)[torch.arange(len(a)), a]
    loss = torch.nn.functional.smooth_l1_loss(q, tgt, reduction="none")
    (loss * w).sum().backward()
    bucket.all_reduce()
    opt.step()


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        torch.set_num_threads(1)
        B = 8
        net = _net(7)
        bucket = FlatGradBucket(list(net.parameters()), dist.group.WORLD)
        assert bucket.world == WORLD
        opt = torch.optim.Adam(net.parameters(), 1e-3, eps=adam_eps(B, WORLD))
        for it in range(3):
            _update(net, bucket, opt, *_batch(100 * it + rank, B))
        for p in net.parameters():                       # gradients are still views of the flat buffer
            assert p.grad.data_ptr() >= bucket.flat.data_ptr()
        flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        # the same three updates with the gradient exchange issued from backward hooks, bucket by bucket
        net2 = _net(7)
        ob = OverlappedGradBucket(list(net2.parameters()), dist.group.WORLD, n_buckets=2, min_bucket_elems=64, split_frac=0.8)
        assert len(ob.buckets) == 2 and ob.buckets[0][3] == ob.flat.numel() and ob.buckets[-1][2] == 0
        opt2 = torch.optim.Adam(net2.parameters(), 1e-3, eps=adam_eps(B, WORLD))
        for it in range(3):
            _update(net2, ob, opt2, *_batch(100 * it + rank, B))
            assert not ob._handles and all(n == 0 for n in ob._left)
        flat2 = torch.cat([p.detach().reshape(-1) for p in net2.parameters()])
        ob.enabled = False                                  # the switch bench.py uses to time the step without the exchange
        g0 = ob.flat.clone()
        _update(net2, ob, opt2, *_batch(999 + rank, B))
        local_only = ob.flat.clone()
        # IS weights normalised over the global batch: one MAX all-reduce of L floats
        gen = torch.Generator().manual_seed(40 + rank)
        prio = torch.rand(3 * B, generator=gen) + 0.05
        s_r = torch.tensor(10.0 + 5 * rank)
        w_glob = global_batch_is_weights(prio, s_r, 0.6, B, dist.group.WORLD)
        w_loc = global_batch_is_weights(prio, s_r, 0.6, B, None)
        # sharding: every stream is owned by exactly one rank; the shards' sampleable counts add up
        from agent0_b200.ring_index import NativeRingIndex
        E, T, n = 6, 40, 3
        mine = local_streams(rank, WORLD, E)
        assert all(shard_of_stream(s, WORLD) == rank for s in mine)
        cap = shard_capacity(4096, WORLD)
        ix = NativeRingIndex(cap, 4 * cap + 64, n)
        for i, s in enumerate(mine):
            ix.set_stack(s, ix.head_fs + 4 * i + np.arange(4))
        ix.plan(np.zeros(0, np.int64), np.zeros((0, 8), np.int64), np.arange(4 * len(mine)), [], [], [])
        for k in range(T):
            st = np.array(mine, dtype=np.int64)
            fs8 = ix.resolve_shift(st, np.ones(len(mine), np.int64))
            ix.plan(st, fs8, np.arange(len(mine)), np.zeros(len(mine)), np.zeros(len(mine)), np.zeros(len(mine), bool))
        tops = torch.tensor([float(ix.top)])
        dist.all_reduce(tops)
        s_g, top_g = global_priority_stats(torch.tensor(1.5 + rank), ix.top, dist.group.WORLD)
        # weight hand-off to actor ranks (launch.py:33-36): one flat broadcast instead of a pickled state_dict
        from agent0_b200.actor import broadcast_model, flat_state
        from agent0_b200.config import make_config
        from agent0_b200.model import DeepQNet
        cfg = make_config("c51", dueling=True, action_dim=4)
        torch.manual_seed(50 + rank)                      # every rank starts from different weights
        model = DeepQNet(cfg)
        model.register_buffer("steps_seen", torch.tensor([3 + rank]))     # an integer buffer rides along
        before = torch.cat([t.reshape(-1).float() for t in flat_state(model)]).clone()
        sent = broadcast_model(model, src=0, process_group=dist.group.WORLD)
        after = torch.cat([t.reshape(-1).float() for t in flat_state(model)])
        out[rank] = dict(model_before=before.numpy(), model_after=after.numpy(), model_bytes=sent,
                         params=flat.numpy(), params_overlapped=flat2.numpy(), grad_local_only=local_only.numpy(),
                         w_glob=w_glob.numpy(), w_loc=w_loc.numpy(), prio=prio.numpy(), s_r=float(s_r),
                         top_sum=float(tops.item()), local_top=ix.top,
                         stats=(float(s_g), float(top_g)), streams=mine)
    finally:
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def gloo_run():
    mp.set_sharing_strategy("file_system")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(_free_port(), out), nprocs=WORLD, join=True)
    return dict(out)


def test_two_ranks_at_batch_b_equal_one_learner_at_batch_2b(gloo_run):
    B = 8
    net = _net(7)
    bucket = FlatGradBucket(list(net.parameters()))
    opt = torch.optim.Adam(net.parameters(), 1e-3, eps=adam_eps(2 * B, 1))
    assert adam_eps(2 * B, 1) == adam_eps(B, 2) == 1e-2 / 16
    for it in range(3):
        parts = [_batch(100 * it + r, B) for r in range(WORLD)]
        _update(net, bucket, opt, *[torch.cat(c) for c in zip(*parts)])
    want = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy()
    np.testing.assert_allclose(gloo_run[0]["params"], want, rtol=2e-5, atol=2e-7)
    assert np.array_equal(gloo_run[0]["params"], gloo_run[1]["params"])        # ranks stay in lock-step


def test_streams_shard_without_overlap_and_counts_add_up(gloo_run):
    E, T, n = 6, 40, 3
    assert sorted(gloo_run[0]["streams"] + gloo_run[1]["streams"]) == list(range(E))
    assert gloo_run[0]["top_sum"] == E * (T - n + 1) == gloo_run[0]["local_top"] + gloo_run[1]["local_top"]
    assert shard_capacity(8_000_000, 8) == 1_000_000 and shard_capacity(10, 4) == 3


def test_global_priority_stats(gloo_run):
    s_g, top_g = gloo_run[0]["stats"]
    assert s_g == pytest.approx(1.5 + 2.5) and top_g == gloo_run[0]["top_sum"]
    assert gloo_run[1]["stats"] == gloo_run[0]["stats"]


def test_broadcast_model_replaces_state_dict_shipping(gloo_run):
    r0, r1 = gloo_run[0], gloo_run[1]
    assert not np.array_equal(r0["model_before"], r1["model_before"])
    assert np.array_equal(r0["model_after"], r0["model_before"])               # the source is unchanged
    assert np.array_equal(r1["model_after"], r0["model_before"])               # the actor rank now holds the learner's net
    assert r0["model_bytes"] == r1["model_bytes"] == (len(r0["model_before"]) - 1) * 4 + 8


def test_overlapped_bucket_equals_the_single_all_reduce(gloo_run):
    """Gradients exchanged bucket by bucket from backward hooks give the parameters of the one-call exchange,
    bit for bit; with the exchange switched off the ranks' gradients differ (so the switch really skips it)."""
    for r in range(WORLD):
        assert np.array_equal(gloo_run[r]["params_overlapped"], gloo_run[r]["params"])
    assert not np.array_equal(gloo_run[0]["grad_local_only"], gloo_run[1]["grad_local_only"])


def test_global_batch_is_weights(gloo_run):
    """w = u / (max over BOTH ranks' batches of u + 1e-8), u = (p / S_r)^-beta (dist.global_batch_is_weights)."""
    B, beta = 8, 0.6
    u = [(gloo_run[r]["prio"] / gloo_run[r]["s_r"]) ** (-beta) for r in range(WORLD)]
    mx = np.maximum(u[0].reshape(3, B).max(1), u[1].reshape(3, B).max(1))
    for r in range(WORLD):
        want = (u[r].reshape(3, B) / (mx[:, None] + 1e-8)).reshape(-1)
        np.testing.assert_allclose(gloo_run[r]["w_glob"], want, rtol=1e-6)
        local = (u[r].reshape(3, B) / (u[r].reshape(3, B).max(1)[:, None] + 1e-8)).reshape(-1)
        np.testing.assert_allclose(gloo_run[r]["w_loc"], local, rtol=1e-6)
    both = np.concatenate([gloo_run[r]["w_glob"].reshape(3, B) for r in range(WORLD)], axis=1)
    assert np.allclose(both.max(1), 1.0, atol=1e-6) and (both <= 1.0 + 1e-6).all()
