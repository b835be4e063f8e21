"""agent0_b200.hotloop.ReplayTargetLoop (the pre-bound, graph-capturable inner loop of Trainer.step,
agent0/deepq/trainer.py:82-104) against the general-purpose API on the same draws: sampled indices,
IS weights, gathered stacks, losses, gradients and the sum-tree after the priority write-back must be
identical; the captured CUDA graph must replay to the same results as the eager launches."""
import numpy as np
import pytest
import torch

from agent0_b200.config import make_config
from agent0_b200.synth import fill_shard_synthetic

pytestmark = pytest.mark.gpu
B, L, A = 16, 3, 4


def _outputs(algo, T, seed=7):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g) * 3.0
    o = {"qsel": rn(T, A)}
    if algo in ("dqn", "mdqn"):
        o.update(online=rn(T, A), tgt_next=rn(T, A), tgt_cur=rn(T, A))
    elif algo == "c51":
        o.update(online=rn(T, A, 51), tgt_next=rn(T, A, 51), atoms=torch.linspace(-10, 10, 51, device="cuda"))
    elif algo == "qr":
        o.update(online=rn(T, A, 200), tgt_next=rn(T, A, 200))
    elif algo == "iqn":
        o.update(online=rn(T, 64, A), tgt_next=rn(T, 64, A), taus=torch.rand(T, 64, device="cuda", generator=g))
    else:
        p = torch.softmax(torch.randn(T, 32, device="cuda", generator=g), -1)
        taus = torch.cat((torch.zeros(T, 1, device="cuda"), torch.cumsum(p, -1)), -1).contiguous()
        o.update(online=rn(T, 32, A), tgt_next=rn(T, 32, A), q_bar=rn(T, 31, A), taus=taus,
                 taus_hat=((taus[:, :-1] + taus[:, 1:]) / 2).contiguous())
    return o


def _shard(algo, seed=3):
    from agent0_b200.replay import ReplayDataset
    cfg = make_config(algo, per=True, n_step=3, batch_size=B, replay_size=4096, double_q=True, dueling=True, num_envs=8,
                      action_dim=A)
    rp = ReplayDataset(cfg, native_nstep=True)
    fill_shard_synthetic(rp, 4096, 8, seed)
    # uneven priorities so that the draw is not uniform
    ids = torch.arange(0, 4000, 7, device="cuda")
    rp.update_priority(ids, torch.rand(len(ids), device="cuda", generator=torch.Generator("cuda").manual_seed(seed)) * 4)
    return rp


def _general_api(rp, algo, o, u):
    from agent0_b200 import losses as LS
    from agent0_b200.replay import split_batches
    b = rp.sample(B, k_batches=L, u=u)
    gam = float(np.float32(0.99 ** 3))
    outs, grads = [], []
    for k, bk in enumerate(split_batches(b, B)):
        s = slice(k * B, (k + 1) * B)
        cm = (bk.actions, bk.rewards_f32, bk.terminals_f32, bk.weights, gam)
        kw = dict(max_p=rp.max_p_tensor)
        if algo == "dqn":
            r = LS.dqn_loss(o["online"][s], o["tgt_next"][s], *cm, qsel=o["qsel"][s], **kw)
        elif algo == "mdqn":
            r = LS.mdqn_loss(o["online"][s], o["tgt_next"][s], o["tgt_cur"][s], *cm, **kw)
        elif algo == "c51":
            r = LS.c51_loss(o["online"][s], o["tgt_next"][s], o["atoms"], *cm, -10.0, 10.0, qsel=o["qsel"][s], **kw)
        elif algo == "qr":
            r = LS.qr_loss(o["online"][s], o["tgt_next"][s], *cm, qsel=o["qsel"][s], **kw)
        elif algo == "iqn":
            r = LS.iqn_loss(o["online"][s], o["taus"][s], o["tgt_next"][s], o["qsel"][s], *cm, **kw)
        else:
            r = LS.fqf_loss(o["online"][s], o["taus"][s], o["taus_hat"][s], o["tgt_next"][s], o["q_bar"][s], o["qsel"][s], *cm, **kw)
        outs.append(r.loss); grads.append(r.grad)
    loss = torch.cat(outs)
    rp.update_priority(b.indices, loss)
    return b, loss, torch.cat(grads)


@pytest.mark.parametrize("algo", ["dqn", "mdqn", "c51", "qr", "iqn", "fqf"])
def test_prebound_loop_equals_general_api_and_graph_replays_identically(algo):
    from agent0_b200.hotloop import ReplayTargetLoop
    T = B * L
    o = _outputs(algo, T)
    u = torch.rand(T, device="cuda", generator=torch.Generator("cuda").manual_seed(11))
    rp_a, rp_b, rp_c = _shard(algo), _shard(algo), _shard(algo)
    assert torch.equal(rp_a.tree, rp_b.tree)
    batch, loss_ref, grad_ref = _general_api(rp_a, algo, o, u)

    def run(rp, graph):
        loop = ReplayTargetLoop(rp, algo, B, L, A, o, n_step=3, double_q=(algo != "mdqn"))
        rp.push_dynamic()
        if graph:
            loop.idx.zero_()
            loop.gather()                       # first-use function attributes are set outside the capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            loop.u.copy_(u)
            # warm-up launches must not touch the tree: capture only records
            with torch.cuda.graph(g):
                loop.sample(); loop.gather()
                for k in range(L):
                    loop.target_loss(k)
                loop.update()
            loop.u.copy_(u)
            g.replay()
        else:
            loop.u.copy_(u)
            loop.sample(); loop.gather()
            for k in range(L):
                loop.target_loss(k)
            loop.update()
        torch.cuda.synchronize()
        return loop

    for rp, graph in ((rp_b, False), (rp_c, True)):
        loop = run(rp, graph)
        assert torch.equal(loop.idx, batch.indices) and torch.equal(loop.w, batch.weights)
        assert torch.equal(loop.frames, batch.frames) and torch.equal(loop.r64, batch.rewards)
        assert torch.equal(loop.act, batch.actions) and torch.equal(loop.d8.bool(), batch.terminals)
        assert torch.equal(loop.loss, loss_ref) and torch.equal(loop.grad, grad_ref)
        assert torch.equal(rp.tree, rp_a.tree) and float(rp.max_p_tensor) == float(rp_a.max_p_tensor)
        assert loop.launches_per_step == 1 + (len(loop.waves) if loop.waves else 1) + L + 1


def test_prebound_loop_step_capture_and_single_launch_k4():
    """step()/capture()/run() on a growing shard, and the one-launch-for-all-batches K4 variant."""
    from agent0_b200.hotloop import ReplayTargetLoop
    T = B * L
    o = _outputs("c51", T)
    rp = _shard("c51")
    loop = ReplayTargetLoop(rp, "c51", B, L, A, o, n_step=3)
    loop.capture()
    root0 = float(rp.tree[1])
    for _ in range(3):
        loop.run()
    torch.cuda.synchronize()
    assert torch.isfinite(loop.loss).all() and int(loop.idx.max()) < rp.size and float(rp.tree[1]) != root0
    tree = rp.tree.cpu().numpy()
    for node in (1, 2, 3, rp.P // 2, rp.P - 1):
        assert tree[node] == np.float32(tree[2 * node] + tree[2 * node + 1])
    # all batches in one K4 launch == one launch per batch
    loop.u.uniform_(); loop.sample(); loop.gather()
    for k in range(L):
        loop.target_loss(k)
    per_batch = (loop.loss.clone(), loop.grad.clone())
    loop.target_loss_all()
    assert torch.equal(loop.loss, per_batch[0]) and torch.equal(loop.grad, per_batch[1])


@pytest.mark.parametrize("count,bulk_min,chunks", [(48, 2048, 1), (640, 2048, 1), (5000, 2048, 1), (48, 2048, 0), (640, 2048, 0),
                                                   (5000, 1, 0), (5000, 1 << 30, 0), (20000, 2048, 1)])
def test_update_report_hands_the_result_to_mapped_host_memory(count, bulk_min, chunks):
    """a0_pt_update_report = a0_pt_update + the step's (indices, losses) stored by the same kernel into
    page-locked host memory: same tree and max_p as the plain update under every K2b schedule
    (one CTA, cluster climb, cluster write + rebuild, one-CTA bulk write), host buffers equal to the
    device inputs once the stream has passed -- duplicates and out-of-range indices included."""
    from agent0_b200 import _lib
    from agent0_b200.replay import ReplayDataset
    lib = _lib.load()
    N = 30000
    cfg = make_config("dqn", per=True, n_step=1, batch_size=8, replay_size=N, num_envs=4)
    rng = np.random.RandomState(count)
    ids = rng.randint(0, N, count).astype(np.int64)
    ids[-3:] = ids[0]
    ids[1] = N + 5                                   # skipped by the update, still reported as given
    loss = np.abs(rng.randn(count)).astype(np.float32) * 3
    trees = []
    try:
        assert lib.a0_set_option(3, bulk_min) == 0 and lib.a0_set_option(8, chunks) == 0
        for report in (False, True):
            rp = ReplayDataset(cfg, native_nstep=True)
            rp.set_priorities(torch.arange(N), torch.as_tensor((np.arange(N) % 13 + 1).astype(np.float32)))
            d_ids, d_loss = torch.as_tensor(ids).cuda(), torch.as_tensor(loss).cuda()
            if report:
                h_ids = torch.full((count,), -7, dtype=torch.int64).pin_memory()
                h_loss = torch.full((count,), -7.0, dtype=torch.float32).pin_memory()
                _lib.check(lib.a0_pt_update_report(rp.h, d_ids.data_ptr(), d_loss.data_ptr(), count, rp.alpha, rp.eps,
                                                   _lib.host_map(h_ids), _lib.host_map(h_loss), _lib.stream_ptr(rp.device)),
                           "a0_pt_update_report")
                ev = torch.cuda.Event()
                ev.record()
                ev.synchronize()
                assert np.array_equal(h_ids.numpy(), ids) and np.array_equal(h_loss.numpy(), loss)
            else:
                rp.update_priority(d_ids, d_loss)
            torch.cuda.synchronize()
            trees.append((rp.tree.clone(), float(rp.max_p_tensor)))
    finally:
        lib.a0_set_option(3, 2048)
        lib.a0_set_option(8, 1)
    assert torch.equal(trees[0][0], trees[1][0]) and trees[0][1] == trees[1][1]
    with pytest.raises(ValueError):
        _lib.host_map(torch.empty(4))                # a pageable tensor
    import ctypes as C
    pageable, out = np.zeros(16, dtype=np.float32), C.c_void_p()
    assert lib.a0_host_map(pageable.ctypes.data, C.byref(out)) != 0 and b"page" in lib.a0_last_error()


def test_reporting_graphs_with_ingest_published_scalars_equal_the_plain_loop():
    """The end-to-end form of the loop (bench.py's `e2e`): the ingest publishes top/beta from its own
    launch, one captured graph per report slot stores losses + indices straight into pinned host
    memory.  Against the plain loop (push_dynamic launch, results read from the device) on a twin
    shard with the same sampler seed: same draws, losses, trees, step after step, and the host
    buffers hold exactly what the device computed."""
    from agent0_b200.hotloop import ReplayTargetLoop
    T = B * L
    o = _outputs("c51", T)
    rp_a, rp_b = _shard("c51"), _shard("c51")
    la = ReplayTargetLoop(rp_a, "c51", B, L, A, o, n_step=3, rng_seed=5)
    lb = ReplayTargetLoop(rp_b, "c51", B, L, A, o, n_step=3, rng_seed=5)
    la.capture()
    rep = lb.bind_report(2)
    lb.capture_reporting()
    rp_a.rng_seek(0); rp_b.rng_seek(0)
    E = 8
    g = torch.Generator().manual_seed(2)
    for s in range(5):
        new = torch.randint(0, 256, (E, 84 * 84), dtype=torch.uint8, generator=g).pin_memory()
        args = (np.arange(E), np.ones(E, dtype=np.int64), new, np.arange(E) % A, np.ones(E) * s, np.zeros(E, dtype=bool))
        rp_a.append_steps(*args)
        la.run()
        rp_b.append_steps(*args, pinned_stable=True, copy_stream=True, publish_dynamic=True)
        lb.run(slot=s % 2, publish=False)
        ev = torch.cuda.Event()
        ev.record()
        ev.synchronize()
        assert torch.equal(rep[s % 2][0], lb.idx.cpu()) and torch.equal(rep[s % 2][1], lb.loss.cpu())
        torch.cuda.synchronize()
        assert torch.equal(la.idx, lb.idx) and torch.equal(la.loss, lb.loss) and torch.equal(la.w, lb.w)
        assert torch.equal(rp_a.tree, rp_b.tree)
    assert len({int(rep[0][0][0]), int(rep[1][0][0])}) == 2 or not torch.equal(rep[0][0], rep[1][0])


@pytest.mark.parametrize("Bn,k,per,waves,pdl,window,prio", [
    (32, 20, True, "auto", False, "auto", True), (32, 20, True, [1, 1, 2, 16], True, 8, True), (512, 4, True, "auto", False, "auto", True),
    (12, 5, True, [2, 3], False, 1, False), (32, 6, False, "auto", False, 40, True), (16, 3, True, [1, 2], True, 0, False),
    (64, 20, True, [1] * 20, False, 700, True), (512, 8, True, "auto", True, "auto", True)])
def test_gather_waves_equal_the_single_gather_launch(Bn, k, per, waves, pdl, window, prio):
    """gather_waves: K2a and the gather cut into waves run on a side stream (a0_rb_sample_mail + a0_rb_gather_mail per wave),
    batch k's K4 waits only for the wave that holds it.  Against the same loop with ONE gather launch on a twin shard with
    the same sampler seed: identical draws, stacks, n-step scalars, losses, gradients and tree -- eagerly, step after step,
    and as a captured graph (the side-stream branch becomes part of the graph).  window: the ordered fetch inside the waves
    (at most that many draws beyond the completed ones in flight; 1 = strictly one after the other); prio: K4 + K2b on the
    loop's high-priority stream."""
    from agent0_b200.hotloop import ReplayTargetLoop
    T = Bn * k
    o = _outputs("c51", T)
    rp_a, rp_b = _shard("c51"), _shard("c51")
    la = ReplayTargetLoop(rp_a, "c51", Bn, k, A, o, n_step=3, per=per, rng_seed=9, gather_waves=None)
    lb = ReplayTargetLoop(rp_b, "c51", Bn, k, A, o, n_step=3, per=per, rng_seed=9, gather_waves=waves, pdl_at_joins=pdl,
                          gather_window=window, k4_priority=prio, waves_when_eager=True)
    if window == "auto":
        assert lb.window == (400 if Bn >= 256 else 128)
    assert la.waves is None and lb.waves is not None and sum(w[1] for w in lb.waves) == k and sum(w[3] for w in lb.waves) == T
    if waves == "auto":
        assert [w[1] for w in lb.waves] == ([1] * k if Bn >= 256 else {20: [1, 1, 2, 4, 12], 6: [1, 1, 2, 2]}[k])
    assert lb.launches_per_step == 1 + len(lb.waves) + k + (1 if per else 0)
    fields = ("idx", "prio", "w", "frames", "act", "r64", "r32", "d8", "d32", "boot", "loss", "grad")

    def same():
        torch.cuda.synchronize()
        for f in fields:
            assert torch.equal(getattr(la, f), getattr(lb, f)), f
        assert torch.equal(rp_a.tree, rp_b.tree) and float(rp_a.max_p_tensor) == float(rp_b.max_p_tensor)

    rp_a.push_dynamic(); rp_b.push_dynamic()
    rp_a.rng_seek(0); rp_b.rng_seek(0)
    for _ in range(3):                                     # eager: the side stream forks from and joins the caller's stream
        lb.frames.zero_(); lb.idx.fill_(-1); lb.loss.fill_(-1.0)
        la.step(); lb.step()
        same()
    lb.step(fused_k4=True); la.step(fused_k4=True)         # one K4 launch for all batches waits for every wave
    same()
    if per:                                                # the result hand-over to mapped host memory
        rep = lb.bind_report(1)
        rep[0][0].fill_(-7); rep[0][1].fill_(-7.0)
        la.step(); lb.step(slot=0)
        same()
        assert torch.equal(rep[0][0], lb.idx.cpu()) and torch.equal(rep[0][1], lb.loss.cpu())
    la.capture(warm=1); lb.capture(warm=1)
    for _ in range(4):
        lb.frames.zero_(); lb.loss.fill_(-1.0)
        la.run(); lb.run()
        same()
    if per:
        w = lb.w.view(k, Bn).max(dim=1)[0]
        assert torch.allclose(w, torch.ones_like(w), atol=1e-6)
    from agent0_b200 import _lib
    mask = __import__("ctypes").c_int64(-1)
    assert _lib.load().a0_get_option(_lib.OPT_PDL, __import__("ctypes").byref(mask)) == 0 and mask.value == 1     # restored after the joins


def test_sampler_units_per_cta_draw_the_same_batches():
    """A0_OPT_K2A_ROUNDS: 20 x 512 draws as 640 CTAs of two 8-draw units instead of 1280 CTAs of one -- the same indices,
    priorities and IS weights, with caller-supplied uniforms and with the sampler's own generator (whose call counter
    advances once per launch either way)."""
    from agent0_b200 import _lib
    from agent0_b200.hotloop import ReplayTargetLoop
    lib = _lib.load()
    rp = _shard("c51")
    Bn, k = 512, 20
    o = _outputs("dqn", Bn * k)
    la = ReplayTargetLoop(rp, "dqn", Bn, k, A, o, n_step=3, gather_waves=None)
    lb = ReplayTargetLoop(rp, "dqn", Bn, k, A, o, n_step=3, gather_waves=None)
    rp.push_dynamic()
    la.u.uniform_(generator=torch.Generator("cuda").manual_seed(3)); lb.u.copy_(la.u)
    try:
        la.sample()
        assert lib.a0_set_option(13, 1) == 0
        lb.sample()
        torch.cuda.synchronize()
        for f in ("idx", "prio", "w"):
            assert torch.equal(getattr(la, f), getattr(lb, f)), f
        la.rng_seed = lb.rng_seed = 77
        for call in (5, 6):
            lib.a0_set_option(13, 0)
            rp.rng_seek(call); la.sample()
            lib.a0_set_option(13, 1)
            rp.rng_seek(call); lb.idx.fill_(-1); lb.sample_gather()
            lb.sample()                                    # the counter advanced by exactly one per launch: this is call + 1
            torch.cuda.synchronize()
            lib.a0_set_option(13, 0)
            rp.rng_seek(call + 1); la.sample()
            torch.cuda.synchronize()
            for f in ("idx", "prio", "w"):
                assert torch.equal(getattr(la, f), getattr(lb, f)), f
    finally:
        lib.a0_set_option(13, 0)


def test_gather_mail_argument_errors():
    from agent0_b200 import _lib
    lib = _lib.load()
    rp = _shard("c51")
    out = torch.empty(64, 8 * rp.F, dtype=torch.uint8, device="cuda")
    st = _lib.stream_ptr(rp.device)
    idx = torch.empty(64, dtype=torch.int64, device="cuda"); pr = torch.empty(64, device="cuda")
    assert lib.a0_rb_sample_mail(None, None, 1, 0, 64, 32, 4096.0, 0.4, 0.0, 0, idx.data_ptr(), pr.data_ptr(), None, st) == -1
    assert lib.a0_rb_sample_mail(rp.h, None, 1, 0, 60, 32, 4096.0, 0.4, 0.0, 0, idx.data_ptr(), pr.data_ptr(), None, st) == -1
    assert lib.a0_rb_sample_mail(rp.h, None, 1, 0, 64, 32, 4096.0, 0.4, 0.0, 0, None, pr.data_ptr(), None, st) == -1
    assert lib.a0_rb_gather_mail(rp.h, 0, 32, 0, 3, 0.99, None, None, None, None, None, None, None, st) == -1
    assert lib.a0_rb_gather_mail(rp.h, -1, 32, 0, 3, 0.99, out.data_ptr(), None, None, None, None, None, None, st) == -1
    assert lib.a0_rb_gather_mail(rp.h, 0, 32, 0, 17, 0.99, out.data_ptr(), None, None, None, None, None, None, st) == -1
    # a wave beyond the mailbox of the last a0_rb_sample_mail (none yet on this handle, or a shorter draw)
    assert lib.a0_rb_gather_mail(rp.h, 1 << 20, 32, 0, 3, 0.99, out.data_ptr(), None, None, None, None, None, None, st) == -1
    assert b"mailbox" in lib.a0_last_error()
    assert lib.a0_rb_gather_mail(rp.h, 0, 32, -1, 3, 0.99, out.data_ptr(), None, None, None, None, None, None, st) == -1
    assert lib.a0_rb_gather_mail(rp.h, 0, 0, 0, 3, 0.99, out.data_ptr(), None, None, None, None, None, None, st) == 0
    v = __import__("ctypes").c_int64(0)
    assert lib.a0_get_option(99, __import__("ctypes").byref(v)) == -1 and lib.a0_get_option(1, None) == -1
    torch.cuda.synchronize()


@pytest.mark.parametrize("Bn,k,per", [(32, 20, True), (512, 4, True), (12, 3, True), (32, 2, False)])
def test_overlapped_sample_gather_equals_the_two_separate_launches(Bn, k, per):
    """a0_rb_sample_gather: the gather runs UNDER the sampler (programmatic launch, record positions
    through a mailbox, griddepcontrol.wait at its end).  Indices, priorities, IS weights, stacks and
    n-step scalars must be those of a0_pt_sample[_rng] followed by a0_rb_gather -- with caller-supplied
    uniforms and with the sampler's own generator, on repeated calls (the mailbox re-arms itself), and
    when captured into a CUDA graph and replayed."""
    from agent0_b200 import _lib
    from agent0_b200.hotloop import ReplayTargetLoop
    lib = _lib.load()
    rp = _shard("c51")
    T = Bn * k
    o = _outputs("dqn", T)
    la = ReplayTargetLoop(rp, "dqn", Bn, k, A, o, n_step=3, per=per, overlap_sample_gather=False)
    lb = ReplayTargetLoop(rp, "dqn", Bn, k, A, o, n_step=3, per=per, overlap_sample_gather=True)
    rp.push_dynamic()
    fields = ("idx", "prio", "w", "frames", "act", "r64", "r32", "d8", "d32", "boot")

    def same():
        torch.cuda.synchronize()
        for f in fields:
            assert torch.equal(getattr(la, f), getattr(lb, f)), f

    g = torch.Generator("cuda").manual_seed(5)
    for rep in range(3):                                   # caller-supplied uniforms
        la.u.uniform_(generator=g)
        lb.u.copy_(la.u)
        la.sample(); la.gather()
        lb.frames.zero_(); lb.idx.fill_(-1)
        lb.sample_gather()
        same()
    la.rng_seed = lb.rng_seed = 1234                       # the sampler's own Philox stream, device call counter
    rp.rng_seek(40)
    la.sample(); la.gather()
    rp.rng_seek(40)
    lb.sample_gather()
    same()
    # captured: every replay advances the counter and refills the mailbox
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        lb.sample_gather()
    for call in (50, 51, 52):
        rp.rng_seek(call)
        la.sample(); la.gather()
        rp.rng_seek(call)                                  # la's launch advanced the counter: step back (stream-ordered)
        gr.replay()
        same()
    # work queued behind the pair sees the sampler complete: weights are normalised (max == 1 per batch)
    if per:
        w = lb.w.view(k, Bn).max(dim=1)[0]
        assert torch.allclose(w, torch.ones_like(w), atol=1e-6)


def test_uniform_replay_stays_uniform_under_the_loop():
    """Uniform replay (BASELINE configs[0]: DQN, n_step=1): the loop's K4 must not raise the shard's running
    max loss, and appended records must get leaf 1.0 whatever max_p holds -- otherwise the tree descent
    draws newer records more often than older ones.  Leaves all equal 1 after steps and appends; the
    draw's empirical law is flat; a NaN / inf loss leaves a prioritized tree finite."""
    from agent0_b200.hotloop import ReplayTargetLoop
    from agent0_b200.replay import ReplayDataset
    cfg = make_config("dqn", per=False, n_step=1, batch_size=B, replay_size=2048, double_q=False, dueling=False, num_envs=8,
                      action_dim=A)
    rp = ReplayDataset(cfg, native_nstep=True)
    fill_shard_synthetic(rp, 1024, 8, 5)
    o = _outputs("dqn", B * L)
    o["online"] *= 50                                        # losses far above 1: what used to inflate max_p
    loop = ReplayTargetLoop(rp, "dqn", B, L, A, o, n_step=1, double_q=False, rng_seed=11)
    rp.push_dynamic()
    for _ in range(4):
        loop.step()
    assert float(loop.loss.max()) > 1.0
    assert rp.max_p == 1.0
    rp.max_p_tensor.fill_(37.0)                              # even a poisoned max_p must not show in new leaves
    streams = np.arange(64, dtype=np.int64) % 8
    frames = torch.randint(0, 256, (64, rp.F), dtype=torch.uint8, device="cuda")
    rp.append_steps(streams, np.ones(64, dtype=np.int64), frames, np.zeros(64, dtype=np.int64), np.zeros(64), np.zeros(64, dtype=bool))
    leaves = rp.priority.leaves()[:rp.size].cpu().numpy()
    live = rp.index.sampleable
    assert (leaves[live] == 1.0).all() and (leaves[~live] == 0.0).all()
    rp.push_dynamic()
    counts = torch.zeros(rp.size, device="cuda")
    for _ in range(200):
        loop.step()
        counts.index_add_(0, loop.idx, torch.ones_like(loop.idx, dtype=torch.float32))
    c = counts.cpu().numpy()[live]
    n_draws, n_live = 200 * B * L, int(live.sum())
    lam = n_draws / n_live
    assert abs(c[:n_live // 2].sum() / n_draws - 0.5) < 0.02 and c.max() < lam + 8 * np.sqrt(lam) + 8
    # a prioritized shard survives a NaN and an infinite loss: the old leaves stay, the root stays finite
    rp2 = _shard("dqn")
    before = rp2.priority.leaves().clone()
    ids = torch.tensor([5, 6, 7, 8], device="cuda")
    rp2.update_priority(ids, torch.tensor([float("nan"), float("inf"), -1.0, 2.0], device="cuda"))
    after = rp2.priority.leaves()
    assert torch.equal(after[:8][5:8], before[:8][5:8]) and abs(float(after[8]) - (2.0 + 0.01) ** 0.5) < 1e-6
    assert np.isfinite(float(rp2.priority_sum())) and np.isfinite(rp2.max_p)


def test_stale_by_one_schedule_keeps_the_tree_consistent_and_the_default_unchanged():
    """hotloop.StaleByOneLoop (opt-in): the write-back of step s runs on a side stream under the draw + gather of
    step s+1.  Every drawn index is a live record, the tree is node == fl32(left + right) after every replay, the
    last step's leaves hold its priorities after flush(), and max_p follows the largest loss seen.  The default
    loop on a twin shard (same fill, same seeds) is untouched by the existence of the schedule: its results equal
    a second default run bit for bit."""
    from agent0_b200.hotloop import ReplayTargetLoop, StaleByOneLoop
    T = B * L
    o = _outputs("c51", T)
    rp = _shard("c51")
    loop = StaleByOneLoop(rp, "c51", B, L, A, o, n_step=3, rng_seed=5)
    loop.capture()
    P = rp.P
    live = torch.as_tensor(np.asarray(rp.index.sampleable), device="cuda")
    seen_max = 1.0
    for it in range(12):
        loop.run()
        torch.cuda.synchronize()
        cur = loop.current()
        assert bool(live[cur.idx].all()) and bool(torch.isfinite(cur.loss).all()) and bool((cur.w > 0).all())
        t = rp.tree
        assert torch.equal(t[1:P], t[2:2 * P:2] + t[3:2 * P:2])
        seen_max = max(seen_max, float(cur.loss.max()))
    last = loop.current()
    before_flush = rp.priority.leaves()[last.idx].clone()
    loop.flush()
    torch.cuda.synchronize()
    idx, loss = last.idx.cpu().numpy(), last.loss.cpu().numpy()
    want = {}
    for i, l in zip(idx, loss):                      # duplicates: the last writer wins
        want[int(i)] = np.sqrt(np.float32(l) + np.float32(0.01))
    leaves = rp.priority.leaves().cpu().numpy()
    np.testing.assert_allclose([leaves[i] for i in want], list(want.values()), rtol=1e-6)
    assert not torch.equal(before_flush, rp.priority.leaves()[last.idx])          # the write-back really was pending
    assert rp.max_p >= seen_max * (1 - 1e-6)          # K4 raises it as it goes (the captured warm-up steps count too)
    t = rp.tree
    assert torch.equal(t[1:P], t[2:2 * P:2] + t[3:2 * P:2])
    # the default schedule: two identical runs on twin shards
    res = []
    for _ in range(2):
        rq = _shard("c51")
        lp = ReplayTargetLoop(rq, "c51", B, L, A, o, n_step=3, rng_seed=5)
        rq.rng_seek(0)
        lp.capture()
        rq.rng_seek(0)
        for _ in range(3):
            lp.run()
        torch.cuda.synchronize()
        res.append((lp.idx.clone(), lp.loss.clone(), rq.tree.clone()))
    assert all(torch.equal(a_, b_) for a_, b_ in zip(res[0], res[1]))
