"""CPU tests of the host index logic (agent0_b200.ring_index) against the oracle's restatement of
the reference's actor packing + deque (golden-pinned in test_oracle_golden.py)."""
import numpy as np
import pytest

from agent0_b200.ring_index import ContentDeduper, NativeContentDeduper, NativeRingIndex, RingIndex, stack_delta
from agent0_b200.synth import record_stream
from oracle import reference_replay as OR
from tests.ring_sim import SimDevice

F = 64   # small frames keep the CPU suite fast; the index logic is frame-size independent

# every test runs against the numpy specification and the C++ index inside libagent0_b200.so
IMPLS = pytest.mark.parametrize("Index", [RingIndex, NativeRingIndex], ids=["spec", "native"])


def _stream(E, T, seed):
    s = record_stream(E, T, seed=seed, frame_hw=(8, 8), p_terminal=0.07, p_life_loss=0.08, p_truncated=0.05)
    return s


def _entries(s, n, discount=0.99):
    return OR.pack_nstep(s["obs"], s["action"], s["reward"], s["done"], n, discount)


def test_stack_delta():
    s = _stream(4, 30, 1)
    for k in range(30):
        d = stack_delta(s["obs"][k], s["obs"][k + 1])
        special = s["terminal"][k] | s["truncated"][k] | s["life_loss"][k]
        assert np.array_equal(d == 4, special) and set(np.unique(d)) <= {1, 4}


@IMPLS
@pytest.mark.parametrize("n", [1, 3])
def test_compat_extend_equals_reference_entries(golden, n, Index):
    """Reference tuples in, gather(i) == reference entry i, bit for bit (full-size golden frames)."""
    g = golden(f"replay_n{n}")
    Fg = 84 * 84
    M = len(g["entry_action"])
    ix = Index(256, 2048, n_step=1)
    dd = ContentDeduper(ix, Fg)
    dev = SimDevice(256, 2048, Fg)
    E = int(g["num_envs"])
    for lo in range(0, M, 30):          # several extend() calls
        hi = min(M, lo + 30)
        frames = np.ascontiguousarray(g["entry_frames"][lo:hi].reshape(hi - lo, 8, Fg))
        streams = np.arange(lo, hi, dtype=np.int64) % E
        fs8, new_src = dd.resolve(streams, frames)
        plan = ix.plan(streams, fs8, new_src, g["entry_action"][lo:hi], g["entry_reward"][lo:hi], g["entry_done"][lo:hi])
        dev.execute(plan, frames.reshape(-1, Fg))
        dd.detach(streams)
    assert ix.top == M
    # de-duplication: about one new frame per entry, not eight
    assert ix.head_fs < 3 * M      # (the golden stream has ~20% stack-replacing events)
    for i in range(M):
        fr, a, r, d, _ = dev.gather(i, 1, 0.99)
        assert np.array_equal(fr, g["entry_frames"][i])
        assert a == g["entry_action"][i] and d == g["entry_done"][i]
        assert np.float64(r).view(np.int64) == g["entry_reward"][i].view(np.int64)


@IMPLS
@pytest.mark.parametrize("n", [1, 3])
def test_native_ingest_equals_reference_entries(golden, n, Index):
    """1-step records in, n-step gather out == the reference actor's folded entries (Appendix C)."""
    g = golden(f"replay_n{n}")
    Fg = 84 * 84
    E, T = int(g["num_envs"]), int(g["steps"])
    obs = g["stream_obs"]
    done = OR.done_rule(g["stream_terminal"], g["stream_life_loss"], g["stream_truncated"])
    ix = Index(256, 2048, n_step=n)
    dev = SimDevice(256, 2048, Fg)
    last4 = {}
    streams = np.arange(E, dtype=np.int64)
    seqs = ix.head_fs + np.arange(4 * E)
    for e in range(E):
        last4[e] = seqs[4 * e:4 * e + 4].copy()
    plan = ix.plan(np.zeros(0, np.int64), np.zeros((0, 8), np.int64), np.arange(4 * E), [], [], [])
    dev.execute(plan, obs[0].reshape(4 * E, Fg))
    for k in range(T):
        kk = stack_delta(obs[k], obs[k + 1])
        new = np.concatenate([obs[k + 1][e, 4 - kk[e]:] for e in range(E)]).reshape(-1, Fg)
        fs8 = ix.resolve_shift(streams, kk, last4)
        plan = ix.plan(streams, fs8, np.arange(len(new)), g["stream_action"][k], g["stream_reward"][k], done[k])
        dev.execute(plan, new)
    assert ix.top == (T - n + 1) * E
    disc = float(g["discount"])
    for k in range(n - 1, T):
        for e in range(E):
            i = k * E + e                       # reference entry index (step-major, env-minor)
            pos = (k - n + 1) * E + e           # our record: stream e, start step k-n+1
            assert dev.leaf[pos] == 1.0
            fr, a, r, d, boot = dev.gather(pos, n, disc)
            assert np.array_equal(fr, g["entry_frames"][i])
            assert a == g["entry_action"][i] and d == g["entry_done"][i]
            assert np.float64(r).view(np.int64) == g["entry_reward"][i].view(np.int64)
            assert boot == ((k + 1) * E + e if k + 1 < T else -1)
    # the newest n-1 records of every stream are not sampleable yet
    assert dev.leaf[(T - n + 1) * E:T * E].sum() == 0


@IMPLS
@pytest.mark.parametrize("n,N,NF", [(1, 50, 160), (3, 64, 200), (2, 40, 4000)])
def test_wraparound_validity(n, N, NF, Index):
    """Long stream through a small ring: every record the index calls sampleable gathers exactly
    what a brute-force history says; evicted records have zero leaves; `top` counts them."""
    E, T = 3, 160
    s = _stream(E, T, seed=9 + n)
    frames_ref, a_ref, r_ref, d_ref = _entries(s, n)
    ix = Index(N, NF, n_step=n, age_limit=16)
    dev = SimDevice(N, NF, F)
    last4 = {}
    seqs = ix.head_fs + np.arange(4 * E)
    for e in range(E):
        last4[e] = seqs[4 * e:4 * e + 4].copy()
    dev.execute(ix.plan(np.zeros(0, np.int64), np.zeros((0, 8), np.int64), np.arange(4 * E), [], [], []),
                s["obs"][0].reshape(4 * E, F))
    streams = np.arange(E, dtype=np.int64)
    seen_full = False
    for k in range(T):
        kk = stack_delta(s["obs"][k], s["obs"][k + 1])
        new = np.concatenate([s["obs"][k + 1][e, 4 - kk[e]:] for e in range(E)]).reshape(-1, F)
        fs8 = ix.resolve_shift(streams, kk, last4)
        dev.execute(ix.plan(streams, fs8, np.arange(len(new)), s["action"][k], s["reward"][k], s["done"][k]), new)
        live = np.flatnonzero(dev.leaf > 0)
        assert len(live) == ix.top == int(ix.sampleable.sum())
        assert np.array_equal(np.flatnonzero(ix.sampleable), live)
        for pos in live:
            # which stream position does this ring slot hold?  seq = latest q < head_q with q % N == pos
            q = ix.head_q - 1 - ((ix.head_q - 1 - pos) % N)
            k0, e = divmod(q, E)
            i = (k0 + n - 1) * E + e
            fr, a, r, d, _ = dev.gather(int(pos), n, 0.99)
            assert np.array_equal(fr, frames_ref[i]), (k, pos)
            assert a == a_ref[i] and d == d_ref[i] and np.float64(r).view(np.int64) == r_ref[i].view(np.int64)
        seen_full |= ix.tail_q > 0
    assert seen_full and ix.top > 0 and ix.top <= N


@IMPLS
def test_compat_wraparound_small_ring(Index):
    """Reference-tuple ingest through a ring that wraps several times."""
    E, T, n = 2, 120, 3
    s = _stream(E, T, seed=4)
    frames_ref, a_ref, r_ref, d_ref = _entries(s, n)
    N, NF = 48, 150
    ix = Index(N, NF, n_step=1, age_limit=12)
    dd = ContentDeduper(ix, F)
    dev = SimDevice(N, NF, F)
    M = len(a_ref)
    for lo in range(0, M, ix.max_chunk):
        hi = min(M, lo + ix.max_chunk)
        fr = np.ascontiguousarray(frames_ref[lo:hi].reshape(hi - lo, 8, F))
        st = np.arange(lo, hi, dtype=np.int64) % E
        fs8, new_src = dd.resolve(st, fr)
        dev.execute(ix.plan(st, fs8, new_src, a_ref[lo:hi], r_ref[lo:hi], d_ref[lo:hi]), fr.reshape(-1, F))
        dd.detach(st)
        for pos in np.flatnonzero(dev.leaf > 0):
            q = ix.head_q - 1 - ((ix.head_q - 1 - pos) % N)
            got, a, r, d, _ = dev.gather(int(pos), 1, 0.99)
            assert np.array_equal(got, frames_ref[q]) and a == a_ref[q] and d == d_ref[q]
    assert ix.tail_q > 0 and 0 < ix.top <= N


@pytest.mark.parametrize("n,N,NF,age,chunk", [(3, 48, 150, 12, 7), (1, 256, 2048, None, 30), (3, 64, 400, 20, 1000), (2, 32, 200, 5, 3)])
def test_native_deduper_equals_the_specification(n, N, NF, age, chunk):
    """a0_dd_resolve (C++) against ContentDeduper (numpy specification) on the same reference entries:
    identical frame sequence numbers and identical new-frame lists, call after call -- ring wrap, frames
    that age out and are stored again, static screens (all eight frames equal), repeated frames inside
    one entry, several entries of a stream per call, and a caller that overwrites its buffer between calls."""
    E, T = 3, 90
    s = _stream(E, T, seed=10 + n)
    frames_ref, a_ref, r_ref, d_ref = _entries(s, n)
    frames_ref = frames_ref.copy()
    M = len(a_ref)
    # static screen: a run of entries of stream 1 whose eight frames are all the same picture; and an
    # entry whose next stack repeats its own first frame
    for i in range(1 + 5 * E, M, E)[:6]:
        frames_ref[i] = np.tile(frames_ref[1, :F], 8)
    frames_ref[2 + 9 * E].reshape(8, F)[5] = frames_ref[2 + 9 * E].reshape(8, F)[0]
    ix_py, ix_cc = NativeRingIndex(N, NF, 1, age), NativeRingIndex(N, NF, 1, age)
    dd_py, dd_cc = ContentDeduper(ix_py, F), NativeContentDeduper(ix_cc, F)
    step = min(chunk, ix_py.max_chunk)
    total_new = 0
    for lo in range(0, M, step):
        hi = min(M, lo + step)
        st = np.arange(lo, hi, dtype=np.int64) % E
        fr_py = np.ascontiguousarray(frames_ref[lo:hi].reshape(hi - lo, 8, F))
        fr_cc = fr_py.copy()
        fs_py, new_py = dd_py.resolve(st, fr_py)
        fs_cc, new_cc = dd_cc.resolve(st, fr_cc)
        assert np.array_equal(fs_py, fs_cc) and np.array_equal(new_py, new_cc)
        for ix, fs8, new in ((ix_py, fs_py, new_py), (ix_cc, fs_cc, new_cc)):
            ix.plan(st, fs8, new, a_ref[lo:hi], r_ref[lo:hi], d_ref[lo:hi])
        dd_py.detach(st)
        dd_cc.detach(st)
        fr_cc[:] = 0                      # the native deduper must not keep pointers into the caller's buffer
        total_new += len(new_cc)
    assert ix_py.head_fs == ix_cc.head_fs == total_new
    if age is None:
        assert total_new < 4 * M          # de-duplication works: far fewer than eight stored frames per entry
    if N < M:
        assert ix_cc.tail_q > 0


@pytest.mark.parametrize("n,N,NF,age", [(1, 64, 300, 16), (3, 48, 220, 12), (4, 200, 5000, 64), (2, 16, 120, 8)])
def test_native_index_emits_the_same_plans_as_the_specification(n, N, NF, age):
    """Differential test: random multi-stream appends (uneven chunks, idle streams that resume
    after the age limit, 0..4 new frames per step) through both implementations; every plan array
    and the whole index state must agree."""
    rng = np.random.RandomState(100 + n)
    spec, nat = RingIndex(N, NF, n, age), NativeRingIndex(N, NF, n, age)
    assert spec.max_chunk == nat.max_chunk
    S = 5
    last4 = {}
    seqs = np.arange(4 * S, dtype=np.int64)
    for ix in (spec, nat):
        p = ix.plan(np.zeros(0, np.int64), np.zeros((0, 8), np.int64), np.arange(4 * S), [], [], [])
        assert len(p.new_frame_pos) == 4 * S
    for sid in range(S):
        last4[sid] = seqs[4 * sid:4 * sid + 4].copy()
        nat.set_stack(sid, last4[sid])
    for it in range(300):
        m = int(rng.randint(1, spec.max_chunk + 1))
        active = rng.choice(S, size=int(rng.randint(1, S + 1)), replace=False)     # some streams idle for a while
        stream = rng.choice(active, size=m).astype(np.int64)
        n_new = rng.choice([0, 1, 1, 1, 2, 4], size=m).astype(np.int64)
        action = rng.randint(0, 18, m)
        reward = rng.randn(m)
        done = rng.rand(m) < 0.1
        fa = spec.resolve_shift(stream, n_new, last4)
        fb = nat.resolve_shift(stream, n_new)
        assert np.array_equal(fa, fb)
        k = int(n_new.sum())
        pa = spec.plan(stream, fa, np.arange(k), action, reward, done)
        pb = nat.plan(stream, fb, np.arange(k), action, reward, done)
        assert np.array_equal(pa.new_frame_pos, pb.new_frame_pos), it
        assert np.array_equal(pa.rec_meta, pb.rec_meta), it
        assert np.array_equal(pa.marks, pb.marks), it
        assert (spec.head_q, spec.tail_q, spec.head_fs, spec.top) == (nat.head_q, nat.tail_q, nat.head_fs, nat.top)
        assert np.array_equal(spec.sampleable, nat.sampleable)
    assert spec.tail_q > 0


def test_native_index_argument_errors():
    nat = NativeRingIndex(64, 300, 1, 16)
    with pytest.raises(RuntimeError, match="too large"):
        m = nat.max_chunk + 1
        nat.plan(np.zeros(m, np.int64), np.zeros((m, 8), np.int64), [], np.zeros(m), np.zeros(m), np.zeros(m, bool))
    with pytest.raises(RuntimeError, match="no observation stack"):
        nat.resolve_shift(np.array([7]), np.array([1]))
    with pytest.raises(RuntimeError):
        NativeRingIndex(64, 20, 1, 16)        # frame ring smaller than twice the age limit
