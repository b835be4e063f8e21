"""oracle/lz4_block.py (the specification of the device decoder K6a) pinned to the system liblz4 --
the library python-lz4 wraps, which is what the reference actor compresses with (agent.py:80) -- and the
label-based host rule a0_ex_resolve against a0_dd_resolve (its byte-level specification).  CPU only."""
import ctypes as C

import numpy as np
import pytest

from agent0_b200 import _lib
from agent0_b200.ring_index import NativeContentDeduper, NativeRingIndex
from agent0_b200.synth import record_stream
from oracle import cpu_path as CP
from oracle import lz4_block as LZ
from oracle import reference_replay as OR


def _liblz4_decode(blob, size):
    z = CP.lz4()
    out = C.create_string_buffer(size)
    k = z.lib.LZ4_decompress_safe(blob[4:], out, len(blob) - 4, size)
    return k, out.raw


def _cases():
    rng = np.random.RandomState(0)
    yield "noise", rng.randint(0, 256, 70000, dtype=np.uint8).tobytes()
    yield "zeros", bytes(56448)
    yield "period3", (b"abc" * 30000)[:56448]
    yield "period40", (bytes(range(40)) * 2000)[:56448]
    s = record_stream(2, 6, seed=3)
    fr, _, _, _ = OR.pack_nstep(s["obs"], s["action"], s["reward"], s["done"], 3, 0.99)
    for i in range(4):
        yield f"entry{i}", fr[i].tobytes()
    yield "short", b"hello world, hello world, hello world"
    yield "tiny", b"x"


@pytest.mark.parametrize("name,data", list(_cases()), ids=[c[0] for c in _cases()])
def test_oracle_decodes_what_liblz4_compresses(name, data):
    blob = CP.lz4().compress(data)
    assert LZ.decode(blob, len(data)) == data
    assert LZ.status(blob, len(data)) == LZ.OK


def test_hand_built_blocks_decode_the_same_under_liblz4():
    """Overlapping matches of every period 1..40 (and 255/256: length-extension boundaries), literal and
    match lengths that need 0, 1 and several extension bytes."""
    rng = np.random.RandomState(1)
    for period in list(range(1, 41)) + [255, 256, 300]:
        head = rng.randint(0, 256, period + rng.randint(0, 20), dtype=np.uint8).tobytes()
        seqs = [(head, period, ml) for ml in (4, 18, 19, 273, 274, 529, 2000)]
        seqs.append((rng.randint(0, 256, 270, dtype=np.uint8).tobytes(), 7, 15 + 4))
        seqs.append((b"", 1, 600))
        seqs.append((rng.randint(0, 256, 15 + 255, dtype=np.uint8).tobytes(), None, None))
        blob, want = LZ.encode_sequences(seqs)
        assert LZ.decode(blob, len(want)) == want
        k, got = _liblz4_decode(blob, len(want))
        assert k == len(want) and got == want, period


def test_malformed_blocks_are_rejected_with_the_documented_status():
    good, want = LZ.encode_sequences([(b"abcdefgh", 8, 40), (b"tail!", None, None)])
    n = len(want)
    assert LZ.status(good, n) == LZ.OK
    assert LZ.status(good, n + 1) == LZ.BAD_SIZE
    assert LZ.status(good[:3], n) == LZ.BAD_SIZE
    assert LZ.status(good[:-2], n) in (LZ.INPUT_OVERRUN, LZ.SHORT_OUTPUT)
    far, _ = LZ.encode_sequences([(b"abcdefgh", 8, 40), (b"tail!", None, None)])
    bad_off = bytearray(far); bad_off[4 + 1 + 8] = 9          # offset 9 > 8 bytes produced
    assert LZ.status(bytes(bad_off), n) == LZ.BAD_OFFSET
    zero_off = bytearray(far); zero_off[4 + 1 + 8] = 0
    assert LZ.status(bytes(zero_off), n) == LZ.BAD_OFFSET
    long_, w2 = LZ.encode_sequences([(b"abcdefgh", 8, 400), (b"", None, None)], size=n)
    assert LZ.status(long_, n) == LZ.OUTPUT_OVERRUN
    short, _ = LZ.encode_sequences([(b"abcdefgh", 8, 4), (b"", None, None)], size=n)
    assert LZ.status(short, n) == LZ.SHORT_OUTPUT
    # liblz4 rejects them too (except offset 0, which the format forbids but LZ4_decompress_safe does not check)
    for blob in (bytes(bad_off), long_, short):
        k, _ = _liblz4_decode(blob, n)
        assert k != n


# ---------------------------------------------------------------------------------- a0_ex_resolve
def canonical_labels(frames, streams, tails):
    """numpy specification of K6b: label = first byte-identical candidate among the 8 frames of the
    stream's previous entry (0..7) and the earlier frames of the own entry (8+c), else 8+j."""
    m = len(streams)
    lab = np.zeros((m, 8), dtype=np.uint8)
    for t in range(m):
        prev = tails.get(int(streams[t]))
        for j in range(8):
            l = 8 + j
            cands = ([(c, prev[c]) for c in range(8)] if prev is not None else []) + [(8 + c, frames[t, c]) for c in range(j)]
            for x, fr in cands:
                if np.array_equal(fr, frames[t, j]):
                    l = x
                    break
            lab[t, j] = l
        tails[int(streams[t])] = frames[t].copy()
    return lab


@pytest.mark.parametrize("n,N,NF,age,chunk", [(3, 48, 150, 12, 7), (1, 256, 2048, None, 30), (3, 64, 400, 20, 1000), (2, 32, 200, 5, 3)])
def test_label_rule_equals_the_byte_level_deduper(n, N, NF, age, chunk):
    """a0_ex_resolve fed with canonical labels takes exactly the decisions a0_dd_resolve takes from the
    bytes: same frame sequence numbers, same new-frame lists, call after call (ring wrap, frames that age
    out and are stored again, static screens, repeats inside one entry)."""
    F = 64
    E, T = 3, 90
    s = record_stream(E, T, seed=10 + n, frame_hw=(8, 8), p_terminal=0.07, p_life_loss=0.08, p_truncated=0.05)
    frames_ref, a_ref, r_ref, d_ref = OR.pack_nstep(s["obs"], s["action"], s["reward"], s["done"], n, 0.99)
    frames_ref = frames_ref.copy()
    M = len(a_ref)
    for i in range(1 + 5 * E, M, E)[:6]:
        frames_ref[i] = np.tile(frames_ref[1, :F], 8)
    frames_ref[2 + 9 * E].reshape(8, F)[5] = frames_ref[2 + 9 * E].reshape(8, F)[0]
    lib = _lib.load()
    ix_a, ix_b = NativeRingIndex(N, NF, 1, age), NativeRingIndex(N, NF, 1, age)
    dd = NativeContentDeduper(ix_a, F)
    ex = C.c_void_p()
    _lib.check(lib.a0_ex_create(C.byref(ex), None, ix_b.h), "a0_ex_create")
    tails = {}
    step = min(chunk, ix_a.max_chunk)
    try:
        for lo in range(0, M, step):
            hi = min(M, lo + step)
            st = np.arange(lo, hi, dtype=np.int64) % E
            fr = np.ascontiguousarray(frames_ref[lo:hi].reshape(hi - lo, 8, F))
            fs_a, new_a = dd.resolve(st, fr)
            lab = canonical_labels(fr, st, tails)
            fs_b = np.empty((hi - lo, 8), dtype=np.int64)
            new_b = np.empty((hi - lo) * 8, dtype=np.int64)
            k = C.c_int32(0)
            _lib.check(lib.a0_ex_resolve(ex, st.ctypes.data, lab.ctypes.data, hi - lo, fs_b.ctypes.data, new_b.ctypes.data,
                                         C.byref(k)), "a0_ex_resolve")
            assert np.array_equal(fs_a, fs_b) and np.array_equal(new_a, new_b[:k.value]), lo
            for ix, fs8, new in ((ix_a, fs_a, new_a), (ix_b, fs_b, new_b[:k.value])):
                ix.plan(st, fs8, new, a_ref[lo:hi], r_ref[lo:hi], d_ref[lo:hi])
        assert ix_a.head_fs == ix_b.head_fs and ix_a.top == ix_b.top
        bad = np.full((1, 8), 3, dtype=np.uint8)            # "equals frame 3 of the previous entry" for a stream without one
        with pytest.raises(RuntimeError, match="not a valid candidate"):
            _lib.check(lib.a0_ex_resolve(ex, np.array([99], dtype=np.int64).ctypes.data, bad.ctypes.data, 1, fs_b.ctypes.data,
                                         new_b.ctypes.data, C.byref(k)), "a0_ex_resolve")
    finally:
        lib.a0_ex_destroy(ex)
