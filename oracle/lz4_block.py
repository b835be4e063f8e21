"""LZ4 block format, restated -- the specification of K6a (a0_k6_lz4_decode).  Test infrastructure only.

The reference actor stores every transition as ``lz4.block.compress(concat(st, st_next))``
(agent0/deepq/agent.py:78-81) and ``ReplayDataset.__getitem__`` decompresses it
(agent0/deepq/replay.py:32-37).  ``lz4`` (python-lz4, ``lz4>=4.3.3`` in the reference's
pyproject.toml:14) is a third-party dependency that is NOT part of /root/reference, so the algorithm is
restated from the published block format (lz4_Block_format.md of the LZ4 project):

* python-lz4's ``block.compress(..., store_size=True)`` (the default the reference uses) prefixes the
  block with the decoded size as a 4-byte little-endian integer;
* a block is a series of sequences: token (high nibble = literal length, low nibble = match length - 4),
  literal-length extension bytes if the nibble is 15 (add bytes until one is < 255), the literals, a
  2-byte little-endian match offset (1..65535, counted back from the current output position), match-length
  extension bytes if that nibble is 15; the match may overlap the bytes it produces (offset < length
  repeats the last ``offset`` bytes); the last sequence stops after its literals.

Pinned (tests/test_lz4_oracle.py) against the system liblz4 -- the library python-lz4 wraps -- in both
directions: blocks made by ``LZ4_compress_default`` decode to the input here, and hand-built blocks
decode identically under ``LZ4_decompress_safe``.
"""

OK, BAD_SIZE, INPUT_OVERRUN, OUTPUT_OVERRUN, BAD_OFFSET, SHORT_OUTPUT = range(6)   # a0_ex_decode status codes


class LZ4Error(ValueError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def decode_block(block, size):
    """One raw LZ4 block -> ``size`` bytes.  Raises LZ4Error with the status code K6a reports."""
    src = memoryview(bytes(block))
    n = len(src)
    out = bytearray()
    ip = 0
    while True:
        if ip >= n:
            raise LZ4Error(INPUT_OVERRUN, "truncated: token expected")
        tok = src[ip]; ip += 1
        ll = tok >> 4
        if ll == 15:
            while True:
                if ip >= n:
                    raise LZ4Error(INPUT_OVERRUN, "truncated literal length")
                x = src[ip]; ip += 1
                ll += x
                if x != 255:
                    break
        if ll:
            if ll > n - ip:
                raise LZ4Error(INPUT_OVERRUN, "literals run past the block")
            if ll > size - len(out):
                raise LZ4Error(OUTPUT_OVERRUN, "literals run past the decoded size")
            out += src[ip:ip + ll]
            ip += ll
        if ip >= n:
            break
        if n - ip < 2:
            raise LZ4Error(INPUT_OVERRUN, "truncated offset")
        off = src[ip] | (src[ip + 1] << 8); ip += 2
        ml = tok & 15
        if ml == 15:
            while True:
                if ip >= n:
                    raise LZ4Error(INPUT_OVERRUN, "truncated match length")
                x = src[ip]; ip += 1
                ml += x
                if x != 255:
                    break
        ml += 4
        if off == 0 or off > len(out):
            raise LZ4Error(BAD_OFFSET, "match offset outside the decoded data")
        if ml > size - len(out):
            raise LZ4Error(OUTPUT_OVERRUN, "match runs past the decoded size")
        start = len(out) - off
        if off >= ml:
            out += out[start:start + ml]
        else:                                   # overlapping: the last `off` bytes repeat
            pat = bytes(out[start:])
            out += (pat * (ml // off + 1))[:ml]
    if len(out) != size:
        raise LZ4Error(SHORT_OUTPUT, f"decoded {len(out)} bytes, expected {size}")
    return bytes(out)


def decode(blob, expect=None):
    """python-lz4 framing: 4-byte little-endian decoded size + raw block."""
    blob = bytes(blob)
    if len(blob) < 5:
        raise LZ4Error(BAD_SIZE, "no size prefix")
    size = int.from_bytes(blob[:4], "little")
    if expect is not None and size != expect:
        raise LZ4Error(BAD_SIZE, f"size prefix {size}, expected {expect}")
    return decode_block(blob[4:], size)


def status(blob, expect):
    """The status code a0_ex_decode reports for this entry."""
    try:
        decode(blob, expect)
        return OK
    except LZ4Error as e:
        return e.code


def _length(extra):
    out = bytearray()
    while extra >= 255:
        out.append(255); extra -= 255
    out.append(extra)
    return bytes(out)


def encode_sequences(seqs, size=None):
    """Builds a python-lz4 blob from explicit sequences [(literals: bytes, offset, match_len) ...]; the
    last one is (literals, None, None).  For hand-made edge cases (overlapping matches of every period,
    255-chains in the length fields); returns (blob, decoded bytes)."""
    block = bytearray()
    out = bytearray()
    for lit, off, ml in seqs:
        lit = bytes(lit)
        ll = len(lit)
        mlx = 0 if ml is None else ml - 4
        assert ml is None or ml >= 4
        tok = (min(ll, 15) << 4) | min(mlx, 15)
        block.append(tok)
        if ll >= 15:
            block += _length(ll - 15)
        block += lit
        out += lit
        if ml is None:
            break
        assert 1 <= off <= len(out) and off < 65536
        block += bytes((off & 255, off >> 8))
        if mlx >= 15:
            block += _length(mlx - 15)
        start = len(out) - off
        for i in range(ml):
            out.append(out[start + i])
    n = len(out) if size is None else size
    return n.to_bytes(4, "little") + bytes(block), bytes(out)
