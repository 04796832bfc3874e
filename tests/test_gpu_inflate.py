"""GPU parity tests of the inflate path through the C ABI: golden streams, the reference's fuzz
corpus error classes, container errors, multi-member inputs, and round trips."""
import io
import zlib

import numpy as np
import pytest

from conftest import read_golden
from test_oracle_golden import DYNAMIC, FIXED, FUZZ, GZ_FTR, GZ_HDR, HELLO, STORED

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import flate_b200
    c = flate_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def o():
    from oracle import oracle
    return oracle


def err_name(ctx, data, container=0):
    import flate_b200
    with pytest.raises(flate_b200.FlateError) as ei:
        ctx.decompress(bytes(data), container)
    if type(ei.value).__name__ == "CudaError":
        raise AssertionError(str(ei.value))
    return type(ei.value).__name__


def test_block_types(ctx):
    """inflate.zig:357-479"""
    assert ctx.decompress(STORED)[0] == HELLO
    assert ctx.decompress(FIXED)[0] == HELLO
    assert ctx.decompress(DYNAMIC)[0] == b"ABCDEABCD ABCDEABCD"
    assert ctx.decompress(GZ_HDR + STORED + GZ_FTR, 1)[0] == HELLO
    assert ctx.decompress(GZ_HDR + DYNAMIC + bytes([0x17, 0x1C, 0x39, 0xB4, 0x13, 0, 0, 0]), 1)[0] == \
        b"ABCDEABCD ABCDEABCD"
    named = bytes([0x1F, 0x8B, 0x08, 0x08, 0xE5, 0x70, 0xB1, 0x65, 0x00, 0x03, 0x68, 0x65, 0x6C, 0x6C, 0x6F, 0x2E,
                   0x74, 0x78, 0x74, 0x00]) + FIXED + GZ_FTR
    assert ctx.decompress(named, 1)[0] == HELLO
    assert ctx.decompress(bytes([0x78, 0x9C]) + STORED + bytes([0x1C, 0xF2, 0x04, 0x47]), 2)[0] == HELLO


@pytest.mark.parametrize("name,out,err", FUZZ)
def test_fuzz_corpus(ctx, name, out, err):
    """inflate.zig:487-526: error class per corrupt input"""
    data = read_golden("fuzz", name + ".input")
    if err:
        assert err_name(ctx, data) == err
    else:
        want = read_golden("fuzz", name + ".expect") if out == "FILE" else out
        assert ctx.decompress(data)[0] == want


def test_container_errors(ctx):
    """src/flate.zig:255-354"""
    assert err_name(ctx, [0x78], 2) == "EndOfStream"
    assert err_name(ctx, [0x79, 0x94], 2) == "BadZlibHeader"
    assert err_name(ctx, [0x88, 0x98], 2) == "BadZlibHeader"
    assert err_name(ctx, [0x78, 0xDA, 0x03, 0x00, 0x00, 0x00, 0x00, 0x00], 2) == "WrongZlibChecksum"
    assert err_name(ctx, [0x78, 0xDA, 0x03, 0x00, 0x00], 2) == "EndOfStream"
    assert err_name(ctx, [0x1F, 0x8B], 1) == "EndOfStream"
    assert err_name(ctx, [0x1F, 0x8B, 0x09, 0, 0, 0, 0, 0, 0, 0x03], 1) == "BadGzipHeader"
    h = [0x1F, 0x8B, 0x08, 0, 0, 0, 0, 0, 0, 0x03]
    assert err_name(ctx, h + [0x03, 0x00, 0, 0, 0, 0x01, 0, 0, 0, 0], 1) == "WrongGzipChecksum"
    assert err_name(ctx, h + [0x03, 0x00, 0, 0, 0], 1) == "EndOfStream"
    assert err_name(ctx, h + [0x03, 0x00, 0, 0, 0, 0, 0, 0, 0, 0x01], 1) == "WrongGzipSize"
    assert err_name(ctx, h + [0x03, 0x00, 0, 0, 0, 0, 0, 0, 0], 1) == "EndOfStream"
    fhcrc = [0x1F, 0x8B, 0x08, 0x12, 0x00, 0x09, 0x6E, 0x88, 0x00, 0xFF, 0x48, 0x65, 0x6C, 0x6C, 0x6F, 0x00,
             0x99, 0xD6, 0x01, 0x00, 0x00, 0xFF, 0xFF, 0, 0, 0, 0, 0, 0, 0, 0]
    assert ctx.decompress(bytes(fhcrc), 1)[0] == b""
    data = bytes([0x08, 0xD7, 0x63, 0xF8, 0xCF, 0xC0, 0xC0, 0x00, 0xC1, 0xFF, 0xFF, 0x43, 0x30, 0x03, 0x03, 0xC3,
                  0xFF, 0xFF, 0xFF, 0x01, 0x83, 0x95, 0x0B, 0xF5])
    want = bytes([0x00, 0xFF, 0x00, 0x00, 0x00, 0xFF, 0x00, 0x00, 0x00, 0xFF, 0x00, 0xFF, 0xFF, 0xFF, 0x00, 0xFF,
                  0xFF, 0xFF, 0x00, 0x00, 0x00, 0x00, 0xFF, 0xFF, 0xFF])
    assert ctx.decompress(data, 2)[0] == want


def test_two_zlib_members_via_reset(ctx):
    """inflate.zig:544-563 'flate bug 18967'"""
    import flate_b200
    data = read_golden("fuzz", "first.input") + read_golden("fuzz", "second.input")
    want = read_golden("fuzz", "first.expect") + read_golden("fuzz", "second.expect")
    out = io.BytesIO()
    d = flate_b200.zlib.decompressor(io.BytesIO(data), ctx=ctx)
    d.decompress(out)
    d.reset()
    d.decompress(out)
    assert out.getvalue() == want


def test_reset_before_end_is_invalid_state(ctx):
    import flate_b200
    d = flate_b200.flate.decompressor(io.BytesIO(FIXED), ctx=ctx)
    with pytest.raises(flate_b200.FlateError) as ei:
        d.reset()
    assert type(ei.value).__name__ == "InvalidState"


def test_roundtrip_all_modes_vs_oracle_and_zlib(ctx, o):
    """bin/roundtrip.zig:14-75 behaviour: every level + huffman + store must round-trip"""
    from flate_b200 import synth
    rng = np.random.default_rng(3)
    datas = [synth.enwik_like(300000, seed=41).tobytes(), synth.mixed_small(250000, seed=42).tobytes(),
             rng.integers(0, 256, 70000, dtype=np.uint8).tobytes(), bytes(100000), b"", b"x"]
    for d in datas:
        for mode in (0, 1, 4, 6, 9):
            for container in (0, 1, 2):
                c = o.compress(d, container, mode)
                plain, used = ctx.decompress(c, container)
                assert plain == d and used == len(c), (len(d), mode, container)
        # streams produced by zlib itself (different block structure, fixed blocks, long distances)
        for lvl in (1, 6, 9):
            co = zlib.compressobj(lvl, zlib.DEFLATED, -15)
            c = co.compress(d) + co.flush()
            assert ctx.decompress(c, 0)[0] == d


def test_reader_interface_and_limits(ctx, o):
    import flate_b200
    d = read_golden("rfc1951.txt")
    c = o.compress(d, 1, 6)
    dec = flate_b200.gzip.decompressor(io.BytesIO(c), ctx=ctx)
    got = bytearray()
    while True:
        b = dec.reader().read(1000)
        if not b:
            break
        assert len(b) <= 1000
        got += b
    assert bytes(got) == d


def test_members_batch(ctx, o):
    """many independent gzip members in one launch (config C3 shape, small)"""
    from flate_b200 import synth
    members, plains = [], []
    for i in range(37):
        p = synth.enwik_like(20000 + 1777 * i, seed=100 + i).tobytes()
        plains.append(p)
        members.append(o.compress(p, 1, 6))
    blob = b"".join(members)
    off = np.cumsum([0] + [len(m) for m in members[:-1]])
    outs, st, used = ctx.decompress_members(blob, off, [len(m) for m in members], [len(p) + 64 for p in plains], 1)
    assert st == [0] * len(members)
    assert outs == plains
    assert used == [len(m) for m in members]
    # corrupt one member: only that one fails, with the reference's class
    bad = bytearray(blob)
    bad[int(off[5]) + len(members[5]) - 6] ^= 0xFF  # CRC byte
    outs, st, _ = ctx.decompress_members(bytes(bad), off, [len(m) for m in members], [len(p) + 64 for p in plains], 1)
    assert st[5] == 12 and all(s == 0 for i, s in enumerate(st) if i != 5)


def test_puff_cross_check(ctx, o):
    """the reference's own differential oracle (bin/fuzz_puff.zig:42-50): output and success must agree"""
    try:
        o.puff(b"\x03\x00")
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libpuff.so absent")
    import flate_b200
    rng = np.random.default_rng(5)
    base = o.compress(read_golden("rfc1951.txt")[:3000], 0, 6)
    for trial in range(60):
        b = bytearray(base)
        for _ in range(1 + trial % 3):
            b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        rc, plain, _ = o.puff(bytes(b), cap=1 << 20)
        try:
            got = ctx.decompress(bytes(b), 0, cap=1 << 20)[0]
            ok = True
        except flate_b200.FlateError:
            ok = False
        assert ok == (rc == 0), trial
        if ok:
            assert got == plain
        # and the error class equals the oracle's
        try:
            want = o.decompress(bytes(b), cap=1 << 20)[0]
            assert ok and want == got
        except o.OracleError as e:
            assert not ok
            assert err_name(ctx, b) == e.name


def test_differential_corruptions_all_containers(ctx, o):
    """Random bit flips, truncations and insertions on valid raw/gzip/zlib streams: the error class (or
    the output) must equal the oracle's for every mutated stream."""
    import flate_b200
    rng = np.random.default_rng(99)
    base_plain = read_golden("rfc1951.txt")[:6000]
    n_err = n_ok = 0
    for trial in range(240):
        container = trial % 3
        mode = (0, 1, 4, 6, 9)[trial % 5]
        b = bytearray(o.compress(base_plain[: 500 + (trial * 37) % 5000], container, mode))
        kind = trial % 4
        if kind == 0:
            for _ in range(1 + trial % 3):
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            del b[int(rng.integers(1, len(b))):]
        elif kind == 2:
            b.insert(int(rng.integers(0, len(b))), int(rng.integers(0, 256)))
        else:
            i = int(rng.integers(0, len(b)))
            b[i] = int(rng.integers(0, 256))
        data = bytes(b)
        try:
            want, want_used = o.decompress(data, container, cap=1 << 20)
            want_err = None
        except o.OracleError as e:
            want_err = e.name
        try:
            got, used = ctx.decompress(data, container, cap=1 << 20)
            got_err = None
        except flate_b200.FlateError as e:
            got_err = type(e).__name__
        assert got_err == want_err, (trial, container, mode, kind, got_err, want_err)
        if want_err is None:
            assert got == want and used == want_used, trial
            n_ok += 1
        else:
            n_err += 1
    assert n_err > 100 and n_ok > 5


def _member_the_reference_accepts(ctx, plain, container, mode):
    """The reference's inflate rejects a dynamic header whose code-length run crosses from the literal to the distance
    lengths (inflate.zig:161-180) although its own block writer emits such runs; about 1 in 40 MiB of level-6 text has
    one.  Shift the data until the one-shot path accepts the member."""
    import flate_b200
    for shift in range(0, 64 * 4099, 4099):
        p = plain[shift:]
        m = ctx.compress(p, container, mode)
        try:
            ctx.decompress(m, container, cap=len(p) + 64)
            return p, m
        except flate_b200.FlateError as e:
            if type(e).__name__ != "InvalidDynamicBlockHeader":
                raise
    raise AssertionError("no acceptable member found")


class _CountingReader:
    """a reader that hands out at most `unit` bytes per call and counts what it was asked for"""

    def __init__(self, data, unit=1 << 20):
        self.data, self.pos, self.unit = data, 0, unit

    def read(self, n):
        n = min(n, self.unit, len(self.data) - self.pos)
        b = self.data[self.pos:self.pos + n]
        self.pos += n
        return b


def test_streaming_decompressor_pulls_on_demand_and_resumes(ctx, o):
    """inflate.zig:313-336 next/get hand out data as it is decoded, from a reader pulled on demand (:283-309): a big
    member is decoded piece by piece (resumed at block boundaries), the first bytes arrive long before the reader is
    drained, reading stops near the member's end and what was read past it is kept."""
    import flate_b200
    from flate_b200 import synth
    base = synth.enwik_like(24 << 20, seed=301).tobytes() + bytes(300000) + np.random.default_rng(9).integers(0, 256, 200000, dtype=np.uint8).tobytes()
    for container, mode in ((1, 6), (2, 9), (0, 1), (1, 0)):
        plain, member = _member_the_reference_accepts(ctx, base, container, mode)
        trailer = b"TRAILING BYTES THAT ARE NOT PART OF THE MEMBER" * 3
        rd = _CountingReader(member + trailer, unit=300000)
        dec = flate_b200.Decompressor(container, rd, ctx=ctx)
        first = dec.next()
        assert first and plain.startswith(first)
        assert rd.pos < len(member) // 4, "the reader was drained before the first byte was handed out"
        got = bytearray(first)
        while True:
            b = dec.next()
            if b is None:
                break
            assert len(b) <= 65536
            got += b
        assert bytes(got) == plain, (container, mode)
        rest = dec.unread_bytes() + rd.data[rd.pos:]
        assert rest == trailer
        dec.close()


def test_streaming_decompressor_members_and_errors(ctx, o):
    """members one after the other through reset() with history kept, truncated input and a wrong checksum: same
    error classes as the one-shot path (inflate.zig:72-78, container.zig:45-51)."""
    import flate_b200
    from flate_b200 import synth
    a, ma = _member_the_reference_accepts(ctx, synth.enwik_like(3 << 20, seed=302).tobytes(), 1, 6)
    b, mb = _member_the_reference_accepts(ctx, synth.mixed_small(400000, seed=303).tobytes(), 1, 4)
    dec = flate_b200.Decompressor(1, _CountingReader(ma + mb, unit=70000), ctx=ctx)
    out = io.BytesIO()
    dec.decompress(out)
    dec.reset()
    dec.decompress(out)
    assert out.getvalue() == a + b
    # truncated inside the deflate stream, and inside the footer
    for cut in (len(ma) // 2, len(ma) - 3):
        dec = flate_b200.Decompressor(1, _CountingReader(ma[:cut], unit=50000), ctx=ctx)
        with pytest.raises(flate_b200.FlateError) as ei:
            dec.decompress(io.BytesIO())
        assert type(ei.value).__name__ == "EndOfStream", cut
    bad = bytearray(ma)
    bad[-6] ^= 0x40
    dec = flate_b200.Decompressor(1, _CountingReader(bytes(bad), unit=50000), ctx=ctx)
    with pytest.raises(flate_b200.FlateError) as ei:
        dec.decompress(io.BytesIO())
    assert type(ei.value).__name__ == "WrongGzipChecksum"
    # a corrupted block in the middle of a long member: the one-shot path's error class
    bad = bytearray(ma)
    for i in range(len(ma) // 2, len(ma) // 2 + 40):
        bad[i] ^= 0xA5
    try:
        ctx.decompress(bytes(bad), 1)
        want = None
    except flate_b200.FlateError as e:
        want = type(e).__name__
    dec = flate_b200.Decompressor(1, _CountingReader(bytes(bad), unit=50000), ctx=ctx)
    try:
        dec.decompress(io.BytesIO())
        got = None
    except flate_b200.FlateError as e:
        got = type(e).__name__
    assert got == want and want is not None


def test_members_batch_with_tiny_members_at_unaligned_offsets(ctx, o):
    """an empty member and a one-byte member in the middle of a batch, at output offsets that are not multiples of 16:
    nothing of a member's first 16-byte line may be written back before the line is complete"""
    from flate_b200 import synth
    items = [synth.enwik_like(200000 + 70001 * i, seed=500 + i).tobytes() for i in range(9)] + [b"", b"x", bytes(100000), b"", b"ab"]
    members = [o.compress(it, 1, 6) for it in items]
    blob = b"".join(members)
    lens = [len(m) for m in members]
    offs = [sum(lens[:i]) for i in range(len(lens))]
    plains, st, used = ctx.decompress_members(blob, offs, lens, [len(it) + 16 for it in items], 1)
    ok = [s == 0 for s in st]
    assert all(p == it for p, it, good in zip(plains, items, ok) if good)
    assert all(ok[9:])


def _sequential_members(ctx, blob):
    """what the reference does with a multi-member file: decompress, reset, decompress ... (inflate.zig:301-309)"""
    import flate_b200
    out, pos, members = bytearray(), 0, 0
    while pos < len(blob):
        plain, used = ctx.decompress(blob[pos:], 1, cap=1 << 24)
        out += plain
        pos += used
        members += 1
    return bytes(out), pos, members


def test_gzip_file_members_found_without_an_index(ctx, o):
    """the members of a concatenated gzip file are found (header look-alikes, validated by
    decoding) and inflated in one launch; same bytes, same member count, same end as the sequential loop, also when the
    compressed data itself contains the magic bytes and when other bytes follow the last member."""
    import flate_b200
    from flate_b200 import synth
    rng = np.random.default_rng(77)
    plains, members = [], []
    for i in range(40):
        p, m = _member_the_reference_accepts(ctx, synth.enwik_like(30000 + 9973 * i, seed=700 + i).tobytes(), 1, 6)
        plains.append(p)
        members.append(m)
    # members whose stored blocks carry header look-alikes, an empty member, a member of random bytes
    decoy = (b"\x1f\x8b\x08\x00" + bytes(20)) * 500
    for p, mode in ((decoy, 0), (b"", 6), (decoy + b"tail", 1), (rng.integers(0, 256, 100000, dtype=np.uint8).tobytes(), 6)):
        plains.insert(7, p)
        members.insert(7, o.compress(p, 1, mode))
    blob = b"".join(members)
    want = b"".join(plains)
    got, used, cnt = ctx.decompress_gzip_file(blob)
    assert got == want and used == len(blob) and cnt == len(members)
    # other bytes after the last member: the members end where the sequential loop would stop
    got, used, cnt = ctx.decompress_gzip_file(blob + b"\x00" * 1000)
    assert got == want and used == len(blob) and cnt == len(members)
    # a corrupt member in the middle: the sequential loop's first error, after the members before it
    bad = bytearray(blob)
    at = sum(len(m) for m in members[:20]) + len(members[20]) // 2
    for i in range(at, at + 30):
        bad[i] ^= 0x5A
    try:
        _sequential_members(ctx, bytes(bad))
        want_err = None
    except flate_b200.FlateError as e:
        want_err = type(e).__name__
    assert want_err is not None
    with pytest.raises(flate_b200.FlateError) as ei:
        ctx.decompress_gzip_file(bytes(bad))
    assert type(ei.value).__name__ == want_err
