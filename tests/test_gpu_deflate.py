"""GPU parity tests of the deflate path, through the C ABI, against the CPU oracle and the
reference's golden vectors.  Bit-exact everywhere (integer/byte work)."""
import io
import json
import os
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, read_golden

pytestmark = pytest.mark.gpu

LEVELS = [4, 5, 6, 7, 8, 9]
WBITS = {0: -15, 1: 31, 2: 15}


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


def first_diff(a, b):
    a = np.frombuffer(a, dtype=np.uint8) if isinstance(a, bytes) else np.asarray(a)
    b = np.frombuffer(b, dtype=np.uint8) if isinstance(b, bytes) else np.asarray(b)
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    return (int(d[0]) if d.size else n), a.size, b.size


def inputs():
    from flate_b200 import synth
    rng = np.random.default_rng(7)
    cases = {
        "empty": b"", "one": b"a", "three": b"abc", "four": b"abcd", "five": b"aaaaa",
        "blah": b"Blah blah blah blah blah!", "abcde": b"ABCDEABCD ABCDEABCD",
        "rfc": read_golden("rfc1951.txt"),
        "zeros70k": bytes(70000), "text64k-1": synth.enwik_like(65535).tobytes(),
        "text64k": synth.enwik_like(65536).tobytes(), "text64k+1": synth.enwik_like(65537).tobytes(),
        "text70000": synth.enwik_like(70000, seed=3).tobytes(),
        "text98304": synth.enwik_like(98304, seed=4).tobytes(),
        "mixed130809": synth.mixed_small(130809, seed=5).tobytes(),
        "mixed130810": synth.mixed_small(130810, seed=6).tobytes(),
        "mixed200001": synth.mixed_small(200001, seed=8).tobytes(),
        "random100k": rng.integers(0, 256, 100000, dtype=np.uint8).tobytes(),
        "lowentropy": rng.integers(0, 3, 150000, dtype=np.uint8).tobytes(),
        "text1M": synth.enwik_like(1 << 20, seed=9).tobytes(),
    }
    return cases


@pytest.fixture(scope="module")
def data_cases():
    return inputs()


@pytest.mark.parametrize("level", LEVELS)
def test_match_tables_equal_oracle(ctx, o, data_cases, level):
    """K1+K2 in isolation: findMatch for every position under both budgets (deflate.zig:233-266)."""
    for name in ("blah", "rfc", "text70000", "mixed130810", "lowentropy"):
        d = data_cases[name]
        if level >= 8 and len(d) > 140000:
            continue
        rf, rq = ctx.debug_match_tables(d, level)
        of, oq = o.match_tables(d, level)
        assert first_diff(rf, of)[0] == len(d), (name, level, "full", first_diff(rf, of))
        assert first_diff(rq, oq)[0] == len(d), (name, level, "quarter", first_diff(rq, oq))


@pytest.mark.parametrize("level", LEVELS)
def test_token_stream_equals_oracle(ctx, o, data_cases, level):
    for name, d in data_cases.items():
        if not d:
            continue
        if level >= 8 and name in ("lowentropy", "text1M"):
            continue
        got = ctx.debug_tokens(d, level)
        want = o.tokenize(d, level)
        fd = first_diff(got, want)
        assert fd[0] == want.size and got.size == want.size, (name, level, fd)


def test_exact_token_goldens(ctx):
    """deflate.zig:539-554"""
    M = lambda d, l: 0x80000000 | ((d - 1) << 8) | (l - 3)
    L = ord
    assert ctx.debug_tokens(b"Blah blah blah blah blah!", 6).tolist() == \
        [L('B'), L('l'), L('a'), L('h'), L(' '), L('b'), M(5, 18), L('!')]
    assert ctx.debug_tokens(b"ABCDEABCD ABCDEABCD", 6).tolist() == \
        [L(c) for c in "ABCDEABCD A"] + [M(10, 8)]


def test_token_count_goldens(ctx):
    """deflate.zig:613-643"""
    table = [
        (("rfc1951.txt",), [7675, 7672, 7599, 7594, 7598, 7599]),
        (("block_writer", "huffman-null-max.input"), [257] * 6),
        (("block_writer", "huffman-pi.input"), [2570, 2564, 2564, 2564, 2564, 2564]),
        (("block_writer", "huffman-text.input"), [235, 234, 234, 234, 234, 234]),
        (("fuzz", "roundtrip1.input"), [333, 331, 331, 331, 331, 331]),
        (("fuzz", "roundtrip2.input"), [334] * 6),
    ]
    for path, counts in table:
        d = read_golden(*path)
        for level, want in zip(LEVELS, counts):
            assert ctx.debug_tokens(d, level).size == want, (path, level)


@pytest.mark.parametrize("kind", ["wb", "dyn", "huff"])
def test_block_writer_goldens(ctx, kind):
    """block_writer.zig:599-706: 43 byte-exact block encodings, eof False and True."""
    cases = json.load(open(os.path.join(GOLDEN, "block_writer_tokens.json")))
    extra = [{"input": "huffman-rand-max.input", "want": "huffman-rand-max.{s}.expect", "want_no_input": "",
              "tokens": []}] if kind == "huff" else []
    n = 0
    for tc in cases + extra:
        toks = np.array([t[0] if len(t) == 1 else (0x80000000 | ((t[0] - 1) << 8) | (t[1] - 3)) for t in tc["tokens"]],
                        dtype=np.uint32)
        variants = []
        if tc["input"] and tc["want"]:
            variants.append((read_golden("block_writer", tc["input"]),
                             read_golden("block_writer", tc["want"].replace("{s}", kind))))
        if kind != "huff" and tc["want_no_input"]:
            variants.append((None, read_golden("block_writer", tc["want_no_input"].replace("{s}", kind))))
        for inp, want in variants:
            got = ctx.debug_block_write(kind, toks, False, inp)
            assert got == want, (tc["want"] or tc["want_no_input"], kind, inp is None, first_diff(got, want))
            got_eof = bytearray(ctx.debug_block_write(kind, toks, True, inp))
            assert got_eof[0] & 1 == 1
            got_eof[0] &= 0xFE
            assert bytes(got_eof) == want
            n += 1
    assert n == {"wb": 17, "dyn": 17, "huff": 9}[kind]


def test_compressed_size_goldens(ctx):
    """src/flate.zig:95-124: exact sizes for 4 inputs x (6 levels + huffman + store) x 3 containers."""
    table = [
        (("rfc1951.txt",), [11513, 11217, 11139, 11126, 11122, 11119], 20287, 36967),
        (("fuzz", "roundtrip1.input"), [373, 370, 370, 370, 370, 370], 393, 393),
        (("fuzz", "roundtrip2.input"), [373] * 6, 394, 394),
        (("fuzz", "deflate-stream.expect"), [351, 347, 347, 347, 347, 347], 498, 747),
    ]
    csize = {0: 0, 1: 18, 2: 6}
    for path, gz, huff, store in table:
        d = read_golden(*path)
        for mode, want in list(zip(LEVELS, gz)) + [(1, huff), (0, store)]:
            for container in (0, 1, 2):
                c = ctx.compress(d, container, mode)
                assert len(c) == want - 18 + csize[container], (path, mode, container)
                assert zlib.decompress(c, WBITS[container]) == d


@pytest.mark.parametrize("mode", [0, 1] + LEVELS)
def test_compress_bit_exact_vs_oracle(ctx, o, data_cases, mode):
    for name, d in data_cases.items():
        if mode >= 8 and name in ("lowentropy", "text1M"):
            continue
        for container in ((0, 1, 2) if len(d) < 100000 else (0,)):
            got = ctx.compress(d, container, mode)
            want = o.compress(d, container, mode)
            assert got == want, (name, mode, container, first_diff(got, want))


def test_kats(ctx):
    """SURVEY appendix A6 + src/flate.zig:356-384"""
    for lvl in LEVELS:
        assert ctx.compress(b"", 0, lvl) == bytes([0x03, 0x00])
    assert ctx.compress(b"a", 0, 6) == bytes([0x4B, 0x04, 0x00])
    assert ctx.compress(b"", 0, 1) == bytes([0x01, 0x00, 0x00, 0xFF, 0xFF])
    hello = b"Hello world\n"
    stored = bytes([1, 12, 0, 0xF3, 0xFF]) + hello
    assert ctx.compress(hello, 0, 0) == stored
    assert ctx.compress(hello, 1, 0) == bytes([0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, 3]) + stored + \
        bytes([0xD5, 0xE0, 0x39, 0xB7, 0x0C, 0, 0, 0])
    assert ctx.compress(hello, 2, 0) == bytes([0x78, 0x9C]) + stored + bytes([0x1C, 0xF2, 0x04, 0x47])
    assert ctx.compress(b"Hello world!", 0, 1) == bytes([1, 12, 0, 0xF3, 0xFF]) + b"Hello world!"


def test_stored_input_quirk(ctx, o):
    """the 32768-th token is a match whose bytes are not in `input`; the reference
    then emits a stored block that drops them.  Bit-exactness means reproducing it."""
    rng = np.random.default_rng(11)
    rnd = rng.integers(0, 256, 32770, dtype=np.uint8).tobytes()
    from flate_b200 import synth
    d = rnd[:32767] + rnd[100:130] + synth.enwik_like(40000, seed=21).tobytes()
    for lvl in (4, 6, 9):
        assert ctx.compress(d, 0, lvl) == o.compress(d, 0, lvl)


def test_huffman_only_and_store_multiblock(ctx, o):
    from flate_b200 import synth
    d = synth.random_zero_mix(65535 * 3 + 17).tobytes()
    for mode in (0, 1):
        for n in (65534, 65535, 65536, 65535 * 2, len(d)):
            got = ctx.compress(d[:n], 0, mode)
            assert got == o.compress(d[:n], 0, mode), (mode, n)
            assert zlib.decompress(got, -15) == d[:n]


def test_public_interface_compressor(ctx, o):
    """src/flate.zig:386-481 testInterface, all API spellings"""
    import flate_b200
    plain = b"Hello world\n" * 50 + read_golden("rfc1951.txt")[:5000]
    for mod, container in ((flate_b200.flate, 0), (flate_b200.gzip, 1), (flate_b200.zlib, 2)):
        w = io.BytesIO()
        mod.compress(io.BytesIO(plain), w, level=flate_b200.Level.default, ctx=ctx)
        assert w.getvalue() == o.compress(plain, container, 6)
        w2 = io.BytesIO()
        c = mod.compressor(w2, level=flate_b200.Level.best, ctx=ctx)
        for i in range(0, len(plain), 777):
            c.writer().write(plain[i:i + 777])
        c.finish()
        assert w2.getvalue() == o.compress(plain, container, 9)
        for simple, mode in ((mod.huffman, 1), (mod.store, 0)):
            w3 = io.BytesIO()
            simple.compress(io.BytesIO(plain), w3, ctx=ctx)
            assert w3.getvalue() == o.compress(plain, container, mode)
            w4 = io.BytesIO()
            sc = simple.compressor(w4, ctx=ctx)
            sc.compress(io.BytesIO(plain))
            sc.finish()
            assert w4.getvalue() == w3.getvalue()


def test_device_resident_compress(ctx, o):
    import torch
    from flate_b200 import synth
    d = synth.enwik_like(3 << 20, seed=31)
    t_in = torch.from_numpy(d).cuda()
    cap = ctx.lib.fb200_compress_bound(d.size, 6) + 64
    t_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n = ctx.compress_device(t_in.data_ptr(), d.size, t_out.data_ptr(), cap, mode=6,
                            stream=torch.cuda.current_stream().cuda_stream)
    got = t_out[:n].cpu().numpy().tobytes()
    assert got == o.compress(d.tobytes(), 0, 6)


def _stream_both(ctx, o, container, mode, pieces):
    """pieces: list of bytes or the string 'flush'.  Returns (gpu_bytes, oracle_bytes)."""
    import flate_b200
    w = io.BytesIO()
    c = flate_b200.Compressor(container, w, mode, ctx=ctx)
    d = o.Deflate(container, mode)
    for p in pieces:
        if isinstance(p, str):
            c.flush()
            d.flush()
        else:
            c.write(p)
            d.write(p)
    c.finish()
    d.finish()
    return w.getvalue(), d.output()


@pytest.mark.parametrize("mode", [0, 1, 4, 6, 9])
def test_streaming_flush_bit_exact(ctx, o, mode):
    """deflate.zig:335-337 flush (sync marker 00 00 ff ff), :351 history kept across flushes; the last 3
    bytes before a flush point are never hashed (Lookup.zig:24)."""
    from flate_b200 import synth
    text = synth.enwik_like(400000, seed=61).tobytes()
    mixed = synth.mixed_small(150000, seed=62).tobytes()
    scripts = [
        [text[:10000], "flush", text[10000:30000]],
        [b"Blah blah blah blah blah!", "flush"],                                   # deflate.zig:556-566
        [text[:3], "flush", text[3:5], "flush", text[5:6], "flush", text[6:5000]],   # flush points closer than 4 bytes
        ["flush", "flush", text[:100], "flush"],                                    # empty segments
        [text[:65536], "flush", text[65536:200000], "flush", text[200000:]],       # across window slides
        [text[:70001], "flush", text[70001:70004], "flush", text[70004:140000]],
        [mixed[:33000], "flush", mixed[33000:99000], "flush", mixed[99000:]],
        [text[:65535], "flush", text[65535:65535 * 2], "flush"],                    # SimpleCompressor slice boundaries
    ]
    for i, sc in enumerate(scripts):
        for container in (0, 1):
            got, want = _stream_both(ctx, o, container, mode, sc)
            assert got == want, (mode, i, container, first_diff(got, want))
    # and the flushed prefix is decodable on its own (the point of a sync flush)
    import flate_b200
    w = io.BytesIO()
    c = flate_b200.Compressor(0, w, mode, ctx=ctx)
    c.write(text[:50000])
    c.flush()
    part = w.getvalue()
    assert part.endswith(b"\x00\x00\xff\xff")
    assert zlib.decompressobj(-15).decompress(part) == text[:50000]
    c.finish()


def test_gzip_zlib_footer_from_device_checksum(ctx, o):
    from flate_b200 import synth
    for n in (0, 1, 4095, 4096, 4097, 1 << 20, (1 << 20) + 13):
        d = synth.mixed_small(n, seed=70 + n % 7).tobytes() if n else b""
        for container in (1, 2):
            got = ctx.compress(d, container, 6)
            assert got[-8:] == o.compress(d, container, 6)[-8:], (n, container)


def _structured(rng, n, kind):
    if kind == 0:      # tiny alphabet: long hash chains, many equal-length candidates (tie-break order matters)
        return rng.integers(0, 2, n, dtype=np.uint8).tobytes()
    if kind == 1:      # periodic with noise: overlapping matches, distance < length
        period = int(rng.integers(1, 40))
        base = rng.integers(97, 123, period, dtype=np.uint8)
        a = np.tile(base, n // period + 1)[:n].copy()
        flips = rng.integers(0, n, max(1, n // 97))
        a[flips] = rng.integers(0, 256, flips.size, dtype=np.uint8)
        return a.tobytes()
    if kind == 2:      # words from a small dictionary
        words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(50)]
        out = bytearray()
        while len(out) < n:
            out += words[int(rng.integers(0, 50))] + b" "
        return bytes(out[:n])
    from flate_b200 import synth
    return synth.mixed_small(n, seed=int(rng.integers(1, 1 << 20))).tobytes()


def test_differential_around_window_boundaries(ctx, o):
    """Sizes straddling every special position of the reference's schedule: the 64 KiB fill, the
    32768-byte slides, the 262-byte look-ahead reserve (SlidingWindow.zig:13), 65535-byte stored/huffman
    slices and the 32768-token block cut."""
    rng = np.random.default_rng(2024)
    specials = [32768, 65274, 65535, 65536, 98042, 98304, 131072]
    sizes = sorted({max(0, s + d) for s in specials for d in (-263, -4, -1, 0, 1, 3, 262)})
    picked = [sizes[i] for i in rng.permutation(len(sizes))[:18]]
    for i, n in enumerate(picked):
        d = _structured(rng, n, i % 4)
        for mode in (4, 6, 9, 1):
            if mode == 9 and i % 4 == 0 and n > 70000:
                continue  # binary alphabet at level 9: the oracle walks 4096-deep chains
            got = ctx.compress(d, 0, mode)
            want = o.compress(d, 0, mode)
            assert got == want, (n, i % 4, mode, first_diff(got, want))


def test_differential_small_random(ctx, o):
    rng = np.random.default_rng(77)
    for i in range(60):
        n = int(rng.integers(0, 3000))
        d = _structured(rng, n, i % 4)
        for mode in (4, 5, 6, 7, 8, 9, 1, 0):
            got = ctx.compress(d, i % 3, mode)
            want = o.compress(d, i % 3, mode)
            assert got == want, (n, i % 4, mode, i % 3, first_diff(got, want))
        if n:
            # and the token seam
            assert ctx.debug_tokens(d, 6).tolist() == o.tokenize(d, 6).tolist()


@pytest.mark.parametrize("level", [4, 6, 9])
def test_position_sharded_stream_equals_whole(ctx, o, level):
    """ranges of one stream searched independently (as different GPUs would), lazy-step
    tables joined, parse + block writer once: byte-identical to the unsharded stream."""
    import torch
    from flate_b200 import synth
    for data in (synth.enwik_like(700001, seed=92), synth.mixed_small(700001, seed=91)):
        _check_sharded_stream(ctx, o, level, data)


def _check_sharded_stream(ctx, o, level, data):
    import torch
    n = data.size
    d_in = torch.from_numpy(data).cuda()
    cap = ctx.lib.fb200_compress_bound(n, level) + 64
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    want = o.compress(data.tobytes(), 1, level)
    import flate_b200
    from flate_b200 import sharding
    ov, align = ctx.shard_overlap, ctx.shard_align
    for parts in (1, 2, 3, 5):
        for mode in (0, 1):                     # sparse parse per range / dense tables per range
            per, ranges = sharding.shard_positions(n, parts, align if mode == 0 else 8192)
            tables = []
            for r in range(parts):
                # a fresh context and a private table per range, like a separate GPU: nothing is inherited
                c = flate_b200.Context(0)
                c.set_parse_mode(mode)
                nx_r = torch.full((parts * per + ov,), 0x5a5a5a5a, dtype=torch.int32, device="cuda")
                lo, hi = ranges[r]
                ok = c.shard_search(d_in.data_ptr(), n, lo, hi, nx_r.data_ptr(), level=level)
                c.close()
                assert ok or mode == 0
                if not ok:
                    break
                tables.append(nx_r)
            if len(tables) < parts:
                continue                        # the sparse parse declined (periodic data): the dense run covers it
            nx = torch.full((parts * per + ov,), -1, dtype=torch.int32, device="cuda")
            for r in range(parts):              # what the all-gather does
                nx[r * per:(r + 1) * per] = tables[r][r * per:(r + 1) * per]
            if mode == 0:
                for r in range(parts - 1):
                    sharding.merge_overlap(nx[(r + 1) * per:(r + 1) * per + ov], tables[r][(r + 1) * per:(r + 1) * per + ov])
            m = ctx.shard_finish(d_in.data_ptr(), n, nx.data_ptr(), d_out.data_ptr(), cap, level=level, container=1)
            got = d_out[:m].cpu().numpy().tobytes()
            assert got == want, (level, parts, mode, first_diff(got, want))


# ---- parse strategies: the speculative sparse parse (default), its repair path and the dense tables ----
def _sparse_cases():
    from flate_b200 import synth
    rng = np.random.default_rng(21)
    text = synth.enwik_like(3 << 20, seed=31)
    return {
        "zeros1M": np.zeros(1 << 20, np.uint8),
        "period7": np.tile(np.arange(7, dtype=np.uint8), 60000),
        "lowentropy600k": rng.integers(0, 4, 600000, dtype=np.uint8),
        "text+zeros+text": np.concatenate([text[:700000], np.zeros(200000, np.uint8), text[700000:1500000]]),
        "text+period+text": np.concatenate([text[:1000000], np.tile(np.arange(11, dtype=np.uint8), 30000), text[1000000:2000000]]),
        "text3M": text,
        "text32768": text[:32768], "text33791": text[:33791], "text33792": text[:33792], "text33793": text[:33793],
        "text65536+1025": text[:65536 + 1025],
    }


@pytest.mark.parametrize("level", [4, 6, 9])
def test_sparse_parse_equals_dense_tables_and_oracle(o, level):
    """Mode 0 (sparse parse; repaired or redone densely when its coverage check fails) and mode 1 (dense
    match tables) both give the oracle's bytes, on inputs chosen to make the speculation fail."""
    import flate_b200
    sparse, dense = flate_b200.Context(0), flate_b200.Context(0)
    dense.set_parse_mode(1)
    try:
        for name, d in _sparse_cases().items():
            if level == 9 and d.size > (2 << 20):
                d = d[:2 << 20]
            want = o.compress(d.tobytes(), o.RAW, level)
            a = sparse.compress(d, flate_b200.RAW, level)
            b = dense.compress(d, flate_b200.RAW, level)
            assert a == want, (name, "sparse", first_diff(a, want))
            assert b == want, (name, "dense", first_diff(b, want))
        assert dense.sparse_fallbacks == 0 and dense.sparse_repairs == 0
        # zeros and the short period never let two parses fall into step: redone densely;
        # a periodic stretch inside text is repaired chunk-wise
        assert sparse.sparse_fallbacks >= 2
        assert sparse.sparse_repairs >= 2
    finally:
        sparse.close()
        dense.close()


def test_sparse_parse_counters_stay_zero_on_text(o):
    import flate_b200
    from flate_b200 import synth
    c = flate_b200.Context(0)
    try:
        d = synth.enwik_like(8 << 20, seed=77)
        got = c.compress(d, flate_b200.GZIP, 6)
        assert got == o.compress(d.tobytes(), o.GZIP, 6)
        assert c.sparse_fallbacks == 0 and c.sparse_repairs == 0
    finally:
        c.close()


def test_dense_mode_tokens_equal_oracle(o):
    import flate_b200
    from flate_b200 import synth
    c = flate_b200.Context(0)
    c.set_parse_mode(1)
    try:
        for n in (5, 4097, 70000, 300001):
            d = synth.enwik_like(n, seed=n).tobytes()
            for level in (4, 6, 9):
                got, want = c.debug_tokens(d, level), o.tokenize(d, level)
                assert got.size == want.size and (got == want).all(), (n, level)
    finally:
        c.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_block_range_sharded_stream_equals_whole(ctx, o, mode):
    """a huffman-only / store stream cut into ranges of 65535-byte slices, every range planned and
    packed on its own ("ranks" run one after the other on this GPU), joined by the exclusive scan of the shard size
    summaries and an OR of the boundary bytes: byte-identical with the oracle's stream (deflate.zig:449-529)."""
    import torch
    from flate_b200 import sharding, synth
    sizes = [0, 1, 65534, 65535, 65536, 131070, 200000, 1500000]
    for n in sizes:
        data = synth.random_zero_mix(n, seed=0x5EED0005 + n) if n else np.zeros(0, dtype=np.uint8)
        if n > 300000:   # text in the middle: dynamic blocks of every size between stored ones
            data[400000:900000] = synth.enwik_like(500000, seed=9)
        for container in (0, 1, 2):
            want = o.compress(data.tobytes(), container, mode)
            for world in (1, 2, 3, 5, 8):
                ranges = sharding.simple_shard_ranges(n, world)
                d_in = torch.from_numpy(data.copy()).cuda() if n else torch.zeros(16, dtype=torch.uint8, device="cuda")
                plans = []
                for r, (lo, hi) in enumerate(ranges):
                    last = r == world - 1
                    if hi > lo or last:
                        plans.append(ctx.simple_shard_plan(d_in.data_ptr() + lo, hi - lo, last, mode=mode, container=container))
                    else:
                        plans.append((0, 0, 0, 1 if container == 2 else 0))
                header = sharding._HEADERS[container]
                starts, end_bit = sharding.shard_start_bits([p[:3] for p in plans], 8 * len(header))
                pad = max(hi - lo for lo, hi in ranges)
                pad = (pad + pad // 8 + 1024 + 255) // 256 * 256
                gathered = torch.zeros(world * pad, dtype=torch.uint8, device="cuda")
                placements = []
                for r, (lo, hi) in enumerate(ranges):
                    last = r == world - 1
                    if hi > lo or last:
                        ctx.simple_shard_plan(d_in.data_ptr() + lo, hi - lo, last, mode=mode, container=container)
                        blo, nb, _ = ctx.simple_shard_pack(starts[r], gathered.data_ptr() + r * pad, pad)
                        placements.append((blo, nb))
                    else:
                        placements.append((starts[r] >> 3, 0))
                body_end = (end_bit + 7) >> 3
                final = torch.zeros(body_end, dtype=torch.uint8, device="cuda")
                if header:
                    final[: len(header)] = torch.frombuffer(bytearray(header), dtype=torch.uint8).cuda()
                sharding.assemble_shards(final, gathered, pad, placements)
                total = plans[0][3]
                for p, (lo, hi) in zip(plans[1:], ranges[1:]):
                    total = ctx.crc32_combine(total, p[3], hi - lo) if container == 1 else ctx.adler32_combine(total, p[3], hi - lo)
                footer = b"" if container == 0 else (
                    total.to_bytes(4, "little") + (n & 0xffffffff).to_bytes(4, "little") if container == 1 else total.to_bytes(4, "big"))
                got = final.cpu().numpy().tobytes() + footer
                assert got == want, (n, container, world, first_diff(got, want))


@pytest.mark.gpu
def test_rolling_sparse_parse_is_bit_exact():
    """The opt-in rolling form of the sparse parse (FB200_SPARSE_ROLL, read once per process) gives the oracle's bytes:
    text with a zero run and a random burst, long enough for several runs per CTA plus a chunk-kernel tail."""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "import flate_b200; from flate_b200 import synth; from oracle import oracle as o\n"
        "d = synth.enwik_like(6 << 20, seed=77).copy()\n"
        "d[1000000:1300000] = 0\n"
        "d[3000000:3100000] = np.random.default_rng(5).integers(0, 256, 100000, dtype=np.uint8)\n"
        "d = d.tobytes()[:6 * 1024 * 1024 - 12345]\n"
        "ctx = flate_b200.Context(0)\n"
        "for lvl in (6, 4, 9):\n"
        "    assert ctx.compress(d, flate_b200.RAW, lvl) == o.compress(d, o.RAW, lvl), lvl\n"
        "print('rolling ok')\n" % ROOT)
    env = dict(os.environ, FB200_SPARSE_ROLL="16,2")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "rolling ok" in r.stdout, r.stderr[-2000:]


def _stream_data(n, seed):
    from flate_b200 import synth
    d = synth.enwik_like(n, seed=seed).copy()
    rng = np.random.default_rng(seed)
    d[n // 5:n // 5 + 300000] = 0                                                  # periodic: the coverage check fails here
    d[n // 2:n // 2 + 200000] = rng.integers(0, 256, 200000, dtype=np.uint8)      # incompressible: stored blocks
    d[n // 2 + 400000:n // 2 + 400000 + 70000] = d[1000:71000]                    # a long repeat from far back
    return d.tobytes()


@pytest.mark.parametrize("mode", [6, 4, 9, 1, 0])
def test_streaming_in_parts_is_bit_exact(ctx, o, mode, monkeypatch):
    """deflate.zig:363-371: write() emits blocks as they fill, and the bytes do not depend on how the input was cut
    into write() calls.  With a small part size the stream below goes through many parts (window slides, carried
    open blocks, carried partial bytes, a part whose speculation fails) and must still equal the one-shot stream."""
    import flate_b200
    monkeypatch.setenv("FB200_STREAM_PART", "256")
    data = _stream_data(5 * 1024 * 1024 + 4321, seed=90 + mode)
    for container in (0, 1, 2):
        want = o.compress(data, container, mode)
        w = io.BytesIO()
        c = flate_b200.Compressor(container, w, mode, ctx=ctx)
        pos, k, emitted_early = 0, 0, 0
        while pos < len(data):
            step = (37 + 7919 * k) % 150000 + 1
            c.write(data[pos:pos + step])
            pos += step
            k += 1
            if pos < len(data) // 2:
                emitted_early = len(w.getvalue())
        c.finish()
        got = w.getvalue()
        assert got == want, (mode, container, first_diff(got, want))
        assert emitted_early > len(want) // 8, "nothing left the compressor before finish()"
        c.close()


@pytest.mark.parametrize("mode", [6, 1])
def test_streaming_parts_with_flushes(ctx, o, mode, monkeypatch):
    """flush points in the middle of a stream that is otherwise compressed part by part (deflate.zig:335-337)."""
    import flate_b200
    monkeypatch.setenv("FB200_STREAM_PART", "128")
    data = _stream_data(2 * 1024 * 1024 + 99, seed=7)
    cuts = [0, 400001, 400003, 1300000, 1300000, 1900000, len(data)]
    w = io.BytesIO()
    c = flate_b200.Compressor(1, w, mode, ctx=ctx)
    d = o.Deflate(1, mode)
    for a, b in zip(cuts[:-1], cuts[1:]):
        for p in range(a, b, 50000):
            c.write(data[p:min(p + 50000, b)])
        d.write(data[a:b])
        if b != len(data):
            c.flush()
            d.flush()
    c.finish()
    d.finish()
    got, want = w.getvalue(), d.output()
    assert got == want, (mode, first_diff(got, want))


def test_pool_batch_over_all_devices(o):
    """fb200_compress_batch / fb200_decompress_members_batch: independent streams and the members of one buffer spread
    over every device present by the library itself; results equal the one-device calls (and the oracle)."""
    import flate_b200
    from flate_b200 import synth
    pool = flate_b200.Pool(0)
    assert pool.devices >= 1
    items = [synth.enwik_like(200000 + 70001 * i, seed=500 + i).tobytes() for i in range(9)] + [b"", b"x", bytes(100000)]
    for container, mode in ((1, 6), (0, 1), (2, 9)):
        got = pool.compress_batch(items, container, mode)
        for g, it in zip(got, items):
            assert g == o.compress(it, container, mode)
    members = [o.compress(it, 1, 6) for it in items]
    blob = b"".join(members)
    lens = [len(m) for m in members]
    offs = [sum(lens[:i]) for i in range(len(lens))]
    plains, st = pool.decompress_members(blob, offs, lens, [len(it) + 16 for it in items], flate_b200.GZIP)
    # the reference's own inflate rejects some streams its writer produces (see test_gpu_inflate): compare with the one-device path
    ctx = flate_b200.Context(0)
    want, wst, _ = ctx.decompress_members(blob, offs, lens, [len(it) + 16 for it in items], flate_b200.GZIP)
    assert st == list(wst) and plains == want
    assert all(p == it for p, it, s in zip(plains, items, st) if s == 0)
    pool.close()
    ctx.close()


def test_huffman_only_many_blocks_split_construction(ctx, o):
    """huffman-only streams of 64 blocks and more build their codes in three passes, the middle one from the eager
    package-merge (block_writer.cu): bytes equal to the oracle's for blocks of every kind of frequency shape -- text,
    uniform random, few symbols, skewed (binding length limit), runs, and blocks that end up stored."""
    from flate_b200 import synth
    rng = np.random.default_rng(2024)
    parts = [synth.enwik_like(3 << 20, seed=81),
             rng.integers(0, 256, 1 << 20, dtype=np.uint8),
             rng.integers(0, 3, 1 << 20, dtype=np.uint8),
             (rng.zipf(1.2, 1 << 20) % 256).astype(np.uint8),
             np.minimum(255, rng.geometric(0.02, 1 << 20)).astype(np.uint8),
             np.repeat(rng.integers(0, 256, 4096, dtype=np.uint8), 256),
             synth.random_zero_mix(2 << 20)]
    # a block whose frequencies double from symbol to symbol: the 15-bit limit binds
    skew = np.concatenate([np.full(1 << min(i, 14), i, dtype=np.uint8) for i in range(24)])
    parts.append(np.tile(skew, 20))
    data = np.concatenate(parts).tobytes()
    assert len(data) // 65535 >= 64
    for container in (0, 1):
        got = ctx.compress(data, container, 1)
        want = o.compress(data, container, 1)
        assert got == want, first_diff(got, want)
