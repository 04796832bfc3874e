"""World-size-2 gloo tests (CPU) of the multi-rank host logic: unit sharding and the ragged
all-gather of per-shard outputs.  The per-shard payloads here come from the CPU oracle (this is a
test of the plumbing, not of the kernels)."""
import os
import socket
import zlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flate_b200 import sharding


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 9, 1024, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flate_b200 import synth
        from oracle import oracle as o
        nmembers = 5
        plains = [synth.enwik_like(20000 + 3000 * i, seed=50 + i).tobytes() for i in range(nmembers)]
        lo, hi = sharding.shard_range(nmembers, rank, world)
        mine = b"".join(o.compress(p, o.GZIP, 6) for p in plains[lo:hi])
        shard = torch.zeros(1 << 17, dtype=torch.uint8)
        shard[: len(mine)] = torch.from_numpy(np.frombuffer(mine, dtype=np.uint8).copy())
        buf, sizes, pad = sharding.all_gather_ragged(shard, len(mine))
        whole = sharding.concat_ragged(buf, sizes, pad)
        # every rank ends up with the same multi-member gzip stream, in member order
        d = zlib.decompressobj(31)
        out, rest = b"", whole
        while rest:
            d = zlib.decompressobj(31)
            out += d.decompress(rest)
            rest = d.unused_data
        q.put((rank, out == b"".join(plains), sizes))
    finally:
        dist.destroy_process_group()


def test_ragged_all_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2] and len(res[0][2]) == 2


# ---- one stream sharded by position: the driver logic with a stub context (no GPU) ----
class _StubCtx:
    """Stands in for flate_b200.Context: 'evaluates' position p as the value p + 1 for every third position
    of its range, and continues 5 entries into the next rank's range like the overlap of the sparse parse."""
    shard_align = 64
    shard_overlap = 16

    def __init__(self, fail_on=None):
        self.mode = 0
        self.fail_on = fail_on
        self.calls = []

    def set_parse_mode(self, mode):
        self.mode = mode

    def shard_search(self, d_in, n, lo, hi, d_nx, level=6, stream=None):
        self.calls.append((self.mode, lo, hi))
        nx = self.table
        if self.mode == 0:
            if self.fail_on is not None and lo <= self.fail_on < hi:
                return False
            nx[lo:min(n, hi + self.shard_overlap)] = -1
            for p in range(lo, hi, 3):
                nx[p] = p + 1
            for p in range(hi, min(n, hi + 5)):
                nx[p] = p + 1
        else:
            for p in range(lo, hi):
                nx[p] = p + 1
        return True

    def shard_finish(self, d_in, n, d_nx, d_out, cap, level=6, container=0, stream=None):
        self.final = self.table[:n].clone()
        return n


def _stream_worker(rank, world, port, q, fail_on):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 300
        ctx = _StubCtx(fail_on)
        real_empty = torch.empty

        def tracking_empty(*a, **k):   # the driver allocates the table itself: keep a handle on it
            t = real_empty(*a, **k)
            if k.get("dtype") == torch.int32 and t.numel() > 2 * _StubCtx.shard_overlap:
                ctx.table = t
            return t
        torch.empty = tracking_empty
        try:
            m = sharding.compress_stream_sharded(ctx, torch.zeros(n, dtype=torch.uint8), n, torch.zeros(16, dtype=torch.uint8))
        finally:
            torch.empty = real_empty
        final = ctx.final.tolist() if rank == 0 else None
        q.put((rank, m, final, ctx.calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fail_on", [None, 200])
def test_stream_sharding_driver_world2(fail_on):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stream_worker, args=(r, 2, port, q, fail_on)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    (r0, m0, final, calls0), (r1, m1, _, calls1) = res
    assert (m0, m1) == (300, 0)
    per, ranges = sharding.shard_positions(300, 2, 64)
    assert ranges == [(0, 192), (192, 300)]
    if fail_on is None:
        assert calls0 == [(0, 0, 192)] and calls1 == [(0, 192, 300)]
        for p in range(300):
            own = (p - (0 if p < 192 else 192)) % 3 == 0
            tail = 192 <= p < 197                       # rank 0's overlap entries, merged into rank 1's range
            assert final[p] == (p + 1 if own or tail else -1), p
    else:
        # one rank declined: both ranks repeat with the dense tables and the overlap merge is skipped
        assert calls0 == [(0, 0, 192), (1, 0, 192)] and calls1 == [(0, 192, 300), (1, 192, 300)]
        assert final == [p + 1 for p in range(300)]


def test_shard_positions_and_merge():
    per, ranges = sharding.shard_positions(1000, 3, 128)
    assert per == 384 and ranges == [(0, 384), (384, 768), (768, 1000)]
    per, ranges = sharding.shard_positions(100, 4, 64)
    assert ranges == [(0, 64), (64, 100), (100, 100), (100, 100)]
    own = torch.tensor([-1, 5, -1, 7], dtype=torch.int32)
    tail = torch.tensor([1, 5, -1], dtype=torch.int32)
    assert sharding.merge_overlap(own, tail).tolist() == [1, 5, -1, 7]


# ---- one huffman-only / store stream sharded by block ranges: the driver logic with a stub context (no GPU) ----
def _bits_to_bytes(bits, phase_bytes=0):
    out = bytearray(phase_bytes + (len(bits) + 7) // 8)
    for i, b in enumerate(bits):
        if b:
            out[phase_bytes + (i >> 3)] |= 1 << (i & 7)
    return bytes(out)


class _StubSimpleCtx:
    """Stands in for flate_b200.Context in compress_simple_sharded: a shard's "compressed form" is a fixed run of
    pseudo-random bits, optionally with one byte re-alignment in the middle (what a stored block does), so the joined
    stream is known in closed form."""

    def __init__(self, rank):
        rng = np.random.default_rng(100 + rank)
        self.pre = [int(b) for b in rng.integers(0, 2, 37 + 11 * rank)]
        self.has = rank % 2 == 0
        self.post = [int(b) for b in rng.integers(0, 2, 8 * (5 + rank))] if self.has else []

    def plain(self, nbytes):
        """The bytes this rank's shard stands for (only their checksum matters to the driver)."""
        return np.random.default_rng(500 + len(self.pre)).integers(0, 256, nbytes, dtype=np.uint8).tobytes()

    def simple_shard_plan(self, d_in, nbytes, is_last, mode=1, container=0, stream=None):
        import zlib
        sm = zlib.crc32(self.plain(nbytes)) if container == 1 else zlib.adler32(self.plain(nbytes)) if container == 2 else 0
        return len(self.pre), int(self.has), len(self.post), sm

    def crc32_combine(self, a, b, len2):
        from flate_b200 import _lib
        return int(_lib.load().fb200_crc32_combine(a, b, len2))

    def adler32_combine(self, a, b, len2):
        from flate_b200 import _lib
        return int(_lib.load().fb200_adler32_combine(a, b, len2))

    def stream_bits(self, x):
        bits = list(self.pre)
        if self.has:
            bits += [0] * ((-(x + len(bits))) % 8) + self.post
        return bits

    def simple_shard_pack(self, start_bit, d_out, cap, stream=None):
        lo = (start_bit >> 3) & ~15
        bits = [0] * (start_bit - 8 * lo) + self.stream_bits(start_bit)
        raw = _bits_to_bytes(bits)
        # like the library: the region is zeroed a little past the shard's end (here: 12 bytes), then the bits go in,
        # at the address the driver passes (its own copy of the stream, or a staging buffer when it runs alone)
        import ctypes
        assert d_out % 16 == 0 and len(raw) + 12 <= cap
        ctypes.memmove(d_out, raw + bytes(12), len(raw) + 12)
        return lo, len(raw), start_bit + len(self.stream_bits(start_bit))


def _simple_worker(rank, world, port, q, n=3 * 65535 + 17):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = _StubSimpleCtx(rank)
        ranges = sharding.simple_shard_ranges(n, world)
        lo, hi = ranges[rank]
        d_shard = torch.zeros(16, dtype=torch.uint8)
        final, total = sharding.compress_simple_sharded(ctx, d_shard, lo, hi, n, mode=1, container=0)
        # expected: the bit strings of the shards that hold anything (a rank without slices emits nothing, the last
        # rank always emits: it owns the final block), one after the other, each re-aligned where it says so
        bits = []
        for r in range(world):
            if ranges[r][1] > ranges[r][0] or r == world - 1:
                bits += _StubSimpleCtx(r).stream_bits(len(bits))
        want = _bits_to_bytes(bits)
        ok = final[:total].numpy().tobytes() == want
        # left distributed: this rank's copy is right in its own byte range and in the bytes shards share
        part, total2, (a, b) = sharding.compress_simple_sharded(ctx, d_shard, lo, hi, n, mode=1, container=0, gather=False)
        ok = ok and total2 == total and part[a:b].numpy().tobytes() == want[a:b]
        # gzip and zlib: header, the same bits behind it, and a footer from the combined per-shard checksums
        import zlib
        plain = b"".join(_StubSimpleCtx(r).plain(ranges[r][1] - ranges[r][0]) for r in range(world))
        for container, hdr, foot in ((1, sharding._HEADERS[1], zlib.crc32(plain).to_bytes(4, "little") + (n & 0xffffffff).to_bytes(4, "little")),
                                    (2, sharding._HEADERS[2], zlib.adler32(plain).to_bytes(4, "big"))):
            bits = []
            for r in range(world):
                if ranges[r][1] > ranges[r][0] or r == world - 1:
                    bits += _StubSimpleCtx(r).stream_bits(8 * len(hdr) + len(bits))
            got, tot = sharding.compress_simple_sharded(ctx, d_shard, lo, hi, n, mode=1, container=container)
            ok = ok and got[:tot].numpy().tobytes() == hdr + _bits_to_bytes(bits) + foot
        q.put((rank, ok, total))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 3 * 65535 + 17), (3, 3 * 65535 + 17), (3, 65535 + 17), (4, 100)])
def test_block_range_sharded_driver(world, n):
    """The driver of the block-range sharded stream over gloo with a stub context: shards of a few bytes that share
    boundary bytes (up to three shards in one byte), ranks without any slice (n = 65535 + 17 over 3 ranks leaves the
    middle rank empty; n = 100 over 4 ranks leaves everything to the last one)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_simple_worker, args=(r, world, port, q, n)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def test_shard_start_bits_realigns_after_stored_blocks():
    starts, end = sharding.shard_start_bits([(13, 0, 0), (5, 1, 80), (3, 0, 0)], 80)
    assert starts == [80, 93, 184] and end == 187
