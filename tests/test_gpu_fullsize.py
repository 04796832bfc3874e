"""Full-size checks at BASELINE.json's config sizes through size-independent properties
(encode -> independent decode round trips, equality on the device), where running the CPU oracle
would take too long.  bench.py additionally compares the whole 256 MiB level-6 output with the
oracle byte for byte in every run."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MIB = 1 << 20


@pytest.fixture(scope="module")
def ctx():
    import flate_b200
    c = flate_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def text256():
    from flate_b200 import synth
    piece = synth.enwik_like(32 * MIB, seed=0x5EED0001)
    return np.tile(piece, 8)  # matches never reach across pieces (window is 32 KiB): same work as unique text


def _device_compress(ctx, data, mode):
    import torch
    t_in = torch.from_numpy(data).cuda()
    cap = ctx.lib.fb200_compress_bound(data.size, mode) + 64
    t_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n = ctx.compress_device(t_in.data_ptr(), data.size, t_out.data_ptr(), cap, mode=mode,
                            stream=torch.cuda.current_stream().cuda_stream)
    return t_out[:n].cpu().numpy().tobytes()


@pytest.mark.parametrize("level", [6, 9])
def test_c2_c4_raw_deflate_256MiB_roundtrip(ctx, text256, level):
    """configs[1] (level 6) and configs[3] (level 9, deep chain walk): one 256 MiB raw stream."""
    out = _device_compress(ctx, text256, level)
    assert 2.4 < text256.size / len(out) < 3.2
    plain = zlib.decompress(out, -15)
    assert len(plain) == text256.size
    assert zlib.crc32(plain) == zlib.crc32(text256.tobytes())


def test_c2_gzip_footer_at_full_size(ctx, text256):
    out = ctx.compress(text256[:128 * MIB], 1, 6)
    want_crc = zlib.crc32(text256[:128 * MIB].tobytes())
    assert int.from_bytes(out[-8:-4], "little") == want_crc
    assert int.from_bytes(out[-4:], "little") == 128 * MIB


def test_host_path_slab_overlap_equals_device_path(ctx, text256):
    """fb200_compress copies the input in 32 MiB slabs and searches slab k while slab k+1 is in flight;
    the result must not depend on that."""
    n = 96 * MIB + 12345
    data = text256[:n]
    assert ctx.compress(data, 0, 6) == _device_compress(ctx, data, 6)


@pytest.mark.parametrize("n", [32 * MIB + 8192 + 1, 40 * MIB + 32768, 41 * MIB - 1, 72 * MIB + 1551])
def test_host_path_sizes_around_slab_and_chunk_boundaries(ctx, text256, n):
    """The sparse parse follows the copy front by its look-ahead and in 32 KiB chunks: no size may matter."""
    data = text256[:n]
    assert ctx.compress(data, 0, 6) == _device_compress(ctx, data, 6)


def test_host_path_repairs_periodic_stretch_like_device_path(text256):
    """Coverage check fails inside the periodic stretches: the repair (and the second parse) must also work when
    the input arrived slab-wise and the output leaves part by part."""
    import flate_b200
    data = text256[:48 * MIB].copy()
    data[5 * MIB:5 * MIB + 300000] = 0
    data[37 * MIB:37 * MIB + 70000] = np.tile(np.arange(13, dtype=np.uint8), 5385)[:70000]
    host, dev, dense = flate_b200.Context(0), flate_b200.Context(0), flate_b200.Context(0)
    dense.set_parse_mode(1)
    try:
        a = host.compress(data, 0, 6)
        b = _device_compress(dev, data, 6)
        c = dense.compress(data, 0, 6)
        assert a == c and b == c
        assert host.sparse_repairs >= 1 and host.sparse_fallbacks == 0
        assert dev.sparse_repairs >= 1 and dev.sparse_fallbacks == 0
        assert zlib.decompress(a, -15) == data.tobytes()
    finally:
        host.close()
        dev.close()
        dense.close()


def test_c5_huffman_only_1GiB_roundtrip(ctx):
    """configs[4] shape (random + zeros mix, stored and dynamic blocks alternating), 1 GiB per GPU."""
    from flate_b200 import synth
    piece = synth.random_zero_mix(256 * MIB)
    data = np.tile(piece, 4)
    out = _device_compress(ctx, data, 1)
    d = zlib.decompressobj(-15)
    crc, total = 0, 0
    view = memoryview(out)
    for i in range(0, len(out), 64 * MIB):
        chunk = d.decompress(view[i:i + 64 * MIB])
        crc = zlib.crc32(chunk, crc)
        total += len(chunk)
    tail = d.flush()
    crc = zlib.crc32(tail, crc)
    total += len(tail)
    assert total == data.size
    assert crc == zlib.crc32(data.tobytes())


def test_c3_inflate_1GiB_of_members_on_device(ctx, text256):
    """configs[2] shape: 1024 gzip members x 1 MiB, inflated in one launch, compared on the device."""
    import torch
    import flate_b200
    uniq = 64
    members, plains = [], []
    for i in range(uniq):
        for shift in range(64):
            lo = (i * MIB + shift * 4099) % (text256.size - MIB + 1)
            m = ctx.compress(text256[lo:lo + MIB], flate_b200.GZIP, 6)
            try:  # the reference rejects code-length runs crossing the lit/dist boundary
                ctx.decompress(m, flate_b200.GZIP, cap=MIB + 64)
                break
            except flate_b200.FlateError:
                continue
        members.append(m)
        plains.append(text256[lo:lo + MIB])
    nmem = 1024
    blob = b"".join(members[i % uniq] for i in range(nmem))
    lens = np.array([len(members[i % uniq]) for i in range(nmem)], dtype=np.uint64)
    offs = np.zeros(nmem, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)[:-1]
    d_blob = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).cuda()
    d_plain = torch.zeros(nmem * MIB + 64, dtype=torch.uint8, device="cuda")
    ooff = np.arange(nmem, dtype=np.uint64) * np.uint64(MIB)
    ocap = np.full(nmem, MIB, dtype=np.uint64)
    rc, ol, used, st = ctx.decompress_members_device(d_blob.data_ptr(), offs, lens, d_plain.data_ptr(), ooff, ocap,
                                                     flate_b200.GZIP)
    assert rc == 0 and (st == 0).all() and (ol == MIB).all() and (used == lens).all()
    want = torch.from_numpy(np.concatenate(plains)).cuda()
    got = d_plain[: nmem * MIB].view(nmem // uniq, uniq * MIB)
    assert bool((got == want.unsqueeze(0)).all())


def test_streaming_compressor_equals_oracle_and_one_shot_at_size(ctx, text256, monkeypatch):
    """SURVEY.md section 8f rank 2 at size: 40 MiB of text + a tar-like tail through Compressor.write in 1 MiB pieces with
    4 MiB parts (ten parts, window slides, a flush in the middle) equals the CPU oracle's streaming Deflate byte for byte;
    and 200 MiB with the default part size (64 MiB: three parts) equals the one-shot stream of the same bytes."""
    import io
    import sys
    import flate_b200
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    import bench
    from oracle import oracle as o
    data = np.concatenate([text256[:40 * MIB], bench.make_tar_like(8 * MIB)]).tobytes()
    monkeypatch.setenv("FB200_STREAM_PART", "4096")
    w = io.BytesIO()
    c = flate_b200.Compressor(1, w, 6, ctx=ctx)
    d = o.Deflate(1, 6)
    cut = 29 * MIB + 17
    for a, b in ((0, cut), (cut, len(data))):
        for p in range(a, b, MIB):
            c.write(data[p:min(p + MIB, b)])
        d.write(data[a:b])
        if b != len(data):
            c.flush()
            d.flush()
    c.finish()
    d.finish()
    assert w.getvalue() == d.output()
    c.close()
    monkeypatch.delenv("FB200_STREAM_PART")
    big = text256[:200 * MIB]
    w = io.BytesIO()
    c = flate_b200.Compressor(0, w, 6, ctx=ctx)
    for p in range(0, big.size, 4 * MIB):
        c.write(big[p:p + 4 * MIB])
    c.finish()
    want = _device_compress(ctx, big, 6)
    assert w.getvalue() == want
    c.close()
    # the same bytes in ONE write() call (larger than the window: the call itself has to run parts and slide)
    w = io.BytesIO()
    c = flate_b200.Compressor(0, w, 6, ctx=ctx)
    c.write(big)
    c.finish()
    assert w.getvalue() == want
    c.close()
