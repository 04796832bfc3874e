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
