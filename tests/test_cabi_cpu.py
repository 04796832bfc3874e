"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/flate_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "flate_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fb200_[a-z0-9_]+)\s*\(", src)) - {"fb200_write_fn", "fb200_read_fn"})


def test_library_exports_every_declared_symbol():
    from flate_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 28
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)


def test_error_names_match_reference_error_set():
    from flate_b200 import _lib, api
    lib = _lib.load()
    # inflate.zig:72-78, huffman_decoder.zig:35-40, container.zig:45-51, bit_writer.zig:35, inflate.zig:302-304
    want = ["Ok", "EndOfStream", "InvalidCode", "InvalidMatch", "InvalidBlockType", "WrongStoredBlockNlen",
            "InvalidDynamicBlockHeader", "OversubscribedHuffmanTree", "IncompleteHuffmanTree", "MissingEndOfBlockCode",
            "BadGzipHeader", "BadZlibHeader", "WrongGzipChecksum", "WrongGzipSize", "WrongZlibChecksum",
            "UnfinishedBits", "InvalidState"]
    for code, name in enumerate(want):
        assert lib.fb200_strerror(code).decode() == name
        if code:
            assert api.ERRORS[code].__name__ == name


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import flate_b200
    with pytest.raises(flate_b200.FlateError) as ei:
        flate_b200.Context(0)
    assert type(ei.value).__name__ == "NoDevice"


def test_compress_bound_is_pure_host_arithmetic():
    from flate_b200 import _lib
    lib = _lib.load()
    for n in (0, 1, 65535, 1 << 20, 1 << 28):
        assert lib.fb200_compress_bound(n, 6) >= n + 5 * (n // 65535 + 1)


def test_product_never_touches_the_oracle():
    """The shipped package must not import, link or open anything under oracle/."""
    pkg = os.path.join(ROOT, "flate_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, f)
    out = os.popen("ldd %s" % os.path.join(pkg, "lib", "libflate_b200.so")).read()
    assert "oracle" not in out


def test_synth_is_deterministic():
    from flate_b200 import synth
    a = synth.enwik_like(100000)
    b = synth.enwik_like(100000)
    assert (a == b).all() and a.size == 100000
    import zlib
    ratio = a.size / len(zlib.compress(a.tobytes(), 6))
    assert 2.2 < ratio < 3.6
    m = synth.random_zero_mix(1 << 20)
    assert m.size == 1 << 20 and (m == 0).mean() > 0.2


def test_zig_binding_declares_every_symbol():
    """zig/flate_b200.zig is the reference-side binding (not compiled here: no zig toolchain): its extern block must
    name exactly the symbols of the header, and the shim must keep the reference's public names (src/gzip.zig:5-66)."""
    src = open(os.path.join(ROOT, "zig", "flate_b200.zig")).read()
    externs = sorted(set(re.findall(r'pub extern "c" fn (fb200_[a-z0-9_]+)\(', src)))
    assert externs == declared_symbols(), set(externs) ^ set(declared_symbols())
    for name in ("pub fn compress(", "pub fn compressor(", "pub fn Compressor(", "pub fn decompress(", "pub fn decompressor(",
                 "pub fn Inflate(", "pub const huffman", "pub const store", "pub fn flush(", "pub fn finish(", "pub fn setWriter(",
                 "pub fn next(", "pub fn get(", "pub fn read(", "pub fn reader(", "pub fn reset(", "pub fn setReader(",
                 "pub const flate = ", "pub const gzip = ", "pub const zlib = "):
        assert name in src, name
    assert "..." not in re.sub(r"//.*", "", src)   # no elided bodies


def test_checksum_combiners_are_pure_host_arithmetic():
    """fb200_crc32_combine / fb200_adler32_combine (used for block-range shards, streaming parts and piecewise inflate):
    the checksum of a concatenation from the checksums of its pieces, against zlib on random splits; and the pool and
    streaming entry points refuse to work without a device instead of falling back."""
    import zlib
    import numpy as np
    from flate_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(11)
    data = rng.integers(0, 256, 300000, dtype=np.uint8).tobytes()
    for _ in range(50):
        a, b = sorted(int(x) for x in rng.integers(0, len(data) + 1, 2))
        left, mid = data[:a], data[a:b]
        assert lib.fb200_crc32_combine(zlib.crc32(left), zlib.crc32(mid), len(mid)) == zlib.crc32(data[:b])
        assert lib.fb200_adler32_combine(zlib.adler32(left), zlib.adler32(mid), len(mid)) == zlib.adler32(data[:b])
    assert lib.fb200_crc32_combine(zlib.crc32(b"abc"), 0, 0) == zlib.crc32(b"abc")
    if lib.fb200_device_count() == 0:
        h = C.c_void_p()
        assert lib.fb200_pool_create(0, C.byref(h)) == 20          # FB200_NO_DEVICE
        assert lib.fb200_pool_devices(None) == 0
