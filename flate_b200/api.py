"""Host-side mirror of the reference's public interface (src/flate.zig, src/gzip.zig, src/zlib.zig)
over the C ABI.  Same names, argument meaning and error behaviour:

    flate_b200.gzip.compress(reader, writer, level=...)        src/gzip.zig:23
    flate_b200.gzip.compressor(writer, level=...) -> Compressor   src/gzip.zig:30  (.write/.flush/.finish/.set_writer)
    flate_b200.gzip.decompress(reader, writer)                 src/gzip.zig:5
    flate_b200.gzip.decompressor(reader) -> Decompressor       src/gzip.zig:13 (.decompress/.next/.get/.read/.reset)
    flate_b200.gzip.huffman.compress / .compressor, .store.compress / .compressor   src/gzip.zig:38-66
  and the same under flate_b200.zlib and flate_b200.flate (raw).

`reader` is anything with .read(n) (or bytes), `writer` anything with .write(b).  Errors are
exceptions named after the Zig error set (FlateError subclasses).
"""
import ctypes as C
import io

import numpy as np

from . import _lib

RAW, GZIP, ZLIB = 0, 1, 2
STORE, HUFFMAN = 0, 1


class Level:  # deflate.zig:23-32
    fast = 4
    level_4 = 4
    level_5 = 5
    default = 6
    level_6 = 6
    level_7 = 7
    level_8 = 8
    best = 9
    level_9 = 9


class FlateError(Exception):
    code = -1


_ERROR_NAMES = ["Ok", "EndOfStream", "InvalidCode", "InvalidMatch", "InvalidBlockType", "WrongStoredBlockNlen",
                "InvalidDynamicBlockHeader", "OversubscribedHuffmanTree", "IncompleteHuffmanTree",
                "MissingEndOfBlockCode", "BadGzipHeader", "BadZlibHeader", "WrongGzipChecksum", "WrongGzipSize",
                "WrongZlibChecksum", "UnfinishedBits", "InvalidState", "NoSpaceLeft", "InvalidArgument", "CudaError",
                "NoDevice", "RetryDense"]
ERRORS = {}
for _i, _n in enumerate(_ERROR_NAMES):
    if _i:
        ERRORS[_i] = type(_n, (FlateError,), {"code": _i})
        globals()[_n] = ERRORS[_i]


def _check(rc):
    if rc:
        cls = ERRORS.get(rc, FlateError)
        msg = cls.__name__
        if rc == 19:
            msg += ": " + _lib.load().fb200_last_cuda_error().decode()
        raise cls(msg)


class Context:
    """One per GPU (fb200_ctx).  Owns the device workspace; not safe for concurrent use."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        _check(self.lib.fb200_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.fb200_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def kernel_launches(self):
        return int(self.lib.fb200_kernel_launches(self.h))

    def set_parse_mode(self, mode):
        """0 = speculative sparse parse with automatic dense redo (default), 1 = dense match tables always."""
        _check(self.lib.fb200_ctx_set_parse_mode(self.h, int(mode)))

    @property
    def sparse_fallbacks(self):
        return int(self.lib.fb200_sparse_fallbacks(self.h))

    @property
    def sparse_repairs(self):
        return int(self.lib.fb200_sparse_repairs(self.h))

    def profile(self, on=True):
        self.lib.fb200_profile_enable(self.h, int(on))

    def profile_read(self):
        """{phase name: (total ms, launches)} accumulated since profile(True)."""
        n = self.lib.fb200_profile_phases()
        ms = np.zeros(n, dtype=np.float64)
        cnt = np.zeros(n, dtype=np.uint64)
        self.lib.fb200_profile_read(self.h, ms.ctypes.data, cnt.ctypes.data, n)
        return {self.lib.fb200_profile_phase_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    # ---- one-shot, host buffers ----
    def compress(self, data, container=RAW, mode=Level.default):
        a = _as_u8(data)
        cap = self.lib.fb200_compress_bound(a.size, mode) + 64
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        _check(self.lib.fb200_compress(self.h, container, mode, _ptr(a), a.size, out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].tobytes()

    def decompress(self, data, container=RAW, cap=None):
        """One member.  Returns (plain, consumed)."""
        a = _as_u8(data)
        if cap is None:
            cap = max(1 << 16, a.size * 8)
        while True:
            out = np.empty(cap, dtype=np.uint8)
            n, used = C.c_size_t(0), C.c_size_t(0)
            rc = self.lib.fb200_decompress(self.h, container, _ptr(a), a.size, out.ctypes.data, cap, C.byref(n),
                                           C.byref(used))
            if rc == 17 and cap < (a.size + 64) * 1100:  # NoSpaceLeft: our output buffer, not the stream
                cap *= 4
                continue
            _check(rc)
            return out[: n.value].tobytes(), used.value

    def decompress_gzip_file(self, data, cap=None):
        """A multi-member gzip file without a member index: members are found and inflated in parallel, the result is
        the sequential loop's (decompress / reset per member).  Returns (plain, consumed, members)."""
        a = _as_u8(data)
        if cap is None:
            cap = max(1 << 16, a.size * 8)
        while True:
            out = np.empty(cap, dtype=np.uint8)
            n, used, mem = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
            rc = self.lib.fb200_decompress_gzip_file(self.h, _ptr(a), a.size, out.ctypes.data, cap, C.byref(n), C.byref(used), C.byref(mem))
            if rc == 17 and cap < (a.size + 64) * 1100:  # NoSpaceLeft: our buffer; n.value is the size the file claims
                cap = max(cap * 4, n.value + 64)
                continue
            _check(rc)
            return out[:n.value].tobytes(), used.value, mem.value

    def decompress_members(self, data, in_off, in_len, out_cap, container=GZIP):
        """k independent members in one launch.  Returns (list of plain bytes, status list)."""
        a = _as_u8(data)
        k = len(in_off)
        io_ = np.asarray(in_off, dtype=np.uint64)
        il = np.asarray(in_len, dtype=np.uint64)
        oc = np.asarray(out_cap, dtype=np.uint64)
        oo = np.zeros(k, dtype=np.uint64)
        if k:
            oo[1:] = np.cumsum(oc)[:-1]
        out = np.empty(int(oc.sum()) + 1, dtype=np.uint8)
        ol = np.zeros(k, dtype=np.uint64)
        used = np.zeros(k, dtype=np.uint64)
        st = np.zeros(k, dtype=np.int32)
        self.lib.fb200_decompress_members(self.h, container, _ptr(a), io_.ctypes.data, il.ctypes.data, k, out.ctypes.data,
                                          oo.ctypes.data, oc.ctypes.data, ol.ctypes.data, used.ctypes.data, st.ctypes.data)
        return [out[int(oo[i]): int(oo[i] + ol[i])].tobytes() for i in range(k)], st.tolist(), used.tolist()

    # ---- device-resident (pointers are raw CUDA device addresses, e.g. torch.Tensor.data_ptr()) ----
    def compress_device(self, d_in, n, d_out, cap, mode=Level.default, container=RAW, stream=None):
        out_len = C.c_size_t(0)
        _check(self.lib.fb200_compress_device(self.h, container, mode, d_in, n, d_out, cap, C.byref(out_len), stream))
        return out_len.value

    def decompress_members_device(self, d_in, in_off, in_len, d_out, out_off, out_cap, container=GZIP, stream=None):
        k = len(in_off)
        arrs = [np.ascontiguousarray(x, dtype=np.uint64) for x in (in_off, in_len, out_off, out_cap)]
        ol = np.zeros(k, dtype=np.uint64)
        used = np.zeros(k, dtype=np.uint64)
        st = np.zeros(k, dtype=np.int32)
        rc = self.lib.fb200_decompress_members_device(self.h, container, d_in, arrs[0].ctypes.data, arrs[1].ctypes.data, k,
                                                      d_out, arrs[2].ctypes.data, arrs[3].ctypes.data, ol.ctypes.data,
                                                      used.ctypes.data, st.ctypes.data, stream)
        return rc, ol, used, st

    # ---- one stream sharded by position over several GPUs ----
    def shard_search(self, d_in, n, lo, hi, d_nx, level=Level.default, stream=None):
        """Stage 1 of the position-sharded stream.  Returns False when the sparse parse cannot vouch for its
        coverage (every rank must then repeat the call after set_parse_mode(1))."""
        rc = self.lib.fb200_deflate_shard_search(self.h, level, d_in, n, lo, hi, d_nx, stream)
        if rc == 21:
            return False
        _check(rc)
        return True

    @property
    def shard_align(self):
        return int(self.lib.fb200_shard_align())

    @property
    def shard_overlap(self):
        return int(self.lib.fb200_shard_overlap())

    def shard_finish(self, d_in, n, d_nx, d_out, cap, level=Level.default, container=RAW, stream=None):
        out_len = C.c_size_t(0)
        _check(self.lib.fb200_deflate_shard_finish(self.h, container, level, d_in, n, d_nx, d_out, cap, C.byref(out_len),
                                                   stream))
        return out_len.value

    # ---- huffman-only / store stream sharded by 65535-byte block ranges over several GPUs ----
    def simple_shard_plan(self, d_in, shard_bytes, is_last, mode=HUFFMAN, container=RAW, stream=None):
        """Stage 1: returns (pre_bits, has_stored, post_bits, checksum) of this rank's range of slices."""
        pre, post, has, sm = C.c_uint64(0), C.c_uint64(0), C.c_int(0), C.c_uint32(0)
        _check(self.lib.fb200_simple_shard_plan(self.h, container, mode, d_in, shard_bytes, int(bool(is_last)), C.byref(pre),
                                                C.byref(has), C.byref(post), C.byref(sm), stream))
        return pre.value, has.value, post.value, sm.value

    def simple_shard_pack(self, start_bit, d_out, cap, stream=None):
        """Stage 2: packs the planned shard at its stream bit offset.  Returns (byte_lo, nbytes, end_bit): d_out[0:nbytes]
        are the stream's bytes from byte_lo on, zero before start_bit."""
        lo, nb, end = C.c_uint64(0), C.c_size_t(0), C.c_uint64(0)
        _check(self.lib.fb200_simple_shard_pack(self.h, start_bit, d_out, cap, C.byref(lo), C.byref(nb), C.byref(end), stream))
        return lo.value, nb.value, end.value

    def crc32_combine(self, crc1, crc2, len2):
        return int(self.lib.fb200_crc32_combine(crc1, crc2, len2))

    def adler32_combine(self, a1, a2, len2):
        return int(self.lib.fb200_adler32_combine(a1, a2, len2))

    # ---- test seams ----
    def debug_tokens(self, data, level=Level.default):
        a = _as_u8(data)
        cap = a.size + 16
        out = np.empty(cap, dtype=np.uint32)
        n = C.c_size_t(0)
        _check(self.lib.fb200_debug_tokens(self.h, level, _ptr(a), a.size, out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].copy()

    def debug_match_tables(self, data, level=Level.default):
        a = _as_u8(data)
        rf = np.zeros(a.size, dtype=np.uint32)
        rq = np.zeros(a.size, dtype=np.uint32)
        _check(self.lib.fb200_debug_match_tables(self.h, level, _ptr(a), a.size, rf.ctypes.data, rq.ctypes.data))
        return rf, rq

    def debug_block_write(self, kind, tokens, eof, input_bytes):
        k = {"wb": 0, "dyn": 1, "huff": 2}[kind]
        t = np.ascontiguousarray(tokens, dtype=np.uint32)
        has = input_bytes is not None
        a = _as_u8(input_bytes if has else b"")
        cap = t.size * 8 + a.size * 2 + 8192
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        _check(self.lib.fb200_debug_block_write(self.h, k, t.ctypes.data if t.size else None, t.size, int(eof), _ptr(a),
                                                a.size, int(has), out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].tobytes()


def _as_u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data if a.size else None


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _read_all(reader):
    if isinstance(reader, (bytes, bytearray, memoryview)):
        return bytes(reader)
    chunks = []
    while True:
        b = reader.read(1 << 20)
        if not b:
            break
        chunks.append(b)
    return b"".join(chunks)


class Pool:
    """Several GPUs of one node behind one call (fb200_pool_*): k independent streams, or the members of a multi-member
    file, spread over the devices by the library (one context and one host thread per device)."""

    def __init__(self, device_mask=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        _check(self.lib.fb200_pool_create(device_mask, C.byref(h)))
        self.h = h

    @property
    def devices(self):
        return int(self.lib.fb200_pool_devices(self.h))

    def compress_batch(self, items, container=RAW, mode=Level.default):
        """items: list of bytes-like.  Returns the list of compressed streams (the reference's compress() per item)."""
        arrs = [_as_u8(x) for x in items]
        k = len(arrs)
        caps = [int(self.lib.fb200_compress_bound(a.size, mode)) for a in arrs]
        outs = [np.empty(c, dtype=np.uint8) for c in caps]
        inp = (C.c_void_p * k)(*[_ptr(a) for a in arrs])
        inl = (C.c_size_t * k)(*[a.size for a in arrs])
        outp = (C.c_void_p * k)(*[o.ctypes.data for o in outs])
        outc = (C.c_size_t * k)(*caps)
        outl = (C.c_size_t * k)()
        st = (C.c_int * k)()
        _check(self.lib.fb200_compress_batch(self.h, container, mode, k, inp, inl, outp, outc, outl, st))
        return [outs[i][:outl[i]].tobytes() for i in range(k)]

    def decompress_members(self, data, in_off, in_len, out_cap, container=GZIP):
        """The members of one buffer, split over the devices by contiguous ranges.  Returns (plains, statuses)."""
        a = _as_u8(data)
        k = len(in_off)
        io_ = np.asarray(in_off, dtype=np.uint64)
        il = np.asarray(in_len, dtype=np.uint64)
        oc = np.asarray(out_cap, dtype=np.uint64)
        oo = np.zeros(k, dtype=np.uint64)
        if k:
            oo[1:] = np.cumsum(oc)[:-1]
        out = np.empty(int(oc.sum()) + 1, dtype=np.uint8)
        ol = np.zeros(k, dtype=np.uint64)
        used = np.zeros(k, dtype=np.uint64)
        st = np.zeros(k, dtype=np.int32)
        self.lib.fb200_decompress_members_batch(self.h, container, _ptr(a), io_.ctypes.data, il.ctypes.data, k, out.ctypes.data,
                                                oo.ctypes.data, oc.ctypes.data, ol.ctypes.data, used.ctypes.data, st.ctypes.data)
        return [out[int(oo[i]):int(oo[i] + ol[i])].tobytes() for i in range(k)], [int(x) for x in st]

    def close(self):
        if getattr(self, "h", None):
            self.lib.fb200_pool_destroy(self.h)
            self.h = None

    __del__ = close


class Compressor:
    """deflate.zig:121-373 Deflate / :449-529 SimpleCompressor behind fb200_deflate_*."""

    def __init__(self, container, writer, mode=Level.default, ctx=None):
        self.ctx = ctx or default_context()
        self.lib = self.ctx.lib
        self._writer = writer
        self._cb = _lib.WRITE_FN(self._on_write)
        h = C.c_void_p()
        _check(self.lib.fb200_deflate_create(self.ctx.h, container, mode, self._cb, None, C.byref(h)))
        self.h = h

    def _on_write(self, _user, data, n):
        try:
            self._writer.write(C.string_at(data, n))
            return 0
        except Exception:  # the writer's error propagates as a failed write
            return 1

    def compress(self, reader):  # deflate.zig:304
        while True:
            b = reader.read(1 << 20) if hasattr(reader, "read") else reader
            if not b:
                break
            self.write(b)
            if not hasattr(reader, "read"):
                break

    def write(self, data):  # deflate.zig:363
        a = _as_u8(data)
        _check(self.lib.fb200_deflate_write(self.h, _ptr(a), a.size))
        return a.size

    def writer(self):  # deflate.zig:369
        return self

    def flush(self):  # deflate.zig:335
        _check(self.lib.fb200_deflate_flush(self.h))

    def finish(self):  # deflate.zig:344
        _check(self.lib.fb200_deflate_finish(self.h))

    def set_writer(self, writer):  # deflate.zig:351
        self._writer = writer

    def close(self):
        if getattr(self, "h", None):
            self.lib.fb200_deflate_destroy(self.h)
            self.h = None

    __del__ = close


class Decompressor:
    """inflate.zig:43-355 Inflate behind fb200_inflate_*."""

    def __init__(self, container, reader, ctx=None):
        self.ctx = ctx or default_context()
        self.lib = self.ctx.lib
        self._reader = io.BytesIO(reader) if isinstance(reader, (bytes, bytearray)) else reader
        self._cb = _lib.READ_FN(self._on_read)
        h = C.c_void_p()
        _check(self.lib.fb200_inflate_create(self.ctx.h, container, self._cb, None, C.byref(h)))
        self.h = h

    def _on_read(self, _user, buf, cap):
        b = self._reader.read(cap)
        if not b:
            return 0
        C.memmove(buf, b, len(b))
        return len(b)

    def get(self, limit=0):  # inflate.zig:326
        p, n = C.c_void_p(), C.c_size_t(0)
        _check(self.lib.fb200_inflate_get(self.h, limit, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value) if n.value else b""

    def next(self):  # inflate.zig:313
        b = self.get(0)
        return b if b else None

    def decompress(self, writer):  # inflate.zig:292
        while True:
            b = self.next()
            if b is None:
                break
            writer.write(b)

    def read(self, n):  # inflate.zig:345
        return self.get(n) if n else b""

    def reader(self):  # inflate.zig:351
        return self

    def reset(self):  # inflate.zig:301
        _check(self.lib.fb200_inflate_reset(self.h))

    def unread_bytes(self):
        """Bytes pulled from the reader past the end of the current member (the reader is read a chunk at a time)."""
        p, n = C.c_void_p(), C.c_size_t(0)
        _check(self.lib.fb200_inflate_unused(self.h, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value) if n.value else b""

    def set_reader(self, reader):  # inflate.zig:283
        self._reader = io.BytesIO(reader) if isinstance(reader, (bytes, bytearray)) else reader

    def close(self):
        if getattr(self, "h", None):
            self.lib.fb200_inflate_destroy(self.h)
            self.h = None

    __del__ = close


class _Simple:
    def __init__(self, container, mode):
        self._c, self._m = container, mode

    def compress(self, reader, writer, ctx=None):
        c = self.compressor(writer, ctx=ctx)
        c.compress(reader)
        c.finish()

    def compressor(self, writer, ctx=None):
        return Compressor(self._c, writer, self._m, ctx=ctx)


class _ContainerModule:
    """One of src/flate.zig (raw), src/gzip.zig, src/zlib.zig."""

    def __init__(self, container):
        self.container = container
        self.huffman = _Simple(container, HUFFMAN)
        self.store = _Simple(container, STORE)

    def compress(self, reader, writer, level=Level.default, ctx=None):
        c = self.compressor(writer, level=level, ctx=ctx)
        c.compress(reader)
        c.finish()

    def compressor(self, writer, level=Level.default, ctx=None):
        return Compressor(self.container, writer, level, ctx=ctx)

    def decompress(self, reader, writer, ctx=None):
        Decompressor(self.container, reader, ctx=ctx).decompress(writer)

    def decompressor(self, reader, ctx=None):
        return Decompressor(self.container, reader, ctx=ctx)

    # bytes -> bytes conveniences
    def compress_bytes(self, data, level=Level.default, ctx=None):
        return (ctx or default_context()).compress(data, self.container, level)

    def decompress_bytes(self, data, ctx=None):
        return (ctx or default_context()).decompress(data, self.container)[0]


flate = _ContainerModule(RAW)
gzip = _ContainerModule(GZIP)
zlib = _ContainerModule(ZLIB)
