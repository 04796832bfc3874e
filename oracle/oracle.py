"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see flate_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The shipped package (flate_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libflate_oracle.so")
PUFF_PATH = os.path.join(HERE, "_ref", "libpuff.so")

RAW, GZIP, ZLIB = 0, 1, 2
STORE, HUFFMAN = 0, 1

ERRORS = ["Ok", "EndOfStream", "InvalidCode", "InvalidMatch", "InvalidBlockType", "WrongStoredBlockNlen",
          "InvalidDynamicBlockHeader", "OversubscribedHuffmanTree", "IncompleteHuffmanTree",
          "MissingEndOfBlockCode", "BadGzipHeader", "BadZlibHeader", "WrongGzipChecksum", "WrongGzipSize",
          "WrongZlibChecksum", "UnfinishedBits", "InvalidState", "NoSpaceLeft", "InvalidArgument"]


class OracleError(Exception):
    def __init__(self, code):
        self.code = code
        self.name = ERRORS[code] if 0 <= code < len(ERRORS) else "Unknown"
        super().__init__(self.name)


def build(native=False):
    """(Re)build the oracle shared library with gcc.  native=True adds -march=native into a
    separate file (used by bench.py on the box it is timed on)."""
    if native:
        out = os.path.join(HERE, "_build", "libflate_oracle_native.so")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-std=c11", "-shared", "-o", out,
                               os.path.join(HERE, "flate_oracle.c")])
        return out
    subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB_PATH


_lib = None


def lib(path=None):
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        build()
    L = C.CDLL(p)
    u8p, szp = C.c_void_p, C.POINTER(C.c_size_t)
    L.fo_compress.argtypes = [C.c_int, C.c_int, u8p, C.c_size_t, u8p, C.c_size_t, szp]
    L.fo_compress_bound.argtypes = [C.c_size_t]
    L.fo_compress_bound.restype = C.c_size_t
    L.fo_decompress.argtypes = [C.c_int, u8p, C.c_size_t, u8p, C.c_size_t, szp, szp]
    L.fo_decompress_hist.argtypes = [C.c_int, u8p, C.c_size_t, u8p, C.c_size_t, C.c_size_t, szp, szp]
    L.fo_deflate_create.argtypes = [C.c_int, C.c_int]
    L.fo_deflate_create.restype = C.c_void_p
    L.fo_deflate_write.argtypes = [C.c_void_p, u8p, C.c_size_t]
    L.fo_deflate_flush.argtypes = [C.c_void_p]
    L.fo_deflate_finish.argtypes = [C.c_void_p]
    L.fo_deflate_output.argtypes = [C.c_void_p, szp]
    L.fo_deflate_output.restype = C.c_void_p
    L.fo_deflate_take.argtypes = [C.c_void_p]
    L.fo_deflate_destroy.argtypes = [C.c_void_p]
    L.fo_tokenize.argtypes = [C.c_int, u8p, C.c_size_t, u8p, C.c_size_t, szp]
    L.fo_match_tables.argtypes = [C.c_int, u8p, C.c_size_t, u8p, u8p]
    L.fo_block_write.argtypes = [C.c_int, u8p, C.c_size_t, C.c_int, u8p, C.c_size_t, C.c_int, u8p, C.c_size_t, szp]
    L.fo_huffman_generate.argtypes = [u8p, C.c_int, C.c_int, u8p, u8p]
    L.fo_crc32.argtypes = [C.c_uint32, u8p, C.c_size_t]
    L.fo_crc32.restype = C.c_uint32
    L.fo_adler32.argtypes = [C.c_uint32, u8p, C.c_size_t]
    L.fo_adler32.restype = C.c_uint32
    if path is None:
        _lib = L
    return L


def _in(data):
    a = np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray, np.ndarray)) else data, dtype=np.uint8)
    return a, (a.ctypes.data if a.size else None)


def compress(data, container=RAW, mode=6, _lib_override=None):
    L = _lib_override or lib()
    a, ap = _in(data)
    cap = L.fo_compress_bound(a.size)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    rc = L.fo_compress(container, mode, ap, a.size, out.ctypes.data, cap, C.byref(n))
    if rc:
        raise OracleError(rc)
    return out[: n.value].tobytes()


def decompress(data, container=RAW, cap=None, hist=b""):
    """Returns (plain, consumed).  Raises OracleError with the reference's error class."""
    L = lib()
    a, ap = _in(data)
    if cap is None:
        cap = max(1 << 16, a.size * 64)
    buf = np.empty(len(hist) + cap, dtype=np.uint8)
    if hist:
        buf[: len(hist)] = np.frombuffer(hist, dtype=np.uint8)
    n, used = C.c_size_t(0), C.c_size_t(0)
    rc = L.fo_decompress_hist(container, ap, a.size, buf.ctypes.data + len(hist), len(hist), cap, C.byref(n),
                              C.byref(used))
    if rc:
        raise OracleError(rc)
    return buf[len(hist): len(hist) + n.value].tobytes(), used.value


class Deflate:
    """Streaming compressor: write / flush / finish, mirroring deflate.zig:304-371."""

    def __init__(self, container=RAW, mode=6):
        self.L = lib()
        self.h = self.L.fo_deflate_create(container, mode)
        if not self.h:
            raise OracleError(18)

    def _chk(self, rc):
        if rc:
            raise OracleError(rc)

    def write(self, data):
        a, ap = _in(data)
        self._chk(self.L.fo_deflate_write(self.h, ap, a.size))

    def flush(self):
        self._chk(self.L.fo_deflate_flush(self.h))

    def finish(self):
        self._chk(self.L.fo_deflate_finish(self.h))

    def output(self, take=False):
        n = C.c_size_t(0)
        p = self.L.fo_deflate_output(self.h, C.byref(n))
        b = C.string_at(p, n.value) if n.value else b""
        if take:
            self.L.fo_deflate_take(self.h)
        return b

    def __del__(self):
        if getattr(self, "h", None):
            self.L.fo_deflate_destroy(self.h)
            self.h = None


def tokenize(data, level=6):
    """Token list (u32 numpy array; literal = byte, match = 0x80000000 | (dist-1)<<8 | (len-3))."""
    L = lib()
    a, ap = _in(data)
    cap = a.size + 16
    out = np.empty(cap, dtype=np.uint32)
    n = C.c_size_t(0)
    rc = L.fo_tokenize(level, ap, a.size, out.ctypes.data, cap, C.byref(n))
    if rc:
        raise OracleError(rc)
    return out[: n.value].copy()


def match_tables(data, level=6):
    L = lib()
    a, ap = _in(data)
    rf = np.zeros(a.size, dtype=np.uint32)
    rq = np.zeros(a.size, dtype=np.uint32)
    rc = L.fo_match_tables(level, ap, a.size, rf.ctypes.data, rq.ctypes.data)
    if rc:
        raise OracleError(rc)
    return rf, rq


def block_write(kind, tokens, eof, input_bytes):
    """kind: 'wb' | 'dyn' | 'huff' (block_writer.zig:622-653)."""
    L = lib()
    k = {"wb": 0, "dyn": 1, "huff": 2}[kind]
    t = np.ascontiguousarray(tokens, dtype=np.uint32)
    has = input_bytes is not None
    a, ap = _in(input_bytes if has else b"")
    cap = 1 << 18
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    rc = L.fo_block_write(k, t.ctypes.data if t.size else None, t.size, int(eof), ap, a.size, int(has),
                          out.ctypes.data, cap, C.byref(n))
    if rc:
        raise OracleError(rc)
    return out[: n.value].tobytes()


def huffman_generate(freq, max_bits):
    L = lib()
    f = np.ascontiguousarray(freq, dtype=np.uint16)
    codes = np.zeros(f.size, dtype=np.uint16)
    lens = np.zeros(f.size, dtype=np.uint16)
    L.fo_huffman_generate(f.ctypes.data, f.size, max_bits, codes.ctypes.data, lens.ctypes.data)
    return codes, lens


def tok_lit(v):
    return v & 0xFF


def tok_match(dist, length):
    return 0x80000000 | ((dist - 1) << 8) | (length - 3)


_puff = None


def puff(data, cap=None):
    """Mark Adler's puff from the reference checkout (bin/puff/puff.c), built into oracle/_ref.
    Returns (rc, plain, consumed).  rc == 0 on success.  Raises FileNotFoundError if absent."""
    global _puff
    if _puff is None:
        if not os.path.exists(PUFF_PATH):
            raise FileNotFoundError(PUFF_PATH)
        _puff = C.CDLL(PUFF_PATH)
        _puff.puff.argtypes = [C.c_void_p, C.POINTER(C.c_ulong), C.c_void_p, C.POINTER(C.c_ulong)]
    a, ap = _in(data)
    if cap is None:
        cap = max(1 << 16, a.size * 64)
    out = np.empty(cap, dtype=np.uint8)
    dl, sl = C.c_ulong(cap), C.c_ulong(a.size)
    rc = _puff.puff(out.ctypes.data, C.byref(dl), ap, C.byref(sl))
    return rc, out[: dl.value].tobytes(), sl.value
