"""ctypes binding of libflate_b200.so (include/flate_b200.h).  Fails loudly if the CUDA library is
missing: there is no CPU fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libflate_b200.so")

WRITE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t)
READ_FN = C.CFUNCTYPE(C.c_size_t, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t)

# every symbol include/flate_b200.h declares: name -> (restype, argtypes)
_P, _SZ, _I = C.c_void_p, C.c_size_t, C.c_int
_SZP = C.POINTER(C.c_size_t)
SIGNATURES = {
    "fb200_ctx_create": (_I, [_I, C.POINTER(_P)]),
    "fb200_ctx_destroy": (None, [_P]),
    "fb200_device_count": (_I, []),
    "fb200_strerror": (C.c_char_p, [_I]),
    "fb200_last_cuda_error": (C.c_char_p, []),
    "fb200_kernel_launches": (C.c_uint64, [_P]),
    "fb200_ctx_set_parse_mode": (_I, [_P, _I]),
    "fb200_sparse_fallbacks": (C.c_uint64, [_P]),
    "fb200_sparse_repairs": (C.c_uint64, [_P]),
    "fb200_profile_enable": (_I, [_P, _I]),
    "fb200_profile_phases": (_I, []),
    "fb200_profile_phase_name": (C.c_char_p, [_I]),
    "fb200_profile_read": (_I, [_P, _P, _P, _I]),
    "fb200_compress_bound": (_SZ, [_SZ, _I]),
    "fb200_compress": (_I, [_P, _I, _I, _P, _SZ, _P, _SZ, _SZP]),
    "fb200_decompress": (_I, [_P, _I, _P, _SZ, _P, _SZ, _SZP, _SZP]),
    "fb200_compress_device": (_I, [_P, _I, _I, _P, _SZ, _P, _SZ, _SZP, _P]),
    "fb200_decompress_members_device": (_I, [_P, _I, _P, _P, _P, _SZ, _P, _P, _P, _P, _P, _P, _P]),
    "fb200_decompress_members": (_I, [_P, _I, _P, _P, _P, _SZ, _P, _P, _P, _P, _P, _P]),
    "fb200_shard_align": (_SZ, []),
    "fb200_shard_overlap": (_SZ, []),
    "fb200_deflate_shard_search": (_I, [_P, _I, _P, _SZ, _SZ, _SZ, _P, _P]),
    "fb200_deflate_shard_finish": (_I, [_P, _I, _I, _P, _SZ, _P, _P, _SZ, _SZP, _P]),
    "fb200_simple_shard_plan": (_I, [_P, _I, _I, _P, _SZ, _I, _P, _P, _P, _P, _P]),
    "fb200_simple_shard_pack": (_I, [_P, C.c_uint64, _P, _SZ, _P, _P, _P, _P]),
    "fb200_crc32_combine": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint64]),
    "fb200_adler32_combine": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint64]),
    "fb200_deflate_create": (_I, [_P, _I, _I, WRITE_FN, _P, C.POINTER(_P)]),
    "fb200_deflate_write": (_I, [_P, _P, _SZ]),
    "fb200_deflate_flush": (_I, [_P]),
    "fb200_deflate_finish": (_I, [_P]),
    "fb200_deflate_set_writer": (None, [_P, WRITE_FN, _P]),
    "fb200_deflate_destroy": (None, [_P]),
    "fb200_inflate_create": (_I, [_P, _I, READ_FN, _P, C.POINTER(_P)]),
    "fb200_inflate_next": (_I, [_P, C.POINTER(_P), _SZP]),
    "fb200_inflate_get": (_I, [_P, _SZ, C.POINTER(_P), _SZP]),
    "fb200_inflate_read": (_I, [_P, _P, _SZ, _SZP]),
    "fb200_inflate_reset": (_I, [_P]),
    "fb200_inflate_set_reader": (None, [_P, READ_FN, _P]),
    "fb200_inflate_rebind": (None, [_P, READ_FN, _P]),
    "fb200_inflate_unused": (_I, [_P, C.POINTER(_P), _SZP]),
    "fb200_decompress_gzip_file": (_I, [_P, _P, _SZ, _P, _SZ, _SZP, _SZP, _SZP]),
    "fb200_pool_create": (_I, [C.c_uint64, C.POINTER(_P)]),
    "fb200_pool_devices": (_I, [_P]),
    "fb200_pool_destroy": (None, [_P]),
    "fb200_compress_batch": (_I, [_P, _I, _I, _SZ, _P, _P, _P, _P, _P, _P]),
    "fb200_decompress_members_batch": (_I, [_P, _I, _P, _P, _P, _SZ, _P, _P, _P, _P, _P, _P]),
    "fb200_inflate_destroy": (None, [_P]),
    "fb200_debug_tokens": (_I, [_P, _I, _P, _SZ, _P, _SZ, _SZP]),
    "fb200_debug_match_tables": (_I, [_P, _I, _P, _SZ, _P, _P]),
    "fb200_debug_block_write": (_I, [_P, _I, _P, _SZ, _I, _P, _SZ, _I, _P, _SZ, _SZP]),
}

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "flate_b200: %s is missing. Build it with `python -m flate_b200.build` (needs nvcc); "
            "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
