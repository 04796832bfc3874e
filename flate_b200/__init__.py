"""flate_b200 -- B200-native DEFLATE engine behind the ianic/flate interface (gzip / zlib / raw).

The compute path is hand-written sm_100a CUDA in csrc/ behind the C ABI in include/flate_b200.h;
this package is the thin host mirror of the reference's public API.  No CPU fallback."""
from .api import (GZIP, HUFFMAN, RAW, STORE, ZLIB, Compressor, Context, Decompressor, ERRORS, FlateError, Level, Pool,
                  default_context, flate, gzip, zlib)

__all__ = ["Context", "Pool", "Compressor", "Decompressor", "FlateError", "ERRORS", "Level", "flate", "gzip", "zlib", "RAW",
           "GZIP", "ZLIB", "STORE", "HUFFMAN", "default_context"]
