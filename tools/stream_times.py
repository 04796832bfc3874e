"""Times the streaming compressor (fb200_deflate_write in 1 MiB pieces from pinned memory) for the part size given by
FB200_STREAM_PART (KiB); development aid."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import flate_b200  # noqa: E402
from flate_b200 import _lib as fb_lib, synth  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = flate_b200.Context(0)
lib = ctx.lib
n = mib << 20
h_in = torch.from_numpy(synth.enwik_like(n, seed=19)).pin_memory()
sink = torch.zeros(n * passes // 2 + (1 << 20), dtype=torch.uint8)
sink_ptr, sink_len, cb_time = sink.data_ptr(), [0], [0.0]


def on_write(_user, data, nbytes):
    t = time.perf_counter()
    C.memmove(sink_ptr + sink_len[0], data, nbytes)
    sink_len[0] += nbytes
    cb_time[0] += time.perf_counter() - t
    return 0


cb = fb_lib.WRITE_FN(on_write)


def once(p):
    sink_len[0] = 0
    cb_time[0] = 0.0
    h = C.c_void_p()
    assert lib.fb200_deflate_create(ctx.h, 0, 6, cb, None, C.byref(h)) == 0
    for _ in range(p):
        for pos in range(0, n, 1 << 20):
            assert lib.fb200_deflate_write(h, h_in.data_ptr() + pos, min(1 << 20, n - pos)) == 0
    assert lib.fb200_deflate_finish(h) == 0
    lib.fb200_deflate_destroy(h)
    return sink_len[0]


once(1)
t0 = time.perf_counter()
got = once(passes)
dt = time.perf_counter() - t0
print("part %s KiB: %d MiB in %.1f ms = %.1f MB/s (writer callback %.1f ms), out %d" % (
    os.environ.get("FB200_STREAM_PART", "32768"), passes * mib, dt * 1e3, passes * n / 1e6 / dt, cb_time[0] * 1e3, got))
