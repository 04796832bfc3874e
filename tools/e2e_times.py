"""Times fb200_compress (pinned host buffers, copies inside) on synthetic text (development aid)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import flate_b200  # noqa: E402
from flate_b200 import synth  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = flate_b200.Context(0)
d = synth.enwik_like(mib << 20, seed=19)
n = d.size
cap = ctx.lib.fb200_compress_bound(n, level) + 64
h_in = torch.from_numpy(d).pin_memory()
h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
ln = C.c_size_t(0)
for _ in range(3):
    rc = ctx.lib.fb200_compress(ctx.h, 0, level, h_in.data_ptr(), n, h_out.data_ptr(), cap, C.byref(ln))
    assert rc == 0, rc
reps = 6
t = time.perf_counter()
for _ in range(reps):
    ctx.lib.fb200_compress(ctx.h, 0, level, h_in.data_ptr(), n, h_out.data_ptr(), cap, C.byref(ln))
dt = (time.perf_counter() - t) / reps
print("SLAB=%s %d MiB L%d e2e: %.3f ms (%.1f MB/s) out=%d" % (os.environ.get("FB200_SLAB", "-"), mib, level, dt * 1e3, n / dt / 1e6, ln.value))
