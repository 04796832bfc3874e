"""Times fb200_compress end to end (pinned host buffers, copies inside) for a slab setting given by FB200_SLAB
("first_MiB,slab_MiB"); development aid."""
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
lib = ctx.lib
n = mib << 20
h_in = torch.from_numpy(synth.enwik_like(n, seed=19)).pin_memory()
cap = lib.fb200_compress_bound(n, level) + 64
h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
ln = C.c_size_t(0)


def once():
    rc = lib.fb200_compress(ctx.h, flate_b200.RAW, level, h_in.data_ptr(), n, h_out.data_ptr(), cap, C.byref(ln))
    assert rc == 0, rc


for _ in range(3):
    once()
best, tot, reps = 1e9, 0.0, 6
for _ in range(reps):
    t = time.perf_counter()
    once()
    dt = time.perf_counter() - t
    best = min(best, dt)
    tot += dt
print("SLAB=%s %d MiB L%d e2e: mean %.3f ms, best %.3f ms, out=%d" % (os.environ.get("FB200_SLAB", "-"), mib, level, tot / reps * 1e3, best * 1e3, ln.value))
