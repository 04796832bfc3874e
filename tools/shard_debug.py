"""Development aid: emulate a 2-rank position-sharded sparse search on one GPU at full size and, if the joined
table does not cover the orbit, report where the orbit meets an unevaluated entry."""
import ctypes
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import flate_b200  # noqa: E402
from flate_b200 import sharding, synth  # noqa: E402

SRC = r"""
#include <stdint.h>
long walk(const uint32_t* nx, long n, long* last_ok) {
    long p = 0, prev = -1;
    while (p < n) {
        uint32_t v = nx[p];
        if (v == 0xFFFFFFFFu) { *last_ok = prev; return p; }
        prev = p;
        p += (v >> 16) ? (v & 255u) + ((v >> 8) & 255u) + 3u : 1u;
    }
    *last_ok = prev;
    return -1;
}
"""
d = tempfile.mkdtemp()
open(os.path.join(d, "w.c"), "w").write(SRC)
subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", os.path.join(d, "w.so"), os.path.join(d, "w.c")])
W = ctypes.CDLL(os.path.join(d, "w.so"))
W.walk.restype = ctypes.c_long
W.walk.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.POINTER(ctypes.c_long)]

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
data = synth.enwik_like(mib << 20, seed=19)
n = data.size
d_in = torch.from_numpy(data).cuda()
ctx = flate_b200.Context(0)
ov, align = ctx.shard_overlap, ctx.shard_align
per, ranges = sharding.shard_positions(n, parts, align)
nx = torch.full((parts * per + ov,), -1, dtype=torch.int32, device="cuda")
tails = []
for r in range(parts):
    c = flate_b200.Context(0)
    t = torch.full((parts * per + ov,), 0x5a5a5a5a, dtype=torch.int32, device="cuda")
    lo, hi = ranges[r]
    ok = c.shard_search(d_in.data_ptr(), n, lo, hi, t.data_ptr(), level=6)
    print("rank", r, "range", lo, hi, "ok", ok, flush=True)
    c.close()
    nx[r * per:(r + 1) * per] = t[r * per:(r + 1) * per]
    tails.append(t[(r + 1) * per:(r + 1) * per + ov].clone())
    del t
for r in range(parts - 1):
    sharding.merge_overlap(nx[(r + 1) * per:(r + 1) * per + ov], tails[r])
h = nx[:n].cpu().numpy().view(np.uint32)
last = ctypes.c_long(0)
bad = W.walk(h.ctypes.data, n, ctypes.byref(last))
print("orbit walk: first unevaluated entry at", bad, "previous arrival", last.value, "per", per)
if bad >= 0:
    lo = max(0, last.value - 8)
    print("entries around:", [(int(p), hex(int(h[p]))) for p in range(lo, min(n, bad + 4))])
    print("valid count in [per-64, per+ov):", int((h[per - 64:per + ov] != 0xFFFFFFFF).sum()))
