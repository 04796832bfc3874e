"""Prints per-phase device times of one level-N compress of synthetic text (development aid)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import flate_b200  # noqa: E402
from flate_b200 import synth  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = flate_b200.Context(0)
ctx.set_parse_mode(int(os.environ.get("FB200_PARSE_MODE", "0")))
if os.environ.get("FB200_DATA") == "tar":   # the bench's tar-like input (text and random members, zero padding, holes)
    import bench
    d = bench.make_tar_like(mib << 20)
else:
    d = synth.enwik_like(mib << 20, seed=19) if level >= 4 else synth.random_zero_mix(mib << 20)
t_in = torch.from_numpy(d).cuda()
cap = ctx.lib.fb200_compress_bound(d.size, level) + 64
t_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
sp = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    n = ctx.compress_device(t_in.data_ptr(), d.size, t_out.data_ptr(), cap, mode=level, stream=sp)
ctx.profile(True)
reps = 3
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(reps):
    n = ctx.compress_device(t_in.data_ptr(), d.size, t_out.data_ptr(), cap, mode=level, stream=sp)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / reps
ph = ctx.profile_read()
print("repairs", ctx.sparse_repairs, "phase counts", {k: v[1] for k, v in ph.items() if v[1]})
print("TUNE=%s SPARSE=%s fallbacks=%d %d MiB L%d: %.3f ms/step (%.1f MB/s) out=%d | %s" % (
    os.environ.get("FB200_TUNE", "-"), os.environ.get("FB200_SPARSE", "-"), ctx.sparse_fallbacks, mib, level, dt * 1e3, d.size / dt / 1e6, n,
    " ".join("%s=%.3f" % (k, v[0] / v[1]) for k, v in ph.items() if v[1])))
