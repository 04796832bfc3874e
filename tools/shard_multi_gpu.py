"""One 256 MiB deflate stream sharded by position over all ranks (run under torchrun); checks the result
against the single-GPU path and prints the strong-scaling time."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.pop("NCCL_DEBUG", None)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import flate_b200  # noqa: E402
from flate_b200 import sharding, synth  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = flate_b200.Context(local)
data = synth.enwik_like(mib << 20, seed=0x5EED0001)   # the same stream on every rank
d_in = torch.from_numpy(data).cuda()
cap = ctx.lib.fb200_compress_bound(data.size, 6) + 64
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
if os.environ.get("FB200_WARM_CTX"):   # what bench.py does before its single-stream leg
    ctx.compress_device(d_in.data_ptr(), data.size, d_out.data_ptr(), cap, mode=6, stream=torch.cuda.current_stream().cuda_stream)
keep = {}
try:
    for _ in range(2):
        m = sharding.compress_stream_sharded(ctx, d_in, data.size, d_out, level=6, keep=keep)
except flate_b200.api.RetryDense:
    import numpy as np
    h = keep["nx"][:data.size].cpu().numpy().view(np.uint32)
    per = keep["per"]
    p, prev = 0, -1
    bad = (h == 0xFFFFFFFF)
    # walk with numpy-free python only near the failure: jump table first
    step = np.where((h >> 16) != 0, (h & 255) + ((h >> 8) & 255) + 3, 1).astype(np.int64)
    while p < data.size and not bad[p]:
        prev = p
        p += int(step[p])
    print("rank", rank, "orbit meets an unevaluated entry at", p, "previous arrival", prev, "per", per,
          "entries", [(q, hex(int(h[q]))) for q in range(max(0, prev - 2), min(data.size, p + 3))], flush=True)
    raise
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    m = sharding.compress_stream_sharded(ctx, d_in, data.size, d_out, level=6)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
if rank == 0:
    sharded = d_out[:m].clone()
    n1 = ctx.compress_device(d_in.data_ptr(), data.size, d_out.data_ptr(), cap, mode=6,
                             stream=torch.cuda.current_stream().cuda_stream)
    same = n1 == m and bool((d_out[:n1] == sharded).all())
    print("single stream %d MiB over %d GPU(s): %.2f ms -> %.1f MB/s, identical to the one-GPU stream: %s"
          % (mib, world, dt * 1e3, data.size / dt / 1e6, same))
if world > 1:
    dist.destroy_process_group()
