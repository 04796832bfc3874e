"""Times the inflate kernel on N gzip members of 1 MiB level-6 text (development aid)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import flate_b200  # noqa: E402
from flate_b200 import synth  # noqa: E402

nmem = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
MB = 1 << 20
ctx = flate_b200.Context(0)
uniq = min(nmem, 64)
text = synth.enwik_like(uniq * MB, seed=5)
members = []
for i in range(uniq):
    for shift in range(1 if os.environ.get('FB200_IGNORE_RC') else 64):
        lo = (i * MB + shift * 4099) % (text.size - MB + 1)
        m = ctx.compress(text[lo:lo + MB], flate_b200.GZIP, 6)
        try:
            ctx.decompress(m, flate_b200.GZIP, cap=MB + 64)
            break
        except flate_b200.FlateError:
            continue
    members.append(m)
blob = b"".join(members[i % uniq] for i in range(nmem))
lens = np.array([len(members[i % uniq]) for i in range(nmem)], dtype=np.uint64)
offs = np.zeros(nmem, dtype=np.uint64)
offs[1:] = np.cumsum(lens)[:-1]
d_blob = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).cuda()
d_plain = torch.empty(nmem * MB + 64, dtype=torch.uint8, device="cuda")
ooff = np.arange(nmem, dtype=np.uint64) * np.uint64(MB)
ocap = np.full(nmem, MB, dtype=np.uint64)
sp = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    rc, ol, used, st = ctx.decompress_members_device(d_blob.data_ptr(), offs, lens, d_plain.data_ptr(), ooff, ocap, flate_b200.GZIP, stream=sp)
assert rc == 0 or os.environ.get('FB200_IGNORE_RC'), rc
torch.cuda.synchronize()
t = time.perf_counter()
reps = 3
for _ in range(reps):
    ctx.decompress_members_device(d_blob.data_ptr(), offs, lens, d_plain.data_ptr(), ooff, ocap, flate_b200.GZIP, stream=sp)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / reps
print("inflate %d members x 1 MiB: %.2f ms -> %.1f MB/s out" % (nmem, dt * 1e3, nmem * MB / dt / 1e6))
