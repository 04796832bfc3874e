"""Times the inflate kernel on N gzip members of 1 MiB level-6 text, for several N (development aid).
Every member's output is compared on the device with the text it was made from."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import flate_b200  # noqa: E402
from flate_b200 import synth  # noqa: E402

counts = [int(a) for a in sys.argv[1:]] or [1, 8, 128, 296, 1024]
MB = 1 << 20
ctx = flate_b200.Context(0)
uniq = 64
text = synth.enwik_like(uniq * MB, seed=5)
members, plains = [], []
for i in range(uniq):
    for shift in range(64):
        lo = (i * MB + shift * 4099) % (text.size - MB + 1)
        m = ctx.compress(text[lo:lo + MB], flate_b200.GZIP, 6)
        try:
            ctx.decompress(m, flate_b200.GZIP, cap=MB + 64)
            break
        except flate_b200.FlateError as e:
            if type(e).__name__ != "InvalidDynamicBlockHeader":
                raise
            continue
    members.append(m)
    plains.append(text[lo:lo + MB])
d_want = torch.from_numpy(np.concatenate(plains)).cuda()
sp = torch.cuda.current_stream().cuda_stream
for nmem in counts:
    blob = b"".join(members[i % uniq] for i in range(nmem))
    lens = np.array([len(members[i % uniq]) for i in range(nmem)], dtype=np.uint64)
    offs = np.zeros(nmem, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)[:-1]
    d_blob = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).cuda()
    d_plain = torch.zeros(nmem * MB + 64, dtype=torch.uint8, device="cuda")
    ooff = np.arange(nmem, dtype=np.uint64) * np.uint64(MB)
    ocap = np.full(nmem, MB, dtype=np.uint64)
    for _ in range(2):
        rc, ol, used, st = ctx.decompress_members_device(d_blob.data_ptr(), offs, lens, d_plain.data_ptr(), ooff, ocap,
                                                         flate_b200.GZIP, stream=sp)
    assert rc == 0, (rc, st[:8])
    assert (ol == MB).all() and (used == lens).all()
    for i in range(0, nmem, uniq):
        k = min(uniq, nmem - i)
        assert torch.equal(d_plain[i * MB:(i + k) * MB], d_want[:k * MB]), "member output differs (batch at %d)" % i
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.profile(True)
    e0.record()
    for _ in range(reps):
        ctx.decompress_members_device(d_blob.data_ptr(), offs, lens, d_plain.data_ptr(), ooff, ocap, flate_b200.GZIP, stream=sp)
    e1.record()
    torch.cuda.synchronize()
    ph = ctx.profile_read()["inflate_members"]
    ctx.profile(False)
    dt = e0.elapsed_time(e1) / reps
    kms = ph[0] / max(1, ph[1])
    print("inflate %4d members x 1 MiB: %.3f ms/call, kernel %.3f ms -> %.1f MB/s out (kernel %.1f MB/s)"
          % (nmem, dt, kms, nmem * MB / dt / 1e3, nmem * MB / kms / 1e3), flush=True)
    if os.environ.get("FB200_PROFILE_RANGE"):  # ncu --profile-from-start off: capture exactly this launch
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        ctx.decompress_members_device(d_blob.data_ptr(), offs, lens, d_plain.data_ptr(), ooff, ocap, flate_b200.GZIP, stream=sp)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    del d_blob, d_plain
