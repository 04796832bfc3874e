"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flate_b200  # noqa: E402
from flate_b200 import synth  # noqa: E402
import io  # noqa: E402
import zlib  # noqa: E402

ctx = flate_b200.Context(0)
text = synth.enwik_like(200000, seed=3).tobytes()
mixed = synth.mixed_small(150000, seed=8).tobytes()
for data in (text, mixed, b"", b"abc"):
    for mode in (0, 1, 4, 6, 9):
        for container in (0, 1, 2):
            c = ctx.compress(data, container, mode)
            assert zlib.decompress(c, {0: -15, 1: 31, 2: 15}[container]) == data
            try:
                p, used = ctx.decompress(c, container)
                assert p == data and used == len(c)
            except flate_b200.FlateError as e:  # the reference's own lit/dist-boundary rejection
                assert type(e).__name__ == "InvalidDynamicBlockHeader"
# sparse parse: several 32 KiB chunks, a periodic stretch (coverage check fails -> dense repair), zeros (-> dense redo)
import numpy as np  # noqa: E402
t3 = synth.enwik_like(150000, seed=4)
rep = np.concatenate([t3[:70000], np.tile(np.arange(9, dtype=np.uint8), 5000), t3[70000:]]).tobytes()
for data in (rep, bytes(120000)):
    for mode in (4, 6):
        c = ctx.compress(data, 0, mode)
        assert zlib.decompress(c, -15) == data
print("sparse repairs", ctx.sparse_repairs, "fallbacks", ctx.sparse_fallbacks)
if os.environ.get("FB200_SANITIZE_QUICK"):   # the compress / decompress matrix and the repairs only
    print("sanitize run ok (quick)")
    sys.exit(0)
w = io.BytesIO()
comp = flate_b200.Compressor(1, w, 6, ctx=ctx)
comp.write(text[:70001]); comp.flush(); comp.write(text[70001:70004]); comp.flush(); comp.write(text[70004:]); comp.finish()
assert zlib.decompress(w.getvalue(), 31) == text
# round 2: streaming compressor in parts (tiny parts: slides, carried open blocks and partial bytes), huffman-only in
# its split code construction (>= 64 blocks), streaming decompressor piece by piece, member discovery, tiny members
os.environ["FB200_STREAM_PART"] = "128"
big = (synth.enwik_like(700000, seed=5).tobytes() + bytes(90000) + text)
for mode in (6, 1):
    w = io.BytesIO()
    comp = flate_b200.Compressor(1, w, mode, ctx=ctx)
    for p in range(0, len(big), 50000):
        comp.write(big[p:p + 50000])
    comp.finish()
    assert zlib.decompress(w.getvalue(), 31) == big
    comp.close()
hb = np.random.default_rng(1).integers(0, 64, 65535 * 70, dtype=np.uint8).tobytes()
c = ctx.compress(hb, 0, 1)
assert zlib.decompress(c, -15) == hb


class Rd:
    def __init__(self, d):
        self.d, self.p = d, 0

    def read(self, n):
        n = min(n, 30000)
        b = self.d[self.p:self.p + n]
        self.p += len(b)
        return b


members = [ctx.compress(x, 1, 6) for x in (text[:60000], b"", b"x", bytes(50000), text[60000:150000])]
out = io.BytesIO()
dec = flate_b200.Decompressor(1, Rd(b"".join(members)), ctx=ctx)
for i in range(len(members)):
    try:
        dec.decompress(out)
    except flate_b200.FlateError as e:
        assert type(e).__name__ == "InvalidDynamicBlockHeader"
        break
    if i + 1 < len(members):
        dec.reset()
try:
    ctx.decompress_gzip_file(b"".join(members))
except flate_b200.FlateError as e:
    assert type(e).__name__ == "InvalidDynamicBlockHeader"
print("sanitize run ok")
