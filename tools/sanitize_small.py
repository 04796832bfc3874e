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
w = io.BytesIO()
comp = flate_b200.Compressor(1, w, 6, ctx=ctx)
comp.write(text[:70001]); comp.flush(); comp.write(text[70001:70004]); comp.flush(); comp.write(text[70004:]); comp.finish()
assert zlib.decompress(w.getvalue(), 31) == text
print("sanitize run ok")
