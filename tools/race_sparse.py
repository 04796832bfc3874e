"""Small sparse-parse run (one dense repair included) for compute-sanitizer --tool racecheck."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, zlib
import flate_b200
from flate_b200 import synth
ctx = flate_b200.Context(0)
t3 = synth.enwik_like(120000, seed=4)
rep = np.concatenate([t3[:50000], np.tile(np.arange(9, dtype=np.uint8), 4000), t3[50000:]]).tobytes()
c = ctx.compress(rep, 0, 6)
assert zlib.decompress(c, -15) == rep
print("repairs", ctx.sparse_repairs, "fallbacks", ctx.sparse_fallbacks, "ok")
