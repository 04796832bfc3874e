import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, flate_b200
from flate_b200 import synth
from oracle import oracle as o
ctx = flate_b200.Context(0)
items = [synth.enwik_like(200000 + 70001 * i, seed=500 + i).tobytes() for i in range(9)] + [b"", b"x", bytes(100000)]
sel = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(range(len(items)))
items = [items[i] for i in sel]
members = [o.compress(it, 1, 6) for it in items]
blob = b"".join(members)
lens = [len(m) for m in members]
offs = [sum(lens[:i]) for i in range(len(lens))]
plains, st, used = ctx.decompress_members(blob, offs, lens, [len(it) + 16 for it in items], flate_b200.GZIP)
print(sel, st, [p == it for p, it in zip(plains, items)], ctx.lib.fb200_last_cuda_error())
