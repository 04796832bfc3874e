"""Development aid: sparse parse (mode 0) against the dense tables (mode 1) on assorted inputs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import flate_b200  # noqa: E402
from flate_b200 import synth  # noqa: E402

sparse = flate_b200.Context(0)
dense = flate_b200.Context(0)
dense.set_parse_mode(1)
rng = np.random.default_rng(5)
text = synth.enwik_like(12 << 20, seed=3)
cases = [("empty", np.zeros(0, np.uint8)), ("zeros1M", np.zeros(1 << 20, np.uint8)),
         ("random1M", rng.integers(0, 256, 1 << 20, dtype=np.uint8)),
         ("lowent", rng.integers(0, 4, 300000, dtype=np.uint8)),
         ("period7", np.tile(np.arange(7, dtype=np.uint8), 100000)),
         ("mixed", synth.mixed_small(3 << 20, seed=4) if hasattr(synth, "mixed_small") else text[:3 << 20]),
         ("text+zeros+text", np.concatenate([text[:700000], np.zeros(200000, np.uint8), text[700000:1500000]])),
         ("text+period+text8M", np.concatenate([text[:3000000], np.tile(np.arange(11, dtype=np.uint8), 30000), text[3000000:8000000]])),
         ("lowent3M", rng.integers(0, 4, 3 << 20, dtype=np.uint8))]
for k in (1, 3, 4, 5, 100, 4095, 4096, 4097, 4608, 4609, 8191, 8192, 32767, 32768, 32769, 33280, 33281, 65536, 65537,
          100000, 1 << 20, (1 << 20) + 17, 5 * (1 << 20) + 4321, 12 << 20):
    cases.append(("text%d" % k, text[:k]))
bad = 0
for name, d in cases:
    for level in (4, 5, 6, 7, 8, 9):
        if level in (8, 9) and d.size > (2 << 20):
            continue
        f0 = sparse.sparse_fallbacks
        r0 = sparse.sparse_repairs
        a = sparse.compress(d, flate_b200.RAW, level)
        b = dense.compress(d, flate_b200.RAW, level)
        ok = a == b
        fb = sparse.sparse_fallbacks - f0
        rp = sparse.sparse_repairs - r0
        if not ok or fb or rp:
            print("%s L%d: %s fallbacks=%d repairs=%d (sizes %d %d)" % (name, level, "OK" if ok else "DIFF", fb, rp, len(a), len(b)), flush=True)
        bad += 0 if ok else 1
    # host path == device path for the sparse mode is covered by the test-suite
print("sparse_check SPARSE=%s: %d cases, %d mismatches, total fallbacks %d, repairs %d" % (
    os.environ.get("FB200_SPARSE", "-"), len(cases), bad, sparse.sparse_fallbacks, sparse.sparse_repairs))
sys.exit(1 if bad else 0)
