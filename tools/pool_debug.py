import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, flate_b200
from flate_b200 import synth, api
from oracle import oracle as o
pool = flate_b200.Pool(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
items = [synth.enwik_like(200000 + 70001 * i, seed=500 + i).tobytes() for i in range(9)] + [b"", b"x", bytes(100000)]
members = [o.compress(it, 1, 6) for it in items]
blob = b"".join(members)
lens = [len(m) for m in members]
offs = [sum(lens[:i]) for i in range(len(lens))]
a = api._as_u8(blob)
k = len(offs)
io_ = np.asarray(offs, dtype=np.uint64); il = np.asarray(lens, dtype=np.uint64)
oc = np.asarray([len(it) + 16 for it in items], dtype=np.uint64)
oo = np.zeros(k, dtype=np.uint64); oo[1:] = np.cumsum(oc)[:-1]
out = np.empty(int(oc.sum()) + 1, dtype=np.uint8)
ol = np.zeros(k, dtype=np.uint64); used = np.zeros(k, dtype=np.uint64); st = np.zeros(k, dtype=np.int32)
rc = pool.lib.fb200_decompress_members_batch(pool.h, 1, api._ptr(a), io_.ctypes.data, il.ctypes.data, k, out.ctypes.data,
                                             oo.ctypes.data, oc.ctypes.data, ol.ctypes.data, used.ctypes.data, st.ctypes.data)
print("rc", rc, "ol", ol.tolist(), "used", used.tolist(), "lens", lens, "st", st.tolist())
print([out[int(oo[i]):int(oo[i] + ol[i])].tobytes() == items[i] for i in range(k)])
print(pool.lib.fb200_last_cuda_error())
