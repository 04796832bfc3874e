"""Stage-by-stage GPU-vs-oracle diagnostic (development aid; run under gpurun).  Every stage is
wrapped so one run reports as much as possible."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flate_b200  # noqa: E402
from flate_b200 import synth  # noqa: E402
from oracle import oracle as o  # noqa: E402


def fd(a, b):
    a = np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else np.asarray(a)
    b = np.frombuffer(b, dtype=np.uint8) if isinstance(b, (bytes, bytearray)) else np.asarray(b)
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    return (int(d[0]) if d.size else (n if a.size != b.size else -1)), a.size, b.size


def stage(name, fn):
    t = time.time()
    try:
        r = fn()
        print("[%s] %s (%.2fs)" % ("OK" if r is None else "FAIL", name, time.time() - t), "" if r is None else r, flush=True)
    except Exception as e:
        print("[EXC] %s: %s %s" % (name, type(e).__name__, e), flush=True)
        traceback.print_exc()


def main():
    ctx = flate_b200.Context(0)
    golden = os.path.join(ROOT, "tests", "golden")
    rfc = open(os.path.join(golden, "rfc1951.txt"), "rb").read()
    datas = {"blah": b"Blah blah blah blah blah!", "rfc": rfc, "text70000": synth.enwik_like(70000, seed=3).tobytes(),
             "mixed200001": synth.mixed_small(200001, seed=8).tobytes()}
    for name, d in datas.items():
        for level in (6, 4, 9):
            def mt():
                rf, rq = ctx.debug_match_tables(d, level)
                of, oq = o.match_tables(d, level)
                a, b = fd(rf, of), fd(rq, oq)
                if a[0] != -1:
                    p = a[0]
                    return "full mismatch at %d: got %s want %s" % (p, [(int(x) & 511, (int(x) >> 9) + 1) for x in rf[p:p + 3]], [(int(x) & 511, (int(x) >> 9) + 1) for x in of[p:p + 3]])
                if b[0] != -1:
                    p = b[0]
                    return "quarter mismatch at %d: got %s want %s" % (p, (int(rq[p]) & 511, (int(rq[p]) >> 9) + 1), (int(oq[p]) & 511, (int(oq[p]) >> 9) + 1))
            stage("match_tables %s L%d" % (name, level), mt)

            def tk():
                g = ctx.debug_tokens(d, level)
                w = o.tokenize(d, level)
                r = fd(g, w)
                if r[0] != -1:
                    i = r[0]
                    return "token mismatch at %d of (%d, %d): got %s want %s" % (i, g.size, w.size, [hex(int(x)) for x in g[i:i + 3]], [hex(int(x)) for x in w[i:i + 3]])
            stage("tokens %s L%d" % (name, level), tk)

            def cp():
                g = ctx.compress(d, 0, level)
                w = o.compress(d, 0, level)
                r = fd(g, w)
                if r[0] != -1:
                    i = r[0]
                    return "byte mismatch at %d of (%d, %d): got %s want %s" % (i, len(g), len(w), g[max(0, i - 2):i + 6].hex(), w[max(0, i - 2):i + 6].hex())
            stage("compress %s L%d" % (name, level), cp)
        for mode in (1, 0):
            def cs():
                g = ctx.compress(d, 1, mode)
                w = o.compress(d, 1, mode)
                r = fd(g, w)
                if r[0] != -1:
                    i = r[0]
                    return "byte mismatch at %d of (%d, %d): got %s want %s" % (i, len(g), len(w), g[max(0, i - 2):i + 6].hex(), w[max(0, i - 2):i + 6].hex())
            stage("compress %s mode%d gzip" % (name, mode), cs)

        def inf():
            for mode in (0, 1, 6):
                for cont in (0, 1, 2):
                    c = o.compress(d, cont, mode)
                    try:
                        p, used = ctx.decompress(c, cont)
                    except flate_b200.FlateError as e:
                        return "mode %d cont %d: %s" % (mode, cont, type(e).__name__)
                    if p != d or used != len(c):
                        return "mode %d cont %d: mismatch %s used %d/%d" % (mode, cont, fd(p, d), used, len(c))
        stage("inflate %s" % name, inf)

    def big():
        d = synth.enwik_like(8 << 20, seed=9).tobytes()
        t = time.time()
        g = ctx.compress(d, 0, 6)
        t1 = time.time() - t
        t = time.time()
        w = o.compress(d, 0, 6)
        t2 = time.time() - t
        r = fd(g, w)
        print("  8MiB L6: gpu %.3fs cpu %.3fs  sizes %d %d" % (t1, t2, len(g), len(w)))
        if r[0] != -1:
            return "byte mismatch at %d" % r[0]
        t = time.time()
        p, _ = ctx.decompress(g, 0, cap=len(d) + 64)
        print("  8MiB inflate: gpu %.3fs" % (time.time() - t))
        if p != d:
            return "inflate mismatch"
    stage("big 8MiB", big)

    def members():
        plains, mem = [], []
        for i in range(40):
            p = synth.enwik_like(300000 + 1777 * i, seed=100 + i).tobytes()
            plains.append(p)
            mem.append(o.compress(p, 1, 6))
        blob = b"".join(mem)
        off = np.cumsum([0] + [len(m) for m in mem[:-1]])
        for rep in range(3):
            t = time.time()
            outs, st, used = ctx.decompress_members(blob, off, [len(m) for m in mem], [len(p) + 64 for p in plains], 1)
            dt = time.time() - t
        print("  members: %.4fs for %d bytes -> %.1f MB/s; status %s" % (dt, sum(map(len, plains)), sum(map(len, plains)) / dt / 1e6, st))
        bad = [i for i in range(40) if outs[i] != plains[i]]
        if bad:
            return "members differ: %s offs %s" % (bad, [int(off[i]) % 16 for i in bad])
    stage("members", members)

    def timing():
        d = synth.enwik_like(32 << 20, seed=19).tobytes()
        g = ctx.compress(d, 0, 6)
        for rep in range(3):
            t = time.time()
            g = ctx.compress(d, 0, 6)
            t1 = time.time() - t
        ctx.profile(True)
        g = ctx.compress(d, 0, 6)
        print("  32MiB L6 e2e: %.4fs (%.1f MB/s) phases %s" % (t1, len(d) / t1 / 1e6, {k: round(v[0], 3) for k, v in ctx.profile_read().items() if v[1]}))
        ctx.profile(False)
        for rep in range(2):
            t = time.time()
            p, _ = ctx.decompress(g, 0, cap=len(d) + 64)
            t2 = time.time() - t
        print("  32MiB single-member inflate: %.4fs (%.1f MB/s)" % (t2, len(d) / t2 / 1e6))
    stage("timing", timing)
    print("launches", ctx.kernel_launches)


if __name__ == "__main__":
    main()
