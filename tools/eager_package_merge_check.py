import sys, numpy as np
sys.path.insert(0, "/root/repo")
from oracle import oracle as o

def eager_bit_counts(freqs_sorted, max_bits):
    """freqs_sorted ascending list of n >= 3 frequencies.  Returns bit_count[1..max_bits] (number of codes per length)."""
    n = len(freqs_sorted)
    L = min(max_bits, n - 1)
    A = list(freqs_sorted)
    seqs = [None] * (L + 1)       # seqs[l] = list of (value, is_leaf)
    seqs[1] = [(v, True) for v in A]
    for l in range(2, L + 1):
        prev = seqs[l - 1]
        pairs = [prev[2 * j][0] + prev[2 * j + 1][0] for j in range(len(prev) // 2)]
        merged, i, j = [], 0, 0
        while (i < n or j < len(pairs)) and len(merged) < 2 * n:
            if j >= len(pairs) or (i < n and A[i] < pairs[j]):   # leaf only if strictly smaller: ties take the pair
                merged.append((A[i], True)); i += 1
            else:
                merged.append((pairs[j], False)); j += 1
        seqs[l] = merged
    T = 2 * n - 2
    a = [0] * (L + 1)
    for l in range(L, 0, -1):
        items = seqs[l][:T]
        a[l] = sum(1 for (_, leaf) in items if leaf)
        T = 2 * (len(items) - a[l])
    # symbol i (sorted ascending) has length = number of levels with i < a[l]; bit_count[len]
    lens = [sum(1 for l in range(1, L + 1) if i < a[l]) for i in range(n)]
    return lens

def oracle_lens(freq_by_symbol, max_bits):
    codes, lens = o.huffman_generate(freq_by_symbol, max_bits)
    return lens

rng = np.random.default_rng(1)
bad = 0
for trial in range(12000):
    n = int(rng.integers(3, 287))
    kind = trial % 8
    if kind == 0: f = rng.integers(1, 5, n)             # many ties
    elif kind == 1: f = rng.integers(1, 65535, n)
    elif kind == 2: f = (rng.zipf(1.3, n) % 60000) + 1
    elif kind == 3: f = np.array([1 << min(i, 15) for i in range(n)]) % 65535 + 1   # deep tree -> length limit binds
    elif kind == 4: f = rng.integers(1, 40, n) ** 3 % 65535 + 1
    elif kind == 5:
        fib = [1, 1]
        while len(fib) < n: fib.append(fib[-1] + fib[-2] if fib[-1] + fib[-2] < 30000 else 1)
        f = np.array(fib[:n])
    elif kind == 6: f = np.where(rng.random(n) < 0.1, rng.integers(1000, 5000, n), 1)
    else: f = np.full(n, int(rng.integers(1, 200)))
    f = f.astype(np.int64)
    if f.sum() > 65000:
        f = np.maximum(1, f * 65000 // f.sum())
    max_bits = 15 if trial % 3 else 7
    if max_bits == 7:
        n = min(n, int(rng.integers(3, 20))); f = f[:n]
    freq = np.zeros(286 if max_bits == 15 else 19, dtype=np.uint16)
    freq[:n] = f
    want = np.array(oracle_lens(freq, max_bits))[:n]
    order = sorted(range(n), key=lambda i: (int(f[i]), i))
    fs = [int(f[i]) for i in order]
    lens_sorted = eager_bit_counts(fs, max_bits)
    # the reference assigns lengths by the histogram over the sorted list: compare histograms
    hist_want = np.bincount(want, minlength=17)
    hist_got = np.bincount(np.array(lens_sorted), minlength=17)
    if not (hist_want == hist_got).all():
        bad += 1
        if bad < 5: print("MISMATCH n", n, "max_bits", max_bits, "kind", kind, hist_want[1:16], hist_got[1:16])
print("trials done, mismatches:", bad)
