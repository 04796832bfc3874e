"""The code construction of the GPU path (block_writer.cu, bit_counts_eager_warp) computes the per-length counts of the
length-limited Huffman code from the EAGER package-merge; the reference's bitCounts (huffman_encoder.zig:122-247) is the
lazy, boundary evaluation of the same construction.  This test pins the claim the kernel rests on: a plain model of the
eager form -- level l = merge of the sorted leaves with the pair sums of level l - 1, a pair first when it ties with a
leaf (:182 takes the leaf only if strictly smaller); the top level takes 2n - 2 items, a level that contributes p pairs
makes the level below contribute 2p items; symbol i gets as many bits as there are levels with more than i leaves taken
-- gives the length histogram of the sequential restatement for every frequency set tried: ties, skewed sets where the
15-bit (7-bit) limit binds, flat sets, Fibonacci-like sets."""
import numpy as np
import pytest

from oracle import oracle as o


def eager_lengths(freqs_sorted, max_bits):
    n = len(freqs_sorted)
    top = min(max_bits, n - 1)
    leaves = list(freqs_sorted)
    seqs = [None, [(v, True) for v in leaves]]
    for _ in range(2, top + 1):
        prev = seqs[-1]
        pairs = [prev[2 * j][0] + prev[2 * j + 1][0] for j in range(len(prev) // 2)]
        merged, i, j = [], 0, 0
        while (i < n or j < len(pairs)) and len(merged) < 2 * n:
            if j >= len(pairs) or (i < n and leaves[i] < pairs[j]):
                merged.append((leaves[i], True))
                i += 1
            else:
                merged.append((pairs[j], False))
                j += 1
        seqs.append(merged)
    take, taken = 2 * n - 2, [0] * (top + 1)
    for lvl in range(top, 0, -1):
        items = seqs[lvl][:take]
        taken[lvl] = sum(1 for _, leaf in items if leaf)
        take = 2 * (len(items) - taken[lvl])
    return [sum(1 for lvl in range(1, top + 1) if i < taken[lvl]) for i in range(n)]


def frequency_sets(count, seed):
    rng = np.random.default_rng(seed)
    for trial in range(count):
        n = int(rng.integers(3, 287))
        kind = trial % 8
        if kind == 0:
            f = rng.integers(1, 5, n)
        elif kind == 1:
            f = rng.integers(1, 65535, n)
        elif kind == 2:
            f = (rng.zipf(1.3, n) % 60000) + 1
        elif kind == 3:
            f = np.array([1 << min(i, 15) for i in range(n)]) % 65535 + 1
        elif kind == 4:
            f = rng.integers(1, 40, n) ** 3 % 65535 + 1
        elif kind == 5:
            fib = [1, 1]
            while len(fib) < n:
                fib.append(fib[-1] + fib[-2] if fib[-1] + fib[-2] < 30000 else 1)
            f = np.array(fib[:n])
        elif kind == 6:
            f = np.where(rng.random(n) < 0.1, rng.integers(1000, 5000, n), 1)
        else:
            f = np.full(n, int(rng.integers(1, 200)))
        f = f.astype(np.int64)
        max_bits = 15 if trial % 3 else 7
        if max_bits == 7:
            n = min(n, int(rng.integers(3, 20)))
            f = f[:n]
        if f.sum() > 65000:   # a block holds at most 65535 symbols + end of block
            f = np.maximum(1, f * 65000 // f.sum())
        yield f, max_bits


def test_eager_package_merge_gives_the_lengths_of_bit_counts():
    for f, max_bits in frequency_sets(1500, seed=1):
        n = f.size
        freq = np.zeros(286 if max_bits == 15 else 19, dtype=np.uint16)
        freq[:n] = f
        _, want = o.huffman_generate(freq, max_bits)
        order = sorted(range(n), key=lambda i: (int(f[i]), i))
        got = eager_lengths([int(f[i]) for i in order], max_bits)
        assert (np.bincount(np.asarray(want[:n], dtype=np.int64), minlength=17) == np.bincount(np.asarray(got), minlength=17)).all(), (n, max_bits)
