"""Deterministic synthetic corpora for tests and bench.py (no datasets on the box).

enwik_like(n, seed): word salad with a Zipf vocabulary, XML-ish markup, wiki links, numerals and
  newlines -- the shape BASELINE.json names for configs C2/C4 ("enwik-like synthetic bytes").
random_zero_mix(n, seed): alternating runs of uniform random bytes and zeros, run lengths uniform
  in [4 KiB, 256 KiB] -- config C5 ("random+zeros mix").
All randomness is splitmix64 on a counter, so the bytes depend only on (n, seed)."""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed, count, start=0):
    """count outputs of splitmix64 for counter values start+1 .. start+count."""
    with np.errstate(over="ignore"):
        i = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_W = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0, 1.9,
                      1.5, 1.0, 0.8, 0.15, 0.15, 0.1, 0.07])


def _vocabulary(seed, nwords=65536):
    r = splitmix64(seed ^ 0xA5A5A5A5, nwords * 13)
    lens = (2 + (r[:nwords] % np.uint64(11))).astype(np.int64)  # 2..12
    cdf = np.cumsum(_LETTER_W) / _LETTER_W.sum()
    u = (r[nwords:] >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    letters = _LETTERS[np.searchsorted(cdf, u).clip(0, 25)].reshape(nwords, 12)
    return lens, letters


_MARKUP = [b"<page><title>", b"</title>", b"<text>", b"</text></page>\n", b"[[", b"]]", b"<ref>", b"</ref>", b"''",
           b"== ", b" ==\n", b"{{cite ", b"}}", b"&quot;", b"<id>", b"</id>\n"]


def enwik_like(n, seed=0x5EED0001, piece=1 << 22):
    """n bytes of enwik-like text."""
    lens, letters = _vocabulary(seed)
    nwords = lens.size
    ranks = np.arange(1, nwords + 1, dtype=np.float64)
    zipf_cdf = np.cumsum(ranks ** -1.1)
    zipf_cdf /= zipf_cdf[-1]
    mk_len = np.array([len(m) for m in _MARKUP], dtype=np.int64)
    mk_pad = np.zeros((len(_MARKUP), 16), dtype=np.uint8)
    for i, m in enumerate(_MARKUP):
        mk_pad[i, : len(m)] = np.frombuffer(m, dtype=np.uint8)
    out = np.empty(n + piece, dtype=np.uint8)
    pos, counter = 0, 0
    while pos < n:
        ntok = piece // 6
        r = splitmix64(seed, ntok * 2, counter)
        counter += ntok * 2
        u = (r[:ntok] >> np.uint64(11)).astype(np.float64) / float(1 << 53)
        w = np.searchsorted(zipf_cdf, u).clip(0, nwords - 1)
        r2 = r[ntok:]
        kind = (r2 % np.uint64(12)).astype(np.int64)           # 0 -> markup token, 1 -> numeral
        sel = ((r2 >> np.uint64(8)) % np.uint64(len(_MARKUP))).astype(np.int64)
        digits = ((r2 >> np.uint64(16)) % np.uint64(10000)).astype(np.int64)
        tlen = lens[w].copy()
        is_mk = kind == 0
        is_num = kind == 1
        tlen[is_mk] = mk_len[sel[is_mk]]
        tlen[is_num] = 4
        # separator: space, or newline roughly every 80 bytes
        sep_nl = ((r2 >> np.uint64(40)) % np.uint64(14)) == 0
        tot = tlen + 1
        ends = np.cumsum(tot)
        starts = ends - tot
        total = int(ends[-1])
        tok_of = np.repeat(np.arange(ntok), tot)
        within = np.arange(total) - starts[tok_of]
        buf = np.empty(total, dtype=np.uint8)
        wl = within.clip(0, 11)
        buf[:] = letters[w[tok_of], wl]
        m = is_mk[tok_of]
        buf[m] = mk_pad[sel[tok_of[m]], within[m].clip(0, 15)]
        m = is_num[tok_of]
        dg = digits[tok_of[m]]
        p10 = np.array([1000, 100, 10, 1])[within[m].clip(0, 3)]
        buf[m] = (48 + (dg // p10) % 10).astype(np.uint8)
        is_sep = within == tlen[tok_of]
        buf[is_sep] = np.where(sep_nl[tok_of[is_sep]], 10, 32).astype(np.uint8)
        take = min(total, n + piece - pos)
        out[pos: pos + take] = buf[:take]
        pos += take
    return out[:n]


def random_zero_mix(n, seed=0x5EED0005):
    """n bytes: alternating runs (even = uniform random bytes, odd = zeros), lengths in [4 KiB, 256 KiB]."""
    out = np.zeros(n, dtype=np.uint8)
    nruns = n // 4096 + 2
    r = splitmix64(seed, nruns)
    run_len = (4096 + (r % np.uint64(256 * 1024 - 4096 + 1))).astype(np.int64)
    pos, i = 0, 0
    ctr = 0
    while pos < n:
        ln = int(min(run_len[i], n - pos))
        if i % 2 == 0:
            words = splitmix64(seed ^ 0x1234567, (ln + 7) // 8, ctr)
            ctr += (ln + 7) // 8
            out[pos: pos + ln] = words.view(np.uint8)[:ln]
        pos += ln
        i += 1
    return out


def mixed_small(n, seed=1):
    """Test input with text, random bursts, zero runs and long repeats (exercises every block type)."""
    parts = []
    total = 0
    r = splitmix64(seed, 4096)
    k = 0
    text = enwik_like(min(max(n, 1 << 16), 1 << 20), seed=seed + 77)
    while total < n:
        kind = int(r[k % 4096] % np.uint64(5))
        ln = int(64 + (r[(k + 1) % 4096] % np.uint64(20000)))
        k += 2
        if kind == 0:
            p = splitmix64(seed + k, (ln + 7) // 8).view(np.uint8)[:ln]
        elif kind == 1:
            p = np.zeros(ln, dtype=np.uint8)
        elif kind == 2:
            pat = text[(k * 131) % 1000: (k * 131) % 1000 + 1 + int(r[(k + 2) % 4096] % np.uint64(40))]
            p = np.tile(pat, ln // pat.size + 1)[:ln]
        else:
            o = int(r[(k + 3) % 4096] % np.uint64(max(1, text.size - ln)))
            p = text[o: o + ln]
        parts.append(p)
        total += p.size
    return np.concatenate(parts)[:n].copy()
