// sim_inflate.cu -- host-side model of the member-parallel inflate rounds (development aid, not product).
// Runs the SAME lane decoder as the kernel (flate_b200/csrc/inflate_span.cuh, __host__ __device__), lanes one
// after the other, with the kernel's round logic (count pass with warm-up, chain validation with retries,
// budget cut, emit pass, in-order match resolution) and checks the bytes against the expected plain file.
// Prints how often lanes fail to fall into step, so the warm-up length V and span S can be chosen.
//   nvcc -O2 -o /tmp/sim_inflate tools/sim/sim_inflate.cu && /tmp/sim_inflate raw.deflate plain.bin [S] [V] [L]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../flate_b200/csrc/inflate_span.cuh"
using namespace fb;

static std::vector<uint8_t> slurp(const char* p) {
    std::vector<uint8_t> v;
    FILE* f = fopen(p, "rb");
    if (!f) { perror(p); exit(1); }
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + n);
    fclose(f);
    return v;
}

struct Bits {
    const uint8_t* p;
    uint64_t n, pos = 0;
    uint32_t get(uint32_t k) {
        uint32_t v = 0;
        for (uint32_t i = 0; i < k; i++, pos++) v |= (uint32_t)((pos >> 3) < n ? (p[pos >> 3] >> (pos & 7)) & 1 : 0) << i;
        return v;
    }
};

static void build_tables(DecTables& T, const uint8_t* lens, uint32_t n, bool is_lit) {
    uint16_t* count = is_lit ? T.lit_count : T.dist_count;
    uint16_t* symbol = is_lit ? T.lit_sym : T.dist_sym;
    uint32_t* fast = is_lit ? T.lit_fast : T.dist_fast;
    const uint32_t fb = is_lit ? kLitFastBits : kDistFastBits;
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (uint32_t i = 0; i < n; i++) if (lens[i]) count[lens[i]]++;
    uint16_t offs[17];
    offs[1] = 0;
    for (int l = 1; l < 16; l++) offs[l + 1] = offs[l] + count[l];
    for (uint32_t i = 0; i < n; i++) if (lens[i]) symbol[offs[lens[i]]++] = (uint16_t)i;
    for (uint32_t i = 0; i < (1u << fb); i++) fast[i] = 0;
    uint32_t code = 0, index = 0;
    for (uint32_t len = 1; len <= 15; len++) {
        for (uint32_t k = 0; k < count[len]; k++) {
            if (len <= fb) {
                uint32_t c = code + k, rev = 0;
                for (uint32_t b = 0; b < len; b++) rev |= ((c >> b) & 1) << (len - 1 - b);
                const uint32_t e = span_entry(symbol[index + k], len, is_lit);
                for (uint32_t x = rev; x < (1u << fb); x += 1u << len) fast[x] = e;
            }
        }
        code = (code + count[len]) << 1;
        index += count[len];
    }
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: sim_inflate raw.deflate plain.bin [S] [V] [L]\n"); return 2; }
    std::vector<uint8_t> in = slurp(argv[1]), want = slurp(argv[2]);
    uint32_t S0 = argc > 3 ? atoi(argv[3]) : 512, V = argc > 4 ? atoi(argv[4]) : 512, L = argc > 5 ? atoi(argv[5]) : 256;
    const uint32_t kRing = 65536, kBudget = 32752, kQ = 3072, kMaxRetry = 4;
    const uint64_t end_bits = (uint64_t)in.size() * 8;
    in.resize(in.size() + 64, 0);  // loads past the end are guarded by `limit`; padding only for the simple bit reader
    // word view (the kernel reads aligned 32-bit words)
    std::vector<uint32_t> words((in.size() + 3) / 4 + 4, 0);
    memcpy(words.data(), in.data(), in.size());
    std::vector<uint8_t> out;
    out.reserve(want.size() + 1024);
    std::vector<uint8_t> ring(kRing);
    std::vector<uint2> queue(kQ);
    Bits br{in.data(), in.size() - 64};
    static DecTables T;
    uint64_t rounds = 0, retries = 0, dirty_lanes = 0, committed_lanes = 0, cut_budget = 0, exact_tokens = 0, fast_tokens_bytes = 0;
    uint64_t count_bits = 0, emit_bits = 0, blocks = 0, sum_S = 0;
    uint32_t S = S0;
    for (;;) {
        const uint32_t bfinal = br.get(1), btype = br.get(2);
        blocks++;
        if (btype == 0) {
            br.pos = (br.pos + 7) & ~7ull;
            const uint32_t len = br.get(16);
            br.get(16);
            for (uint32_t i = 0; i < len; i++) out.push_back((uint8_t)br.get(8));
        } else {
            uint8_t ll[320];
            memset(ll, 0, sizeof ll);
            uint32_t hlit = 288, hdist = 32;
            if (btype == 2) {
                hlit = br.get(5) + 257;
                hdist = br.get(5) + 1;
                const uint32_t hclen = br.get(4) + 4;
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t cl[19] = {0};
                for (uint32_t i = 0; i < hclen; i++) cl[order[i]] = (uint8_t)br.get(3);
                static DecTables C;
                build_tables(C, cl, 19, false);  // distance slot doubles as the code-length decoder
                for (uint32_t i = 0; i < hlit + hdist;) {
                    uint32_t sym = 0, nb = 0;
                    const uint64_t save = br.pos;
                    const uint32_t peek = br.get(15);
                    br.pos = save;
                    if (!span_slow_find(C.dist_count, C.dist_sym, peek, sym, nb)) { fprintf(stderr, "bad codegen\n"); return 1; }
                    br.pos += nb;
                    if (sym < 16) ll[i++] = (uint8_t)sym;
                    else if (sym == 16) { uint32_t r = 3 + br.get(2); while (r--) { ll[i] = ll[i - 1]; i++; } }
                    else if (sym == 17) i += 3 + br.get(3);
                    else i += 11 + br.get(7);
                }
            } else {
                for (uint32_t i = 0; i < 288; i++) ll[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
                for (uint32_t i = 0; i < 32; i++) ll[288 + i] = 5;
            }
            build_tables(T, ll, hlit, true);
            build_tables(T, ll + hlit, hdist, false);
            span_long_tables(T, 0, T.lit_count);
            span_long_tables(T, 1, T.dist_count);
            // ---- the block's body: rounds ----
            uint64_t cur = br.pos;
            bool done = false;
            while (!done) {
                bool need_exact = true;
                if (end_bits - cur >= 4096) {
                    rounds++;
                    sum_S += S;
                    const uint64_t rb_bits = cur & ~31ull;
                    const uint32_t* wb = words.data() + (rb_bits >> 5);
                    const uint32_t c0 = (uint32_t)(cur - rb_bits);
                    const int64_t lim64 = (int64_t)end_bits - (int64_t)rb_bits - 128;
                    const int32_t limit = (int32_t)(lim64 > (1 << 30) ? (1 << 30) : lim64);
                    std::vector<SpanResult> R(L);
                    for (uint32_t j = 0; j < L; j++) {
                        const uint32_t sj = c0 + j * S, se = c0 + (j + 1) * S;
                        const uint32_t ws = (j == 0 || sj < c0 + V) ? c0 : sj - V;
                        decode_span<false>(T, wb, ws, sj, se, limit, R[j]);
                        count_bits += (R[j].end != kSpanInvalid ? R[j].end : se) - ws;
                    }
                    uint32_t nvalid = 0;
                    for (uint32_t it = 0;; it++) {
                        nvalid = L;
                        for (uint32_t j = 1; j < L; j++) {
                            const bool ok = R[j - 1].flag == kSpanNone && R[j].start != kSpanInvalid && R[j].start == R[j - 1].end;
                            if (!ok) { nvalid = j; break; }
                        }
                        if (nvalid == L || R[nvalid - 1].flag != kSpanNone || it == kMaxRetry) break;
                        retries++;
                        std::vector<SpanResult> N = R;
                        for (uint32_t j = nvalid; j < L; j++) {
                            if (R[j - 1].flag != kSpanNone) continue;
                            if (R[j].start != kSpanInvalid && R[j].start == R[j - 1].end) continue;
                            dirty_lanes++;
                            const uint32_t se = c0 + (j + 1) * S;
                            decode_span<false>(T, wb, R[j - 1].end, R[j - 1].end, se, limit, N[j]);
                        }
                        R = N;
                    }
                    if (R[0].flag == kSpanDead) nvalid = 0;
                    // budget cut
                    uint32_t ncommit = 0, cb = 0, cm = 0;
                    const uint64_t cap_left = want.size() + 64 - out.size();
                    for (uint32_t j = 0; j < nvalid; j++) {
                        if (cb + R[j].bytes > kBudget || cm + R[j].nm > kQ || cb + R[j].bytes > cap_left) { cut_budget++; break; }
                        cb += R[j].bytes;
                        cm += R[j].nm;
                        ncommit = j + 1;
                    }
                    if (ncommit) {
                        const uint64_t pos0 = out.size();
                        uint32_t ob = 0, om = 0, last = ncommit - 1;
                        uint32_t tot_b = 0, tot_m = 0;
                        for (uint32_t j = 0; j < ncommit; j++) {
                            SpanEmit em;
                            em.ring = ring.data();
                            em.ring_mask = kRing - 1;
                            em.slot0 = (uint32_t)(pos0 + ob);
                            em.rel0 = ob;
                            em.queue = queue.data();
                            em.q0 = om;
                            const uint64_t reach = pos0 + ob;
                            em.reach = (uint32_t)(reach > 0xffff0000ull ? 0xffff0000ull : reach);
                            SpanResult E;
                            decode_span<true>(T, wb, R[j].start, R[j].start, R[j].end, limit, E, &em);
                            emit_bits += E.end - E.start;
                            if (E.flag == kSpanBad) {
                                last = j;
                                R[j] = E;
                                R[j].flag = kSpanIrreg;
                                tot_b = ob + E.bytes;
                                tot_m = om + E.nm;
                                break;
                            }
                            if (E.end != R[j].end || E.bytes != R[j].bytes || E.nm != R[j].nm) {
                                fprintf(stderr, "emit/count mismatch lane %u: end %u/%u bytes %u/%u nm %u/%u flag %u/%u\n", j, E.end,
                                        R[j].end, E.bytes, R[j].bytes, E.nm, R[j].nm, E.flag, R[j].flag);
                                return 1;
                            }
                            ob += R[j].bytes;
                            om += R[j].nm;
                            tot_b = ob;
                            tot_m = om;
                        }
                        // in-order match resolution
                        for (uint32_t k = 0; k < tot_m; k++) {
                            const uint32_t rel = queue[k].x, len = queue[k].y >> 16, dist = (queue[k].y & 0xffff) + 1;
                            for (uint32_t i = 0; i < len; i++)
                                ring[(pos0 + rel + i) & (kRing - 1)] = ring[(pos0 + rel + i - dist) & (kRing - 1)];
                        }
                        for (uint32_t i = 0; i < tot_b; i++) out.push_back(ring[(pos0 + i) & (kRing - 1)]);
                        fast_tokens_bytes += tot_b;
                        committed_lanes += last + 1;
                        cur = rb_bits + R[last].end;
                        done = R[last].flag == kSpanEob;
                        need_exact = R[last].flag == kSpanIrreg;
                        const uint64_t bits = R[last].end - c0;
                        if (tot_b && bits) {  // next span: aim at ~90% of the budget
                            uint64_t s = (uint64_t)(kBudget * 9 / 10) * bits / ((uint64_t)L * tot_b);
                            s = s < 64 ? 64 : s > 1024 ? 1024 : s;
                            S = (uint32_t)s;
                        }
                    }
                }
                if (!done && need_exact) {
                    // exact sequential path: a few tokens (all of them near the end of the input)
                    br.pos = cur;
                    const bool unlimited = end_bits - cur < 4096;
                    for (uint32_t t = 0; unlimited || t < 8; t++) {
                        uint32_t sym = 0, nb = 0;
                        uint64_t save = br.pos;
                        uint32_t peek = br.get(15);
                        br.pos = save;
                        if (!span_slow_find(T.lit_count, T.lit_sym, peek, sym, nb)) { fprintf(stderr, "invalid code\n"); return 1; }
                        br.pos += nb;
                        exact_tokens++;
                        if (sym < 256) { ring[out.size() & (kRing - 1)] = (uint8_t)sym; out.push_back((uint8_t)sym); continue; }
                        if (sym == 256) { done = true; break; }
                        const uint32_t length = span_len_base(sym - 257) + br.get(span_len_extra(sym - 257));
                        save = br.pos;
                        peek = br.get(15);
                        br.pos = save;
                        uint32_t dsym = 0;
                        if (!span_slow_find(T.dist_count, T.dist_sym, peek, dsym, nb)) { fprintf(stderr, "invalid dist\n"); return 1; }
                        br.pos += nb;
                        const uint32_t dist = span_dist_base(dsym) + br.get(span_dist_extra(dsym));
                        for (uint32_t i = 0; i < length; i++) {
                            const uint8_t b = ring[(out.size() - dist) & (kRing - 1)];
                            ring[out.size() & (kRing - 1)] = b;
                            out.push_back(b);
                        }
                    }
                    cur = br.pos;
                }
            }
            br.pos = cur;
        }
        // stored blocks bypass the ring in this model: refresh it
        if (btype == 0)
            for (size_t i = out.size() > kRing ? out.size() - kRing : 0; i < out.size(); i++) ring[i & (kRing - 1)] = out[i];
        if (bfinal) break;
    }
    const bool ok = out.size() == want.size() && memcmp(out.data(), want.data(), out.size()) == 0;
    printf("%s: out %zu bytes, blocks %llu, rounds %llu (mean S %.0f), retries %llu, dirty lanes %llu, committed lanes/round %.1f, budget cuts %llu\n",
           ok ? "OK" : "MISMATCH", out.size(), (unsigned long long)blocks, (unsigned long long)rounds, rounds ? (double)sum_S / rounds : 0.0,
           (unsigned long long)retries, (unsigned long long)dirty_lanes, rounds ? (double)committed_lanes / rounds : 0.0,
           (unsigned long long)cut_budget);
    printf("  exact-path tokens %llu, fast bytes %llu; decode work: count %.2fx, emit %.2fx of the stream's bits\n",
           (unsigned long long)exact_tokens, (unsigned long long)fast_tokens_bytes, (double)count_bits / end_bits, (double)emit_bits / end_bits);
    return ok ? 0 : 1;
}
