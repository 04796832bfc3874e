// inflate_span.cuh -- the lane-level symbol decoder of the member-parallel inflate kernel.
//
// One deflate block's Huffman body is decoded by many lanes at once: lane j starts a little before
// bit offset s_j = cur + j * S, runs until it has fallen into step with the real token sequence
// (Huffman/deflate streams self-synchronise after a few tokens), and from the first token boundary
// at or past s_j counts what the tokens up to the first boundary at or past s_(j+1) produce.  Lane 0
// starts at the true cursor; a lane's span is accepted only if it starts exactly where its
// predecessor's span ended, so whatever is committed is the token sequence the sequential loop of
// inflate.zig:220-249 would have decoded.  Only regular tokens are ever committed here (literal,
// or length + distance with valid symbols and all bits present); end of block, invalid symbols,
// short input and capacity are left to the exact sequential path of the kernel, which owns the
// error classes.
//
// The function is __host__ __device__ so that tools/sim/sim_inflate.cpp can run the same code on
// the CPU (lanes simulated one after the other) against zlib.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD inline
#endif

namespace fb {

constexpr uint32_t kLitFastBits = 10, kDistFastBits = 8;
constexpr uint32_t kSpanInvalid = 0xffffffffu;

// direct-table entry: code length (4 bits, 0 = not in the direct table) | symbol << 4 (9 bits) |
// extra-bit count << 13 (4 bits) | base value << 17 (length or distance base; 15 bits)
struct DecTables {
    uint32_t lit_fast[1u << kLitFastBits];    // the two direct tables are adjacent: one load serves both kinds of step
    uint32_t dist_fast[1u << kDistFastBits];
    // codes longer than the direct tables, [0] literal/length, [1] distance: for length l, lim = (first + count)
    // left-justified to 15 bits, first = first canonical code, index = symbols with shorter codes
    uint32_t long_lim[2][16];
    uint16_t long_first[2][16], long_index[2][16];
    uint16_t lit_count[16], dist_count[16];
    uint16_t lit_sym[288], dist_sym[32];
    uint8_t lit_lens[288], dist_lens[32];
};

// fills long_* of one alphabet from its per-length counts (one thread)
FB_HD void span_long_tables(DecTables& T, uint32_t which, const uint16_t* count) {
    uint32_t first = 0, index = 0;
    for (uint32_t len = 1; len <= 15; len++) {
        T.long_first[which][len] = (uint16_t)first;
        T.long_index[which][len] = (uint16_t)index;
        T.long_lim[which][len] = (first + count[len]) << (15 - len);
        index += count[len];
        first = (first + count[len]) << 1;
    }
}

FB_HD uint32_t span_len_base(uint32_t code) {  // RFC 1951 3.2.5, code index 0..28
    if (code < 8) return 3 + code;
    if (code == 28) return 258;
    const uint32_t eb = (code - 4) >> 2;
    return 3 + ((4 + (code & 3)) << eb);
}
FB_HD uint32_t span_len_extra(uint32_t code) { return (code < 8 || code == 28) ? 0 : (code - 4) >> 2; }
FB_HD uint32_t span_dist_base(uint32_t code) {  // code 0..29
    if (code < 4) return 1 + code;
    const uint32_t eb = (code - 2) >> 1;
    return 1 + ((2 + (code & 1)) << eb);
}
FB_HD uint32_t span_dist_extra(uint32_t code) { return code < 4 ? 0 : (code - 2) >> 1; }

FB_HD uint32_t span_entry(uint32_t sym, uint32_t nbits, bool is_lit) {
    uint32_t eb = 0, base_v = 0;
    if (is_lit) {
        if (sym >= 257 && sym <= 285) {
            eb = span_len_extra(sym - 257);
            base_v = span_len_base(sym - 257);
        }
    } else if (sym <= 29) {
        eb = span_dist_extra(sym);
        base_v = span_dist_base(sym);
    }
    return nbits | (sym << 4) | (eb << 13) | (base_v << 17);
}

// canonical decode on an LSB-first peek (huffman_decoder.zig:156-175 find semantics); false = no code matches
FB_HD bool span_slow_find(const uint16_t* count, const uint16_t* symbol, uint32_t peek, uint32_t& sym, uint32_t& nbits) {
    int code = 0, first = 0, index = 0;
    for (uint32_t len = 1; len <= 15; len++) {
        code |= (int)(peek & 1);
        peek >>= 1;
        const int cnt = count[len];
        if (code - cnt < first) {
            sym = symbol[index + (code - first)];
            nbits = len;
            return true;
        }
        index += cnt;
        first += cnt;
        first <<= 1;
        code <<= 1;
    }
    return false;
}

enum SpanFlag : uint32_t {
    kSpanNone = 0,   // ran to the end of its span
    kSpanEob = 1,    // end-of-block symbol decoded inside the counted part; `end` is the bit after it
    kSpanIrreg = 2,  // a token the fast path does not commit starts at `end`
    kSpanDead = 3,   // never reached a usable state (irregular token while still synchronising, or out of input)
    kSpanBad = 4     // emit pass only: match reaches before the start of the output (InvalidMatch candidate) at `end`
};

struct SpanResult {
    uint32_t start;  // first token boundary at or past count_from (kSpanInvalid if dead)
    uint32_t end;    // where the counted part stopped (a token boundary)
    uint32_t bytes;  // output bytes of the counted tokens
    uint32_t nm;     // matches among them
    uint32_t flag;
};

struct SpanEmit {       // emit pass: where this lane's literals and matches go
    uint8_t* ring;      // output window (shared memory on the device)
    uint32_t ring_mask;
    uint32_t slot0;     // ring slot of the lane's first output byte
    uint32_t rel0;      // position of the lane's first output byte relative to the round's first
    uint2* queue;       // match descriptors of the round: x = position relative to the round's first byte,
                        // y = length << 16 | distance - 1
    uint32_t q0;        // this lane's first queue slot
    uint32_t reach;     // bytes a match may reach back from the lane's first output byte (saturated)
};

#if defined(__CUDA_ARCH__)
#define FB_FUNNEL_R(lo, hi, s) __funnelshift_r((lo), (hi), (s))
#define FB_LDG32(p) __ldg(p)
#define FB_STQ(p, v) __stcg((p), (v))
#else
#define FB_STQ(p, v) (*(p) = (v))
#define FB_FUNNEL_R(lo, hi, s) ((uint32_t)(((((uint64_t)(hi)) << 32) | (lo)) >> ((s) & 31)))
#define FB_LDG32(p) (*(p))
#endif

// Decodes tokens from bit `bp` (relative to word pointer `wb`) with the block's tables.
//   count pass (EMIT = false): tokens before the first boundary at or past `count_from` only serve to fall
//     into step; counting stops at the first boundary at or past `span_end`.
//   emit pass (EMIT = true): `bp` is a true boundary, `span_end` the end found by the count pass; literals go
//     to the ring, matches to the queue.
// `limit`: last bit offset at which a token may start (all loads below stay inside the member's input).
// One loop iteration decodes ONE Huffman symbol with its extra bits -- a literal/length symbol, or the distance
// symbol of the match whose length was decoded by the previous iteration -- so that the lanes of a warp, which
// sit at unrelated places of the stream, execute the same instructions whatever kind of token they are in.
template <bool EMIT>
FB_HD void decode_span(const DecTables& T, const uint32_t* __restrict__ wb, uint32_t bp, uint32_t count_from, uint32_t span_end,
                       int32_t limit, SpanResult& r, const SpanEmit* em = nullptr) {
    r.start = kSpanInvalid;
    r.end = kSpanInvalid;
    r.bytes = 0;
    r.nm = 0;
    r.flag = kSpanDead;
    if ((int32_t)bp > limit) return;
    uint32_t idx = bp >> 5, bo = bp & 31;
    uint32_t w0 = FB_LDG32(wb + idx), w1 = FB_LDG32(wb + idx + 1), w2 = FB_LDG32(wb + idx + 2);
    bool counting = EMIT;
    if (EMIT) r.start = bp;
    uint32_t bytes = 0, nm = 0;
    uint32_t want_dist = 0, length = 0, tok = bp;
    uint32_t flag = kSpanDead, end = kSpanInvalid;
    for (;;) {
        if (!want_dist) {
            tok = (idx << 5) + bo;
            if (!EMIT && !counting && tok >= count_from) {
                counting = true;
                r.start = tok;
                bytes = 0;
                nm = 0;
            }
            if (tok >= span_end) {  // count_from <= span_end, so counting is on here
                end = tok;
                flag = kSpanNone;
                break;
            }
            if ((int32_t)tok > limit) {
                end = tok;
                flag = kSpanIrreg;
                break;
            }
        }
        const uint32_t win = FB_FUNNEL_R(w0, w1, bo);
        uint32_t e = T.lit_fast[want_dist ? (1u << kLitFastBits) + (win & ((1u << kDistFastBits) - 1)) : (win & ((1u << kLitFastBits) - 1))];
        uint32_t nb = e & 15;
        if (nb == 0) {  // code longer than the direct table, or no code at all
#if defined(__CUDA_ARCH__)
            const uint32_t code15 = __brev(win) >> 17;
#else
            uint32_t code15 = 0;
            for (int b = 0; b < 15; b++) code15 |= ((win >> b) & 1u) << (14 - b);
#endif
            uint32_t len = (want_dist ? kDistFastBits : kLitFastBits) + 1;
            while (len <= 15 && code15 >= T.long_lim[want_dist][len]) len++;
            uint32_t off = 0xffffffffu, cnt = 0;
            if (len <= 15) {
                off = (code15 >> (15 - len)) - T.long_first[want_dist][len];
                cnt = (want_dist ? T.dist_count : T.lit_count)[len];
            }
            if (off >= cnt) {  // no code matches these bits: the exact path decides what that means
                end = tok;
                flag = kSpanIrreg;
                break;
            }
            const uint32_t sym = (want_dist ? T.dist_sym : T.lit_sym)[T.long_index[want_dist][len] + off];
            e = span_entry(sym, len, !want_dist);
            nb = len;
        }
        const uint32_t sym = (e >> 4) & 511u;
        const uint32_t eb = (e >> 13) & 15u;
        const uint32_t val = (e >> 17) + ((win >> nb) & ((1u << eb) - 1));
        bo += nb + eb;  // at most 15 + 13 bits
        if (bo >= 32) {
            bo -= 32;
            idx++;
            w0 = w1;
            w1 = w2;
            w2 = FB_LDG32(wb + idx + 2);
        }
        if (!want_dist) {
            if (sym < 256) {
                if (EMIT) em->ring[(em->slot0 + bytes) & em->ring_mask] = (uint8_t)sym;
                bytes++;
            } else if (sym == 256) {
                end = tok + nb;
                flag = kSpanEob;
                break;
            } else if (sym > 285) {
                end = tok;
                flag = kSpanIrreg;
                break;
            } else {
                length = val;
                want_dist = 1;
            }
        } else {
            if (sym > 29) {
                end = tok;
                flag = kSpanIrreg;
                break;
            }
            if (EMIT) {
                if (val > em->reach + bytes) {  // CircularBuffer.zig:45: left to the exact path
                    end = tok;
                    flag = kSpanBad;
                    break;
                }
                FB_STQ(em->queue + em->q0 + nm, make_uint2(em->rel0 + bytes, (length << 16) | (val - 1)));
            }
            bytes += length;
            nm++;
            want_dist = 0;
        }
    }
    if (!counting) {  // an irregular token (or the end of the block) while still falling into step
        r.start = kSpanInvalid;
        return;
    }
    r.end = end;
    r.flag = flag;
    r.bytes = bytes;
    r.nm = nm;
}

}  // namespace fb
