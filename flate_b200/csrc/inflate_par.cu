// inflate_par.cu -- inflate on sm_100a, one CTA (8 warps) per member, symbol decode parallel INSIDE the member.
//
// The sequential hot loop of the reference (inflate.zig:220-249 dynamicBlock, :104-124 fixedBlock:
// decode a symbol, emit a literal or copy a match, repeat) is re-derived as rounds over the current
// deflate block:
//   1. count pass   every lane decodes a span of S bits of the block body starting a warm-up of V bits
//                   earlier, so that by the time it reaches its span it has fallen into step with the real
//                   token sequence (inflate_span.cuh); it records where its first token at or past the
//                   span start begins, where its last one ends, and how many bytes / matches they produce;
//   2. validation   lane 0 started at the true cursor; a lane is accepted iff it starts exactly where its
//                   predecessor ended.  Lanes that did not fall into step in time decode again from the
//                   predecessor's end (a few per round); the accepted prefix is what the sequential loop
//                   would have decoded, token for token;
//   3. budget cut   prefix sums of bytes and matches; the longest prefix of lanes that fits the output
//                   window, the match queue and the caller's capacity is committed;
//   4. emit pass    the committed lanes decode once more, writing literals straight to their final
//                   place in the 64 KiB output window (shared memory) and queueing matches;
//   5. resolution   the queue is resolved in stream order, 256 matches at a time: a match copies as soon as
//                   every byte it reads is final (CircularBuffer.zig:44-75 writeMatch, incl. overlapping copies);
//   6. drain        the window is written to HBM in 16-byte coalesced stores.
// Everything that is not a regular token -- block headers, end of block when it is not found by a lane,
// invalid symbols, input running out, capacity, matches reaching before the start of the output -- goes
// through the exact sequential path (warp 0, the bit cursor of bit_reader.zig), which owns the error
// classes and their order.  The fast path never commits a token the exact path
// would not have decoded identically.
#include "inflate_dev.cuh"

namespace fb {

namespace par {

constexpr uint32_t kLanes = 512, kWarps = kLanes / 32;
constexpr uint32_t kRing = 131072;                    // output window: 32 KiB of history + one round
constexpr uint32_t kBudget = kRing - 32768 - 16;      // output bytes a round may commit
constexpr uint32_t kQueue = 16384;                    // matches a round may commit (queue in HBM/L2, one per resident CTA)
constexpr uint32_t kWarm = 640;                       // V: bits decoded before a lane's span to fall into step
constexpr uint32_t kSpanMin = 64, kSpanMax = 2048, kSpanInit = 512;
constexpr uint32_t kMaxRetry = 4;
constexpr uint32_t kFastMinBits = 4096;               // closer to the end of the input only the exact path runs
constexpr uint32_t kTailSlack = 256;                  // bits before the end of the input where lanes stop
constexpr uint32_t kFlushAt = 8192;                   // exact path: drain when this many bytes are pending
constexpr uint32_t kThreadCopy = 16;                  // matches up to this length are copied by one thread

struct Ctrl {                 // written by one thread, read by all after a barrier
    unsigned long long cur;   // cursor as an absolute bit address; always a token / block boundary
    unsigned long long pos;   // bytes produced
    unsigned long long flushed;
    int status;
    uint32_t done;            // end of block seen
    uint32_t bfinal, btype;
    uint32_t span;            // S of the next round
    uint32_t holdoff;         // exact-path invocations to run before the next fast attempt
    uint32_t holdoff_next;
    uint32_t need_exact;
    uint32_t nvalid, ncommit, first_bad;
    uint32_t tot_b, tot_m, last_flag;
    uint32_t stored_len;
    unsigned long long stored_src;
    uint32_t rounds, retries, exact_calls;
    uint32_t member;          // index of the member this CTA works on
    unsigned long long blk_cur, blk_pos;  // cursor and output position at the start of the current block (resume point)
};

struct Shared {
    DecTables T;
    uint32_t l_start[kLanes], l_end[kLanes], l_flag[kLanes];
    uint32_t w_a[kWarps], w_b[kWarps], w_c[kWarps];
    uint32_t pend[(kBudget + 31) / 32 + 1];  // pending-byte bitmap of the round being resolved
    uint32_t cg_fast[128];                   // direct table of the code-length code (7 bits)
    uint32_t cnt32[16], offs32[16];          // build_decoder_warp scratch
    uint32_t crc_tab[256];
    Ctrl c;
};

struct Window {
    uint8_t* ring;
    uint8_t* out;
    uint2* queue;  // this CTA's match queue (global memory; written in the emit pass, read back through L2)
    uint32_t A;
    __device__ __forceinline__ uint32_t slot(uint64_t p) const { return (uint32_t)(p + A) & (kRing - 1); }
};

// drain [flushed, upto) to HBM; all threads.  upto is 16-byte aligned in (p + A) or final.
__device__ void drain_cta(const Window& W, uint64_t flushed, uint64_t upto) {
    const uint32_t tid = threadIdx.x;
    uint64_t f = flushed;
    if (upto <= f) return;
    const uint64_t head_end = min(upto, (f + W.A + 15) / 16 * 16 - W.A);
    for (uint64_t p = f + tid; p < head_end; p += kLanes) W.out[p] = W.ring[W.slot(p)];
    f = head_end;
    const uint64_t nvec = (upto - f) / 16;
    for (uint64_t v = tid; v < nvec; v += kLanes) {
        const uint64_t p = f + v * 16;
        *reinterpret_cast<uint4*>(W.out + p) = *reinterpret_cast<const uint4*>(W.ring + W.slot(p));
    }
    f += nvec * 16;
    for (uint64_t p = f + tid; p < upto; p += kLanes) W.out[p] = W.ring[W.slot(p)];
}

// inclusive scan of two counters over the CTA; s_a/s_b: kWarps words of scratch.  Ends with a barrier.
__device__ __forceinline__ void scan2_cta(uint32_t& a, uint32_t& b, uint32_t* s_a, uint32_t* s_b) {
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b, o);
        if (lane >= (uint32_t)o) { a += ta; b += tb; }
    }
    if (lane == 31) { s_a[w] = a; s_b[w] = b; }
    __syncthreads();
    uint32_t pa = 0, pb = 0;
    for (uint32_t i = 0; i < w; i++) { pa += s_a[i]; pb += s_b[i]; }
    a += pa;
    b += pb;
    __syncthreads();
}

// first thread of the CTA whose predicate is false (kLanes if none); s: kWarps words.  Ends with a barrier.
__device__ __forceinline__ uint32_t first_false_cta(bool ok, uint32_t* s) {
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t m = __ballot_sync(0xffffffffu, !ok);
    if (lane == 0) s[w] = m ? (w * 32 + (uint32_t)__ffs(m) - 1) : kLanes;
    __syncthreads();
    uint32_t r = kLanes;
#pragma unroll
    for (uint32_t i = 0; i < kWarps; i++) r = min(r, s[i]);
    __syncthreads();
    return r;
}

// ---- pending-byte bitmap of a round: bit p is set while output byte p (relative to the round's first byte) is the
// destination of a match that has not been copied yet.  A match may copy as soon as no byte of its source is pending.
__device__ __forceinline__ void pend_update(uint32_t* bm, uint32_t a, uint32_t b, bool set) {  // bits [a, b), a < b
    const uint32_t wa = a >> 5, wb = (b - 1) >> 5;
    for (uint32_t w = wa; w <= wb; w++) {
        uint32_t m = 0xffffffffu;
        if (w == wa) m &= 0xffffffffu << (a & 31);
        if (w == wb) m &= 0xffffffffu >> (31 - ((b - 1) & 31));
        if (set) atomicOr(bm + w, m);
        else atomicAnd(bm + w, ~m);
    }
}
__device__ __forceinline__ bool pend_any(const volatile uint32_t* bm, uint32_t a, uint32_t b) {  // any bit of [a, b) set; a < b
    const uint32_t wa = a >> 5, wb = (b - 1) >> 5;
    uint32_t acc = 0;
    for (uint32_t w = wa; w <= wb; w++) {
        uint32_t m = 0xffffffffu;
        if (w == wa) m &= 0xffffffffu << (a & 31);
        if (w == wb) m &= 0xffffffffu >> (31 - ((b - 1) & 31));
        acc |= bm[w] & m;
    }
    return acc != 0;
}

// Resolve queue[0, M) (stream order) of a round that produced tot_b bytes starting at output position pos0.
// Groups of 32 consecutive matches are dealt round-robin to the warps; a warp copies a match as soon as its source
// bytes are final (pending bitmap), so only true dependency chains serialise.  The earliest unfinished match is
// always ready and its warp is always working on its group (a warp's earlier groups are earlier in the stream),
// hence progress.  Bytes before the member's output (history of earlier members) come from HBM.
__device__ void resolve_matches(Shared& S, const Window& W, uint32_t M, uint64_t pos0, uint32_t tot_b) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t base = (uint32_t)(pos0 + W.A);
    const int64_t floor_rel = -(int64_t)min(pos0, (uint64_t)0x40000000u);  // sources below this are before the member's output
    uint32_t* bm = S.pend;
    for (uint32_t i = tid; i < (tot_b + 31) / 32; i += kLanes) bm[i] = 0;
    __syncthreads();
    for (uint32_t k = tid; k < M; k += kLanes) {
        const uint2 q = __ldcg(W.queue + k);
        pend_update(bm, q.x, q.x + (q.y >> 16), true);
    }
    __syncthreads();
    for (uint32_t g0 = w * 32; g0 < M; g0 += kLanes) {
        const bool have = g0 + lane < M;
        uint32_t dst = 0, len = 0, dist = 1;
        if (have) {
            const uint2 q = __ldcg(W.queue + g0 + lane);
            dst = q.x;
            len = q.y >> 16;
            dist = (q.y & 0xffffu) + 1;
        }
        const int32_t src = (int32_t)dst - (int32_t)dist;
        const int32_t need_end = src + (int32_t)min(len, dist);  // source bytes that must exist before the copy starts
        const bool far = have && (int64_t)src < floor_rel;       // reaches before the member's output
        bool pending = have;
        while (__any_sync(0xffffffffu, pending)) {
            const bool ready = pending && (need_end <= 0 || !pend_any(bm, (uint32_t)max(src, 0), (uint32_t)need_end));
            if (!__any_sync(0xffffffffu, ready)) continue;  // the bytes we wait for belong to another warp's group
            __threadfence_block();  // the bitmap was read before the bytes are
            // short copies: one thread each, byte by byte in order (so overlapping copies replicate)
            if (ready && !far && len <= kThreadCopy) {
                for (uint32_t i = 0; i < len; i++)
                    W.ring[(base + dst + i) & (kRing - 1)] = W.ring[(base + (uint32_t)src + i) & (kRing - 1)];
            }
            // long copies and history copies: the warp together, one match after the other
            uint32_t lm = __ballot_sync(0xffffffffu, ready && (far || len > kThreadCopy));
            while (lm) {
                const int b = __ffs(lm) - 1;
                lm &= lm - 1;
                const uint32_t c_dst = __shfl_sync(0xffffffffu, dst, b), c_len = __shfl_sync(0xffffffffu, len, b);
                const uint32_t c_dist = __shfl_sync(0xffffffffu, dist, b);
                const int32_t c_src = (int32_t)c_dst - (int32_t)c_dist;
                if ((int64_t)c_src < floor_rel) {
                    // source (partly) before the member's output
                    for (uint32_t i = lane; i < c_len; i += 32) {
                        const uint32_t k = c_dist >= c_len ? i : i % c_dist;
                        const int64_t sp = (int64_t)pos0 + c_src + k;
                        const uint8_t v = sp < 0 ? __ldcg(W.out + sp) : W.ring[W.slot((uint64_t)sp)];
                        W.ring[(base + c_dst + i) & (kRing - 1)] = v;
                    }
                } else if (c_dist >= 32 || c_dist >= c_len) {
                    // 32 bytes per step; a step only reads bytes that existed before it (distance >= 32) or the copy does not overlap
                    for (uint32_t i0 = 0; i0 < c_len; i0 += 32) {
                        const uint32_t i = i0 + lane;
                        if (i < c_len) W.ring[(base + c_dst + i) & (kRing - 1)] = W.ring[(base + (uint32_t)c_src + i) & (kRing - 1)];
                        __syncwarp();
                    }
                } else {
                    // short period: every byte is a copy of one of the first c_dist source bytes
                    for (uint32_t i = lane; i < c_len; i += 32)
                        W.ring[(base + c_dst + i) & (kRing - 1)] = W.ring[(base + (uint32_t)c_src + i % c_dist) & (kRing - 1)];
                }
                __syncwarp();
            }
            __threadfence_block();  // the bytes are written before the bitmap says so
            if (ready) {
                pend_update(bm, dst, dst + len, false);
                pending = false;
            }
        }
    }
    __syncthreads();
}

// CRC-32 / Adler-32 of the member's output (in HBM, written by this CTA), all threads; result in every thread
__device__ uint32_t cta_crc32(Shared& S, const uint8_t* p, uint64_t n) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint64_t per = ((n + kLanes - 1) / kLanes + 15) & ~15ull;
    const uint64_t b = min(per * tid, n), e = min(b + per, n);
    uint32_t term = 0;
    if (e > b) term = multmodp(x8nmodp(n - e), crc32_chunk(S.crc_tab, p + b, e - b));
    for (int o = 16; o > 0; o >>= 1) term ^= __shfl_xor_sync(0xffffffffu, term, o);
    if (lane == 0) S.w_a[w] = term;
    __syncthreads();
    uint32_t r = 0;
    for (uint32_t i = 0; i < kWarps; i++) r ^= S.w_a[i];
    __syncthreads();
    return r;
}
__device__ uint32_t cta_adler32(Shared& S, const uint8_t* p, uint64_t n) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint64_t per = (n + kLanes - 1) / kLanes;
    const uint64_t b = min(per * tid, n), e = min(b + per, n);
    uint64_t a = 0, s = 0;
    for (uint64_t i = b; i < e; i++) {
        a += p[i];
        s += a;
        if ((i & 2047) == 2047) { a %= 65521; s %= 65521; }
    }
    a %= 65521;
    s %= 65521;
    uint32_t A = (uint32_t)a, B = (uint32_t)((s + a * ((n - e) % 65521)) % 65521);
    for (int o = 16; o > 0; o >>= 1) {
        A += __shfl_xor_sync(0xffffffffu, A, o);
        B += __shfl_xor_sync(0xffffffffu, B, o);
    }
    if (lane == 0) { S.w_a[w] = A % 65521; S.w_b[w] = B % 65521; }
    __syncthreads();
    uint64_t ta = 0, tb = 0;
    for (uint32_t i = 0; i < kWarps; i++) { ta += S.w_a[i]; tb += S.w_b[i]; }
    __syncthreads();
    ta = (ta + 1) % 65521;
    tb = (tb + n % 65521) % 65521;
    return (uint32_t)((tb << 16) | ta);
}

#define BCAST(x) __shfl_sync(0xffffffffu, (x), 0)

// ---- exact sequential path, warp 0 (inflate.zig:220-239 / :104-124 with the cursor rules of bit_reader.zig).
// Decodes from S.c.cur until: end of block, an error, `max_tokens` tokens, or kFlushAt bytes pending.
__device__ void exact_tokens(Shared& S, const Window& W, BitCursor& bc, const MemberDesc& md, uint32_t max_tokens) {
    const uint32_t lane = threadIdx.x & 31;
    uint64_t pos = S.c.pos;
    const uint64_t flushed = S.c.flushed, cap = md.out_cap;
    int status = FB200_OK;
    bool done = false;
    uint32_t tokens = 0;
    if (lane == 0) bc.seek(S.c.cur);
    for (;;) {
        uint32_t ev_len = 0, ev_dist = 0;
        bool yield = false;
        if (lane == 0) {
            for (;;) {
                if (pos - flushed >= kFlushAt || tokens >= max_tokens) { yield = true; break; }
                if (bc.empty()) { status = FB200_END_OF_STREAM; break; }  // fill(15) / fill(7+2)
                uint32_t sym, nb;
                const uint32_t e = S.T.lit_fast[bc.peek(kLitFast)];
                if (e & 15) {
                    sym = (e >> 4) & 511u;
                    nb = e & 15;
                } else {
                    status = slow_find(S.T.lit_count, S.T.lit_sym, 15, bc.peek(15), sym, nb);
                    if (status) break;
                }
                if (!bc.shift(nb)) { status = FB200_END_OF_STREAM; break; }
                tokens++;
                if (sym < 256) {
                    if (pos >= cap) { status = FB200_NO_SPACE_LEFT; break; }
                    W.ring[W.slot(pos)] = (uint8_t)sym;
                    pos++;
                    continue;
                }
                if (sym == 256) { done = true; break; }
                // match: fill(5+15+13), decodeLength, distance symbol, decodeDistance.  Symbols 286/287 only exist in
                // the fixed code, where the reference rejects them before any fill (inflate.zig:111).
                const uint32_t lcode = sym - 257;
                if (lcode > 28) { status = FB200_INVALID_CODE; break; }
                if (bc.empty()) { status = FB200_END_OF_STREAM; break; }
                uint32_t length = c_len_base[lcode];
                const uint32_t leb = length_extra_bits(lcode);
                if (leb) {
                    const uint32_t x = bc.peek(leb);
                    if (!bc.shift(leb)) { status = FB200_END_OF_STREAM; break; }
                    length += x;
                }
                uint32_t dsym, dnb;
                bc.refill();
                const uint32_t de = S.T.dist_fast[bc.peek(kDistFast)];
                if (de & 15) {
                    dsym = (de >> 4) & 511u;
                    dnb = de & 15;
                } else {
                    status = slow_find(S.T.dist_count, S.T.dist_sym, 15, bc.peek(15), dsym, dnb);
                    if (status) break;
                }
                if (!bc.shift(dnb)) { status = FB200_END_OF_STREAM; break; }
                if (dsym > 29) { status = FB200_INVALID_CODE; break; }
                uint32_t distance = c_dist_base[dsym];
                const uint32_t deb = distance_extra_bits(dsym);
                if (deb) {
                    const uint32_t x = bc.peek(deb);
                    if (!bc.shift(deb)) { status = FB200_END_OF_STREAM; break; }
                    distance += x;
                }
                if (md.hist + pos < distance) { status = FB200_INVALID_MATCH; break; }  // CircularBuffer.zig:45-50
                if (pos + length > cap) { status = FB200_NO_SPACE_LEFT; break; }
                ev_len = length;
                ev_dist = distance;
                break;
            }
        }
        __syncwarp();
        status = BCAST(status);
        done = BCAST(done);
        yield = BCAST(yield);
        pos = __shfl_sync(0xffffffffu, (unsigned long long)pos, 0);
        ev_len = BCAST(ev_len);
        if (ev_len) {
            ev_dist = BCAST(ev_dist);
            if (ev_dist <= pos) {
                const uint64_t from = pos - ev_dist;
                if (ev_dist >= 32 || ev_dist >= ev_len) {
                    for (uint32_t i0 = 0; i0 < ev_len; i0 += 32) {
                        const uint32_t i = i0 + lane;
                        if (i < ev_len) W.ring[W.slot(pos + i)] = W.ring[W.slot(from + i)];
                        __syncwarp();
                    }
                } else {
                    for (uint32_t i = lane; i < ev_len; i += 32) W.ring[W.slot(pos + i)] = W.ring[W.slot(from + i % ev_dist)];
                }
            } else {
                // reaches into an earlier member's output (HBM, before W.out)
                for (uint32_t i = lane; i < ev_len; i += 32) {
                    const int64_t sp = (int64_t)pos - ev_dist + (ev_dist >= ev_len ? i : i % ev_dist);
                    W.ring[W.slot(pos + i)] = sp < 0 ? __ldcg(W.out + sp) : W.ring[W.slot((uint64_t)sp)];
                }
            }
            pos += ev_len;
            __syncwarp();
        }
        if (status || done || yield) break;
    }
    if (lane == 0) {
        S.c.pos = pos;
        S.c.status = status;
        S.c.done = done ? 1u : 0u;
        S.c.cur = bc.bit_address();
    }
}

// huffman_decoder.zig:126-153 checkCompletnes + canonical tables, warp-parallel: per-length counts with shared-memory
// atomics, the validation on one lane (15 steps), then a stable counting sort of the symbols 32 at a time
// (__match_any_sync ranks the lanes that share a code length).  Returns the status (uniform).  Same validation order
// as build_decoder (inflate_dev.cuh): MissingEndOfBlockCode, Oversubscribed, Incomplete.
__device__ int build_decoder_warp(Shared& S, const uint8_t* lens, uint32_t n, bool is_lit, uint32_t max_code_bits, uint16_t* count,
                                  uint16_t* symbol, uint32_t* fast, uint32_t fast_bits) {
    const uint32_t lane = threadIdx.x & 31;
    int status = FB200_OK;
    uint32_t* cnt32 = S.cnt32;
    uint32_t* offs = S.offs32;
    if (lane < 16) cnt32[lane] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) {
        const uint32_t l = lens[i];
        if (l) atomicAdd(cnt32 + l, 1u);
    }
    __syncwarp();
    if (lane == 0) {
        if (is_lit && lens[256] == 0) status = FB200_MISSING_END_OF_BLOCK_CODE;  // :127-128
        if (status == FB200_OK) {
            uint32_t mx = 0;
            for (uint32_t len = 1; len < 16; len++)
                if (cnt32[len]) mx = len;
            if (mx != 0) {
                int left = 1;
                for (uint32_t len = 1; len <= max_code_bits; len++) {
                    left <<= 1;
                    if ((int)cnt32[len] > left) { status = FB200_OVERSUBSCRIBED_HUFFMAN_TREE; break; }
                    left -= (int)cnt32[len];
                }
                if (status == FB200_OK && left > 0) {
                    // incomplete is tolerated only for a single one-bit code in the literal and distance alphabets
                    if (!(max_code_bits > 7 && mx == cnt32[1])) status = FB200_INCOMPLETE_HUFFMAN_TREE;  // :148-151
                }
            }
            count[0] = 0;
            offs[0] = 0;
            offs[1] = 0;
            for (uint32_t len = 1; len < 16; len++) {
                count[len] = (uint16_t)cnt32[len];
                if (len < 15) offs[len + 1] = offs[len] + cnt32[len];
            }
        }
    }
    __syncwarp();
    status = __shfl_sync(0xffffffffu, status, 0);
    if (status != FB200_OK) return status;
    for (uint32_t c0 = 0; c0 < n; c0 += 32) {
        const uint32_t i = c0 + lane;
        const uint32_t l = i < n ? lens[i] : 0;
        const uint32_t same = __match_any_sync(0xffffffffu, l);
        const uint32_t rank = __popc(same & ((1u << lane) - 1));
        if (l) symbol[offs[l] + rank] = (uint16_t)i;
        __syncwarp();
        if (l && rank == 0) offs[l] += __popc(same);
        __syncwarp();
    }
    if (fast == nullptr) return status;
    for (uint32_t i = lane; i < (1u << fast_bits); i += 32) fast[i] = 0;
    __syncwarp();
    uint32_t code = 0, index = 0;
    for (uint32_t len = 1; len <= fast_bits && len <= max_code_bits; len++) {
        const uint32_t cnt = count[len];
        // symbols symbol[index .. index+cnt) have codes code .. code+cnt-1 (MSB-first)
        for (uint32_t k = lane; k < cnt; k += 32) {
            const uint32_t rev = __brev(code + k) >> (32 - len);
            const uint32_t entry = span_entry(symbol[index + k], len, is_lit);
            for (uint32_t e = rev; e < (1u << fast_bits); e += (1u << len)) fast[e] = entry;
        }
        code = (code + cnt) << 1;
        index += cnt;
    }
    __syncwarp();
    return status;
}

// ---- block header, warp 0 (inflate.zig:251-268 step, :144-185 dynamicBlockHeader); leaves the tables in S.T
__device__ void block_header(Shared& S, BitCursor& bc, const MemberDesc& md, bool& fixed_ready) {
    const uint32_t lane = threadIdx.x & 31;
    int status = FB200_OK;
    uint32_t bfinal = 0, btype = 0;
    DecTables& T = S.T;
    if (lane == 0) {
        bc.seek(S.c.cur);
        status = bc.read(1, bfinal);
        if (!status) status = bc.read(2, btype);
    }
    status = BCAST(status);
    bfinal = BCAST(bfinal);
    btype = BCAST(btype);
    if (!status && btype == 0) {  // stored block, inflate.zig:89-102
        if (lane == 0) {
            bc.align_to_byte();
            uint32_t len = 0, nlen = 0;
            status = bc.read(16, len);
            if (!status) status = bc.read(16, nlen);
            if (!status && len != ((~nlen) & 0xffffu)) status = FB200_WRONG_STORED_BLOCK_NLEN;
            if (!status) {
                const uint8_t* src = bc.byte_pos();
                if ((uint64_t)(bc.end - src) < len) status = FB200_END_OF_STREAM;
                else if (S.c.pos + len > md.out_cap) status = FB200_NO_SPACE_LEFT;
                S.c.stored_len = len;
                S.c.stored_src = (unsigned long long)(uintptr_t)src;
            }
        }
        status = BCAST(status);
    } else if (!status && btype == 2) {
        fixed_ready = false;
        uint32_t hlit = 0, hdist = 0;
        if (lane == 0) {
            uint32_t v = 0, hclen = 0;
            status = bc.read(5, v); hlit = v + 257;
            if (!status) { status = bc.read(5, v); hdist = v + 1; }
            if (!status) { status = bc.read(4, v); hclen = v + 4; }
            if (!status && (hlit > 286 || hdist > 30)) status = FB200_INVALID_DYNAMIC_BLOCK_HEADER;
            if (!status) {  // code-length code lengths go to dist_lens[0..19) temporarily
                for (uint32_t i = 0; i < 19; i++) T.dist_lens[i] = 0;
                for (uint32_t i = 0; i < hclen && !status; i++) {
                    status = bc.read(3, v);
                    T.dist_lens[c_cl_order[i]] = (uint8_t)v;
                }
            }
        }
        status = BCAST(status);
        // CodegenDecoder(19, 7, 7): built into dist_count / dist_sym (no fast table)
        if (!status) status = build_decoder_warp(S, T.dist_lens, 19, false, 7, T.dist_count, T.dist_sym, S.cg_fast, 7);
        if (!status) {
            if (lane == 0) {
                // two passes: literal lengths then distance lengths (inflate.zig:161-180)
                for (uint32_t i = 0; i < kNumLit; i++) T.lit_lens[i] = 0;
                uint8_t dl[kNumDist];
                for (uint32_t i = 0; i < kNumDist; i++) dl[i] = 0;
                for (int pass = 0; pass < 2 && !status; pass++) {
                    uint8_t* lens = pass == 0 ? T.lit_lens : dl;
                    const uint32_t lens_len = pass == 0 ? kNumLit : kNumDist;
                    const uint32_t want = pass == 0 ? hlit : hdist;
                    uint32_t p = 0;
                    while (p < want && !status) {
                        if (bc.empty()) { status = FB200_END_OF_STREAM; break; }  // peekF(u7): fill(7)
                        const uint32_t ce = S.cg_fast[bc.peek(7)];  // every code of the 7-bit alphabet is in the direct table
                        const uint32_t sym = (ce >> 4) & 511u, nb = ce & 15;
                        if (nb == 0) { status = FB200_INVALID_CODE; break; }
                        if (!bc.shift(nb)) { status = FB200_END_OF_STREAM; break; }
                        if (p >= lens_len) { status = FB200_INVALID_DYNAMIC_BLOCK_HEADER; break; }  // inflate.zig:189-216
                        uint32_t v = 0;
                        if (sym <= 15) {
                            lens[p] = (uint8_t)sym;
                            p += 1;
                        } else if (sym == 16) {
                            status = bc.read(2, v);
                            if (status) break;
                            const uint32_t rep = v + 3;
                            if (p == 0 || p + rep > lens_len) { status = FB200_INVALID_DYNAMIC_BLOCK_HEADER; break; }
                            for (uint32_t i = 0; i < rep; i++) lens[p + i] = lens[p + i - 1];
                            p += rep;
                        } else if (sym == 17) {
                            status = bc.read(3, v);
                            if (status) break;
                            p += v + 3;
                        } else {
                            status = bc.read(7, v);
                            if (status) break;
                            p += v + 11;
                        }
                    }
                    if (!status && p > want) status = FB200_INVALID_DYNAMIC_BLOCK_HEADER;
                }
                for (uint32_t i = 0; i < kNumDist; i++) T.dist_lens[i] = dl[i];
            }
            status = BCAST(status);
            __syncwarp();
        }
        if (!status) status = build_decoder_warp(S, T.lit_lens, kNumLit, true, 15, T.lit_count, T.lit_sym, T.lit_fast, kLitFast);
        if (!status) status = build_decoder_warp(S, T.dist_lens, kNumDist, false, 15, T.dist_count, T.dist_sym, T.dist_fast, kDistFast);
    } else if (!status && btype == 1) {
        if (!fixed_ready) {
            // fixed block: the reference decodes by arithmetic (bit_reader.zig:205-217); the same symbols come out of
            // the canonical code with lengths 8/9/7/8 over 288 symbols and 32 five-bit distance codes.  286/287 and
            // 30/31 decode and are then rejected (inflate.zig:111,136).
            build_fixed_lens(T.lit_lens, T.dist_lens);
            status = build_decoder_warp(S, T.lit_lens, 288, true, 15, T.lit_count, T.lit_sym, T.lit_fast, kLitFast);
            if (!status) status = build_decoder_warp(S, T.dist_lens, 32, false, 15, T.dist_count, T.dist_sym, T.dist_fast, kDistFast);
            fixed_ready = !status;
        }
    } else if (!status) {
        status = FB200_INVALID_BLOCK_TYPE;  // inflate.zig:267
    }
    if (!status && btype != 0 && lane < 2) span_long_tables(T, lane, lane ? T.dist_count : T.lit_count);
    if (lane == 0) {
        S.c.status = status;
        S.c.bfinal = bfinal;
        S.c.btype = btype;
        S.c.done = 0;
        S.c.cur = bc.bit_address();
    }
}

// ---- one round of the lane-parallel symbol decode; all threads.  Updates S.c.{cur,pos,done,need_exact,span,...}.
__device__ void fast_round(Shared& S, const Window& W, const MemberDesc& md, uint64_t end_bits) {
    const uint32_t tid = threadIdx.x;
    const unsigned long long cur = S.c.cur;
    const uint64_t pos0 = S.c.pos;
    const uint32_t span = S.c.span;
    const unsigned long long rb_bits = cur & ~31ull;
    const uint32_t* wb = reinterpret_cast<const uint32_t*>((uintptr_t)(rb_bits >> 3));
    const uint32_t c0 = (uint32_t)(cur - rb_bits);
    const long long lim64 = (long long)end_bits - (long long)rb_bits - (long long)kTailSlack;
    const int32_t limit = (int32_t)(lim64 > (1ll << 30) ? (1ll << 30) : lim64);
    const uint32_t sj = c0 + tid * span, se = sj + span;
    const uint32_t ws = (tid == 0 || sj < c0 + kWarm) ? c0 : sj - kWarm;
    __syncthreads();  // everybody has read the control block

    // 1. count pass
    SpanResult r;
    decode_span<false>(S.T, wb, ws, sj, se, limit, r);
    S.l_start[tid] = r.start;
    S.l_end[tid] = r.end;
    S.l_flag[tid] = r.flag;
    __syncthreads();

    // 2. validation with retries
    uint32_t nvalid = 0;
    for (uint32_t it = 0;; it++) {
        const uint32_t pend = tid ? S.l_end[tid - 1] : 0, pflag = tid ? S.l_flag[tid - 1] : 0;
        const bool ok = tid == 0 ? r.flag != kSpanDead : (pflag == kSpanNone && r.start != kSpanInvalid && r.start == pend);
        nvalid = first_false_cta(ok, S.w_a);
        if (nvalid == kLanes || nvalid == 0 || it == kMaxRetry) break;
        if (S.l_flag[nvalid - 1] != kSpanNone) break;  // the chain ends at a flagged lane
        const bool redo = !ok && tid >= nvalid && pflag == kSpanNone && pend != kSpanInvalid;
        if (redo) decode_span<false>(S.T, wb, pend, pend, se, limit, r);
        __syncthreads();  // predecessors' values were read above
        if (redo) {
            S.l_start[tid] = r.start;
            S.l_end[tid] = r.end;
            S.l_flag[tid] = r.flag;
        }
        if (tid == 0) S.c.retries++;
        __syncthreads();
    }

    // 3. budget cut
    uint32_t ib = tid < nvalid ? r.bytes : 0, im = tid < nvalid ? r.nm : 0;
    scan2_cta(ib, im, S.w_a, S.w_b);
    const uint64_t cap_left = md.out_cap - pos0;
    const uint32_t budget = (uint32_t)min((uint64_t)kBudget, cap_left);
    const bool fits = tid < nvalid && ib <= budget && im <= kQueue;
    const uint32_t ncommit = first_false_cta(fits, S.w_a);
    if (tid == 0) {
        S.c.rounds++;
        S.c.first_bad = kLanes;
        S.c.ncommit = ncommit;
        if (ncommit == 0) {
            // nothing fits: either lane 0 is unusable (exact path decides why) or its span alone is too productive
            S.c.need_exact = 1;
            if (nvalid > 0 && r.flag != kSpanDead) {
                if (span > kSpanMin) {
                    S.c.span = max(kSpanMin, span / 4);
                } else {
                    S.c.holdoff = S.c.holdoff_next;
                    S.c.holdoff_next = min(64u, S.c.holdoff_next * 2);
                }
            }
        }
    }
    __syncthreads();
    if (ncommit == 0) return;

    // 4. emit pass
    const uint32_t eb = ib - (tid < nvalid ? r.bytes : 0), em_ = im - (tid < nvalid ? r.nm : 0);  // exclusive
    SpanResult e = r;
    if (tid < ncommit) {
        SpanEmit em;
        em.ring = W.ring;
        em.ring_mask = kRing - 1;
        em.slot0 = (uint32_t)(pos0 + W.A) + eb;
        em.rel0 = eb;
        em.queue = W.queue;
        em.q0 = em_;
        const uint64_t reach = md.hist + pos0 + eb;
        em.reach = (uint32_t)min(reach, (uint64_t)0xffff0000u);
        decode_span<true>(S.T, wb, r.start, r.start, r.end, limit, e, &em);
        if (e.flag == kSpanBad) atomicMin(&S.c.first_bad, tid);
    }
    __syncthreads();
    const uint32_t last = min(S.c.first_bad, ncommit - 1);
    if (tid == last) {
        const bool bad = e.flag == kSpanBad;
        S.c.tot_b = eb + (bad ? e.bytes : r.bytes);
        S.c.tot_m = em_ + (bad ? e.nm : r.nm);
        S.c.last_flag = bad ? (uint32_t)kSpanIrreg : r.flag;
        const uint32_t end_rel = bad ? e.end : r.end;
        S.c.cur = rb_bits + end_rel;
        S.c.pos = pos0 + S.c.tot_b;
        S.c.done = (!bad && r.flag == kSpanEob) ? 1u : 0u;
        S.c.need_exact = (bad || r.flag == kSpanIrreg) ? 1u : 0u;
        S.c.holdoff_next = 1;
        const uint32_t bits = end_rel - c0;
        if (S.c.tot_b && bits) {  // next span: aim at 90 % of the budget
            unsigned long long s = (unsigned long long)(kBudget * 9 / 10) * bits / ((unsigned long long)kLanes * S.c.tot_b);
            S.c.span = (uint32_t)(s < kSpanMin ? kSpanMin : s > kSpanMax ? kSpanMax : s);
        }
    }
    __syncthreads();

    // 5. resolution
    resolve_matches(S, W, S.c.tot_m, pos0, S.c.tot_b);
}

}  // namespace par

using namespace par;

// one member, all threads of the CTA
// kPieces: the member may be one piece of a member that a streaming decompressor decodes bit by bit (MemberDesc::flags,
// start_bit); the batch path is compiled without any of that
template <bool kPieces>
static __device__ void inflate_one_member(Shared& S, uint8_t* ring, uint2* queue, int container, const uint8_t* __restrict__ d_in,
                                          const MemberDesc md, uint8_t* d_out, MemberResult* __restrict__ result) {
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    Window W;
    W.ring = ring;
    W.queue = queue;
    W.out = d_out + md.out_off;
    W.A = (uint32_t)((uintptr_t)W.out & 15);
    const uint8_t* in_begin = d_in + md.in_off;
    const uint64_t end_bits = (uint64_t)(uintptr_t)(in_begin + md.in_len) * 8ull;

    BitCursor bc;  // meaningful in thread 0 only
    bc.next = in_begin;
    bc.end = in_begin + md.in_len;
    bc.buf = 0;
    bc.cnt = 0;

    const bool resume = kPieces && (md.flags & kMemberResume) != 0, partial = kPieces && (md.flags & kMemberPartial) != 0;
    const bool nofooter = kPieces && (md.flags & kMemberNoFooter) != 0;
    // ---- container header (container.zig:111-152), thread 0 ----
    if (tid == 0) {
        int status = FB200_OK;
        S.c.blk_cur = (unsigned long long)(uintptr_t)in_begin * 8ull + (kPieces ? md.start_bit : 0u);
        S.c.blk_pos = 0;
        if (kPieces && md.start_bit) {
            uint32_t skip;
            status = bc.read(md.start_bit, skip);
        }
        if (container != FB200_RAW && !resume && !status) {
            uint32_t v;
            if (container == FB200_GZIP) {
                uint32_t magic1 = 0, magic2 = 0, method = 0, flags = 0;
                if (!status) status = bc.read(8, magic1);
                if (!status) status = bc.read(8, magic2);
                if (!status) status = bc.read(8, method);
                if (!status) status = bc.read(8, flags);
                for (int i = 0; i < 6 && !status; i++) status = bc.read(8, v);
                if (!status && (magic1 != 0x1f || magic2 != 0x8b || method != 0x08)) status = FB200_BAD_GZIP_HEADER;
                if (!status && (flags & 0x04)) {
                    uint32_t xlen = 0;
                    status = bc.read(16, xlen);
                    for (uint32_t i = 0; i < xlen && !status; i++) status = bc.read(8, v);
                }
                if (!status && (flags & 0x08)) do { status = bc.read(8, v); } while (!status && v != 0);
                if (!status && (flags & 0x10)) do { status = bc.read(8, v); } while (!status && v != 0);
                if (!status && (flags & 0x02)) {
                    status = bc.read(8, v);
                    if (!status) status = bc.read(8, v);
                }
            } else {
                uint32_t cm = 0, cinfo = 0;
                status = bc.read(4, cm);
                if (!status) status = bc.read(4, cinfo);
                if (!status) status = bc.read(8, v);
                if (!status && (cm != 8 || cinfo > 7)) status = FB200_BAD_ZLIB_HEADER;
            }
        }
        S.c.cur = bc.bit_address();
        S.c.pos = 0;
        S.c.flushed = 0;
        S.c.status = status;
        S.c.done = 0;
        S.c.span = kSpanInit;
        S.c.holdoff = 0;
        S.c.holdoff_next = 1;
        S.c.need_exact = 0;
        S.c.ncommit = 0;
        S.c.rounds = 0;
        S.c.retries = 0;
        S.c.exact_calls = 0;
    }
    __syncthreads();

    bool fixed_ready = false;
    int status = S.c.status;
    while (status == FB200_OK) {  // inflate.zig:251-280 step: one deflate block per iteration
        __syncthreads();  // every warp has read the control block of the previous block before warp 0 rewrites it
        if (kPieces && tid == 0) {   // a block boundary: where a partial call can be resumed
            S.c.blk_cur = S.c.cur;
            S.c.blk_pos = S.c.pos;
        }
        if (warp == 0) block_header(S, bc, md, fixed_ready);
        __syncthreads();
        status = S.c.status;
        if (status) break;
        const uint32_t bfinal = S.c.bfinal, btype = S.c.btype;
        if (btype == 0) {
            const uint32_t len = S.c.stored_len;
            const uint8_t* src = reinterpret_cast<const uint8_t*>((uintptr_t)S.c.stored_src);
            uint64_t pos = S.c.pos, flushed = S.c.flushed;
            __syncthreads();
            for (uint32_t done = 0; done < len;) {  // through the window in pieces so that history stays valid
                const uint32_t piece = min(len - done, 16384u);
                for (uint32_t i = tid; i < piece; i += kLanes) W.ring[W.slot(pos + i)] = src[done + i];
                __syncthreads();
                pos += piece;
                done += piece;
                const uint64_t tail = (pos + W.A) & 15;  // bytes of the last, incomplete 16-byte line stay in the window
                const uint64_t upto = pos > tail ? pos - tail : 0;  // (a member that starts unaligned and has produced less than a line)
                if (upto > flushed) {
                    drain_cta(W, flushed, upto);
                    flushed = upto;
                }
                __syncthreads();
            }
            if (tid == 0) {
                S.c.pos = pos;
                S.c.flushed = flushed;
                S.c.cur = (unsigned long long)(uintptr_t)(src + len) * 8ull;
            }
            __syncthreads();
        } else {
            // ---- symbol loop ----
            for (;;) {
                const bool fast_ok = end_bits - S.c.cur >= kFastMinBits && S.c.holdoff == 0;
                __syncthreads();
                if (fast_ok) {
                    fast_round(S, W, md, end_bits);
                    __syncthreads();
                } else if (tid == 0) {
                    S.c.need_exact = 1;
                    if (S.c.holdoff) S.c.holdoff--;
                }
                __syncthreads();
                if (!S.c.done && S.c.need_exact) {
                    const bool long_run = !fast_ok || S.c.ncommit == 0;
                    __syncthreads();
                    if (warp == 0) exact_tokens(S, W, bc, md, long_run ? 0xffffffffu : 4u);
                    if (tid == 0) { S.c.need_exact = 0; S.c.exact_calls++; }
                    __syncthreads();
                }
                // drain what is complete
                {
                    const uint64_t pos = S.c.pos, flushed = S.c.flushed;
                    const uint64_t tail = (pos + W.A) & 15;  // bytes of the last, incomplete 16-byte line stay in the window
                    const uint64_t upto = pos > tail ? pos - tail : 0;  // (nothing complete yet when pos + A < 16)
                    __syncthreads();
                    if (upto > flushed) {
                        drain_cta(W, flushed, upto);
                        if (tid == 0) S.c.flushed = upto;
                    }
                    __syncthreads();
                }
                status = S.c.status;
                if (status || S.c.done) break;
            }
            if (status) break;
        }
        if (bfinal) break;
    }
    __syncthreads();
    uint32_t info = status == FB200_OK ? 1u : 0u;  // the loop only ends without an error after the final block
    if (status != FB200_OK && partial) {
        // back to the start of the block that could not be completed: the caller resumes there with more input / room
        info = ((uint32_t)status & 0xffu) << 8;
        status = FB200_OK;
        if (tid == 0) {
            S.c.pos = S.c.blk_pos;
            S.c.cur = S.c.blk_cur;
        }
        __syncthreads();
    }
    const uint64_t pos = S.c.pos;
    drain_cta(W, S.c.flushed, pos);  // whatever was produced, also on error (the caller sees out_len and the status)
    __syncthreads();

    // ---- checksum of what was produced, then the protocol footer (inflate.zig:271-275, container.zig:154-166) ----
    uint32_t sum = 0;
    if (status == FB200_OK && container != FB200_RAW) {
        if (container == FB200_GZIP) {
            for (uint32_t i = tid; i < 256; i += kLanes) {
                uint32_t c = i;
                for (int b = 0; b < 8; b++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
                S.crc_tab[i] = c;
            }
            __syncthreads();
            sum = cta_crc32(S, W.out, pos);
        } else {
            sum = cta_adler32(S, W.out, pos);
        }
    }
    const uint32_t part_sum = nofooter ? sum : 0;  // kMemberNoFooter: the caller combines the pieces' sums and reads the footer itself
    if (tid == 0) {
        bc.seek(S.c.cur);
        if (status == FB200_OK && container != FB200_RAW && !nofooter) {
            bc.align_to_byte();
            uint32_t v = 0;
            status = bc.read(32, v);
            if (container == FB200_GZIP) {
                if (!status && v != sum) status = FB200_WRONG_GZIP_CHECKSUM;
                if (!status) status = bc.read(32, v);
                if (!status && v != (uint32_t)pos) status = FB200_WRONG_GZIP_SIZE;
            } else {
                const uint32_t be = __byte_perm(sum, 0, 0x0123);
                if (!status && v != be) status = FB200_WRONG_ZLIB_CHECKSUM;
            }
        }
    }
    if (tid == 0) {
        MemberResult res{};
        res.resume_bits = S.c.cur - (unsigned long long)(uintptr_t)in_begin * 8ull;  // before the alignment below
        if (status == FB200_OK) bc.align_to_byte();
        res.out_len = pos;
        // bytes consumed: everything handed to the cursor minus whole bytes still buffered
        res.consumed = (uint64_t)((bc.next - (bc.cnt >> 3)) - in_begin);
        res.status = (uint32_t)status;
        res.pad = S.c.rounds | (S.c.retries << 12) | (S.c.exact_calls << 22);
        res.sum = part_sum;
        res.info = info;
        *result = res;
    }
}

// Persistent CTAs: each takes the next member from a work counter until none is left, so that members of very
// different sizes balance and the match queues (one per CTA) stay L2-resident.
template <bool kPieces>
__global__ void __launch_bounds__(par::kLanes, 1)
inflate_members_par_kernel(int container, const uint8_t* __restrict__ d_in, const MemberDesc* __restrict__ descs, uint32_t k,
                           uint8_t* d_out, MemberResult* __restrict__ results, uint2* __restrict__ queues, uint32_t* work_counter) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Shared& S = *reinterpret_cast<Shared*>(smem_raw + kRing);
    uint2* queue = queues + (size_t)blockIdx.x * par::kQueue;
    for (;;) {
        __syncthreads();  // the previous member's last reads of the control block are done
        if (threadIdx.x == 0) S.c.member = atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t m = S.c.member;
        if (m >= k) break;
        inflate_one_member<kPieces>(S, smem_raw, queue, container, d_in, descs[m], d_out, results + m);
    }
}

size_t inflate_par_scratch_bytes(uint32_t k, int sm_count) {
    const uint32_t grid = k < (uint32_t)sm_count ? k : (uint32_t)sm_count;
    return 256 + (size_t)grid * par::kQueue * sizeof(uint2);
}

cudaError_t inflate_members_par(int container, const uint8_t* d_in, const MemberDesc* d_desc, uint32_t k, uint8_t* d_out,
                                MemberResult* d_res, void* d_scratch, int sm_count, cudaStream_t st, bool pieces) {
    if (k == 0) return cudaSuccess;
    static bool attr_set[64] = {};  // per device
    const size_t smem = par::kRing + sizeof(par::Shared);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaFuncSetAttribute(inflate_members_par_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(inflate_members_par_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    // scratch: a work counter (first 256 bytes), then one match queue per CTA
    const uint32_t grid = k < (uint32_t)sm_count ? k : (uint32_t)sm_count;
    uint32_t* counter = reinterpret_cast<uint32_t*>(d_scratch);
    uint2* queues = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(d_scratch) + 256);
    cudaError_t e = cudaMemsetAsync(counter, 0, 4, st);
    if (e != cudaSuccess) return e;
    if (pieces) inflate_members_par_kernel<true><<<grid, par::kLanes, smem, st>>>(container, d_in, d_desc, k, d_out, d_res, queues, counter);
    else inflate_members_par_kernel<false><<<grid, par::kLanes, smem, st>>>(container, d_in, d_desc, k, d_out, d_res, queues, counter);
    return cudaGetLastError();
}

}  // namespace fb
