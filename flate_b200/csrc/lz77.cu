// lz77.cu -- LZ77 side of the deflate path on sm_100a: hash-chain links, match search,
// lazy-parse selection and token emission.
//
// The reference (deflate.zig:154-266) is a sequential loop.  Here its result is re-derived as
// data-parallel phases with identical output:
//   1. hash chains are parse independent: every position is inserted once, in order
//      (deflate.zig:236 lookup.add, :209 bulkAdd)  ->  link[p] = distance to the previous position
//      with the same 15-bit hash  (hash_link_kernel)
//   2. findMatch depends on the parse only through the budget (chain vs chain>>2) and the final
//      "> min_len" filter  ->  per position R_full(p), R_quarter(p)  (match_search_kernel)
//   3. the slide schedule is a pure function of position  ->  slide_base()
//   4. the lazy parse restricted to "no pending match" arrivals is a function f(p) > p with
//      f(p)-p <= 515  ->  lazy_step_kernel + chunk exit tables + orbit marking + compaction
#include "common.cuh"
#include "pipeline.cuh"

namespace fb {

// ------------------------------------------------------------------------------------------
// K1: hash links.  Lookup.zig:23-84.  One warp walks a run of tiles in position order; the
// head table (hash -> most recent position + 1) lives in shared memory.  Within a group of 32
// consecutive positions the predecessor is found with __match_any_sync; the last lane of each
// hash group publishes the new head.  A warm-up tile primes the head table so runs are
// independent (links farther than 32768 are dropped anyway, deflate.zig:250).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kLinkTile = 32768;
constexpr uint32_t kLinkRun = 8;  // tiles per block (plus one warm-up tile)

__device__ __forceinline__ uint32_t hash4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
    uint32_t v = (b0 << 24) | (b1 << 16) | (b2 << 8) | b3;  // Lookup.zig:75-80 (big-endian read)
    return (v * 0x9E3779B1u) >> 17;                       // Lookup.zig:12,82-84
}

__global__ void __launch_bounds__(32, 1)
hash_link_kernel(const uint8_t* __restrict__ in, uint32_t n, uint16_t* __restrict__ link) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t* head = reinterpret_cast<uint32_t*>(smem_raw);             // 32768 * 4
    uint8_t* bytes = smem_raw + 32768 * 4;                                // kLinkTile + 16
    const uint32_t lane = threadIdx.x;
    const uint32_t ntiles = (n + kLinkTile - 1) / kLinkTile;
    const uint32_t first = blockIdx.x * kLinkRun;
    if (first >= ntiles) return;
    const uint32_t last = min(first + kLinkRun, ntiles);
    for (uint32_t i = lane; i < 32768; i += 32) head[i] = 0;
    __syncwarp();
    const uint32_t t0 = first > 0 ? first - 1 : 0;
    for (uint32_t t = t0; t < last; t++) {
        const bool emit = t >= first;
        const uint32_t base = t * kLinkTile;
        const uint32_t cnt = min(kLinkTile, n - base);
        // stage the tile (+3 bytes look-ahead) in shared memory
        const uint32_t need = min(cnt + 3, n - base);
        if (((uintptr_t)(in + base) & 15) == 0) {
            const uint4* src = reinterpret_cast<const uint4*>(in + base);
            uint4* dst = reinterpret_cast<uint4*>(bytes);
            const uint32_t nv = need / 16;
            for (uint32_t i = lane; i < nv; i += 32) dst[i] = src[i];
            for (uint32_t i = nv * 16 + lane; i < need; i += 32) bytes[i] = in[base + i];
        } else {
            for (uint32_t i = lane; i < need; i += 32) bytes[i] = in[base + i];
        }
        __syncwarp();
        for (uint32_t it = 0; it < cnt; it += 32) {
            const uint32_t off = it + lane;
            const uint32_t p = base + off;
            const bool valid = off < cnt && (uint64_t)p + 4 <= n;  // Lookup.zig:24 needs 4 bytes
            uint32_t h = 0x10000u | lane;                          // unique key for idle lanes
            if (valid) h = hash4(bytes[off], bytes[off + 1], bytes[off + 2], bytes[off + 3]);
            const uint32_t peers = __match_any_sync(0xffffffffu, h);
            uint32_t lnk = 0;
            if (valid) {
                const uint32_t lower = peers & ((1u << lane) - 1);
                uint32_t q1;  // previous position + 1, 0 = none
                if (lower) q1 = base + it + (31 - __clz(lower)) + 1;
                else q1 = head[h];
                if (q1 != 0) {
                    const uint32_t d = p + 1 - q1;
                    if (d <= kMaxDist) lnk = d;
                }
                if ((peers >> lane) == 1u) head[h] = p + 1;  // highest lane of the hash group
            }
            if (emit && off < cnt) link[p] = (uint16_t)lnk;  // 32768 wraps to 0x8000, fits
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2: match search.  deflate.zig:233-266 findMatch + SlidingWindow.zig:81-104 match.
// One thread per position.  A block owns kSearchTile new positions and stages
//   bytes [s-32768, s+tile+258+pad)  and  links [s-32768, s+tile)
// in shared memory, so chain walks and compares never leave the SM.
// For each position it produces the result of the walk with min_len = 0 under the full budget
// and, as a snapshot after chain>>2 candidates, under the quarter budget.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kSearchTile = 4096;
constexpr uint32_t kSearchThreads = 512;
constexpr uint32_t kSearchBytes = kHist + kSearchTile + 272;   // window + look-ahead, 16B multiple
constexpr uint32_t kSearchLinks = kHist + kSearchTile;
constexpr uint32_t kSearchSmem = kSearchBytes + kSearchLinks * 2;

// unaligned 4-byte little-endian load from shared memory (two aligned words + funnel shift)
__device__ __forceinline__ uint32_t lds_u32_unaligned(const uint8_t* base, uint32_t idx) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (idx >> 2);
    return __funnelshift_r(w[0], w[1], (idx & 3) * 8);
}

__global__ void __launch_bounds__(kSearchThreads, 2)
match_search_kernel(const uint8_t* __restrict__ in, uint32_t n, const uint16_t* __restrict__ link,
                    LevelArgs lv, uint32_t* __restrict__ r_full, uint32_t* __restrict__ r_quarter) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* sb = smem_raw;                                                 // bytes
    uint16_t* sl = reinterpret_cast<uint16_t*>(smem_raw + kSearchBytes);    // links
    const uint32_t s = blockIdx.x * kSearchTile;                            // first new position
    const int64_t wb = (int64_t)s - kHist;                                  // window base (may be < 0)
    const uint32_t lo = wb < 0 ? (uint32_t)(-wb) : 0;                       // first valid smem index

    // ---- stage window ----
    {
        const uint32_t byte_hi = (uint32_t)min((int64_t)kSearchBytes, (int64_t)n - wb);  // exclusive
        // 16-byte vector body (wb is a multiple of 4096, input assumed 16B aligned) + scalar tail
        const bool aligned = ((uintptr_t)in & 15) == 0;
        const uint32_t v_lo = lo / 16, v_hi = aligned ? byte_hi / 16 : v_lo;
        const uint4* src = reinterpret_cast<const uint4*>(in + wb);
        uint4* dst = reinterpret_cast<uint4*>(sb);
        for (uint32_t i = v_lo + threadIdx.x; i < v_hi; i += kSearchThreads) dst[i] = src[i];
        for (uint32_t i = max(lo, v_hi * 16) + threadIdx.x; i < byte_hi; i += kSearchThreads) sb[i] = in[wb + i];
        for (uint32_t i = byte_hi + threadIdx.x; i < kSearchBytes; i += kSearchThreads) sb[i] = 0;
        const uint32_t link_hi = (uint32_t)min((int64_t)kSearchLinks, (int64_t)n - wb);
        const uint32_t lv_lo = lo / 8, lv_hi = link_hi / 8;
        const uint4* lsrc = reinterpret_cast<const uint4*>(link + wb);
        uint4* ldst = reinterpret_cast<uint4*>(sl);
        for (uint32_t i = lv_lo + threadIdx.x; i < lv_hi; i += kSearchThreads) ldst[i] = lsrc[i];
        for (uint32_t i = max(lo, lv_hi * 8) + threadIdx.x; i < link_hi; i += kSearchThreads) sl[i] = link[wb + i];
    }
    __syncthreads();

    const uint32_t quarter = lv.chain >> 2;
    for (uint32_t k = threadIdx.x; k < kSearchTile; k += kSearchThreads) {
        const uint32_t p = s + k;
        if (p >= n) break;
        uint32_t best_len = 0, best_dist = 0, snap = 0;
        bool snapped = false;
        const uint32_t remaining = n - p;
        if (remaining >= kMinMatch) {  // Lookup.zig:24: no insertion / search with < 4 bytes left
            const uint32_t max_len = min(remaining, kMaxMatch);  // SlidingWindow.zig:82
            const uint32_t base = slide_base(p, n);
            const uint32_t pi = p - (uint32_t)wb;  // smem index of p  (wb <= p always)
            const uint32_t first4 = lds_u32_unaligned(sb, pi);
            uint32_t qi = pi;
            uint32_t cnt = 0;
            while (true) {  // deflate.zig:248 "Hot path loop!"
                const uint32_t l = sl[qi];
                if (l == 0) break;
                qi -= l;
                const uint32_t dist = pi - qi;
                const int64_t q = wb + qi;
                if (dist > kMaxDist || q <= (int64_t)base) break;  // deflate.zig:250, :248 (pos 0 = none)
                cnt++;
                // ---- SlidingWindow.match with the running best as min_len ----
                // a candidate only matters if it is strictly longer than best_len, i.e. bytes
                // [0, best_len] all agree; test the first word and the byte at best_len first.
                if (lds_u32_unaligned(sb, qi) == first4 && (best_len == 0 || sb[qi + best_len] == sb[pi + best_len])) {
                    uint32_t i = 4;
                    while (i < max_len) {
                        const uint32_t x = lds_u32_unaligned(sb, qi + i) ^ lds_u32_unaligned(sb, pi + i);
                        if (x) {
                            i += (__ffs(x) - 1) >> 3;
                            break;
                        }
                        i += 4;
                    }
                    if (i > max_len) i = max_len;
                    if (i > best_len) {
                        best_len = i;
                        best_dist = dist;
                        if (i >= lv.nice) break;    // deflate.zig:256-259
                        if (i >= max_len) {          // nothing can be strictly longer: the rest of the
                            break;                   // walk cannot change either result
                        }
                    }
                }
                if (cnt == quarter) {
                    snap = best_len ? pack_match(best_len, best_dist) : 0;
                    snapped = true;
                }
                if (cnt >= lv.chain) break;
            }
        }
        const uint32_t full = best_len ? pack_match(best_len, best_dist) : 0;
        r_full[p] = full;
        r_quarter[p] = snapped ? snap : full;
    }
}

// ------------------------------------------------------------------------------------------
// K3a: lazy step.  deflate.zig:160-193 restricted to arrivals with no pending match.
// From such an arrival at p the reference emits k literals p..p+k-1 (each displaced by a strictly
// longer match one byte later) and then one match at p+k, or a single literal if nothing matches.
// ------------------------------------------------------------------------------------------
__global__ void lazy_step_kernel(const uint32_t* __restrict__ r_full, const uint32_t* __restrict__ r_quarter, uint32_t n,
                                 LevelArgs lv, uint32_t* __restrict__ nx) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t r = r_full[p];
    uint32_t out = 0;
    if (r != 0) {
        uint32_t cur = p;
        while (true) {
            const uint32_t len = match_len_of(r);
            if (len >= lv.lazy) break;  // deflate.zig:171
            const uint32_t nxt = cur + 1;
            if (nxt >= n) break;
            const uint32_t r2 = (len >= lv.good) ? r_quarter[nxt] : r_full[nxt];  // deflate.zig:241-245
            if (match_len_of(r2) > len) {  // better match one byte later: p becomes a literal
                cur = nxt;
                r = r2;
            } else {
                break;  // deflate.zig:182-184: emit the pending match
            }
        }
        out = (cur - p) | ((match_len_of(r) - 3) << 8) | (match_dist_of(r) << 16);
    }
    nx[p] = out;
}

// ------------------------------------------------------------------------------------------
// K3b: chunk exit tables by pointer jumping.  For every possible entry offset e < 516 of a chunk
// of kChunk positions, the offset (into the next chunk) of the first arrival past the chunk end.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
chunk_exit_kernel(const uint32_t* __restrict__ nx, uint32_t n, uint16_t* __restrict__ exits) {
    __shared__ uint16_t nxt[kChunk];
    const uint32_t c = blockIdx.x;
    const uint32_t cs = c * kChunk;
    for (uint32_t i = threadIdx.x; i < kChunk; i += blockDim.x) {
        const uint32_t p = cs + i;
        uint32_t t = kChunk;  // positions past the end of data exit immediately
        if (p < n) t = i + nx_step(nx[p]);
        nxt[i] = (uint16_t)t;
    }
    __syncthreads();
    while (true) {
        bool pending = false;
        for (uint32_t i = threadIdx.x; i < kChunk; i += blockDim.x) {
            const uint32_t t = nxt[i];
            if (t < kChunk) {
                nxt[i] = nxt[t];  // racing reads see some power of f: still correct
                pending = true;
            }
        }
        if (!__syncthreads_or(pending)) break;
    }
    for (uint32_t i = threadIdx.x; i < kEntries; i += blockDim.x) exits[(size_t)c * kEntries + i] = nxt[i] - kChunk;
}

// K3c: resolve the true entry offset of every chunk.  Two-level: groups of kGroup chunks.
__global__ void group_exit_kernel(const uint16_t* __restrict__ exits, uint32_t nchunks, uint16_t* __restrict__ gexits) {
    const uint32_t g = blockIdx.x;
    const uint32_t e = threadIdx.x;
    if (e >= kEntries) return;
    const uint32_t c0 = g * kGroup, c1 = min(c0 + kGroup, nchunks);
    uint32_t cur = e;
    for (uint32_t c = c0; c < c1; c++) cur = exits[(size_t)c * kEntries + cur];
    gexits[(size_t)g * kEntries + e] = (uint16_t)cur;
}
__global__ void group_entry_kernel(const uint16_t* __restrict__ gexits, uint32_t ngroups, uint16_t* __restrict__ gentry) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t cur = 0;
    for (uint32_t g = 0; g < ngroups; g++) {
        gentry[g] = (uint16_t)cur;
        cur = gexits[(size_t)g * kEntries + cur];
    }
}
__global__ void chunk_entry_kernel(const uint16_t* __restrict__ exits, const uint16_t* __restrict__ gentry,
                                   uint32_t nchunks, uint16_t* __restrict__ entry) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c0 = g * kGroup;
    if (c0 >= nchunks) return;
    const uint32_t c1 = min(c0 + kGroup, nchunks);
    uint32_t cur = gentry[g];
    for (uint32_t c = c0; c < c1; c++) {
        entry[c] = (uint16_t)cur;
        cur = exits[(size_t)c * kEntries + cur];
    }
}

// ------------------------------------------------------------------------------------------
// K3d: walk the orbit of each chunk from its true entry; record arrivals in a bitmap and count
// the tokens they emit (k literals + 1 match, or 1 literal).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kMarkThreads = 128;
__global__ void __launch_bounds__(kMarkThreads)
orbit_mark_kernel(const uint32_t* __restrict__ nx, uint32_t n, const uint16_t* __restrict__ entry,
                  uint32_t* __restrict__ bitmap, uint32_t* __restrict__ chunk_tokens) {
    __shared__ uint16_t step[kChunk];   // step | 0x8000.. no: plain step (<= 515)
    __shared__ uint16_t ntok[kChunk];   // tokens emitted by an arrival here
    __shared__ uint32_t bits[kChunk / 32];
    const uint32_t c = blockIdx.x;
    const uint32_t cs = c * kChunk;
    for (uint32_t i = threadIdx.x; i < kChunk; i += kMarkThreads) {
        const uint32_t p = cs + i;
        uint32_t s = 1, t = 0;
        if (p < n) {
            const uint32_t v = nx[p];
            s = nx_step(v);
            t = (v >> 16) ? (v & 255u) + 1 : 1;
        }
        step[i] = (uint16_t)s;
        ntok[i] = (uint16_t)t;
    }
    for (uint32_t i = threadIdx.x; i < kChunk / 32; i += kMarkThreads) bits[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t i = entry[c];
        uint32_t total = 0;
        const uint32_t lim = min(kChunk, n > cs ? n - cs : 0u);
        while (i < lim) {
            bits[i >> 5] |= 1u << (i & 31);
            total += ntok[i];
            i += step[i];
        }
        chunk_tokens[c] = total;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kChunk / 32; i += kMarkThreads) bitmap[(size_t)c * (kChunk / 32) + i] = bits[i];
}

// exclusive scan of per-chunk token counts (single block; nchunks is at most ~1M)
__global__ void __launch_bounds__(1024)
scan_tokens_kernel(const uint32_t* __restrict__ counts, uint32_t nchunks, uint32_t* __restrict__ offsets,
                   uint32_t* __restrict__ total_out) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nchunks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nchunks ? counts[i] : 0;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sums[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t wprefix = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
        const uint32_t incl = carry + wprefix + x;
        if (i < nchunks) offsets[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

// ------------------------------------------------------------------------------------------
// K3e: token emission (compaction).  Each arrival writes its k literals and its match (or its one
// literal) at the chunk's token offset + the prefix of earlier arrivals in the chunk.  The thread
// that writes the last token of a 32768-token block also records the reference's `rp` at that
// moment (deflate.zig:227-230, SlidingWindow.zig:119-123).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kEmitThreads = 256;
__global__ void __launch_bounds__(kEmitThreads)
emit_tokens_kernel(const uint8_t* __restrict__ in, const uint32_t* __restrict__ nx, uint32_t n,
                   const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ tok_offset, LevelArgs lv,
                   uint32_t* __restrict__ tokens, uint32_t* __restrict__ cut_rp) {
    __shared__ uint32_t warp_sums[kEmitThreads / 32];
    const uint32_t c = blockIdx.x;
    const uint32_t cs = c * kChunk;
    constexpr uint32_t kPer = kChunk / kEmitThreads;  // 16 positions per thread
    const uint32_t i0 = threadIdx.x * kPer;
    const uint32_t word = bitmap[(size_t)c * (kChunk / 32) + (i0 >> 5)];
    const uint32_t mask = (word >> (i0 & 31)) & ((1u << kPer) - 1);
    // tokens of my arrivals
    uint32_t mine = 0;
    for (uint32_t m = mask; m; m &= m - 1) {
        const uint32_t v = nx[cs + i0 + (__ffs(m) - 1)];
        mine += (v >> 16) ? (v & 255u) + 1 : 1;
    }
    uint32_t x = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t wprefix = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) wprefix += warp_sums[w];
    uint32_t t = tok_offset[c] + wprefix + x - mine;
    for (uint32_t m = mask; m; m &= m - 1) {
        const uint32_t p = cs + i0 + (__ffs(m) - 1);
        const uint32_t v = nx[p];
        if ((v >> 16) == 0) {
            tokens[t] = in[p];
            if ((t & (kTokensPerBlock - 1)) == kTokensPerBlock - 1) cut_rp[t >> 15] = p + 1;
            t++;
        } else {
            const uint32_t k = v & 255u, len = ((v >> 8) & 255u) + 3, dist = v >> 16;
            for (uint32_t j = 0; j < k; j++) {
                tokens[t] = in[p + j];
                if ((t & (kTokensPerBlock - 1)) == kTokensPerBlock - 1) cut_rp[t >> 15] = p + j + 1;
                t++;
            }
            tokens[t] = tok_match(dist, len);
            // immediate match (len >= lazy) is added while processing its own position, a deferred
            // one while processing the next position (deflate.zig:171-184)
            if ((t & (kTokensPerBlock - 1)) == kTokensPerBlock - 1) cut_rp[t >> 15] = p + k + (len >= lv.lazy ? 0 : 1);
            t++;
        }
    }
}

// ------------------------------------------------------------------------------------------
// host-side launcher
// ------------------------------------------------------------------------------------------
cudaError_t lz77_tokenize(const Lz77Buffers& b, const uint8_t* d_in, uint32_t n, const LevelArgs& lv, cudaStream_t st,
                          PhaseTimer* pt) {
    PhaseTimer dummy;
    if (!pt) pt = &dummy;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(hash_link_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4 + kLinkTile + 16);
        cudaFuncSetAttribute(match_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSearchSmem);
        attr_set = true;
    }
    if (n == 0) {
        cudaMemsetAsync(b.total_tokens, 0, sizeof(uint32_t), st);
        return cudaGetLastError();
    }
    const uint32_t ntiles = (n + kLinkTile - 1) / kLinkTile;
    hash_link_kernel<<<(ntiles + kLinkRun - 1) / kLinkRun, 32, 32768 * 4 + kLinkTile + 16, st>>>(d_in, n, b.link);
    pt->mark(st, kPhLink);
    match_search_kernel<<<(n + kSearchTile - 1) / kSearchTile, kSearchThreads, kSearchSmem, st>>>(d_in, n, b.link, lv,
                                                                                                b.r_full, b.r_quarter);
    pt->mark(st, kPhSearch);
    lazy_step_kernel<<<(n + 255) / 256, 256, 0, st>>>(b.r_full, b.r_quarter, n, lv, b.nx);
    pt->mark(st, kPhLazy);
    const uint32_t nchunks = (n + kChunk - 1) / kChunk;
    const uint32_t ngroups = (nchunks + kGroup - 1) / kGroup;
    chunk_exit_kernel<<<nchunks, 1024, 0, st>>>(b.nx, n, b.exits);
    pt->mark(st, kPhChunkExit);
    group_exit_kernel<<<ngroups, 544, 0, st>>>(b.exits, nchunks, b.gexits);
    group_entry_kernel<<<1, 32, 0, st>>>(b.gexits, ngroups, b.gentry);
    chunk_entry_kernel<<<(ngroups + 127) / 128, 128, 0, st>>>(b.exits, b.gentry, nchunks, b.entry);
    pt->mark(st, kPhResolve);
    orbit_mark_kernel<<<nchunks, kMarkThreads, 0, st>>>(b.nx, n, b.entry, b.bitmap, b.chunk_tokens);
    pt->mark(st, kPhMark);
    scan_tokens_kernel<<<1, 1024, 0, st>>>(b.chunk_tokens, nchunks, b.tok_offset, b.total_tokens);
    pt->mark(st, kPhScan);
    emit_tokens_kernel<<<nchunks, kEmitThreads, 0, st>>>(d_in, b.nx, n, b.bitmap, b.tok_offset, lv, b.tokens, b.cut_rp);
    pt->mark(st, kPhEmit);
    return cudaGetLastError();
}

}  // namespace fb
