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

#include <cstdio>
#include <cstdlib>

namespace fb {

// ------------------------------------------------------------------------------------------
// K1: hash links.  Lookup.zig:23-84.
// "Previous position with the same hash" is an ordered-predecessor problem.  A block walks a run
// of 8192-position tiles; inside a tile the 15-bit hash space is split over the block's 16 warps
// (warp w owns hashes with top 4 bits == w), so 16 sequential chains advance in parallel:
//   1. all threads hash the tile into shared memory (u16 per position, 0xFFFF = not insertable)
//   2. the tile's positions are split by owner, stably (per-owner counts, then an ordered scatter: one ballot per
//      bit of the owner number -- MATCH.ANY's latency grows with the number of distinct values, 16 here), and every
//      warp resolves its own list 32 entries at a time: predecessor inside the group via __match_any_sync,
//      otherwise from its slice of the head table; the last lane of each hash group publishes the new head
//   3. links are collected in shared memory and flushed coalesced; the head table is rebased by one
//      tile (the reference's Lookup.slide, Lookup.zig:43-51, is the same saturating subtract)
// Head entries are u16 codes c = p - (tile_base - 32768) + 1 in [1, 40960], 0 = none.
// Four warm-up tiles (32768 positions) prime the table so runs are independent (links farther than
// 32768 are dropped anyway, deflate.zig:250).
// ------------------------------------------------------------------------------------------
// link[p] = distance to the previous position with the same hash; kNoLink when there is none within
// 32768 (or p was never inserted).  0xFFFF makes the walk's single bound test catch it: q - 65535 is
// always below the lowest admissible candidate.
constexpr uint16_t kNoLink = 0xFFFF;
constexpr uint32_t kLinkTile = 8192;
constexpr uint32_t kLinkWarm = kHist / kLinkTile;  // warm-up tiles
constexpr uint32_t kLinkWarps = 16;
constexpr uint32_t kLinkGroup = 3;  // tiles between two rebases of the head table
constexpr uint32_t kRunFlag = 0x8000u;   // on a hash: the position continues a run of one byte (its link is 1)
constexpr uint32_t kRunLink = 0xFFFEu;   // in the link slot of such a position until the tile is flushed (real links are <= 0x8000)
constexpr uint32_t kLinkThreads = kLinkWarps * 32;
constexpr uint32_t kLinkPerWarp = kLinkTile / kLinkWarps;  // positions each warp splits
constexpr uint32_t kLinkSmem = 32768 * 2 /*head*/ + kLinkTile * 2 /*partition lists*/ + kLinkTile * 2 /*hash, then link*/ +
                               (kLinkWarps * 17 + 32 + kLinkTile / 32) * 4;

__device__ __forceinline__ uint32_t hash_be(uint32_t le32) {
    // Lookup.zig:75-84: big-endian read of 4 bytes, times 0x9E3779B1, top 15 bits
    return (__byte_perm(le32, 0, 0x0123) * 0x9E3779B1u) >> 17;
}

__global__ void __launch_bounds__(kLinkThreads, 2)
hash_link_kernel(const uint8_t* __restrict__ in, uint32_t begin, uint32_t range_end, uint32_t n, uint32_t run,
                 const uint32_t* __restrict__ skip, uint32_t nskip, uint16_t* __restrict__ link) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint16_t* head = reinterpret_cast<uint16_t*>(smem_raw);
    uint16_t* lists = reinterpret_cast<uint16_t*>(smem_raw + 65536);          // tile offsets grouped by owner warp, in position order
    uint16_t* hl = reinterpret_cast<uint16_t*>(smem_raw + 65536 + kLinkTile * 2);  // hash per position, later its link
    uint32_t* cnt = reinterpret_cast<uint32_t*>(hl + kLinkTile);              // [warp][17] counts -> running bases
    uint32_t* pstart = cnt + kLinkWarps * 17;                                 // [17] partition starts
    uint32_t* lastbits = pstart + 32;                                         // [kLinkTile / 32] last position of a run of one byte
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t ltmask = (1u << lane) - 1;
    // [begin, n) is the segment being compressed (begin > 0 after a sync flush); earlier positions are
    // history: hashed again to prime the table, never re-linked.  `skip` lists history positions that
    // the reference never inserted (the last 3 bytes before each flush point, Lookup.zig:24).
    // links are produced for positions [begin, range_end) (range_end is tile aligned or n); validity of a
    // position is judged against the stream end n
    const uint32_t ntiles = (range_end + kLinkTile - 1) / kLinkTile;
    const uint32_t first = begin / kLinkTile + blockIdx.x * run;
    if (first >= ntiles) return;
    const uint32_t last = min(first + run, ntiles);
    for (uint32_t i = tid; i < 32768 / 2; i += kLinkThreads) reinterpret_cast<uint32_t*>(head)[i] = 0;
    const uint32_t t0 = first > kLinkWarm ? first - kLinkWarm : 0;
    const bool aligned = ((uintptr_t)in & 3) == 0;
    // head entries are codes relative to a base that moves every kLinkGroup tiles (one pass over the 64 KiB table per
    // group instead of per tile): code = position - group_base + 32768 + 1, at most 32768 + 3 * 8192 = 57344
    uint32_t grp = 0;
    for (uint32_t t = t0; t < last; t++) {
        const bool emit = t >= first;
        const uint32_t base = t * kLinkTile;
        const uint32_t cnt_pos = min(kLinkTile, n - base);
        // ---- 1. hashes of the tile (0xFFFF = not insertable) ----
        bool any_run = false;
        if (aligned) {
            const uint32_t* words = reinterpret_cast<const uint32_t*>(in + base);
            const uint32_t nwords_total = (n - base + 3) / 4, nwords_full = (n - base) / 4, room = n - base;
            for (uint32_t i = tid; i < kLinkTile / 4; i += kLinkThreads) {
                // whole words inside the stream; the last, partial word is assembled from its valid bytes
                uint32_t w0 = 0, w1 = 0;
                if (i < nwords_full) w0 = words[i];
                else if (i < nwords_total)
                    for (uint32_t j = 0; i * 4 + j < n - base; j++) w0 |= (uint32_t)in[base + i * 4 + j] << (8 * j);
                if (i + 1 < nwords_full) w1 = words[i + 1];
                else if (i + 1 < nwords_total)
                    for (uint32_t j = 0; (i + 1) * 4 + j < n - base; j++) w1 |= (uint32_t)in[base + (i + 1) * 4 + j] << (8 * j);
                uint32_t hh[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; k++) {
                    const uint32_t off = i * 4 + k;
                    const bool valid = off + 4 <= room;  // Lookup.zig:24 needs 4 bytes (room = bytes from the tile's start to the stream's end)
                    hh[k] = valid ? hash_be(__funnelshift_r(w0, w1, 8 * k)) : 0xFFFFu;
                }
                any_run |= w0 == __byte_perm(w0, 0, 0x0321) && i < nwords_full;  // four equal bytes: a run of one byte may pass here
                reinterpret_cast<uint2*>(hl)[i] = make_uint2(hh[0] | (hh[1] << 16), hh[2] | (hh[3] << 16));
            }
        } else {
            for (uint32_t off = tid; off < kLinkTile; off += kLinkThreads) {
                uint32_t h = 0xFFFFu;
                if (off < cnt_pos && (uint64_t)base + off + 4 <= n) {
                    const uint8_t* b = in + base + off;
                    h = hash_be((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24));
                    any_run |= b[0] == b[1] && b[1] == b[2] && b[2] == b[3];
                }
                hl[off] = (uint16_t)h;
            }
        }
        for (uint32_t i = tid; i < kLinkWarps * 17; i += kLinkThreads) cnt[i] = 0;
        if (nskip) {
            __syncthreads();
            for (uint32_t j = tid; j < nskip; j += kLinkThreads) {
                const uint32_t q = skip[j];
                if (q >= base && q < base + cnt_pos) hl[q - base] = 0xFFFFu;
            }
        }
        // Runs of one byte.  A position whose four bytes equal the four bytes one position earlier has its predecessor's
        // hash, so its link is 1 whatever the table says.  In a tile with long runs (zero padding, holes) such positions
        // are flagged (kRunFlag), stay out of the ordered resolve, and only the last one of a run takes part, to leave the
        // right head behind: without this a run of zeros puts a whole tile on one warp's list, 32 conflicts per step.
        // Tiles without four equal bytes in a row at 16 places or more (text) skip all of it.
        const bool has_runs = __syncthreads_count(any_run) >= 16;
        if (has_runs) {
            static_assert(kLinkTile / kLinkThreads <= 32, "one decision bit per round of the loop");
            uint32_t flagged = 0;  // decided for all positions first, written after the barrier: the test reads the neighbour's hash
            for (uint32_t off = tid, r = 0; off < cnt_pos; off += kLinkThreads, r++) {
                const uint32_t h = hl[off];
                if (h == 0xFFFFu || h == 0x7FFFu || (base | off) == 0) continue;
                // the position before must have been inserted (the three before a flush point never are, Lookup.zig:24)
                bool pred_in = off ? hl[off - 1] != 0xFFFFu : true;
                if (off == 0)
                    for (uint32_t j = 0; j < nskip; j++) pred_in = pred_in && skip[j] != base - 1;
                const uint8_t* b = in + base + off;
                if (pred_in && b[-1] == b[0] && b[0] == b[1] && b[1] == b[2] && b[2] == b[3]) flagged |= 1u << r;
            }
            __syncthreads();
            for (uint32_t off = tid, r = 0; off < cnt_pos; off += kLinkThreads, r++)
                if ((flagged >> r) & 1u) hl[off] = (uint16_t)(hl[off] | kRunFlag);
            __syncthreads();
        }
        // ---- 2. stable split of the tile's positions by owner warp (hash >> 11) ----
        // 2a. every warp counts, per owner, the positions of its own 512-position slice
        // (counting needs no order: one shared-memory atomic per position; the ordered ranks are only needed in 2c)
        for (uint32_t it = 0; it < kLinkPerWarp / 32; it++) {
            const uint32_t off = w * kLinkPerWarp + it * 32 + lane;
            const uint32_t h = hl[off];
            if (!has_runs) {
                if ((h >> 11) < kLinkWarps) atomicAdd(&cnt[w * 17 + (h >> 11)], 1u);  // 0..15, or 31 for not insertable
                continue;
            }
            const uint32_t hn = off + 1 < cnt_pos ? (uint32_t)hl[off + 1] : 0u;  // (the neighbour may belong to the next warp: read only)
            const bool in_run = (h & kRunFlag) && h != 0xFFFFu, next_in_run = (hn & kRunFlag) && hn != 0xFFFFu;
            const bool run_last = in_run && !next_in_run;
            const uint32_t lb = __ballot_sync(0xffffffffu, run_last);
            if (lane == 0) lastbits[off >> 5] = lb;
            const uint32_t part = (in_run && !run_last) ? 31u : (h & 0x7FFFu) >> 11;  // 0..15; run interior and not insertable: none
            if (h != 0xFFFFu && part < kLinkWarps) atomicAdd(&cnt[w * 17 + part], 1u);
        }
        __syncthreads();
        // 2b. exclusive prefix down each owner's column, then over the owners
        if (tid < kLinkWarps) {
            uint32_t run_sum = 0;
            for (uint32_t ww = 0; ww < kLinkWarps; ww++) {
                const uint32_t c = cnt[ww * 17 + tid];
                cnt[ww * 17 + tid] = run_sum;
                run_sum += c;
            }
            pstart[tid + 1] = run_sum;  // totals for now
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t acc = 0;
            pstart[0] = 0;
            for (uint32_t k = 1; k <= kLinkWarps; k++) {
                acc += pstart[k];
                pstart[k] = acc;
            }
        }
        __syncthreads();
        // 2c. scatter (off, hash) into the owner's list, keeping position order.  Lane l < 16 carries, in a register,
        // where the next entry of owner l from this warp's slice goes; a step needs one ballot per bit of the owner
        // number, from which every lane derives the lanes with ITS position's owner (its rank among them is its place)
        // and lane l the lanes with owner l (their number moves its cursor): no shared-memory counter on the way.
        {
            uint32_t cursor = lane < kLinkWarps ? pstart[lane] + cnt[w * 17 + lane] : 0u;
            uint32_t own_x[4];  // all-ones where this lane's number has a 0 bit
#pragma unroll
            for (uint32_t b = 0; b < 4; b++) own_x[b] = ((lane >> b) & 1u) - 1u;
            for (uint32_t it = 0; it < kLinkPerWarp / 32; it++) {
                const uint32_t off = w * kLinkPerWarp + it * 32 + lane;
                const uint32_t h = hl[off];
                uint32_t part = h >> 11;
                if (has_runs) {
                    const bool in_run = (h & kRunFlag) && h != 0xFFFFu, run_last = in_run && ((lastbits[off >> 5] >> lane) & 1u);
                    part = h == 0xFFFFu || (in_run && !run_last) ? 31u : (h & 0x7FFFu) >> 11;
                    if (in_run && !run_last) hl[off] = kRunLink;  // its link is 1: nothing else to find out (only this lane reads hl[off] here)
                }
                const uint32_t listed = __ballot_sync(0xffffffffu, part < kLinkWarps);
                uint32_t peers = listed, mine = listed;
#pragma unroll
                for (uint32_t b = 0; b < 4; b++) {
                    const uint32_t bal = __ballot_sync(0xffffffffu, (part >> b) & 1u);
                    peers &= bal ^ (((part >> b) & 1u) - 1u);
                    mine &= bal ^ own_x[b];
                }
                const uint32_t at = __shfl_sync(0xffffffffu, cursor, part & (kLinkWarps - 1));
                if (part < kLinkWarps) lists[at + __popc(peers & ltmask)] = (uint16_t)off;
                cursor += __popc(mine);  // (lanes 16..31 count owners that do not exist: unused)
            }
        }
        __syncthreads();
        // ---- 3. every warp resolves its own list in order, 32 entries per step ----
        {
            const uint32_t lo = pstart[w], hi = pstart[w + 1];
            for (uint32_t sidx = lo; sidx < hi; sidx += 32) {
                uint32_t h = 0x10000u | lane, off = 0;  // unique key for idle lanes
                const bool have = sidx + lane < hi;
                bool run_last = false;
                if (have) {
                    off = lists[sidx + lane];
                    h = hl[off];
                    run_last = (h & kRunFlag) != 0;  // a listed position with the flag is the last one of a run
                    h &= 0x7FFFu;
                }
                // Lanes with the same hash are ordered with MATCH.ANY: a lane's predecessor is the next lower lane of
                // its group, or, for the group's first lane, the head entry; the group's last lane leaves the new head.
                // One writer per entry and step, and every lane has read the old head before anybody publishes.  (An
                // optimistic variant -- everybody publishes, reads back, and only a group that saw a loser is ordered --
                // saved the MATCH on data without repeats but took the ordered path on nearly every step of text, and
                // its same-address stores were a race by the letter.)
                const uint32_t code = off + grp * kLinkTile + kHist + 1;
                uint32_t e = 0;
                if (have) e = head[h];
                const uint32_t peers = __match_any_sync(0xffffffffu, h);
                const uint32_t lower = peers & ltmask;
                const uint32_t src = lower ? 31 - __clz(lower) : lane;
                const uint32_t off_prev = __shfl_sync(0xffffffffu, off, src);
                uint32_t d = 0;
                if (lower) {
                    d = off - off_prev;  // predecessor inside the group, same tile
                } else if (e) {
                    d = code - e;
                    if (d > kMaxDist) d = 0;
                }
                __syncwarp();
                if (have && (peers >> lane) == 1u) head[h] = (uint16_t)code;
                if (have) hl[off] = (uint16_t)(run_last ? 1u : d);  // only this lane ever needed the hash of this position
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- 4. flush links (0xFFFF = never inserted -> no link), rebase the head table by one tile ----
        if (emit) {
            uint16_t* dst = link + base;
            if (base >= begin) {
                const uint32_t nv = cnt_pos / 8;
                for (uint32_t i = tid; i < nv; i += kLinkThreads) {
                    uint4 v = reinterpret_cast<const uint4*>(hl)[i];
                    uint32_t* x = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
                    for (int k = 0; k < 4; k++) {  // no predecessor (0) and never inserted (0xFFFF) both become kNoLink
                        if ((x[k] & 0xffffu) == 0) x[k] |= 0x0000ffffu;
                        if ((x[k] >> 16) == 0) x[k] |= 0xffff0000u;
                        if (has_runs) {  // inside a run of one byte
                            if ((x[k] & 0xffffu) == kRunLink) x[k] = (x[k] & 0xffff0000u) | 1u;
                            if ((x[k] >> 16) == kRunLink) x[k] = (x[k] & 0x0000ffffu) | 0x00010000u;
                        }
                    }
                    reinterpret_cast<uint4*>(dst)[i] = v;
                }
                for (uint32_t i = nv * 8 + tid; i < cnt_pos; i += kLinkThreads) dst[i] = hl[i] == 0 ? kNoLink : hl[i] == kRunLink ? (uint16_t)1 : hl[i];
            } else {  // the tile straddles the segment start: keep the links of earlier segments
                for (uint32_t i = tid; i < cnt_pos; i += kLinkThreads)
                    if (base + i >= begin) dst[i] = hl[i] == 0 ? kNoLink : hl[i] == kRunLink ? (uint16_t)1 : hl[i];
            }
        }
        if (++grp == kLinkGroup && t + 1 < last) {
            constexpr uint32_t kShift = kLinkGroup * kLinkTile;  // entries at or below it are more than 32768 behind the new base
            for (uint32_t i = tid; i < 32768 / 2; i += kLinkThreads) {
                const uint32_t v = reinterpret_cast<uint32_t*>(head)[i];
                const uint32_t lo = v & 0xffffu, hi = v >> 16;
                reinterpret_cast<uint32_t*>(head)[i] = (lo > kShift ? lo - kShift : 0u) | ((hi > kShift ? hi - kShift : 0u) << 16);
            }
            grp = 0;
        }
        __syncthreads();
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
constexpr uint32_t kSearchOff = 16;                                          // index shift: slot 0 means "none"
constexpr uint32_t kSearchBytes = kSearchOff + kHist + kSearchTile + 272;   // window + look-ahead, 16B multiple
constexpr uint32_t kSearchLinks = kSearchOff + kHist + kSearchTile;
constexpr uint32_t kSearchSmem = kSearchBytes + kSearchLinks * 2;

// unaligned 4-byte little-endian load from shared memory (two aligned words + funnel shift)
__device__ __forceinline__ uint32_t lds_u32_unaligned(const uint8_t* base, uint32_t idx) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (idx >> 2);
    return __funnelshift_r(w[0], w[1], (idx & 3) * 8);
}

__device__ __forceinline__ uint32_t lds_shared_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_shared_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

struct SearchTune {
    uint32_t pend_at;    // run the batched full compares once this many lanes wait for one
    uint32_t refill_at;  // re-arm finished lanes once this many are idle
};

// Lane states of the walk
enum : uint32_t { kIdle = 0, kStepping = 1, kPending = 2, kDone = 3 };

template <int kStepsPerRound>
__global__ void __launch_bounds__(kSearchThreads, 2)
match_search_kernel(const uint8_t* __restrict__ in, uint32_t seg_begin, uint32_t begin, uint32_t range_end, uint32_t n,
                    const uint16_t* __restrict__ link, LevelArgs lv, SearchTune tune, uint32_t* __restrict__ r_full,
                    uint32_t* __restrict__ r_quarter) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint32_t tile_next;
    __shared__ __align__(8) uint64_t stage_bar;
    uint8_t* sb = smem_raw;                                                 // bytes, slot i+16 = position wb+i
    uint16_t* sl = reinterpret_cast<uint16_t*>(smem_raw + kSearchBytes);    // link (distance, kNoLink = none) per slot
    const uint32_t s = (begin / kSearchTile + blockIdx.x) * kSearchTile;    // first new position (tiles are absolute)
    r_full -= seg_begin;                                                    // result tables are segment relative
    r_quarter -= seg_begin;
    const int64_t wb = (int64_t)s - kHist;                                  // window base (may be < 0)
    const uint32_t lo = wb < 0 ? (uint32_t)(-wb) : 0;                       // first valid window index
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ltmask = (1u << lane) - 1;

    // ---- stage window ----
    // Interior tiles: two TMA bulk copies (cp.async.bulk global -> shared, completion on an mbarrier)
    // issued by one thread: 36 KiB of bytes and 72 KiB of links land in shared memory without passing
    // through registers.  Edge tiles (stream start / end) use a guarded copy loop.
    const bool tma_ok = wb >= 0 && (uint64_t)s + kSearchTile + 272 <= n && (((uintptr_t)in | (uintptr_t)link) & 15) == 0;
    if (tma_ok) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&stage_bar);
        constexpr uint32_t kBytesTx = kSearchBytes - kSearchOff, kLinksTx = (kSearchLinks - kSearchOff) * 2;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kBytesTx + kLinksTx) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(sb + kSearchOff)), "l"(in + wb), "r"(kBytesTx), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(sl + kSearchOff)), "l"(link + wb), "r"(kLinksTx), "r"(bar) : "memory");
            tile_next = kSearchThreads;  // the first kSearchThreads positions are pre-assigned
        }
        __syncthreads();  // barrier initialised and armed before anyone polls it
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_STAGE:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
            "@p bra DONE_STAGE;\n"
            "bra WAIT_STAGE;\n"
            "DONE_STAGE:\n"
            "}\n" ::"r"(bar) : "memory");
    } else {
        const uint32_t byte_hi = (uint32_t)min((int64_t)(kSearchBytes - kSearchOff), (int64_t)n - wb);  // exclusive
        const bool aligned = ((uintptr_t)in & 15) == 0;  // wb is a multiple of 4096
        const uint32_t v_lo = lo / 16, v_hi = aligned ? byte_hi / 16 : v_lo;
        const uint4* src = reinterpret_cast<const uint4*>(in + wb);
        uint4* dst = reinterpret_cast<uint4*>(sb + kSearchOff);
        for (uint32_t i = v_lo + threadIdx.x; i < v_hi; i += kSearchThreads) dst[i] = src[i];
        for (uint32_t i = max(lo, v_hi * 16) + threadIdx.x; i < byte_hi; i += kSearchThreads) sb[kSearchOff + i] = in[wb + i];
        for (uint32_t i = byte_hi + threadIdx.x; i < kSearchBytes - kSearchOff; i += kSearchThreads) sb[kSearchOff + i] = 0;
        const uint32_t link_hi = (uint32_t)min((int64_t)(kSearchLinks - kSearchOff), (int64_t)n - wb);
        for (uint32_t i = lo + threadIdx.x; i < link_hi; i += kSearchThreads) sl[kSearchOff + i] = link[wb + i];
        if (threadIdx.x == 0) tile_next = kSearchThreads;
        __syncthreads();
    }

    const uint32_t tile_cnt = min(kSearchTile, range_end - s);  // positions of this tile that are searched in this launch
    const uint32_t slide_J = max(n >> 15, 1u) - 1;             // slides that ever happen for a stream of n bytes
    const uint32_t quarter = lv.chain >> 2;
    // 32-bit shared-window addresses: explicit ld.shared keeps the address arithmetic out of the hot loop
    uint32_t sb_addr = (uint32_t)__cvta_generic_to_shared(sb);
    asm volatile("mov.u32 %0, %0;" : "+r"(sb_addr));  // opaque: keep it in a register instead of rematerialising it per step
    const uint32_t sl_addr = sb_addr + kSearchBytes;

    // Per-lane walk state.  The warp alternates between
    //   phase A: kStepsPerRound chain steps for every stepping lane (link load + one-byte reject test)
    //   phase B: the full compares of all pending lanes together
    // and re-arms finished lanes in batches from a tile-wide counter, so all phases run with many
    // lanes busy although chain lengths differ wildly between neighbouring positions.
    // `left` = candidates this lane may still visit before its next budget event; 0 whenever the lane
    // is not stepping, so the step guard is a single test.
    uint32_t st = kIdle, left = 0, saved = 0;
    uint32_t pi = 0, qi = 0, lim = 0, best_len = 0, best_dist = 0, snap = 0, ro_addr = 0, cb = 0, first4 = 0, max_len = 0;
    bool snapped = false;

    auto arm = [&](uint32_t k) {  // start the walk of tile position k (deflate.zig:233-245)
        if (k >= tile_cnt) return;
        const uint32_t p = s + k;
        if (p < begin) return;                // belongs to an earlier segment
        const uint32_t remaining = n - p;
        pi = kSearchOff + kHist + k;          // slot of p
        best_len = 0;
        best_dist = 0;
        snapped = false;
        snap = 0;
        if (remaining < kMinMatch) {  // Lookup.zig:24: no insertion / search with < 4 bytes left
            st = kDone;
            return;
        }
        max_len = min(remaining, kMaxMatch);  // SlidingWindow.zig:82
        qi = pi;
        // candidates must satisfy  p - q <= 32768 (deflate.zig:250)  and  q > slide base (:248, pos 0 = none)
        // slide_base(p, n) in 32-bit arithmetic (p, n < 2^31): 32768 * min(max((p + 262) >> 15, 1) - 1, J)
        const uint32_t jj = max((p + kMinLookahead) >> 15, 1u) - 1;
        const int32_t base_slot = (int32_t)(min(jj, slide_J) << 15) - (int32_t)wb + (int32_t)kSearchOff;
        lim = (uint32_t)max((int32_t)pi - (int32_t)kMaxDist, base_slot + 1);
        left = quarter;  // first budget event: the quarter snapshot (deflate.zig:241-245)
        first4 = lds_u32_unaligned(sb, pi);
        ro_addr = sb_addr + 3;  // a useful candidate agrees on byte `ro` = max(best_len, 3)
        cb = sb[pi + 3];
        st = kStepping;
    };
    // budget events (deflate.zig:241-248): after chain>>2 candidates the quarter-budget result is
    // snapshotted, after `chain` candidates the walk ends
    auto event = [&]() {
        if (!snapped) {
            snap = best_len ? pack_match(best_len, best_dist) : 0;
            snapped = true;
            left = lv.chain - quarter;
        } else {
            st = kDone;
        }
    };

    arm(threadIdx.x);
    bool exhausted = false;  // warp-uniform: the tile has no unassigned position left
    while (true) {
        // ---- phase A: chain steps (deflate.zig:248 "Hot path loop!") ----
#pragma unroll
        for (int u = 0; u < kStepsPerRound; u++) {
            if (left) {
                qi -= lds_shared_u16(sl_addr + 2 * qi);
                if ((int32_t)qi < (int32_t)lim) {  // end of chain (kNoLink), too far, or at/below the slide base
                    st = kDone;
                    left = 0;
                } else if (lds_shared_u8(ro_addr + qi) == cb) {  // may beat the best so far: needs the full compare
                    st = kPending;
                    saved = left;
                    left = 0;
                } else {
                    left--;
                }
            }
        }
        if (st == kStepping && left == 0) event();
        // ---- phase B: full compares (SlidingWindow.match with the running best as min_len) ----
        const uint32_t pend = __ballot_sync(0xffffffffu, st == kPending);
        if (pend) {
            const uint32_t stepping = __ballot_sync(0xffffffffu, st == kStepping);
            if (__popc(pend) >= tune.pend_at || stepping == 0) {
                if (st == kPending) {
                    st = kStepping;
                    left = saved - 1;  // this candidate is paid for either way
                    if (lds_u32_unaligned(sb, qi) == first4) {
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
                            best_dist = pi - qi;
                            // deflate.zig:256-259 nice; or nothing can be strictly longer: neither result can change
                            if (i >= lv.nice || i >= max_len) {
                                st = kDone;
                                left = 0;
                            } else {
                                ro_addr = sb_addr + i;
                                cb = sb[pi + i];
                            }
                        }
                    }
                    if (st == kStepping && left == 0) event();
                }
            }
        }
        // ---- results of finished lanes, re-arm ----
        const uint32_t idle = __ballot_sync(0xffffffffu, st == kIdle || st == kDone);
        if (idle) {
            const uint32_t nidle = __popc(idle);
            if ((!exhausted && nidle >= tune.refill_at) || idle == 0xffffffffu) {
                if (st == kDone) {
                    const uint32_t p = (uint32_t)(wb + (int64_t)(pi - kSearchOff));
                    const uint32_t full = best_len ? pack_match(best_len, best_dist) : 0;
                    r_full[p] = full;
                    r_quarter[p] = snapped ? snap : full;
                    st = kIdle;
                }
                if (!exhausted) {
                    uint32_t base_k = 0;
                    if (lane == 0) base_k = atomicAdd(&tile_next, nidle);
                    base_k = __shfl_sync(0xffffffffu, base_k, 0);
                    if (base_k >= tile_cnt) exhausted = true;
                    else if (st == kIdle) arm(base_k + __popc(idle & ltmask));
                }
                if (exhausted && __ballot_sync(0xffffffffu, st != kIdle) == 0) break;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2 (rolling form).  Same walk as match_search_kernel, but one persistent CTA per SM rolls a
// 64 KiB ring of bytes + links over a long run of positions, loading 4 KiB epochs ahead of the
// walks.  Lanes claim positions from a run-wide counter, so the pool of work never drains the way
// a 4096-position tile does (the tile kernel spends a large part of its time in the tail where only
// the longest chains are still walking).  Ring slot of position p is p mod 65536; links are kept as
// distances with 0xFFFF = none.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kRing = 65536;
constexpr uint32_t kEpoch = 4096;
constexpr uint32_t kRollThreads = 1024;
constexpr uint32_t kLookahead = 272;  // bytes a walk may read past its position (258 + word slack)
constexpr uint32_t kRollSmem = kRing + kRing * 2 + 64;
enum : uint32_t { kClaimed = 4 };  // lane holds a position whose data is not loaded yet

__device__ __forceinline__ uint32_t ring_u32_unaligned(uint32_t sb_addr, uint32_t p) {
    const uint32_t a0 = (p & ~3u) & (kRing - 1), a1 = (a0 + 4) & (kRing - 1);
    uint32_t w0, w1;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(sb_addr + a0));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(sb_addr + a1));
    return __funnelshift_r(w0, w1, (p & 3) * 8);
}

template <int kStepsPerRound>
__global__ void __launch_bounds__(kRollThreads, 1)
match_search_roll_kernel(const uint8_t* __restrict__ in, uint32_t begin, uint32_t n, const uint16_t* __restrict__ link,
                         LevelArgs lv, SearchTune tune, uint32_t run_epochs, uint32_t* __restrict__ r_full,
                         uint32_t* __restrict__ r_quarter) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* sb = smem_raw;
    uint16_t* sl = reinterpret_cast<uint16_t*>(smem_raw + kRing);
    volatile uint32_t* ctl = reinterpret_cast<volatile uint32_t*>(smem_raw + kRing * 3);
    // ctl[0] next position to hand out, ctl[1] avail (positions < avail have their data), ctl[2] lb (load front),
    // ctl[3] want_load, ctl[4] pmin
    const uint32_t lane = threadIdx.x & 31, tid = threadIdx.x;
    const uint32_t ltmask = (1u << lane) - 1;
    r_full -= begin;
    r_quarter -= begin;

    const uint32_t r0 = (begin / kEpoch + blockIdx.x * run_epochs) * kEpoch;
    if (r0 >= n) return;
    const uint32_t r1 = (uint32_t)min((uint64_t)n, (uint64_t)r0 + (uint64_t)run_epochs * kEpoch);  // positions [r0, r1)
    const uint32_t load_end = (uint32_t)min((uint64_t)((n + kEpoch - 1) / kEpoch) * kEpoch,
                                            (uint64_t)(r1 + kEpoch - 1) / kEpoch * kEpoch + kEpoch);

    // cooperative load of positions [a, b) (multiples of 16) into the ring; bytes past n are zero, links none
    auto load_range = [&](uint32_t a, uint32_t b) {
        const bool aligned = ((uintptr_t)in & 15) == 0;
        for (uint32_t p = a + tid * 16; p < b; p += kRollThreads * 16) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (aligned && p + 16 <= n) v = *reinterpret_cast<const uint4*>(in + p);
            else {
                uint32_t w[4] = {0, 0, 0, 0};
                for (uint32_t j = 0; j < 16 && p + j < n; j++) w[j >> 2] |= (uint32_t)in[p + j] << (8 * (j & 3));
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
            *reinterpret_cast<uint4*>(sb + (p & (kRing - 1))) = v;
        }
        for (uint32_t p = a + tid * 8; p < b; p += kRollThreads * 8) {
            uint32_t w[4];
            if (p + 8 <= n) {
                const uint4 v = *reinterpret_cast<const uint4*>(link + p);
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            } else {
                for (int k = 0; k < 4; k++) {
                    const uint32_t l0 = p + 2 * k < n ? link[p + 2 * k] : 0xffffu, l1 = p + 2 * k + 1 < n ? link[p + 2 * k + 1] : 0xffffu;
                    w[k] = l0 | (l1 << 16);
                }
            }
            *reinterpret_cast<uint4*>(sl + (p & (kRing - 1))) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    };
    {
        const uint32_t hs = r0 >= kHist ? r0 - kHist : 0;
        const uint32_t lb0 = min(load_end, r0 + 3 * kEpoch);
        load_range(hs, lb0);
        if (tid == 0) {
            ctl[0] = r0;
            ctl[2] = lb0;
            ctl[1] = lb0 >= n ? r1 : min(r1, lb0 - kLookahead);
            ctl[3] = 0;
            ctl[4] = 0xffffffffu;
        }
    }
    __syncthreads();

    const uint32_t quarter = lv.chain >> 2;
    uint32_t sb_addr = (uint32_t)__cvta_generic_to_shared(sb);
    asm volatile("mov.u32 %0, %0;" : "+r"(sb_addr));
    const uint32_t sl_addr = sb_addr + kRing;

    uint32_t st = kIdle, left = 0, saved = 0;
    uint32_t p = 0, q = 0, best_len = 0, best_dist = 0, snap = 0, ro = 3, cb = 0, first4 = 0, max_len = 0;
    int32_t lim = 0;
    bool snapped = false;

    auto arm = [&]() {  // start the walk of position p (deflate.zig:233-245); its data is in the ring
        best_len = 0;
        best_dist = 0;
        snapped = false;
        snap = 0;
        if (p < begin) {  // belongs to an earlier segment
            st = kIdle;
            return;
        }
        const uint32_t remaining = n - p;
        if (remaining < kMinMatch) {  // Lookup.zig:24
            st = kDone;
            return;
        }
        max_len = min(remaining, kMaxMatch);  // SlidingWindow.zig:82
        q = p;
        // candidates must satisfy  p - q <= 32768 (deflate.zig:250)  and  q > slide base (:248, pos 0 = none)
        lim = (int32_t)max((int64_t)p - (int64_t)kMaxDist, (int64_t)slide_base(p, n) + 1);
        left = quarter;
        first4 = ring_u32_unaligned(sb_addr, p);
        ro = 3;
        cb = lds_shared_u8(sb_addr + ((p + 3) & (kRing - 1)));
        st = kStepping;
    };
    auto event = [&]() {
        if (!snapped) {
            snap = best_len ? pack_match(best_len, best_dist) : 0;
            snapped = true;
            left = lv.chain - quarter;
        } else {
            st = kDone;
        }
    };

    while (true) {
        // ---- epoch loads: every warp passes here once per round, so the barriers line up ----
        if (ctl[3]) {
            __syncthreads();
            {
                uint32_t mine = (st == kStepping || st == kPending) ? p : 0xffffffffu;
                for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
                if (lane == 0 && mine != 0xffffffffu) atomicMin((uint32_t*)&ctl[4], mine);
            }
            __syncthreads();
            const uint32_t lb = ctl[2], pmin = ctl[4];
            // the new epoch overwrites positions [lb - 65536, lb - 61440): every walk in flight must be past them
            const bool safe = lb < load_end && (pmin == 0xffffffffu || (uint64_t)pmin + 28672 >= lb);
            if (safe) load_range(lb, lb + kEpoch);
            __syncthreads();
            if (tid == 0) {
                if (safe) {
                    const uint32_t nlb = lb + kEpoch;
                    ctl[2] = nlb;
                    ctl[1] = nlb >= n ? r1 : min(r1, nlb - kLookahead);
                }
                ctl[3] = 0;
                ctl[4] = 0xffffffffu;
            }
            __syncthreads();
        }
        // ---- phase A: chain steps (deflate.zig:248 "Hot path loop!") ----
#pragma unroll
        for (int u = 0; u < kStepsPerRound; u++) {
            if (left) {
                const uint32_t l = lds_shared_u16(sl_addr + ((q & (kRing - 1)) << 1));
                q -= l;
                if ((int32_t)q < lim) {  // end of chain (l = 0xFFFF), too far, or at/below the slide base
                    st = kDone;
                    left = 0;
                } else if (lds_shared_u8(sb_addr + ((q + ro) & (kRing - 1))) == cb) {
                    st = kPending;
                    saved = left;
                    left = 0;
                } else {
                    left--;
                }
            }
        }
        if (st == kStepping && left == 0) event();
        // ---- phase B: full compares (SlidingWindow.match with the running best as min_len) ----
        const uint32_t pend = __ballot_sync(0xffffffffu, st == kPending);
        if (pend) {
            const uint32_t stepping = __ballot_sync(0xffffffffu, st == kStepping);
            if (__popc(pend) >= tune.pend_at || stepping == 0) {
                if (st == kPending) {
                    st = kStepping;
                    left = saved - 1;
                    if (ring_u32_unaligned(sb_addr, q) == first4) {
                        uint32_t i = 4;
                        while (i < max_len) {
                            const uint32_t x = ring_u32_unaligned(sb_addr, q + i) ^ ring_u32_unaligned(sb_addr, p + i);
                            if (x) {
                                i += (__ffs(x) - 1) >> 3;
                                break;
                            }
                            i += 4;
                        }
                        if (i > max_len) i = max_len;
                        if (i > best_len) {
                            best_len = i;
                            best_dist = p - q;
                            if (i >= lv.nice || i >= max_len) {  // deflate.zig:256-259, or nothing can be longer
                                st = kDone;
                                left = 0;
                            } else {
                                ro = i;
                                cb = lds_shared_u8(sb_addr + ((p + i) & (kRing - 1)));
                            }
                        }
                    }
                    if (st == kStepping && left == 0) event();
                }
            }
        }
        // ---- results of finished lanes, claims, re-arm ----
        const uint32_t avail = ctl[1];
        if (st == kClaimed && p < avail) arm();
        const uint32_t idle = __ballot_sync(0xffffffffu, st == kIdle || st == kDone);
        if (idle) {
            const uint32_t nidle = __popc(idle);
            const uint32_t nxt = ctl[0];
            const bool exhausted = nxt >= r1;
            if ((!exhausted && nidle >= tune.refill_at) || idle == 0xffffffffu) {
                if (st == kDone) {
                    const uint32_t full = best_len ? pack_match(best_len, best_dist) : 0;
                    r_full[p] = full;
                    r_quarter[p] = snapped ? snap : full;
                    st = kIdle;
                }
                if (!exhausted && nxt < avail) {
                    uint32_t base_p = 0;
                    if (lane == 0) base_p = atomicAdd((uint32_t*)&ctl[0], nidle);
                    base_p = __shfl_sync(0xffffffffu, base_p, 0);
                    if (st == kIdle) {
                        const uint32_t mine = base_p + __popc(idle & ltmask);
                        if (mine < r1) {
                            p = mine;
                            if (p < avail) arm();
                            else st = kClaimed;
                        }
                    }
                }
            }
        }
        // ask for the next epoch when the hand-out front gets close to the loaded front (every warp, every
        // round: waiting lanes must not depend on somebody else being idle)
        if (lane == 0 && ctl[2] < load_end && ctl[0] + 2048 >= avail) ctl[3] = 1;
        if (ctl[2] >= load_end && ctl[0] >= r1 && __ballot_sync(0xffffffffu, st != kIdle) == 0) break;
    }
}

// ------------------------------------------------------------------------------------------
// K2s: speculative sparse parse (findMatch + the lazy rule, only where a parse can arrive).
//
// The step f(p) of the lazy parse from a clean arrival p (deflate.zig:160-193: k deferred literals,
// then one match, or one literal) is a pure function of p, and the reference only ever evaluates it
// on the orbit of 0.  Parses started at different positions fall into step after a few tokens, so it
// is enough to evaluate f on the orbits of many seeds: a CTA owns kT positions plus an overlap of kW
// positions of the next CTA's range, puts a seed every kG positions, and every lane follows one
// seed's orbit, searching on demand (with the real min_len and the real budget of deflate.zig:241-245)
// and claiming each arrival in a bitmap.  A lane stops when it reaches an arrival somebody else has
// claimed: that lane carries the orbit on.  So the set of evaluated arrivals is closed under f up to
// the end of the span, every seed is in it, and position 0 is a seed: the true orbit stays inside it
// as long as every orbit that started in a CTA's own range has joined the orbit of one of the overlap
// seeds (which the next CTA evaluates as well) before the span ends.  That is checked exactly (`safe`
// bitmap); if it ever fails the host falls back to the dense tables (match_search_kernel), so the
// result never depends on the speculation.  nx[] is pre-filled with kNxInvalid by the caller; the
// token emitter reports any invalid entry it meets on the true orbit as a second line of defence.
// On text this evaluates ~8x fewer chain candidates than searching every position.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kSpHalo = 256;  // a lazy run searches up to 255 positions past its arrival
struct SparseTune {
    uint32_t pend_at;    // run the batched full compares once this many lanes wait for one
    uint32_t done_at;    // run the lazy decisions once this many lanes finished a walk
    uint32_t refill_at;  // hand out new seeds once this many lanes are idle
};
template <uint32_t kT, uint32_t kW>
struct SparseCfg {
    static constexpr uint32_t kSpan = kT + kW;
    static constexpr uint32_t kBytes = kSearchOff + kHist + kSpan + kSpHalo + 272;
    static constexpr uint32_t kLinks = kSearchOff + kHist + kSpan + kSpHalo;
    static constexpr uint32_t kWords = kSpan / 32;
    static constexpr uint32_t kSmem = kBytes + kLinks * 2 + kWords * 8;
};
enum : uint32_t { kSearchDone = 4, kArrive = 5, kStart = 6 };
constexpr uint32_t kMaxCross = 32;

template <uint32_t kT, uint32_t kW, uint32_t kG, uint32_t kThreads, int kMinBlocks, int kSteps, bool kDense>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
sparse_parse_kernel(const uint8_t* __restrict__ in, uint32_t begin, uint32_t first_chunk, const uint32_t* __restrict__ chunk_list,
                    uint32_t n, const uint16_t* __restrict__ link, LevelArgs lv, SparseTune tune, uint32_t* __restrict__ nx,
                    uint32_t* __restrict__ chunk_fail, uint32_t* __restrict__ flags) {
    // begin: first position of the segment being compressed (0 for a whole stream; > 0 after a sync flush, when
    // earlier bytes are history only).  The parse restarts clean at `begin`, so `begin` is a seed; positions before
    // it are never evaluated.  nx is indexed by stream position (the caller passes its table shifted by -begin).
    // kDense: every position of the CTA's own range is a seed (the overlap keeps the regular seeds, the ones
    // the next CTA evaluates too).  Used to repair after a failed coverage check: once every entry of chunk
    // c+1 is evaluated, it no longer matters where the orbits leaving chunk c arrive.
    // chunk_list (may be null): the chunks to process, one per CTA; else chunk = first_chunk + blockIdx.x.
    using C = SparseCfg<kT, kW>;
    static_assert(C::kBytes % 16 == 0 && (C::kLinks * 2) % 16 == 0 && kT % kG == 0 && kW % kG == 0, "layout");
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint32_t seed_next, cross_cnt, failed;
    __shared__ uint32_t cross[kMaxCross];
    __shared__ __align__(8) uint64_t stage_bar;
    uint8_t* sb = smem_raw;                                               // bytes, slot i+16 = position wb+i
    uint16_t* sl = reinterpret_cast<uint16_t*>(smem_raw + C::kBytes);     // link (distance, kNoLink = none) per slot
    uint32_t* valid = reinterpret_cast<uint32_t*>(smem_raw + C::kBytes + C::kLinks * 2);  // arrivals claimed by some lane
    uint32_t* safe = valid + C::kWords;                                   // arrivals on the orbit of an overlap seed
    const uint32_t chunk = chunk_list ? chunk_list[blockIdx.x] : first_chunk + blockIdx.x;
    const uint32_t s = chunk * kT;                                        // first own position
    const int64_t wb = (int64_t)s - kHist;
    const uint32_t lo = wb < 0 ? (uint32_t)(-wb) : 0;
    const uint32_t span_len = min(C::kSpan, n - s);                       // arrivals evaluated here: offsets < span_len
    const bool open_end = (uint64_t)s + C::kSpan < n;                     // the stream goes on past the span
    const uint32_t own_len = min(kT, span_len);
    const uint32_t n_ov = (span_len - own_len + kG - 1) / kG;              // regular seeds in the overlap
    const uint32_t nreg = kDense ? n_ov + own_len : (span_len + kG - 1) / kG;  // regular seeds
    const bool has_begin = begin > s && begin - s < own_len;                   // the segment starts inside this chunk
    const uint32_t nseeds = nreg + (has_begin ? 1u : 0u);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ltmask = (1u << lane) - 1;

    for (uint32_t i = threadIdx.x; i < 2 * C::kWords; i += kThreads) valid[i] = 0;
    if (threadIdx.x == 0) {
        seed_next = 0;
        cross_cnt = 0;
        failed = 0;
    }
    // ---- stage window (same scheme as match_search_kernel) ----
    constexpr uint32_t kBytesTx = C::kBytes - kSearchOff, kLinksTx = (C::kLinks - kSearchOff) * 2;
    const bool tma_ok = wb >= 0 && (uint64_t)s + C::kSpan + kSpHalo + 272 <= n && (((uintptr_t)in | (uintptr_t)link) & 15) == 0;
    if (tma_ok) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&stage_bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kBytesTx + kLinksTx) : "memory");
            constexpr uint32_t kPiece = 32768;  // bulk copies in pieces of at most 32 KiB
            for (uint32_t o = 0; o < kBytesTx; o += kPiece) {
                const uint32_t len = min(kPiece, kBytesTx - o);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"((uint32_t)__cvta_generic_to_shared(sb + kSearchOff + o)), "l"(in + wb + o), "r"(len), "r"(bar) : "memory");
            }
            const uint8_t* lsrc = reinterpret_cast<const uint8_t*>(link + wb);
            uint8_t* ldst = reinterpret_cast<uint8_t*>(sl + kSearchOff);
            for (uint32_t o = 0; o < kLinksTx; o += kPiece) {
                const uint32_t len = min(kPiece, kLinksTx - o);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"((uint32_t)__cvta_generic_to_shared(ldst + o)), "l"(lsrc + o), "r"(len), "r"(bar) : "memory");
            }
        }
        __syncthreads();
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_SP:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
            "@p bra DONE_SP;\n"
            "bra WAIT_SP;\n"
            "DONE_SP:\n"
            "}\n" ::"r"(bar) : "memory");
    } else {
        const uint32_t byte_hi = (uint32_t)min((int64_t)kBytesTx, (int64_t)n - wb);  // exclusive
        const bool aligned = ((uintptr_t)in & 15) == 0;
        const uint32_t v_lo = (lo + 15) / 16, v_hi = aligned ? byte_hi / 16 : v_lo;
        const uint4* src = reinterpret_cast<const uint4*>(in + wb);
        uint4* dst = reinterpret_cast<uint4*>(sb + kSearchOff);
        for (uint32_t i = v_lo + threadIdx.x; i < v_hi; i += kThreads) dst[i] = src[i];
        for (uint32_t i = lo + threadIdx.x; i < min(v_lo * 16, byte_hi); i += kThreads) sb[kSearchOff + i] = in[wb + i];
        for (uint32_t i = max(lo, v_hi * 16) + threadIdx.x; i < byte_hi; i += kThreads) sb[kSearchOff + i] = in[wb + i];
        for (uint32_t i = byte_hi + threadIdx.x; i < kBytesTx; i += kThreads) sb[kSearchOff + i] = 0;
        const uint32_t link_hi = (uint32_t)min((int64_t)(C::kLinks - kSearchOff), (int64_t)n - wb);
        for (uint32_t i = lo + threadIdx.x; i < link_hi; i += kThreads) sl[kSearchOff + i] = link[wb + i];
        __syncthreads();
    }

    const uint32_t slide_J = max(n >> 15, 1u) - 1;
    const uint32_t quarter = lv.chain >> 2;
    uint32_t sb_addr = (uint32_t)__cvta_generic_to_shared(sb);
    asm volatile("mov.u32 %0, %0;" : "+r"(sb_addr));
    const uint32_t sl_addr = sb_addr + C::kBytes;

    uint32_t st = kIdle, left = 0, saved = 0;
    uint32_t pi = 0, qi = 0, lim = 0, best_len = 0, best_dist = 0, ro_addr = 0, cb = 0, first4 = 0, max_len = 0;
    uint32_t p0 = 0, curk = 0, cur_len = 0, cur_dist = 0;  // the arrival being evaluated and its pending match
    bool from_overlap = false;

    // Lane states beyond the walk itself: kSearchDone (walk ended, lazy rule pending), kArrive (the orbit
    // reaches the clean arrival a_rel), kStart (findMatch at a_rel with min_len = best_len and `saved`
    // candidates is about to begin).  Every transition has exactly one copy of its code in the loop below.
    uint32_t a_rel = 0;
    // findMatch(pos = s + a_rel, min_len = best_len) with `saved` candidates (deflate.zig:233-266)
    auto start_search = [&]() {
        const uint32_t rel = a_rel, min_len = best_len, budget = saved;
        best_dist = 0;
        left = 0;
        st = kSearchDone;
        const uint32_t p = s + rel;
        if (p >= n) return;
        const uint32_t remaining = n - p;
        if (remaining < kMinMatch) return;     // Lookup.zig:24
        max_len = min(remaining, kMaxMatch);   // SlidingWindow.zig:82
        if (min_len >= max_len || budget == 0) return;
        pi = kSearchOff + kHist + rel;
        qi = pi;
        const uint32_t jj = max((p + kMinLookahead) >> 15, 1u) - 1;
        const int32_t base_slot = (int32_t)(min(jj, slide_J) << 15) - (int32_t)wb + (int32_t)kSearchOff;
        lim = (uint32_t)max((int32_t)pi - (int32_t)kMaxDist, base_slot + 1);
        left = budget;
        first4 = lds_u32_unaligned(sb, pi);
        const uint32_t ro = max(min_len, 3u);  // a longer match agrees on byte max(best_len, 3)
        ro_addr = sb_addr + ro;
        cb = sb[pi + ro];
        st = kStepping;
    };
    // the orbit reaches the clean arrival s + a_rel (p0 still holds the previous arrival)
    auto arrive = [&]() {
        const uint32_t rel = a_rel;
        st = kIdle;
        if (rel >= span_len) {
            if (!from_overlap && open_end) {  // must have joined an overlap seed's orbit: checked after the loop
                const uint32_t idx = atomicAdd(&cross_cnt, 1u);
                if (idx < kMaxCross) cross[idx] = p0;
            }
            return;
        }
        const uint32_t w = rel >> 5, bit = 1u << (rel & 31);
        const uint32_t old = atomicOr(&valid[w], bit);
        if (from_overlap) {
            if (atomicOr(&safe[w], bit) & bit) return;  // another overlap seed's lane carries on from here
        } else if (old & bit) {
            return;                                      // somebody carries on from here
        }
        p0 = rel;
        curk = 0;
        cur_len = 0;
        cur_dist = 0;
        best_len = 0;
        saved = lv.chain;
        st = kStart;
    };
    // a walk ended: apply the lazy rule (deflate.zig:166-190)
    auto decide = [&]() {
        const bool found = best_dist != 0;
        bool fin = false;
        if (cur_len == 0) {
            if (found) {
                cur_len = best_len;
                cur_dist = best_dist;
            } else {
                fin = true;  // plain literal
            }
        } else if (found) {  // better match one byte later: the pending one becomes a literal
            curk++;
            cur_len = best_len;
            cur_dist = best_dist;
        } else {
            fin = true;      // emit the pending match
        }
        if (!fin && cur_len >= lv.lazy) fin = true;  // deflate.zig:171
        if (!fin) {
            a_rel = p0 + curk + 1;
            best_len = cur_len;
            saved = cur_len >= lv.good ? quarter : lv.chain;  // deflate.zig:241-245
            st = kStart;
        } else {
            uint32_t out = 0, step = 1;
            if (cur_len) {
                out = curk | ((cur_len - 3) << 8) | (cur_dist << 16);
                step = curk + cur_len;
            }
            nx[s + p0] = out;
            a_rel = p0 + step;
            st = kArrive;
        }
    };

    bool exhausted = false;
    while (true) {
        // ---- phase A: chain steps (deflate.zig:248 "Hot path loop!") ----
        // The link of the next candidate is fetched together with the reject byte of this one, before it is
        // known whether the walk goes on: one shared-memory latency per step on the critical path, not two.
        uint32_t nl = left ? lds_shared_u16(sl_addr + 2 * qi) : 0;
#pragma unroll
        for (int u = 0; u < kSteps; u++) {
            if (left) {
                qi -= nl;
                const bool in_range = (int32_t)qi >= (int32_t)lim;
                const uint32_t qs = in_range ? qi : pi;  // a slot that is always safe to read
                const uint32_t b = lds_shared_u8(ro_addr + qs);
                nl = lds_shared_u16(sl_addr + 2 * qs);
                if (!in_range) {  // end of chain (kNoLink), too far, or at/below the slide base
                    st = kSearchDone;
                    left = 0;
                } else if (b == cb) {  // may beat the best so far: needs the full compare
                    st = kPending;
                    saved = left;
                    left = 0;
                } else {
                    left--;
                }
            }
        }
        if (st == kStepping && left == 0) st = kSearchDone;
        // ---- phase B: full compares (SlidingWindow.match with the running best as min_len) ----
        const uint32_t pend = __ballot_sync(0xffffffffu, st == kPending);
        if (pend) {
            const uint32_t stepping = __ballot_sync(0xffffffffu, st == kStepping);
            if (__popc(pend) >= tune.pend_at || stepping == 0) {
                if (st == kPending) {
                    st = kStepping;
                    left = saved - 1;  // this candidate is paid for either way
                    if (lds_u32_unaligned(sb, qi) == first4) {
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
                            best_dist = pi - qi;
                            if (i >= lv.nice || i >= max_len) {  // deflate.zig:256-259, or nothing can be longer
                                st = kSearchDone;
                                left = 0;
                            } else {
                                ro_addr = sb_addr + i;
                                cb = sb[pi + i];
                            }
                        }
                    }
                    if (st == kStepping && left == 0) st = kSearchDone;
                }
            }
        }
        // ---- phase C: lazy decisions of finished walks, next arrival of the orbit ----
        const uint32_t done = __ballot_sync(0xffffffffu, st == kSearchDone);
        if (done) {
            const uint32_t busy = __ballot_sync(0xffffffffu, st == kStepping || st == kPending);
            if (__popc(done) >= tune.done_at || busy == 0) {
                if (st == kSearchDone) decide();
            }
        }
        // ---- phase D: new seeds for idle lanes, last segment first (a lane soon meets the trail of the
        // segment ahead) ----
        const uint32_t idle = __ballot_sync(0xffffffffu, st == kIdle);
        if (idle && !exhausted && (__popc(idle) >= tune.refill_at || idle == 0xffffffffu)) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&seed_next, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= nseeds) {
                exhausted = true;
            } else if (st == kIdle) {
                const uint32_t t = base + __popc(idle & ltmask);
                if (t < nseeds) {
                    if (t == nreg) a_rel = begin - s;  // the segment's first position: where the true orbit starts
                    else if (kDense) a_rel = t < n_ov ? kT + (n_ov - 1 - t) * kG : own_len - 1 - (t - n_ov);
                    else a_rel = (nreg - 1 - t) * kG;
                    from_overlap = a_rel >= kT;
                    if (s + a_rel >= begin) st = kArrive;  // positions before the segment are history, not seeds
                }
            }
        }
        // ---- phase E: arrivals (claim, or stop at somebody's trail), phase F: walks begin ----
        if (__any_sync(0xffffffffu, st == kArrive)) {
            if (st == kArrive) arrive();
        }
        if (__any_sync(0xffffffffu, st == kStart)) {
            if (st == kStart) start_search();
        }
        if (exhausted && __ballot_sync(0xffffffffu, st != kIdle) == 0) break;
    }
    __syncthreads();
    if (open_end) {
        const uint32_t cnt = cross_cnt;
        if (threadIdx.x < min(cnt, kMaxCross)) {
            const uint32_t r = cross[threadIdx.x];
            if (!((safe[r >> 5] >> (r & 31)) & 1u)) failed = 1;
        }
        if (threadIdx.x == 0 && cnt > kMaxCross) failed = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        chunk_fail[chunk] = failed;  // orbits of this chunk's own range may arrive anywhere in the next chunk
        if (failed) atomicOr(flags, 1u);
    }
}

// ------------------------------------------------------------------------------------------
// K2s, rolling form (opt-in with FB200_SPARSE_ROLL=16; measured slower than the chunk kernel).
//
// Same orbits, same claims, same result as sparse_parse_kernel, but a persistent CTA walks a RUN of many
// chunks with one window that rolls forward, so lanes never wait for the slowest orbit of a chunk: a lane
// whose orbit ended takes the next seed at once.  The window is a ring of 65536 positions addressed by the
// low 16 bits of the stream position (bytes, links, one claim bit per position); it is loaded by TMA bulk
// copies in epochs of 4096 positions.  Seeds are handed out in ascending order, one every kG positions.
//   resident positions      [lb - 65536, lb)          lb = load front (shared, only ever grows)
//   a search at p needs     [p - 32768, p + 272)      (candidates reach 32768 back, compares 258 + slack ahead)
// An orbit is registered in the epoch of the arrival it is evaluating (live[epoch & 15]; the seeds of an
// epoch are counted in when the epoch is loaded), `tail` is the oldest epoch that still has a registered
// orbit or an unclaimed seed, and the front moves by one epoch when tail * 4096 - 32768 >= lb - 65536 + 4096.
// A lane whose next position is not resident yet simply waits in place (its registration is within two
// epochs of the front, so it never holds the tail back); the slowest orbits are far behind the front and
// always free to move, hence progress.  Inside a run there is no hand-over: all orbits share the claim
// bitmap.  At the end of the run the overlap of kW positions is evaluated as in the chunk kernel and the
// same exact check decides chunk_fail of the run's last chunk.
// Preconditions (the launcher sends everything else to the chunk kernel): `in` and `link` 16-byte aligned,
// every run at least two chunks long, run end + kW + 256 + 272 rounded up to an epoch <= n.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kSrRing = 65536, kSrEpoch = 4096, kSrLook = 272, kSrMirror = 288;
constexpr uint32_t kSrBytes = kSrRing + kSrMirror;
constexpr uint32_t kSrSmem = kSrBytes + kSrRing * 2 + kSrRing / 8;
constexpr uint32_t kSrChunk = 32768;  // granularity of runs and of chunk_fail (= kSparseT)

struct SrShared {
    uint32_t live[16];
    uint32_t cross[kMaxCross];
    uint32_t safe[32];  // kW / 32 words: arrivals of the run's overlap that lie on an overlap seed's orbit
    uint32_t seed_next, cross_cnt, failed, run, lb, tail, inflight, parity, warps_done;
    // the run being worked on
    uint32_t s0, s1, span_end, load_end, seed_lo, has_begin, nseeds;
};

// seeds of epoch e that will be handed out (regular ones at or past seed_lo, plus the segment start)
__device__ __forceinline__ uint32_t sr_epoch_seeds(const SrShared& S, uint32_t e, uint32_t kG, uint32_t begin) {
    const uint32_t lo = max(e * kSrEpoch, S.seed_lo), hi = min((e + 1) * kSrEpoch, S.span_end);
    uint32_t c = hi > lo ? (hi - lo + kG - 1) / kG : 0;  // lo is a multiple of kG
    if (S.has_begin && (begin >> 12) == e) c++;
    return c;
}

__device__ __forceinline__ void sr_load_epoch(uint32_t bar, uint8_t* sb, uint16_t* sl, const uint8_t* in, const uint16_t* link, uint32_t pos) {
    const uint32_t slot = pos & (kSrRing - 1);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(sb + slot)), "l"(in + pos), "r"(kSrEpoch), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(sl + slot)), "l"(link + pos), "r"(kSrEpoch * 2), "r"(bar) : "memory");
    if (slot == 0)  // the first bytes of the ring again behind its end: compares read across the wrap
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(sb + kSrRing)), "l"(in + pos), "r"(kSrMirror), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t sr_epoch_tx(uint32_t pos) { return kSrEpoch * 3 + ((pos & (kSrRing - 1)) == 0 ? kSrMirror : 0); }
__device__ __forceinline__ bool sr_bar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// Load front duty, warp 0 once per round: retire finished epochs, start the next epoch's load when its ring
// slots are free, publish it when it has landed.
__device__ __noinline__ void sr_loader(SrShared& S, uint32_t bar, uint8_t* sb, uint16_t* sl, uint32_t* valid, const uint8_t* in,
                                       const uint16_t* link, uint32_t kG, uint32_t begin) {
    const uint32_t lane = threadIdx.x & 31;
    volatile SrShared& V = S;
    const uint32_t lb = V.lb;
    uint32_t go = 0;
    if (lane == 0) {
        if (V.inflight) {
            if (sr_bar_test(bar, V.parity)) {
                V.parity = V.parity ^ 1u;
                V.inflight = 0;
                __threadfence_block();
                V.lb = lb + kSrEpoch;  // publish: the epoch's bytes, links, cleared claim bits and seed count are in place
            }
        } else if (lb < V.load_end) {
            uint32_t tail = V.tail;
            while (tail * kSrEpoch + 32768 < lb + kSrEpoch && V.live[tail & 15] == 0) tail++;
            V.tail = tail;
            if (tail * kSrEpoch + 32768 >= lb + kSrEpoch) go = 1;
        }
    }
    go = __shfl_sync(0xffffffffu, go, 0);
    if (go) {
        // epoch lb / 4096 replaces epoch lb / 4096 - 16, which nobody can reach any more
        const uint32_t w0 = (lb & (kSrRing - 1)) >> 5;
        for (uint32_t i = lane; i < kSrEpoch / 32; i += 32) valid[w0 + i] = 0;
        __syncwarp();
        if (lane == 0) {
            V.live[(lb >> 12) & 15] = sr_epoch_seeds(S, lb >> 12, kG, begin);
            __threadfence_block();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(sr_epoch_tx(lb)) : "memory");
            sr_load_epoch(bar, sb, sl, in, link, lb);
            V.inflight = 1;
        }
    }
}

template <uint32_t kW, uint32_t kG, uint32_t kThreads, int kSteps>
__global__ void __launch_bounds__(kThreads, 1)
sparse_roll_kernel(const uint8_t* __restrict__ in, uint32_t begin, uint32_t first_chunk, uint32_t end_chunk, uint32_t nruns, uint32_t n,
                   const uint16_t* __restrict__ link, LevelArgs lv, uint32_t* __restrict__ nx, uint32_t* __restrict__ chunk_fail,
                   uint32_t* __restrict__ flags, uint32_t* __restrict__ run_counter, uint32_t prio) {
    // prio > 0: a warp none of whose orbits is registered within `prio` epochs of the tail yields issue slots (short sleep)
    static_assert(kW == 1024 && kSrEpoch % kG == 0 && kSrChunk % kSrEpoch == 0, "layout");
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ SrShared S;
    __shared__ __align__(8) uint64_t stage_bar;
    uint8_t* sb = smem_raw;                                                    // byte of position p at p & 0xffff
    uint16_t* sl = reinterpret_cast<uint16_t*>(smem_raw + kSrBytes);           // link of position p at p & 0xffff
    uint32_t* valid = reinterpret_cast<uint32_t*>(smem_raw + kSrBytes + kSrRing * 2);  // claim bit of position p at p & 0xffff
    volatile SrShared& V = S;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ltmask = (1u << lane) - 1;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&stage_bar);
    constexpr uint32_t kWarps = kThreads / 32;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.parity = 0;
    }
    const uint32_t slide_J = max(n >> 15, 1u) - 1;
    const uint32_t quarter = lv.chain >> 2;
    uint32_t sb_addr = (uint32_t)__cvta_generic_to_shared(sb);
    asm volatile("mov.u32 %0, %0;" : "+r"(sb_addr));
    const uint32_t sl_addr = sb_addr + kSrBytes;
    const uint32_t nch = end_chunk - first_chunk;

    for (;;) {
        // ---- next run ----
        __syncthreads();  // the previous run is over for every thread (and the barrier is initialised)
        if (threadIdx.x == 0) {
            const uint32_t r = atomicAdd(run_counter, 1u);
            S.run = r;
            if (r < nruns) {
                const uint32_t c0 = first_chunk + (uint32_t)(((uint64_t)nch * r) / nruns);
                const uint32_t c1 = first_chunk + (uint32_t)(((uint64_t)nch * (r + 1)) / nruns);
                S.s0 = c0 * kSrChunk;
                S.s1 = c1 * kSrChunk;
                S.span_end = S.s1 + kW;
                S.load_end = (S.span_end + kSpHalo + kSrLook + kSrEpoch - 1) / kSrEpoch * kSrEpoch;
                S.has_begin = (begin > S.s0 && begin < S.s1) ? 1u : 0u;
                S.seed_lo = begin > S.s0 ? (begin + kG - 1) / kG * kG : S.s0;  // regular seeds before the segment start are not seeds
                S.nseeds = (S.span_end - S.s0) / kG + S.has_begin;
                S.seed_next = 0;
                S.cross_cnt = 0;
                S.failed = 0;
                S.warps_done = 0;
                S.inflight = 0;
            }
        }
        __syncthreads();
        if (S.run >= nruns) break;
        const uint32_t s0 = S.s0, s1 = S.s1, span_end = S.span_end, nseeds = S.nseeds, has_begin = S.has_begin;
        const uint32_t lb0 = max(s0 + 32768u, kSrRing);  // first load: the whole ring
        for (uint32_t i = threadIdx.x; i < kSrRing / 32; i += kThreads) valid[i] = 0;
        if (threadIdx.x < 32) S.safe[threadIdx.x] = 0;
        if (threadIdx.x < 16) {
            const uint32_t e = (lb0 >> 12) - 16 + threadIdx.x;
            S.live[e & 15] = sr_epoch_seeds(S, e, kG, begin);
        }
        if (threadIdx.x == 0) {
            S.tail = s0 >> 12;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            uint32_t tx = 0;
            for (uint32_t e = 0; e < 16; e++) tx += sr_epoch_tx(lb0 - kSrRing + e * kSrEpoch);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
            for (uint32_t e = 0; e < 16; e++) sr_load_epoch(bar, sb, sl, in, link, lb0 - kSrRing + e * kSrEpoch);
        }
        __syncthreads();
        {
            const uint32_t par = S.parity;
            while (!sr_bar_test(bar, par)) {}
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            S.parity ^= 1u;
            S.lb = lb0;
        }
        __syncthreads();

        uint32_t st = kIdle, left = 0, saved = 0;
        uint32_t pi = 0, qi = 0, lim = 0, best_len = 0, best_dist = 0, ro_addr = 0, cb = 0, first4 = 0, max_len = 0;
        uint32_t p0 = 0, curk = 0, cur_len = 0, cur_dist = 0;  // the arrival being evaluated and its pending match
        uint32_t a_pos = 0;                                    // stream position of the next arrival / search
        bool from_overlap = false;

        // findMatch(pos = a_pos, min_len = best_len) with `saved` candidates (deflate.zig:233-266)
        auto start_search = [&]() {
            const uint32_t p = a_pos;
            if (p + kSrLook > V.lb) return;  // not resident yet: stays in kStart
            const uint32_t min_len = best_len, budget = saved;
            best_dist = 0;
            left = 0;
            st = kSearchDone;
            if (p >= n) return;
            const uint32_t remaining = n - p;
            if (remaining < kMinMatch) return;     // Lookup.zig:24
            max_len = min(remaining, kMaxMatch);   // SlidingWindow.zig:82
            if (min_len >= max_len || budget == 0) return;
            pi = p;
            qi = p;
            const uint32_t jj = max((p + kMinLookahead) >> 15, 1u) - 1;
            const int32_t base = (int32_t)(min(jj, slide_J) << 15);
            lim = (uint32_t)max((int32_t)p - (int32_t)kMaxDist, base + 1);
            left = budget;
            const uint32_t ps = p & (kSrRing - 1);
            first4 = lds_u32_unaligned(sb, ps);
            const uint32_t ro = max(min_len, 3u);  // a longer match agrees on byte max(best_len, 3)
            ro_addr = sb_addr + ro;
            cb = sb[ps + ro];
            st = kStepping;
        };
        // the orbit reaches the clean arrival a_pos (p0 still holds the arrival it is registered at)
        auto arrive = [&]() {
            const uint32_t a = a_pos;
            bool ended = false;
            if (a >= span_end) {
                if (!from_overlap) {  // must have joined an overlap seed's orbit: checked after the run
                    const uint32_t idx = atomicAdd(&S.cross_cnt, 1u);
                    if (idx < kMaxCross) S.cross[idx] = p0;
                }
                ended = true;
            } else {
                if (a + kSrLook > V.lb) return;  // not resident yet: stays in kArrive
                const uint32_t w = (a & (kSrRing - 1)) >> 5, bit = 1u << (a & 31);
                const uint32_t old = atomicOr(&valid[w], bit);
                if (from_overlap) {
                    const uint32_t o = a - s1;  // an overlap seed's orbit never leaves [s1, span_end) alive
                    if (atomicOr(&S.safe[o >> 5], 1u << (o & 31)) & (1u << (o & 31))) ended = true;  // another overlap seed's lane carries on
                } else if (old & bit) {
                    ended = true;  // somebody carries on from here
                }
            }
            if (ended) {
                atomicSub(&S.live[(p0 >> 12) & 15], 1u);
                st = kIdle;
                return;
            }
            if ((a ^ p0) >> 12) {  // register in the new epoch first, so that the orbit is never unaccounted for
                atomicAdd(&S.live[(a >> 12) & 15], 1u);
                atomicSub(&S.live[(p0 >> 12) & 15], 1u);
            }
            p0 = a;
            curk = 0;
            cur_len = 0;
            cur_dist = 0;
            best_len = 0;
            saved = lv.chain;
            st = kStart;
        };
        // a walk ended: apply the lazy rule (deflate.zig:166-190)
        auto decide = [&]() {
            const bool found = best_dist != 0;
            bool fin = false;
            if (cur_len == 0) {
                if (found) {
                    cur_len = best_len;
                    cur_dist = best_dist;
                } else {
                    fin = true;  // plain literal
                }
            } else if (found) {  // better match one byte later: the pending one becomes a literal
                curk++;
                cur_len = best_len;
                cur_dist = best_dist;
            } else {
                fin = true;      // emit the pending match
            }
            if (!fin && cur_len >= lv.lazy) fin = true;  // deflate.zig:171
            if (!fin) {
                a_pos = p0 + curk + 1;
                best_len = cur_len;
                saved = cur_len >= lv.good ? quarter : lv.chain;  // deflate.zig:241-245
                st = kStart;
            } else {
                uint32_t out = 0, step = 1;
                if (cur_len) {
                    out = curk | ((cur_len - 3) << 8) | (cur_dist << 16);
                    step = curk + cur_len;
                }
                nx[p0] = out;
                a_pos = p0 + step;
                st = kArrive;
            }
        };

        bool exhausted = false, counted = false;
        while (true) {
            // ---- phase A: chain steps (deflate.zig:248 "Hot path loop!") ----
            uint32_t nl = left ? lds_shared_u16(sl_addr + 2 * (qi & (kSrRing - 1))) : 0;
#pragma unroll
            for (int u = 0; u < kSteps; u++) {
                if (left) {
                    qi -= nl;
                    const bool in_range = (int32_t)qi >= (int32_t)lim;
                    const uint32_t qs = (in_range ? qi : pi) & (kSrRing - 1);  // a slot that is always safe to read
                    const uint32_t b = lds_shared_u8(ro_addr + qs);
                    nl = lds_shared_u16(sl_addr + 2 * qs);
                    if (!in_range) {  // end of chain (kNoLink), too far, or at/below the slide base
                        st = kSearchDone;
                        left = 0;
                    } else if (b == cb) {  // may beat the best so far: needs the full compare
                        st = kPending;
                        saved = left;
                        left = 0;
                    } else {
                        left--;
                    }
                }
            }
            if (st == kStepping && left == 0) st = kSearchDone;
            // ---- phase B: full compares (SlidingWindow.match with the running best as min_len) ----
            if (__any_sync(0xffffffffu, st == kPending)) {
                if (st == kPending) {
                    st = kStepping;
                    left = saved - 1;  // this candidate is paid for either way
                    const uint32_t qs = qi & (kSrRing - 1), ps = pi & (kSrRing - 1);
                    if (lds_u32_unaligned(sb, qs) == first4) {
                        uint32_t i = 4;
                        while (i < max_len) {
                            const uint32_t x = lds_u32_unaligned(sb, qs + i) ^ lds_u32_unaligned(sb, ps + i);
                            if (x) {
                                i += (__ffs(x) - 1) >> 3;
                                break;
                            }
                            i += 4;
                        }
                        if (i > max_len) i = max_len;
                        if (i > best_len) {
                            best_len = i;
                            best_dist = pi - qi;
                            if (i >= lv.nice || i >= max_len) {  // deflate.zig:256-259, or nothing can be longer
                                st = kSearchDone;
                                left = 0;
                            } else {
                                ro_addr = sb_addr + i;
                                cb = sb[ps + i];
                            }
                        }
                    }
                    if (st == kStepping && left == 0) st = kSearchDone;
                }
            }
            // ---- phase C: lazy decisions of finished walks, next arrival of the orbit ----
            if (__any_sync(0xffffffffu, st == kSearchDone)) {
                if (st == kSearchDone) decide();
            }
            // ---- phase D: new seeds for idle lanes, ascending ----
            const uint32_t idle = __ballot_sync(0xffffffffu, st == kIdle);
            if (idle && !exhausted) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&S.seed_next, (uint32_t)__popc(idle));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= nseeds) {
                    exhausted = true;
                } else if (st == kIdle) {
                    const uint32_t t = base + __popc(idle & ltmask);
                    if (t < nseeds) {
                        // the segment's first position (where the true orbit starts) goes first: it lies in the run's first chunk
                        const uint32_t x = (has_begin && t == 0) ? begin : s0 + (t - has_begin) * kG;
                        if (x >= begin) {  // positions before the segment are history, not seeds
                            a_pos = x;
                            p0 = x;  // registered where its epoch counted it in
                            from_overlap = x >= s1;
                            st = kArrive;
                        }
                    }
                }
            }
            // ---- phase E: arrivals (claim, or stop at somebody's trail), phase F: walks begin ----
            if (__any_sync(0xffffffffu, st == kArrive)) {
                if (st == kArrive) arrive();
            }
            if (__any_sync(0xffffffffu, st == kStart)) {
                if (st == kStart) start_search();
            }
            if (warp == 0) sr_loader(S, bar, sb, sl, valid, in, link, kG, begin);
            const uint32_t working = __ballot_sync(0xffffffffu, st == kStepping || st == kPending || st == kSearchDone);
            if (exhausted && __ballot_sync(0xffffffffu, st != kIdle) == 0) {
                if (warp != 0) {
                    if (lane == 0) atomicAdd(&S.warps_done, 1u);
                    break;
                }
                if (V.warps_done == kWarps - 1) break;  // warp 0 serves the load front until everybody is done
                __nanosleep(200);
            } else if (working == 0) {
                __nanosleep(100);  // everybody here waits for the load front (or for seeds): leave the issue slots to the others
            } else if (prio) {
                const uint32_t near_tail = __ballot_sync(0xffffffffu, st != kIdle && (p0 >> 12) < V.tail + prio);
                if (near_tail == 0) __nanosleep(prio >> 8 ? prio >> 8 : 200);
            }
            (void)counted;
        }
        __syncthreads();
        if (threadIdx.x == 0 && S.inflight) {  // a look-ahead epoch nobody needed is still landing
            const uint32_t par = S.parity;
            while (!sr_bar_test(bar, par)) {}
            S.parity ^= 1u;
            S.inflight = 0;
        }
        {
            const uint32_t cnt = S.cross_cnt;
            if (threadIdx.x < min(cnt, kMaxCross)) {
                const uint32_t o = S.cross[threadIdx.x] - s1;  // the last arrival before the span end lies in the overlap
                if (o >= kW || !((S.safe[o >> 5] >> (o & 31)) & 1u)) S.failed = 1;
            }
            if (threadIdx.x == 0 && cnt > kMaxCross) S.failed = 1;
        }
        __syncthreads();
        for (uint32_t c = s0 / kSrChunk + threadIdx.x; c < s1 / kSrChunk; c += kThreads)
            chunk_fail[c] = (c + 1 == s1 / kSrChunk) ? S.failed : 0u;  // orbits of the run's own range may arrive anywhere in the next chunk
        if (threadIdx.x == 0 && S.failed) atomicOr(flags, 1u);
    }
}

// ------------------------------------------------------------------------------------------
// K3a: lazy step.  deflate.zig:160-193 restricted to arrivals with no pending match.
// From such an arrival at p the reference emits k literals p..p+k-1 (each displaced by a strictly
// longer match one byte later) and then one match at p+k, or a single literal if nothing matches.
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// K3b: chunk exit tables by pointer jumping.  For every possible entry offset e < 516 of a chunk
// of kChunk positions, the offset (into the next chunk) of the first arrival past the chunk end.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kLazyHalo = 256;  // a lazy run looks at most 255 positions ahead of its arrival
constexpr uint32_t kLazyThreads = 256;
__global__ void __launch_bounds__(kLazyThreads)
lazy_exit_kernel(const uint32_t* __restrict__ r_full, const uint32_t* __restrict__ r_quarter, uint32_t n, LevelArgs lv,
                 uint32_t* __restrict__ nx, uint16_t* __restrict__ exits) {
    // K3a + K3b fused: the chunk's match tables are staged once in shared memory, every thread
    // evaluates the lazy rule for 4 arrivals (deflate.zig:160-193), then the chunk's exit table is
    // built by pointer jumping.
    __shared__ uint32_t sf[kChunk + kLazyHalo];
    __shared__ uint32_t sq[kChunk + kLazyHalo];
    __shared__ uint16_t nxt[kChunk];
    const uint32_t c = blockIdx.x;
    const uint32_t cs = c * kChunk;
    for (uint32_t i = threadIdx.x; i < kChunk + kLazyHalo; i += blockDim.x) {
        const uint32_t p = cs + i;
        sf[i] = p < n ? r_full[p] : 0;
        sq[i] = p < n ? r_quarter[p] : 0;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kChunk; i += blockDim.x) {
        const uint32_t p = cs + i;
        uint32_t t = kChunk;  // positions past the end of data exit immediately
        if (p < n) {
            uint32_t r = sf[i];
            uint32_t out = 0;
            if (r != 0) {
                uint32_t cur = i;
                while (true) {
                    const uint32_t len = match_len_of(r);
                    if (len >= lv.lazy) break;  // deflate.zig:171
                    const uint32_t nb = cur + 1;  // stays inside the halo: at most 254 deferrals
                    const uint32_t r2 = (len >= lv.good) ? sq[nb] : sf[nb];  // deflate.zig:241-245
                    if (match_len_of(r2) > len) {  // better match one byte later: cur becomes a literal
                        cur = nb;
                        r = r2;
                    } else {
                        break;  // deflate.zig:182-184: emit the pending match
                    }
                }
                out = (cur - i) | ((match_len_of(r) - 3) << 8) | (match_dist_of(r) << 16);
            }
            nx[p] = out;
            t = i + nx_step(out);
        }
        nxt[i] = (uint16_t)t;
    }
    __syncthreads();
    while (true) {  // pointer jumping: a read phase and a write phase per round (kLazyThreads threads)
        bool pending = false;
        uint32_t nv[kChunk / kLazyThreads];
#pragma unroll
        for (uint32_t k = 0; k < kChunk / kLazyThreads; k++) {
            uint32_t t = nxt[threadIdx.x + k * kLazyThreads];
            if (t < kChunk) {
                t = nxt[t];
                pending = true;
            }
            nv[k] = t;
        }
        __syncthreads();
#pragma unroll
        for (uint32_t k = 0; k < kChunk / kLazyThreads; k++) nxt[threadIdx.x + k * kLazyThreads] = (uint16_t)nv[k];
        if (!__syncthreads_or(pending)) break;
    }
    for (uint32_t i = threadIdx.x; i < kEntries; i += blockDim.x) exits[(size_t)c * kEntries + i] = nxt[i] - kChunk;
}

// K3a alone, for a range of positions whose match tables start at `r_base` (used when a stream is
// sharded by position over several GPUs: every rank produces nx for its own range).
__global__ void lazy_step_range_kernel(const uint32_t* __restrict__ r_full, const uint32_t* __restrict__ r_quarter,
                                       uint32_t count, uint32_t avail, LevelArgs lv, uint32_t* __restrict__ nx) {
    // r_full / r_quarter / nx are indexed from the range start; `avail` (>= count) entries of the tables are valid
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t r = r_full[i];
    uint32_t out = 0;
    if (r != 0) {
        uint32_t cur = i;
        while (true) {
            const uint32_t len = match_len_of(r);
            if (len >= lv.lazy) break;  // deflate.zig:171
            const uint32_t nb = cur + 1;
            if (nb >= avail) break;     // only at the end of the stream (the caller provides 256 positions of slack)
            const uint32_t r2 = (len >= lv.good) ? r_quarter[nb] : r_full[nb];  // deflate.zig:241-245
            if (match_len_of(r2) > len) {
                cur = nb;
                r = r2;
            } else {
                break;
            }
        }
        out = (cur - i) | ((match_len_of(r) - 3) << 8) | (match_dist_of(r) << 16);
    }
    nx[i] = out;
}

// K3b alone: exit tables from nx.  jump[i] = first arrival at or past the end of i's 256-position sub-chunk.  Because
// every step goes forward, that is a right-to-left recurrence: jump[i] = i + step(i) if that leaves the piece, else
// jump[i + step(i)], which is already final.  One lane per 64-position piece runs it (64 lanes per chunk, 64 dependent
// steps), then two parallel passes widen the pieces to 128 and 256 positions (a position of the first half looks up
// where its exit lands in the second half, whose entries are final for the wider piece too): O(n) work where pointer
// jumping over the chunk does O(n log n), and no racing accesses.  The 516 possible entries then hop from sub-chunk to
// sub-chunk (at most 16 hops).  The jump table is kept in HBM (2 B per position): orbit_mark needs exactly this table
// and would otherwise rebuild it.
constexpr uint32_t kSub = 256;                    // sub-chunk walked by one lane of orbit_mark
constexpr uint32_t kSubs = kChunk / kSub;         // 16 walkers per chunk
constexpr uint32_t kPiece = 64, kPieces = kChunk / kPiece;
// the pieces are 66 entries (33 words) apart in shared memory, so that the lanes that walk them in step hit different banks
__device__ __forceinline__ uint32_t exit_slot(uint32_t i) { return i + 2 * (i >> 6); }
__global__ void __launch_bounds__(1024)
chunk_exit_kernel(const uint32_t* __restrict__ nx, uint32_t n, uint16_t* __restrict__ exits, uint16_t* __restrict__ jumps) {
    __shared__ __align__(16) uint16_t jump[kChunk + 2 * kPieces];
    const uint32_t c = blockIdx.x;
    const uint32_t cs = c * kChunk;
    if (cs + kChunk <= n && ((uintptr_t)(nx + cs) & 15) == 0) {  // whole chunk, aligned table: 16-byte loads (more bytes in flight)
        const uint4* nx4 = reinterpret_cast<const uint4*>(nx + cs);
        for (uint32_t q = threadIdx.x; q < kChunk / 4; q += blockDim.x) {
            const uint4 v = nx4[q];
            const uint32_t i = q * 4;  // four entries of one piece: consecutive slots
            uint16_t* dst = jump + exit_slot(i);
            dst[0] = (uint16_t)(i + nx_step(nx_clean(v.x)));
            dst[1] = (uint16_t)(i + 1 + nx_step(nx_clean(v.y)));
            dst[2] = (uint16_t)(i + 2 + nx_step(nx_clean(v.z)));
            dst[3] = (uint16_t)(i + 3 + nx_step(nx_clean(v.w)));
        }
    } else {
        for (uint32_t i = threadIdx.x; i < kChunk; i += blockDim.x) {
            const uint32_t p = cs + i;
            jump[exit_slot(i)] = (uint16_t)(p < n ? i + nx_step(nx_clean(nx[p])) : kChunk);  // <= 4095 + 515
        }
    }
    __syncthreads();
    for (uint32_t piece = threadIdx.x; piece < kPieces; piece += blockDim.x) {
        const uint32_t lo = piece * kPiece, end = lo + kPiece;
        for (uint32_t i = end; i-- > lo;) {
            const uint32_t t = jump[exit_slot(i)];
            if (t < end) jump[exit_slot(i)] = jump[exit_slot(t)];  // t > i: final already
        }
    }
    __syncthreads();
#pragma unroll
    for (uint32_t width = 2 * kPiece; width <= kSub; width *= 2) {
        // positions in the first half of every `width` block: an exit that lands in the second half goes on from there
        for (uint32_t k = threadIdx.x; k < kChunk / 2; k += blockDim.x) {
            const uint32_t i = (k / (width / 2)) * width + (k % (width / 2));
            const uint32_t end = (i | (width - 1)) + 1;
            const uint32_t t = jump[exit_slot(i)];
            if (t < end) jump[exit_slot(i)] = jump[exit_slot(t)];
        }
        __syncthreads();
    }
    for (uint32_t e = threadIdx.x; e < kEntries; e += blockDim.x) {
        uint32_t cur = e;
        while (cur < kChunk) cur = jump[exit_slot(cur)];
        exits[(size_t)c * kEntries + e] = (uint16_t)(cur - kChunk);
    }
    // back to the linear layout, two entries per word (a piece is 32 words, its slot 33)
    uint32_t* dst = reinterpret_cast<uint32_t*>(jumps + (size_t)c * kChunk);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(jump);
    for (uint32_t w = threadIdx.x; w < kChunk / 2; w += blockDim.x) dst[w] = src[w + (w >> 5)];
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
    // gentry[ngroups] = where the orbit leaves the last group: the first arrival past the chunk grid, relative to its end
    // (a streaming compressor continues the parse from there with the next part of the stream)
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t cur = 0;
    for (uint32_t g = 0; g < ngroups; g++) {
        gentry[g] = (uint16_t)cur;
        cur = gexits[(size_t)g * kEntries + cur];
    }
    gentry[ngroups] = (uint16_t)cur;
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
constexpr uint32_t kMarkThreads = 256;
__global__ void __launch_bounds__(kMarkThreads)
orbit_mark_kernel(const uint32_t* __restrict__ nx, uint32_t n, const uint16_t* __restrict__ entry,
                  const uint16_t* __restrict__ jumps, uint32_t* __restrict__ bitmap, uint32_t* __restrict__ chunk_tokens) {
    // jumps (may be null): the sub-chunk jump table chunk_exit_kernel left in HBM
    // The orbit inside a chunk is sequential, but once the first arrival in every 256-position
    // sub-chunk is known the 16 sub-chunks can be walked at the same time.  Those arrivals come from
    // pointer jumping restricted to sub-chunks (jump[i] = first arrival at or past the end of i's
    // sub-chunk), chained from the chunk's true entry.
    __shared__ __align__(16) uint32_t sn[kChunk];    // step | tokens emitted by an arrival here << 16
    __shared__ __align__(16) uint16_t jump[kChunk];
    __shared__ uint32_t bits[kChunk / 32];
    __shared__ uint32_t sub_entry[kSubs];
    __shared__ uint32_t sub_tokens[kSubs];
    const uint32_t c = blockIdx.x;
    const uint32_t cs = c * kChunk;
    if (jumps && cs + kChunk <= n && ((uintptr_t)(nx + cs) & 15) == 0) {  // whole chunk, aligned table: 16-byte loads
        const uint4* nx4 = reinterpret_cast<const uint4*>(nx + cs);
#pragma unroll 4
        for (uint32_t q = threadIdx.x; q < kChunk / 4; q += kMarkThreads) {
            const uint4 v4 = nx4[q];
            const uint32_t vv[4] = {nx_clean(v4.x), nx_clean(v4.y), nx_clean(v4.z), nx_clean(v4.w)};
            uint4 o;
            uint32_t* ov = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int k = 0; k < 4; k++) ov[k] = nx_step(vv[k]) | (((vv[k] >> 16) ? (vv[k] & 255u) + 1 : 1u) << 16);
            reinterpret_cast<uint4*>(sn)[q] = o;
        }
    } else
    for (uint32_t i = threadIdx.x; i < kChunk; i += kMarkThreads) {
        const uint32_t p = cs + i;
        uint32_t s = 1, t = 0;
        if (p < n) {
            const uint32_t v = nx_clean(nx[p]);
            s = nx_step(v);
            t = (v >> 16) ? (v & 255u) + 1 : 1;
        }
        sn[i] = s | (t << 16);
        if (!jumps) jump[i] = (uint16_t)(i + s);  // <= 4095 + 515
    }
    if (jumps) {
        const uint4* src = reinterpret_cast<const uint4*>(jumps + (size_t)c * kChunk);
        for (uint32_t i = threadIdx.x; i < kChunk / 8; i += kMarkThreads) reinterpret_cast<uint4*>(jump)[i] = src[i];
    }
    for (uint32_t i = threadIdx.x; i < kChunk / 32; i += kMarkThreads) bits[i] = 0;
    __syncthreads();
    while (!jumps) {  // one read phase and one write phase per round
        bool pending = false;
        uint32_t nv[kChunk / kMarkThreads];
#pragma unroll
        for (uint32_t k = 0; k < kChunk / kMarkThreads; k++) {
            const uint32_t i = threadIdx.x + k * kMarkThreads;
            uint32_t t = jump[i];
            if (t < (i | (kSub - 1)) + 1) {  // still inside i's sub-chunk
                t = jump[t];
                pending = true;
            }
            nv[k] = t;
        }
        __syncthreads();
#pragma unroll
        for (uint32_t k = 0; k < kChunk / kMarkThreads; k++) jump[threadIdx.x + k * kMarkThreads] = (uint16_t)nv[k];
        if (!__syncthreads_or(pending)) break;
    }
    if (threadIdx.x == 0) {
        uint32_t e = entry[c];
        for (uint32_t k = 0; k < kSubs; k++) {
            sub_entry[k] = e;
            if (e < (k + 1) * kSub) e = jump[e];  // else: a long match jumps over this sub-chunk entirely
        }
    }
    __syncthreads();
    if (threadIdx.x < kSubs) {
        const uint32_t k = threadIdx.x;
        const uint32_t lim = min(min((k + 1) * kSub, kChunk), n > cs ? n - cs : 0u);
        uint32_t i = sub_entry[k];
        uint32_t total = 0, word = 0xffffffffu, acc = 0;
        while (i < lim) {
            const uint32_t v = sn[i];
            if ((i >> 5) != word) {
                if (word != 0xffffffffu) bits[word] = acc;  // words belong to exactly one sub-chunk
                word = i >> 5;
                acc = 0;
            }
            acc |= 1u << (i & 31);
            total += v >> 16;
            i += v & 0xffffu;
        }
        if (word != 0xffffffffu) bits[word] = acc;
        sub_tokens[k] = total;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (uint32_t k = 0; k < kSubs; k++) total += sub_tokens[k];
        chunk_tokens[c] = total;
    }
    for (uint32_t i = threadIdx.x; i < kChunk / 32; i += kMarkThreads) bitmap[(size_t)c * (kChunk / 32) + i] = bits[i];
}

// exclusive scan of per-chunk token counts (single block; nchunks is at most ~1M)
__global__ void __launch_bounds__(1024)
scan_tokens_kernel(const uint32_t* __restrict__ counts, uint32_t nchunks, uint32_t* __restrict__ offsets,
                   uint32_t* __restrict__ total_out, uint32_t carry0) {
    // carry0: tokens already in the token buffer (the open block a streaming compressor carries from part to part)
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = carry0;
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
                   uint32_t* __restrict__ tokens, uint32_t* __restrict__ cut_rp, uint32_t* __restrict__ flags) {
    // flags (may be null): bit 1 is set when the orbit meets an entry the sparse parse never evaluated
    __shared__ uint32_t warp_sums[kEmitThreads / 32];
    const uint32_t c = blockIdx.x;
    const uint32_t cs = c * kChunk;
    constexpr uint32_t kPer = kChunk / kEmitThreads;  // 16 positions per thread
    const uint32_t i0 = threadIdx.x * kPer;
    const uint32_t word = bitmap[(size_t)c * (kChunk / 32) + (i0 >> 5)];
    const uint32_t mask = kPer >= 32 ? word : (word >> (i0 & 31)) & ((1u << (kPer & 31)) - 1);
    // tokens of my arrivals (scalar loads, twice: the arrivals are a quarter of the positions; fetching all 16 entries of a
    // thread as vectors reads the whole table (0.84 ms against 0.59 ms), keeping the first pass's values in registers for
    // the second costs more than the reloads from L1/L2 (0.64 ms))
    uint32_t mine = 0;
    for (uint32_t m = mask; m; m &= m - 1) {
        const uint32_t v = nx_clean(nx[cs + i0 + (__ffs(m) - 1)]);
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
        uint32_t v = nx[p];
        if (v == kNxInvalid) {
            if (flags) atomicOr(flags, 2u);
            v = 0;
        }
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
static int g_num_sms = 148;
static int g_search_steps = 8;
static SearchTune g_tune{3, 8};
static int g_use_roll = 0;
constexpr uint32_t kSparseT = 32768, kSparseW = 1024, kSparseThreads = 1024;
static int g_sparse_variant = 0;  // see lz77_sparse_range
static uint32_t g_link_run = 128;  // longest run of hash tiles per CTA (FB200_LINK_RUN): each run pays 4 warm-up tiles
static int g_exit_threads = 128;   // chunk_exit_kernel block size (FB200_EXIT_THREADS): small blocks hide the load latency better
static SparseTune g_sparse_tune{1, 1, 1};
static int g_sparse_roll = 0;      // FB200_SPARSE_ROLL: 0 = chunk kernel only (default: faster), 16 / 32 = rolling kernel with a seed every 16 / 32 positions
static uint32_t g_sparse_roll_k = 1;  // runs per SM (dynamic hand-out when > 1)
static uint32_t g_sparse_roll_prio = 0;  // see sparse_roll_kernel
static int g_sparse_long = 1;      // FB200_SPARSE_LONG=0: levels 8 and 9 use the 8-step kernel too
static void lz77_init_once() {
    // function attributes are per device: a process may hold contexts on several GPUs
    static bool done[64] = {};
    static bool knobs_read = false;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && done[dev]) return;
    if (dev >= 0 && dev < 64) done[dev] = true;
    cudaFuncSetAttribute(hash_link_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLinkSmem);
    cudaFuncSetAttribute(match_search_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSearchSmem);
    cudaFuncSetAttribute(match_search_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSearchSmem);
    cudaFuncSetAttribute(match_search_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSearchSmem);
    cudaFuncSetAttribute(match_search_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSearchSmem);
    cudaFuncSetAttribute(match_search_roll_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRollSmem);
    cudaFuncSetAttribute(match_search_roll_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRollSmem);
#define FB_SPARSE_ATTR(G, ST, D) cudaFuncSetAttribute(sparse_parse_kernel<kSparseT, kSparseW, G, kSparseThreads, 1, ST, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SparseCfg<kSparseT, kSparseW>::kSmem)
    FB_SPARSE_ATTR(32, 8, false);
    FB_SPARSE_ATTR(32, 4, false);
    FB_SPARSE_ATTR(16, 8, false);
    FB_SPARSE_ATTR(32, 8, true);
    FB_SPARSE_ATTR(32, 16, false);
#undef FB_SPARSE_ATTR
    cudaFuncSetAttribute(sparse_roll_kernel<kSparseW, 16, kSparseThreads, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSrSmem);
    cudaFuncSetAttribute(sparse_roll_kernel<kSparseW, 32, kSparseThreads, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSrSmem);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (knobs_read) return;
    knobs_read = true;
    // development knobs: FB200_TUNE="steps,pend_at,refill_at", FB200_SEARCH=roll (persistent rolling-window kernel)
    if (const char* e = getenv("FB200_TUNE")) {
        int a = 0, p = 0, r = 0;
        if (sscanf(e, "%d,%d,%d", &a, &p, &r) == 3) {
            g_search_steps = a;
            g_tune.pend_at = (uint32_t)p;
            g_tune.refill_at = (uint32_t)r;
        }
    }
    const char* e = getenv("FB200_SEARCH");
    g_use_roll = (e && e[0] == 'r') ? 1 : 0;
    // FB200_SPARSE="variant[,pend_at,done_at,refill_at]": shape and pacing of the sparse parse kernel
    if (const char* lr = getenv("FB200_LINK_RUN")) g_link_run = atoi(lr) >= 4 && atoi(lr) <= 4096 ? (uint32_t)atoi(lr) : 128;
    if (const char* et = getenv("FB200_EXIT_THREADS")) g_exit_threads = atoi(et) >= 128 && atoi(et) <= 1024 ? atoi(et) : 128;
    if (const char* sr = getenv("FB200_SPARSE_ROLL")) {  // "G[,runs per SM]"
        int g = 16, k = 1, pr = 0;
        const int got = sscanf(sr, "%d,%d,%d", &g, &k, &pr);
        if (got == 3 && pr >= 0) g_sparse_roll_prio = (uint32_t)pr;
        if (got >= 1) g_sparse_roll = (g == 0 || g == 16 || g == 32) ? g : 16;
        if (got >= 2 && k >= 1 && k <= 16) g_sparse_roll_k = (uint32_t)k;
    }
    if (const char* sl = getenv("FB200_SPARSE_LONG")) g_sparse_long = atoi(sl) != 0;
    if (const char* sp = getenv("FB200_SPARSE")) {
        int v = 0, p = 0, d = 0, r = 0;
        const int got = sscanf(sp, "%d,%d,%d,%d", &v, &p, &d, &r);
        if (got >= 1) g_sparse_variant = v;
        if (got == 4) g_sparse_tune = SparseTune{(uint32_t)p, (uint32_t)d, (uint32_t)r};
    }
}

// K1 + K2 for stream positions [from, range_end) of the segment that starts at seg_begin.  `from` and
// `range_end` must be multiples of 8192 unless they are seg_begin / n.  Lets the caller overlap the
// host-to-device copy of later parts of the input with the search of earlier parts.
cudaError_t lz77_search_range(const Lz77Buffers& b, const uint8_t* d_in, uint32_t seg_begin, uint32_t from, uint32_t range_end,
                              uint32_t n, const uint32_t* d_skip, uint32_t nskip, const LevelArgs& lv, cudaStream_t st,
                              PhaseTimer* pt, uint32_t link_from) {
    // link_from (<= from, default = from): first position whose link is (re)computed.  A caller that has not
    // linked the 32 KiB of history before `from` (a rank searching the middle of a sharded stream) passes
    // from - 32768.
    lz77_init_once();
    PhaseTimer dummy;
    if (!pt) pt = &dummy;
    if (range_end <= from) return cudaSuccess;
    if (link_from > from) link_from = from;
    const uint32_t ntiles = (range_end + kLinkTile - 1) / kLinkTile - link_from / kLinkTile;
    // run length: long enough to amortise the 4 warm-up tiles (<= g_link_run tiles), and chosen so that the grid is
    // close to a whole number of waves of 2 CTAs per SM
    const uint32_t slots = 2 * (uint32_t)g_num_sms;
    const uint32_t waves = (ntiles + slots * g_link_run - 1) / (slots * g_link_run);
    const uint32_t run = max(1u, (ntiles + slots * waves - 1) / (slots * waves));
    hash_link_kernel<<<(ntiles + run - 1) / run, kLinkThreads, kLinkSmem, st>>>(d_in, link_from, range_end, n, run, d_skip, nskip, b.link);
    pt->mark(st, kPhLink);
    if (g_use_roll && from == seg_begin && range_end == n) {
        const uint32_t epochs = (n + kEpoch - 1) / kEpoch - from / kEpoch;
        const uint32_t run_epochs = (epochs + g_num_sms - 1) / g_num_sms;
        const uint32_t rgrid = (epochs + run_epochs - 1) / run_epochs;
        if (g_search_steps == 4) match_search_roll_kernel<4><<<rgrid, kRollThreads, kRollSmem, st>>>(d_in, from, n, b.link, lv, g_tune, run_epochs, b.r_full, b.r_quarter);
        else match_search_roll_kernel<8><<<rgrid, kRollThreads, kRollSmem, st>>>(d_in, from, n, b.link, lv, g_tune, run_epochs, b.r_full, b.r_quarter);
    } else {
        const uint32_t grid = (range_end + kSearchTile - 1) / kSearchTile - from / kSearchTile;
#define FB_LAUNCH_SEARCH(K) match_search_kernel<K><<<grid, kSearchThreads, kSearchSmem, st>>>(d_in, seg_begin, from, range_end, n, b.link, lv, g_tune, b.r_full, b.r_quarter)
        if (g_search_steps == 2) FB_LAUNCH_SEARCH(2);
        else if (g_search_steps == 4) FB_LAUNCH_SEARCH(4);
        else if (g_search_steps == 16) FB_LAUNCH_SEARCH(16);
        else FB_LAUNCH_SEARCH(8);
#undef FB_LAUNCH_SEARCH
    }
    pt->mark(st, kPhSearch);
    return cudaGetLastError();
}

// K1 alone for positions [link_from, range_end).
cudaError_t lz77_link_range(const Lz77Buffers& b, const uint8_t* d_in, uint32_t link_from, uint32_t range_end, uint32_t n,
                            cudaStream_t st, PhaseTimer* pt, const uint32_t* d_skip, uint32_t nskip) {
    lz77_init_once();
    PhaseTimer dummy;
    if (!pt) pt = &dummy;
    if (range_end <= link_from) return cudaSuccess;
    const uint32_t ntiles = (range_end + kLinkTile - 1) / kLinkTile - link_from / kLinkTile;
    const uint32_t slots = 2 * (uint32_t)g_num_sms;
    const uint32_t waves = (ntiles + slots * g_link_run - 1) / (slots * g_link_run);
    const uint32_t run = max(1u, (ntiles + slots * waves - 1) / (slots * waves));
    hash_link_kernel<<<(ntiles + run - 1) / run, kLinkThreads, kLinkSmem, st>>>(d_in, link_from, range_end, n, run, d_skip, nskip, b.link);
    pt->mark(st, kPhLink);
    return cudaGetLastError();
}

// K2s for the sparse-parse chunks [first_chunk, end_chunk) of a stream of n positions (begin = 0, no skip list).
// b.link must be complete up to min(n, end_chunk * T + W + 256); b.nx must have been filled with kNxInvalid.
// chunk_fail[c] is set when orbits leaving chunk c were not seen to join the next chunk's seeds.
uint32_t lz77_sparse_chunk() { return kSparseT; }
uint32_t lz77_sparse_lookahead() { return kSparseW + kSpHalo + 272; }
uint32_t lz77_sparse_link_ahead() { return kSparseW + kSpHalo + kSrLook + kSrEpoch; }
cudaError_t lz77_sparse_range(const Lz77Buffers& b, const uint8_t* d_in, uint32_t first_chunk, uint32_t end_chunk, uint32_t n,
                              const LevelArgs& lv, uint32_t* chunk_fail, uint32_t* flags, cudaStream_t st, PhaseTimer* pt,
                              uint32_t begin, uint32_t* run_counter) {
    // begin > 0: b.nx is relative to the segment start (like every table after the match search)
    lz77_init_once();
    PhaseTimer dummy;
    if (!pt) pt = &dummy;
    if (end_chunk <= first_chunk) return cudaSuccess;
    // The interior of the stream goes to the rolling kernel (runs of whole chunks, one persistent CTA per SM); the
    // chunks whose window would reach past the end of the stream, unaligned buffers and very short ranges keep the
    // chunk kernel.  Both evaluate the same seeds with the same hand-over check, so they mix freely.
    if (g_sparse_roll && run_counter && (((uintptr_t)d_in | (uintptr_t)b.link) & 15) == 0) {
        const uint64_t need = (uint64_t)kSparseW + kSpHalo + kSrLook + kSrEpoch;  // keeps the last epoch load inside the stream
        const uint32_t roll_end = n > need ? min(end_chunk, (uint32_t)((n - need) / kSparseT)) : 0;
        if (roll_end >= first_chunk + 2) {
            const uint32_t nch = roll_end - first_chunk;
            uint32_t nruns = min(nch / 2, (uint32_t)g_num_sms * g_sparse_roll_k);
            if (nruns > (uint32_t)g_num_sms && nch / nruns < 8) nruns = min(nch / 2, (uint32_t)g_num_sms);  // short runs: one per SM
            cudaMemsetAsync(run_counter, 0, sizeof(uint32_t), st);
            const uint32_t rgrid = min(nruns, (uint32_t)g_num_sms);
            if (g_sparse_roll == 32) sparse_roll_kernel<kSparseW, 32, kSparseThreads, 8><<<rgrid, kSparseThreads, kSrSmem, st>>>(d_in, begin, first_chunk, roll_end, nruns, n, b.link, lv, b.nx - begin, chunk_fail, flags, run_counter, g_sparse_roll_prio);
            else sparse_roll_kernel<kSparseW, 16, kSparseThreads, 8><<<rgrid, kSparseThreads, kSrSmem, st>>>(d_in, begin, first_chunk, roll_end, nruns, n, b.link, lv, b.nx - begin, chunk_fail, flags, run_counter, g_sparse_roll_prio);
            first_chunk = roll_end;
            if (end_chunk <= first_chunk) {
                pt->mark(st, kPhSparse);
                return cudaGetLastError();
            }
        }
    }
    const uint32_t grid = end_chunk - first_chunk;
#define FB_SPARSE(G, ST) sparse_parse_kernel<kSparseT, kSparseW, G, kSparseThreads, 1, ST, false><<<grid, kSparseThreads, SparseCfg<kSparseT, kSparseW>::kSmem, st>>>(d_in, begin, first_chunk, nullptr, n, b.link, lv, g_sparse_tune, b.nx - begin, chunk_fail, flags)
    // the seeds of a run's overlap must be seeds of whatever evaluates the next chunk: same spacing as the rolling kernel
    // long chains (levels 8 and 9: 1024 / 4096 candidates): 16 chain steps per round, the phases around the walk matter less
    const int variant = g_sparse_variant == 0 && g_sparse_roll == 16 && run_counter ? 2
                        : g_sparse_variant == 0 && lv.chain >= 1024 && g_sparse_long ? 3 : g_sparse_variant;
    switch (variant) {
        case 1: FB_SPARSE(32, 4); break;
        case 2: FB_SPARSE(16, 8); break;
        case 3: FB_SPARSE(32, 16); break;
        default: FB_SPARSE(32, 8); break;
    }
#undef FB_SPARSE
    pt->mark(st, kPhSparse);
    return cudaGetLastError();
}
// Repair: evaluates every position of the listed chunks (device array of `count` chunk numbers).
cudaError_t lz77_sparse_dense_chunks(const Lz77Buffers& b, const uint8_t* d_in, const uint32_t* chunk_list, uint32_t count, uint32_t n,
                                     const LevelArgs& lv, uint32_t* chunk_fail, uint32_t* flags, cudaStream_t st, PhaseTimer* pt,
                                     uint32_t begin) {
    lz77_init_once();
    PhaseTimer dummy;
    if (!pt) pt = &dummy;
    if (count == 0) return cudaSuccess;
    sparse_parse_kernel<kSparseT, kSparseW, 32, kSparseThreads, 1, 8, true><<<count, kSparseThreads, SparseCfg<kSparseT, kSparseW>::kSmem, st>>>(
        d_in, begin, 0, chunk_list, n, b.link, lv, g_sparse_tune, b.nx - begin, chunk_fail, flags);
    pt->mark(st, kPhSparse);
    return cudaGetLastError();
}

// K3: lazy parse + token emission of the segment [begin, n) from the match tables.
cudaError_t lz77_parse(const Lz77Buffers& b, const uint8_t* d_in, uint32_t begin, uint32_t n, const LevelArgs& lv,
                       cudaStream_t st, PhaseTimer* pt) {
    PhaseTimer dummy;
    if (!pt) pt = &dummy;
    if (n == begin) {
        cudaMemsetAsync(b.total_tokens, 0, sizeof(uint32_t), st);
        return cudaGetLastError();
    }
    // everything after the match search works in segment-relative positions
    const uint8_t* d_seg = d_in + begin;
    n -= begin;
    const uint32_t nchunks = (n + kChunk - 1) / kChunk;
    const uint32_t ngroups = (nchunks + kGroup - 1) / kGroup;
    lazy_exit_kernel<<<nchunks, kLazyThreads, 0, st>>>(b.r_full, b.r_quarter, n, lv, b.nx, b.exits);
    pt->mark(st, kPhChunkExit);
    group_exit_kernel<<<ngroups, 544, 0, st>>>(b.exits, nchunks, b.gexits);
    group_entry_kernel<<<1, 32, 0, st>>>(b.gexits, ngroups, b.gentry);
    chunk_entry_kernel<<<(ngroups + 127) / 128, 128, 0, st>>>(b.exits, b.gentry, nchunks, b.entry);
    pt->mark(st, kPhResolve);
    orbit_mark_kernel<<<nchunks, kMarkThreads, 0, st>>>(b.nx, n, b.entry, nullptr, b.bitmap, b.chunk_tokens);
    pt->mark(st, kPhMark);
    scan_tokens_kernel<<<1, 1024, 0, st>>>(b.chunk_tokens, nchunks, b.tok_offset, b.total_tokens, 0);
    pt->mark(st, kPhScan);
    emit_tokens_kernel<<<nchunks, kEmitThreads, 0, st>>>(d_seg, b.nx, n, b.bitmap, b.tok_offset, lv, b.tokens, b.cut_rp, nullptr);
    pt->mark(st, kPhEmit);
    return cudaGetLastError();
}

// Position-sharded single stream.  Stage 1 on every rank: links, match search
// and lazy step for stream positions [from, to); nx_out[0 .. to-from) receives the packed lazy steps.
// `from` must be 0 or a multiple of 8192.  d_in must be readable from max(0, from - 32768) to
// min(n, to + 8192 + 272).
cudaError_t lz77_shard_search(const Lz77Buffers& b, const uint8_t* d_in, uint32_t from, uint32_t to, uint32_t n,
                              const LevelArgs& lv, uint32_t* nx_out, cudaStream_t st, PhaseTimer* pt) {
    PhaseTimer dummy;
    if (!pt) pt = &dummy;
    if (to <= from) return cudaSuccess;
    // the lazy rule looks up to 255 positions past its arrival: search one extra hash tile
    const uint32_t range_end = (uint32_t)min((uint64_t)n, ((uint64_t)to + 256 + kLinkTile - 1) / kLinkTile * kLinkTile);
    cudaError_t e = lz77_search_range(b, d_in, from, from, range_end, n, nullptr, 0, lv, st, pt, from >= kHist ? from - kHist : 0);
    if (e != cudaSuccess) return e;
    lazy_step_range_kernel<<<(to - from + 255) / 256, 256, 0, st>>>(b.r_full, b.r_quarter, to - from, range_end - from, lv, nx_out);
    pt->mark(st, kPhLazy);
    return cudaGetLastError();
}

// Stage 2 on one rank: parse + token emission from a complete nx table (b.nx).
cudaError_t lz77_parse_from_nx(const Lz77Buffers& b, const uint8_t* d_in, uint32_t n, const LevelArgs& lv, cudaStream_t st,
                               PhaseTimer* pt, uint32_t* flags, uint32_t tok_carry) {
    // tok_carry: tokens already at the front of b.tokens (they open the first block); the new ones follow them.
    // b.gentry[number of groups] receives the first arrival past the chunk grid (relative to its end).
    PhaseTimer dummy;
    if (!pt) pt = &dummy;
    if (n == 0) {
        cudaMemcpyAsync(b.total_tokens, &tok_carry, sizeof(uint32_t), cudaMemcpyHostToDevice, st);  // pageable source: staged at once
        cudaMemsetAsync(b.gentry, 0, sizeof(uint16_t), st);
        return cudaGetLastError();
    }
    const uint32_t nchunks = (n + kChunk - 1) / kChunk;
    const uint32_t ngroups = (nchunks + kGroup - 1) / kGroup;
    chunk_exit_kernel<<<nchunks, g_exit_threads, 0, st>>>(b.nx, n, b.exits, b.jumps);
    pt->mark(st, kPhChunkExit);
    group_exit_kernel<<<ngroups, 544, 0, st>>>(b.exits, nchunks, b.gexits);
    group_entry_kernel<<<1, 32, 0, st>>>(b.gexits, ngroups, b.gentry);
    chunk_entry_kernel<<<(ngroups + 127) / 128, 128, 0, st>>>(b.exits, b.gentry, nchunks, b.entry);
    pt->mark(st, kPhResolve);
    orbit_mark_kernel<<<nchunks, kMarkThreads, 0, st>>>(b.nx, n, b.entry, b.jumps, b.bitmap, b.chunk_tokens);
    pt->mark(st, kPhMark);
    scan_tokens_kernel<<<1, 1024, 0, st>>>(b.chunk_tokens, nchunks, b.tok_offset, b.total_tokens, tok_carry);
    pt->mark(st, kPhScan);
    emit_tokens_kernel<<<nchunks, kEmitThreads, 0, st>>>(d_in, b.nx, n, b.bitmap, b.tok_offset, lv, b.tokens, b.cut_rp, flags);
    pt->mark(st, kPhEmit);
    return cudaGetLastError();
}

cudaError_t lz77_tokenize(const Lz77Buffers& b, const uint8_t* d_in, uint32_t begin, uint32_t n, const uint32_t* d_skip,
                          uint32_t nskip, const LevelArgs& lv, cudaStream_t st, PhaseTimer* pt) {
    cudaError_t e = lz77_search_range(b, d_in, begin, begin, n, n, d_skip, nskip, lv, st, pt, begin);
    if (e != cudaSuccess) return e;
    return lz77_parse(b, d_in, begin, n, lv, st, pt);
}

}  // namespace fb
