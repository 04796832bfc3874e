// block_writer.cu -- Huffman side of the deflate path on sm_100a: token/byte histograms, bit-exact
// length-limited code construction, block-type choice, header generation and the prefix-scan
// bit-pack of the coded stream.
//
// Replaces block_writer.zig (indexTokens :444, generateCodegen :78, dynamicSize :179, fixedSize
// :206, storedSizeFits :221, write :307, dynamicBlock :395, huffmanBlock :524, dynamicHeader :237,
// writeTokens :492, storedBlock :385), huffman_encoder.zig (generate :62, bitCounts :122,
// assignEncodingAndSize :251) and bit_writer.zig (only the byte stream is observable).
#include <cstdlib>

#include "common.cuh"
#include "pipeline.cuh"

namespace fb {

__constant__ uint8_t c_codegen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// ------------------------------------------------------------------------------------------
// block plans for the level modes: block b holds tokens [32768 b, 32768 (b+1)); the stored-input
// candidate is the reference's window[fp .. rp) at the moment of the cut.
// ------------------------------------------------------------------------------------------
__global__ void plan_level_blocks_kernel(const uint32_t* __restrict__ total_tokens, const uint32_t* __restrict__ cut_rp,
                                         uint32_t begin, uint32_t n, uint32_t max_blocks, uint32_t final_flush, BlockPlan* __restrict__ plans,
                                         uint32_t* __restrict__ nblocks_out, uint32_t fp0) {
    const uint32_t T = *total_tokens;
    // last one may be empty (deflate.zig:227-230,344); a part of a stream only writes the blocks that are complete
    const uint32_t ntok_blocks = final_flush == 2 ? T / kTokensPerBlock : T / kTokensPerBlock + 1;
    const uint32_t nblocks = ntok_blocks + (final_flush ? 0 : 1);
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *nblocks_out = nblocks < max_blocks ? nblocks : max_blocks;
    if (b >= nblocks || b >= max_blocks) return;
    BlockPlan pl;
    if (b < ntok_blocks) {
        pl.tok_begin = b * kTokensPerBlock;
        pl.tok_count = min(kTokensPerBlock, T - pl.tok_begin);
        // cut_rp is relative to the segment start; fp after a flush is the flush point (SlidingWindow.zig:113)
        const uint32_t rp = (b + 1 < ntok_blocks || final_flush == 2) ? begin + cut_rp[b] : n;
        const uint32_t fp = b == 0 ? fp0 : begin + cut_rp[b - 1];
        pl.in_begin = fp;
        pl.in_len = rp - fp;
        pl.has_input = fp >= slide_base(rp, n);  // fp < 0 after a slide => null (SlidingWindow.zig:121)
        pl.eof = (b + 1 == ntok_blocks) && final_flush == 1;
        pl.kind = kWrite;
    } else {  // sync marker: empty stored block (deflate.zig:276-278)
        pl.tok_begin = 0; pl.tok_count = 0; pl.in_begin = 0; pl.in_len = 0; pl.has_input = 1; pl.eof = 0; pl.kind = 3;
    }
    plans[b] = pl;
}

// ------------------------------------------------------------------------------------------
// K4: histograms.  block_writer.zig:444-463 (tokens), :575-585 (bytes).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kHistThreads = 512;
constexpr uint32_t kHistCopies = 8;  // sub-histograms to spread shared-memory atomic contention

__global__ void __launch_bounds__(kHistThreads)
histogram_tokens_kernel(const uint32_t* __restrict__ tokens, const BlockPlan* __restrict__ plans,
                        const uint32_t* __restrict__ nblocks_dev, uint32_t* __restrict__ lit_freq,
                        uint32_t* __restrict__ dist_freq) {
    __shared__ uint32_t h[kHistCopies][320];
    const uint32_t b = blockIdx.x;
    if (b >= *nblocks_dev) return;
    for (uint32_t i = threadIdx.x; i < kHistCopies * 320; i += kHistThreads) (&h[0][0])[i] = 0;
    __syncthreads();
    const BlockPlan pl = plans[b];
    uint32_t* mine = h[(threadIdx.x >> 5) & (kHistCopies - 1)];
    const uint32_t* tk = tokens + pl.tok_begin;
    for (uint32_t i = threadIdx.x; i < pl.tok_count; i += kHistThreads) {
        const uint32_t t = tk[i];
        if (t & kTokMatch) {
            uint32_t lc, eb, ev, dc;
            length_code(t & 255u, lc, eb, ev);
            distance_code((t >> 8) & 0x7fffu, dc, eb, ev);
            atomicAdd(&mine[257 + lc], 1u);
            atomicAdd(&mine[288 + dc], 1u);
        } else {
            atomicAdd(&mine[t & 255u], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 320; i += kHistThreads) {
        uint32_t s = 0;
        for (uint32_t k = 0; k < kHistCopies; k++) s += h[k][i];
        if (i < kNumLit) lit_freq[(size_t)b * kNumLit + i] = s;
        else if (i >= 288 && i < 288 + kNumDist) dist_freq[(size_t)b * kNumDist + (i - 288)] = s;
    }
}

__global__ void __launch_bounds__(kHistThreads)
histogram_bytes_kernel(const uint8_t* __restrict__ in, const BlockPlan* __restrict__ plans, uint32_t nblocks,
                       uint32_t* __restrict__ lit_freq) {
    __shared__ uint32_t h[kHistCopies][256];
    const uint32_t b = blockIdx.x;
    if (b >= nblocks) return;
    for (uint32_t i = threadIdx.x; i < kHistCopies * 256; i += kHistThreads) (&h[0][0])[i] = 0;
    __syncthreads();
    const BlockPlan pl = plans[b];
    uint32_t* mine = h[(threadIdx.x >> 5) & (kHistCopies - 1)];
    const uint8_t* src = in + pl.in_begin;
    const uint32_t len = pl.in_len;
    // head to 4-byte alignment, word body, tail
    const uint32_t mis = (uint32_t)((4 - ((uintptr_t)src & 3)) & 3);
    const uint32_t head = min(mis, len);
    if (threadIdx.x < head) atomicAdd(&mine[src[threadIdx.x]], 1u);
    const uint32_t nwords = (len - head) / 4;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(src + head);
    for (uint32_t i = threadIdx.x; i < nwords; i += kHistThreads) {
        const uint32_t v = w[i];
        atomicAdd(&mine[v & 255u], 1u);
        atomicAdd(&mine[(v >> 8) & 255u], 1u);
        atomicAdd(&mine[(v >> 16) & 255u], 1u);
        atomicAdd(&mine[v >> 24], 1u);
    }
    const uint32_t done = head + nwords * 4;
    if (threadIdx.x < len - done) atomicAdd(&mine[src[done + threadIdx.x]], 1u);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kNumLit; i += kHistThreads) {
        uint32_t s = 0;
        if (i < 256)
            for (uint32_t k = 0; k < kHistCopies; k++) s += h[k][i];
        lit_freq[(size_t)b * kNumLit + i] = s;
    }
}

// ------------------------------------------------------------------------------------------
// K5: code construction + block-type choice.  One warp per deflate block; lane 0 runs the
// sequential parts (the algorithm is a literal restatement of the Go-lineage bitCounts and must
// stay bit-exact), all lanes share the rank sort.
// ------------------------------------------------------------------------------------------
struct __align__(16) LevelInfo {
    uint32_t last_freq, next_char_freq, next_pair_freq, needed;
};
struct HuffScratch {
    uint16_t s_lit[288], s_freq[288];  // list sorted by (freq, literal)
    uint16_t t_lit[288], t_freq[288];  // compacted list in literal order
    LevelInfo levels[17];
    uint32_t leaf_counts[17][16];
    uint32_t bit_count[17];
    uint32_t next_code[17];
    uint32_t count;
};

__device__ __forceinline__ uint32_t bit_reverse(uint32_t v, uint32_t nbits) { return __brev(v) >> (32 - nbits); }

// huffman_encoder.zig:122-247.  list = s.s_freq[0..n) ascending; n >= 3.
// The lazy boundary package-merge as the reference runs it, step for step, by lane 0 alone (it is inherently
// sequential).  Only short lists come here (fewer than 16 symbols: small distance and code-length alphabets); longer
// ones take the eager form below, which gives the same length histogram with all lanes at work.
__device__ uint32_t bit_counts_warp(HuffScratch& s, uint32_t n, uint32_t max_bits) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t kMaxI32 = 0x7fffffffu;
    if (max_bits > n - 1) max_bits = n - 1;  // :131
    for (uint32_t i = lane; i < 17 * 16; i += 32) (&s.leaf_counts[0][0])[i] = 0;
    for (uint32_t l = lane; l < 17; l += 32) s.levels[l] = LevelInfo{0, 0, 0, 0};
    __syncwarp();
    if (lane == 0) {
        const uint32_t f0 = s.s_freq[0], f1 = s.s_freq[1], f2 = s.s_freq[2];
        for (uint32_t level = 1; level <= max_bits; level++) {  // :144-161
            s.levels[level] = LevelInfo{f1, f2, level == 1 ? kMaxI32 : f0 + f1, 0};
            s.leaf_counts[level][level] = 2;
        }
        s.levels[max_bits].needed = 2 * n - 4;  // :164
        uint32_t level = max_bits;
        while (true) {  // :168-224
            LevelInfo l = s.levels[level];
            if (l.next_pair_freq == kMaxI32 && l.next_char_freq == kMaxI32) {  // :170 (leaf sentinel is 65535: not taken)
                s.levels[level].needed = 0;
                s.levels[level + 1].next_pair_freq = kMaxI32;
                level += 1;
                continue;
            }
            const uint32_t prev_freq = l.last_freq;
            if (l.next_char_freq < l.next_pair_freq) {  // :182 next item is a leaf
                const uint32_t next = s.leaf_counts[level][level] + 1;
                l.last_freq = l.next_char_freq;
                s.leaf_counts[level][level] = next;
                l.next_char_freq = (next >= n) ? 65535u : (uint32_t)s.s_freq[next];  // :188-192, maxNode :282
            } else {  // :193 next item is a pair from the level below
                l.last_freq = l.next_pair_freq;
                for (uint32_t j = 0; j < level; j++) s.leaf_counts[level][j] = s.leaf_counts[level - 1][j];  // :199
                s.levels[level - 1].needed = 2;
            }
            l.needed -= 1;
            s.levels[level] = l;
            if (l.needed == 0) {  // :204
                if (level == max_bits) break;
                s.levels[level + 1].next_pair_freq = prev_freq + l.last_freq;
                level += 1;
            } else {
                while (s.levels[level - 1].needed > 0) {  // :217
                    level -= 1;
                    if (level == 0) break;
                }
            }
        }
        uint32_t bits = 1;
        for (uint32_t lv = max_bits; lv > 0; lv--) {  // :235-245
            s.bit_count[bits] = s.leaf_counts[max_bits][lv] - s.leaf_counts[max_bits][lv - 1];
            bits++;
        }
    }
    __syncwarp();
    return max_bits;
}

constexpr uint32_t kEagerItems = 576;  // 2 * 288
constexpr uint32_t kEagerSlots = 288;
struct EagerShared {
    uint32_t leaf[kEagerSlots];      // sorted leaf frequencies
    uint32_t pair[kEagerSlots];      // pair sums of the level below
    uint32_t item[2][kEagerItems];   // items of the level below / of this level
    uint32_t mask[16][kEagerItems / 32];  // level l: bit i set = item i is a leaf
    uint32_t len[16];
};
// Per-length counts of the length-limited code of n >= 3 sorted frequencies (S.leaf[0..n) filled by the caller, warp
// converged): out[1..L] with L = min(max_bits, n - 1), the same numbers bitCounts (huffman_encoder.zig:122-247) computes.
// See bit_counts_eager_kernel for the construction.  Returns L.
template <typename OutT>
__device__ uint32_t bit_counts_eager_warp(EagerShared& S, uint32_t n, uint32_t max_bits, OutT* out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t L = min(max_bits, n - 1);  // :131
    const uint32_t T = 2 * n - 2;             // items the top level takes; no level contributes more
    for (uint32_t i = lane; i < n; i += 32) S.item[0][i] = S.leaf[i];
    for (uint32_t i = lane; i < 16 * (kEagerItems / 32); i += 32) (&S.mask[0][0])[i] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < (n + 31) / 32; i += 32) S.mask[1][i] = (i + 1) * 32 <= n ? 0xffffffffu : (1u << (n & 31)) - 1;
    if (lane == 0) S.len[1] = n;
    __syncwarp();
    uint32_t cur = 0, len_prev = n;
    for (uint32_t l = 2; l <= L; l++) {
        const uint32_t* below = S.item[cur];
        uint32_t* here = S.item[cur ^ 1];
        const uint32_t m = len_prev / 2;
        for (uint32_t j = lane; j < m; j += 32) S.pair[j] = below[2 * j] + below[2 * j + 1];
        __syncwarp();
        // leaves: position = i + number of pairs <= leaf (a pair that ties goes first)
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t v = S.leaf[i];
            uint32_t lo = 0, hi = m;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (S.pair[mid] <= v) lo = mid + 1;
                else hi = mid;
            }
            const uint32_t pos = i + lo;
            if (pos < T) {
                here[pos] = v;
                atomicOr(&S.mask[l][pos >> 5], 1u << (pos & 31));
            }
        }
        // pairs: position = j + number of leaves < pair
        for (uint32_t j = lane; j < m; j += 32) {
            const uint32_t v = S.pair[j];
            uint32_t lo = 0, hi = n;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (S.leaf[mid] < v) lo = mid + 1;
                else hi = mid;
            }
            const uint32_t pos = j + lo;
            if (pos < T) here[pos] = v;
        }
        len_prev = min(n + m, T);
        if (lane == 0) S.len[l] = len_prev;
        cur ^= 1;
        __syncwarp();
    }
    // top-down: how many leaves each level contributes
    uint32_t take = T, a_above = 0;
    if (lane < 16) out[lane] = 0;
    __syncwarp();
    for (uint32_t l = L; l >= 1; l--) {
        take = min(take, S.len[l]);
        uint32_t a = 0;
        for (uint32_t wd = lane; wd * 32 < take; wd += 32) {
            uint32_t bits = S.mask[l][wd];
            if ((wd + 1) * 32 > take) bits &= (1u << (take & 31)) - 1;
            a += __popc(bits);
        }
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        // bit_count[L - l + 1] = a_l - a_(l-1): written when the level below is known
        if (l < L && lane == 0) out[L - l] = (OutT)(a_above - a);
        a_above = a;
        take = 2 * (take - a);
    }
    if (lane == 0) out[L] = (OutT)a_above;  // level 1: a_1 - a_0
    __syncwarp();
    return L;
}

// huffman_encoder.zig:62-95 generate + :251-278 assignEncodingAndSize.  Whole warp.
// out[i] = code | len << 16 (code bit-reversed, ready for LSB-first packing); len 0 for unused.
// Split form (huffman-only streams: tens of thousands of blocks with full alphabets): the serial bitCounts of 32 blocks
// runs on the 32 lanes of a warp in a kernel of its own (bit_counts_lanes_kernel), between a pass that sorts
// (kSortOnly: sorted list to global memory) and a pass that takes up again from the per-length counts (kFromCounts).
enum HuffPhase : int { kHuffFull = 0, kHuffSortOnly = 1, kHuffFromCounts = 2 };
struct HuffSplit {         // per-block slots in global memory
    uint16_t* slit;        // [288] symbols sorted by (freq, literal)
    uint16_t* sfreq;       // [288] their frequencies
    uint32_t* count;       // number of used symbols; kHuffDone: the block was finished by the sort pass (stored)
    uint16_t* bit_count;   // [16] codes per length
};
constexpr uint32_t kHuffDone = 0xffffffffu;

template <int kPhase>
__device__ void huff_generate_warp(HuffScratch& s, const uint16_t* freq, uint32_t n, uint32_t max_bits, uint32_t* out,
                                   const HuffSplit* split = nullptr, EagerShared* eager = nullptr) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t count = 0;
    for (uint32_t base = 0; base < n; base += 32) {  // compact non-zero symbols, literal order
        const uint32_t i = base + lane;
        const bool nz = i < n && freq[i] != 0;
        const uint32_t bal = __ballot_sync(0xffffffffu, nz);
        if (nz) {
            const uint32_t pos = count + __popc(bal & ((1u << lane) - 1));
            s.t_lit[pos] = (uint16_t)i;
            s.t_freq[pos] = freq[i];
        } else if (i < n) {
            out[i] = 0;
        }
        count += __popc(bal);
    }
    __syncwarp();
    if (kPhase == kHuffSortOnly && lane == 0) *split->count = count;
    if (count <= 2) {  // :79-87
        if (lane < count) out[s.t_lit[lane]] = lane | (1u << 16);
        __syncwarp();
        return;
    }
    uint32_t mb;
    if (kPhase != kHuffFromCounts) {
        // rank sort by (freq, literal): a total order, so any sort agrees with std.mem.sort (:89, :355-361)
        for (uint32_t i = lane; i < count; i += 32) {
            const uint32_t key = ((uint32_t)s.t_freq[i] << 16) | s.t_lit[i];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < count; j++) rank += ((((uint32_t)s.t_freq[j] << 16) | s.t_lit[j]) < key);
            s.s_lit[rank] = s.t_lit[i];
            s.s_freq[rank] = s.t_freq[i];
        }
        __syncwarp();
        if (kPhase == kHuffSortOnly) {
            for (uint32_t i = lane; i < count; i += 32) {
                split->slit[i] = s.s_lit[i];
                split->sfreq[i] = s.s_freq[i];
            }
            return;
        }
        if (eager && count >= 16) {  // large alphabets: the data-parallel form of the same construction
            for (uint32_t i = lane; i < count; i += 32) eager->leaf[i] = s.s_freq[i];
            __syncwarp();
            mb = bit_counts_eager_warp(*eager, count, max_bits, s.bit_count);
        } else {
            mb = bit_counts_warp(s, count, max_bits);
        }
    } else {
        for (uint32_t i = lane; i < count; i += 32) s.s_lit[i] = split->slit[i];
        if (lane < 16) s.bit_count[lane] = split->bit_count[lane];
        mb = max_bits > count - 1 ? count - 1 : max_bits;  // :131
        __syncwarp();
    }
    if (lane == 0) {
        // lengths: the last bit_count[1] symbols of the sorted list get 1 bit, the next bit_count[2] get 2, ...
        uint32_t remaining = count, code = 0;
        for (uint32_t nb = 1; nb <= mb; nb++) {
            code <<= 1;  // :256 (the nb = 0 iteration shifts zero)
            s.next_code[nb] = code;
            const uint32_t bits = s.bit_count[nb];
            for (uint32_t k = remaining - bits; k < remaining; k++) out[s.s_lit[k]] = nb << 16;
            remaining -= bits;
            code += bits;
        }
        // codes in literal order within each length (:267-275)
        for (uint32_t k = 0; k < count; k++) {
            const uint32_t sym = s.t_lit[k];
            const uint32_t nb = out[sym] >> 16;
            out[sym] = bit_reverse(s.next_code[nb]++, nb) | (nb << 16);
        }
    }
    __syncwarp();
}

// (the split passes leave the counts to bit_counts_eager_kernel: they carry no scratch for them)
template <bool kWithEager>
struct EagerHolder {
    EagerShared eager;
    __device__ EagerShared* eager_ptr() { return &eager; }
};
template <>
struct EagerHolder<false> {
    __device__ EagerShared* eager_ptr() { return nullptr; }
};
template <bool kWithEager>
struct BuildSharedT : EagerHolder<kWithEager> {
    HuffScratch hs;
    uint16_t lit_freq[kNumLit];
    uint16_t dist_freq[kNumDist];
    uint16_t codegen_freq[kNumCodegen];
    uint8_t codegen[kNumLit + kNumDist + 2];
    uint32_t lit_code[kNumLit];
    uint32_t dist_code[kNumDist];
    uint32_t codegen_code[kNumCodegen];
    uint32_t hdr[kHdrWords];
    uint32_t type, hdr_bits;
    uint64_t body_bits;
};

__device__ __forceinline__ uint32_t fixed_lit_code(uint32_t ch) {  // huffman_encoder.zig:298-330
    uint32_t bits, size;
    if (ch <= 143) { bits = ch + 48; size = 8; }
    else if (ch <= 255) { bits = ch + 400 - 144; size = 9; }
    else if (ch <= 279) { bits = ch - 256; size = 7; }
    else { bits = ch + 192 - 280; size = 8; }
    return bit_reverse(bits, size) | (size << 16);
}

// header bit writer into desc.hdr (lane 0 only)
struct HdrWriter {
    uint32_t* w;
    uint32_t nbits;
    __device__ void put(uint32_t v, uint32_t nb) {
        if (nb == 0) return;
        const uint32_t wi = nbits >> 5, sh = nbits & 31;
        w[wi] |= v << sh;
        if (sh + nb > 32) w[wi + 1] |= v >> (32 - sh);
        nbits += nb;
    }
};

constexpr uint32_t kBuildWarps = 2;  // deflate blocks per CTA

constexpr uint32_t kSplitSlots = 288;
template <int kPhase>
__global__ void __launch_bounds__(kBuildWarps * 32)
build_blocks_kernel(const BlockPlan* __restrict__ plans, const uint32_t* __restrict__ nblocks_dev,
                    const uint32_t* __restrict__ lit_freq_g, const uint32_t* __restrict__ dist_freq_g,
                    BlockDesc* __restrict__ descs, uint16_t* __restrict__ g_slit, uint16_t* __restrict__ g_sfreq,
                    uint32_t* __restrict__ g_count, uint16_t* __restrict__ g_bit_count) {
    __shared__ BuildSharedT<kPhase == kHuffFull> sh_all[kBuildWarps];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * kBuildWarps + (threadIdx.x >> 5);
    if (b >= *nblocks_dev) return;
    HuffSplit split{nullptr, nullptr, nullptr, nullptr};
    if (kPhase != kHuffFull) {
        split = HuffSplit{g_slit + (size_t)b * kSplitSlots, g_sfreq + (size_t)b * kSplitSlots, g_count + b, g_bit_count + (size_t)b * 16};
        if (kPhase == kHuffFromCounts && *split.count == kHuffDone) return;  // finished by the sort pass
    }
    auto& sh = sh_all[threadIdx.x >> 5];
    const BlockPlan pl = plans[b];
    BlockDesc& d = descs[b];

    if (lane == 0) {
        d.eof = pl.eof;
        d.tok_begin = pl.tok_begin;
        d.tok_count = pl.tok_count;
        d.in_begin = pl.in_begin;
        d.in_len = pl.in_len;
        d.bit_offset = 0;
    }
    if (pl.kind == 3) {  // verbatim stored block (store mode, sync marker): block_writer.zig:385-388
        if (lane == 0) { d.type = kStored; d.hdr_bits = 3; d.body_bits = 0; d.hdr[0] = pl.eof ? 1 : 0; }
        if (kPhase == kHuffSortOnly && lane == 0) *split.count = kHuffDone;
        return;
    }
    for (uint32_t i = lane; i < kHdrWords; i += 32) sh.hdr[i] = 0;

    // ---- frequencies (u16 in the reference; counts never exceed 65535) ----
    for (uint32_t i = lane; i < kNumLit; i += 32) sh.lit_freq[i] = (uint16_t)lit_freq_g[(size_t)b * kNumLit + i];
    for (uint32_t i = lane; i < kNumDist; i += 32)
        sh.dist_freq[i] = pl.kind == kHuffmanBlock ? 0 : (uint16_t)dist_freq_g[(size_t)b * kNumDist + i];
    __syncwarp();
    uint32_t num_literals, num_distances;
    // the reference counts one distance symbol that is never written when a block has no match
    // (:476-481, :531): it is part of the size estimates but not of the emitted body
    bool phantom_dist = pl.kind == kHuffmanBlock;
    if (pl.kind == kHuffmanBlock) {  // block_writer.zig:528-532
        if (lane == 0) { sh.lit_freq[kEndBlock] = 1; sh.dist_freq[0] = 1; }
        num_literals = kEndBlock + 1;
        num_distances = 1;
    } else {  // block_writer.zig:464-481
        if (lane == 0) sh.lit_freq[kEndBlock] += 1;
        __syncwarp();
        num_literals = kNumLit;
        while (sh.lit_freq[num_literals - 1] == 0) num_literals--;
        num_distances = kNumDist;
        while (num_distances > 0 && sh.dist_freq[num_distances - 1] == 0) num_distances--;
        __syncwarp();  // every lane has finished reading dist_freq before lane 0 writes the phantom symbol
        if (num_distances == 0) {
            if (lane == 0) sh.dist_freq[0] = 1;
            num_distances = 1;
            phantom_dist = true;
        }
    }
    __syncwarp();

    // ---- stored without building the code, when that is provable ----
    // dynamicBlock / huffmanBlock store a block iff stored_size < size + (size >> 4) (:418-426, :553-561), monotone in
    // size.  Every prefix code costs at least the entropy of the literal frequencies, so when the inequality already
    // holds for that lower bound the reference's choice is known without constructing the length-limited code (the
    // serial part of this kernel; incompressible slices of a huffman-only stream are the case that matters).
    if (pl.kind != kWrite && pl.has_input && pl.in_len <= kMaxStore) {
        double acc = 0.0;
        uint32_t tot = 0;
        for (uint32_t i = lane; i < kNumLit; i += 32) {
            const uint32_t f = sh.lit_freq[i];
            if (f) {
                acc += (double)f * log2((double)f);
                tot += f;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            tot += __shfl_xor_sync(0xffffffffu, tot, o);
        }
        const double h = (double)tot * log2((double)tot) - acc;  // bits; the margin below dwarfs any rounding error
        const uint64_t lower = h > 256.0 ? (uint64_t)h - 128 : 0;
        const uint64_t stored_size = ((uint64_t)pl.in_len + 5) * 8;
        if (stored_size < lower + (lower >> 4)) {  // uniform over the warp
            if (lane == 0) {
                d.type = kStored;
                d.hdr_bits = 3;  // :283-286
                d.body_bits = 0;
            }
            for (uint32_t i = lane; i < kHdrWords; i += 32) d.hdr[i] = i == 0 ? (pl.eof ? 1u : 0u) : 0u;
            if (kPhase == kHuffSortOnly && lane == 0) *split.count = kHuffDone;
            return;
        }
    }

    // ---- code construction ----
    huff_generate_warp<kPhase>(sh.hs, sh.lit_freq, kNumLit, 15, sh.lit_code, &split, sh.eager_ptr());
    if (kPhase == kHuffSortOnly) return;  // bit_counts_lanes_kernel and the kHuffFromCounts pass go on from here
    if (pl.kind == kHuffmanBlock) {  // huffmanDistanceEncoder, huffman_encoder.zig:340-348: one 1-bit code
        for (uint32_t i = lane; i < kNumDist; i += 32) sh.dist_code[i] = i == 0 ? (1u << 16) : 0;
    } else {
        huff_generate_warp<kHuffFull>(sh.hs, sh.dist_freq, kNumDist, 15, sh.dist_code, nullptr, sh.eager_ptr());
    }
    __syncwarp();

    // ---- generateCodegen, block_writer.zig:78-171 (lane 0) ----
    if (lane == 0) {
        for (uint32_t i = 0; i < kNumCodegen; i++) sh.codegen_freq[i] = 0;
        uint8_t* cg = sh.codegen;
        for (uint32_t i = 0; i < num_literals; i++) cg[i] = (uint8_t)(sh.lit_code[i] >> 16);
        for (uint32_t i = 0; i < num_distances; i++) cg[num_literals + i] = (uint8_t)(sh.dist_code[i] >> 16);
        cg[num_literals + num_distances] = 255;
        uint32_t size = cg[0];
        int count = 1;
        uint32_t out_index = 0;
        for (uint32_t in_index = 1; size != 255; in_index++) {
            const uint32_t next_size = cg[in_index];
            if (next_size == size) { count++; continue; }
            if (size != 0) {
                cg[out_index++] = (uint8_t)size;
                sh.codegen_freq[size]++;
                count--;
                while (count >= 3) {
                    const int nrep = count < 6 ? count : 6;
                    cg[out_index++] = 16;
                    cg[out_index++] = (uint8_t)(nrep - 3);
                    sh.codegen_freq[16]++;
                    count -= nrep;
                }
            } else {
                while (count >= 11) {
                    const int nrep = count < 138 ? count : 138;
                    cg[out_index++] = 18;
                    cg[out_index++] = (uint8_t)(nrep - 11);
                    sh.codegen_freq[18]++;
                    count -= nrep;
                }
                if (count >= 3) {
                    cg[out_index++] = 17;
                    cg[out_index++] = (uint8_t)(count - 3);
                    sh.codegen_freq[17]++;
                    count = 0;
                }
            }
            count--;
            for (; count >= 0; count--) {
                cg[out_index++] = (uint8_t)size;
                sh.codegen_freq[size]++;
            }
            size = next_size;
            count = 1;
        }
        cg[out_index] = 255;
    }
    __syncwarp();
    huff_generate_warp<kHuffFull>(sh.hs, sh.codegen_freq, kNumCodegen, 7, sh.codegen_code);
    __syncwarp();

    if (lane == 0) {
    // ---- sizes and the block-type choice (lane 0) ----
    uint32_t true_extra = 0;  // extra bits actually written with the tokens
    for (uint32_t lc = 8; lc < 29; lc++) true_extra += (uint32_t)sh.lit_freq[257 + lc] * length_extra_bits(lc);
    if (pl.kind != kHuffmanBlock)
        for (uint32_t dc = 4; dc < kNumDist; dc++) true_extra += (uint32_t)sh.dist_freq[dc] * distance_extra_bits(dc);
    const bool storable = pl.has_input && pl.in_len <= kMaxStore;  // block_writer.zig:221-229
    const uint32_t stored_size = storable ? (pl.in_len + 5) * 8 : 0;
    // the reference only adds the extra-bit cost when the block is storable (:317-334), and never in
    // dynamicBlock / huffmanBlock (:414, :549)
    const uint32_t extra_bits = (pl.kind == kWrite && storable) ? true_extra : 0;

    uint32_t num_codegens = kNumCodegen;  // :185-188
    while (num_codegens > 4 && sh.codegen_freq[c_codegen_order[num_codegens - 1]] == 0) num_codegens--;
    uint32_t cg_bits = 0;
    for (uint32_t i = 0; i < kNumCodegen; i++) cg_bits += (uint32_t)sh.codegen_freq[i] * (sh.codegen_code[i] >> 16);
    uint32_t lit_bits = 0, dist_bits = 0, fixed_bits = 0;
    for (uint32_t i = 0; i < kNumLit; i++) {
        lit_bits += (uint32_t)sh.lit_freq[i] * (sh.lit_code[i] >> 16);
        fixed_bits += (uint32_t)sh.lit_freq[i] * (fixed_lit_code(i) >> 16);
    }
    for (uint32_t i = 0; i < kNumDist; i++) {
        dist_bits += (uint32_t)sh.dist_freq[i] * (sh.dist_code[i] >> 16);
        fixed_bits += (uint32_t)sh.dist_freq[i] * 5;
    }
    const uint32_t dyn_header = 3 + 5 + 5 + 4 + 3 * num_codegens + cg_bits + (uint32_t)sh.codegen_freq[16] * 2 +
                                (uint32_t)sh.codegen_freq[17] * 3 + (uint32_t)sh.codegen_freq[18] * 7;
    const uint32_t dyn_size = dyn_header + lit_bits + dist_bits + extra_bits;  // :179-203
    const uint32_t fixed_size = 3 + fixed_bits + extra_bits;                     // :206-211

    uint32_t type;
    if (pl.kind == kWrite) {  // :336-372
        uint32_t size = fixed_size;
        type = kFixed;
        if (dyn_size < size) { size = dyn_size; type = kDynamic; }
        if (storable && stored_size < size) type = kStored;
    } else {  // :418-426, :553-561
        type = kDynamic;
        if (storable && stored_size < (dyn_size + (dyn_size >> 4))) type = kStored;
    }
    sh.type = type;
    HdrWriter hw{sh.hdr, 0};
    if (type == kStored) {
        hw.put(pl.eof ? 1 : 0, 3);  // :283-286
        sh.body_bits = 0;
    } else if (type == kFixed) {
        hw.put(pl.eof ? 3 : 2, 3);  // :293-300
        sh.body_bits = (uint64_t)fixed_bits + true_extra - (phantom_dist ? 5 : 0);
    } else {  // dynamicHeader, :237-281
        hw.put(pl.eof ? 5 : 4, 3);
        hw.put(num_literals - 257, 5);
        hw.put(num_distances - 1, 5);
        hw.put(num_codegens - 4, 4);
        for (uint32_t i = 0; i < num_codegens; i++) hw.put(sh.codegen_code[c_codegen_order[i]] >> 16, 3);
        for (uint32_t i = 0;;) {
            const uint32_t cw = sh.codegen[i++];
            if (cw == 255) break;
            hw.put(sh.codegen_code[cw] & 0xffffu, sh.codegen_code[cw] >> 16);
            if (cw == 16) hw.put(sh.codegen[i++], 2);
            else if (cw == 17) hw.put(sh.codegen[i++], 3);
            else if (cw == 18) hw.put(sh.codegen[i++], 7);
        }
        sh.body_bits = (uint64_t)lit_bits + dist_bits + true_extra - (phantom_dist ? (sh.dist_code[0] >> 16) : 0);
    }
    sh.hdr_bits = hw.nbits;
    }  // lane 0
    __syncwarp();
    const uint32_t type = sh.type;
    if (lane == 0) { d.type = type; d.hdr_bits = sh.hdr_bits; d.body_bits = sh.body_bits; }
    for (uint32_t i = lane; i < kHdrWords; i += 32) d.hdr[i] = sh.hdr[i];
    if (type == kFixed) {
        for (uint32_t i = lane; i < kNumLit; i += 32) d.lit_code[i] = fixed_lit_code(i);
        for (uint32_t i = lane; i < kNumDist; i += 32) d.dist_code[i] = bit_reverse(i, 5) | (5u << 16);
    } else if (type == kDynamic) {
        for (uint32_t i = lane; i < kNumLit; i += 32) d.lit_code[i] = sh.lit_code[i];
        for (uint32_t i = lane; i < kNumDist; i += 32) d.dist_code[i] = sh.dist_code[i];
    }
}

// ------------------------------------------------------------------------------------------
// K5 split, middle pass: huffman_encoder.zig:122-247 bitCounts, one block per lane.  The algorithm is a serial walk
// over about 2 n (levels) list items; a warp that runs it for ONE block spends 32 lanes on one instruction stream
// (bit_counts_warp), here the 32 lanes carry 32 blocks.  The per-block state (level records and the leaf-count
// triangle) lives in shared memory, element e of lane l at e * 32 + l, so the lanes never collide on a bank.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kLaneWarps = 7;
constexpr uint32_t kLaneList = 258;                      // sorted frequencies of a block (at most 257 symbols in a huffman-only block)
constexpr uint32_t kLaneTri = 135;                       // leaf counts (level, j <= level) of levels 1..15
constexpr uint32_t kLaneSmemPerWarp = (17 * 2 * 4 + 17 * 2 * 2 + kLaneTri * 2 + kLaneList * 2) * 32;  // per block: 990 bytes
__global__ void __launch_bounds__(kLaneWarps * 32, 1)
bit_counts_lanes_kernel(const uint32_t* __restrict__ nblocks_dev, const uint16_t* __restrict__ g_sfreq, const uint32_t* __restrict__ g_count,
                        uint16_t* __restrict__ g_bit_count) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t* lv32 = reinterpret_cast<uint32_t*>(smem_raw + (size_t)w * kLaneSmemPerWarp);  // [level][last | next_pair][lane]
    uint16_t* lv16 = reinterpret_cast<uint16_t*>(lv32 + 17 * 2 * 32);                       // [level][next_char | needed][lane]
    uint16_t* lc = lv16 + 17 * 2 * 32;                                                      // triangle [(level, j)][lane]
    uint16_t* sl = lc + kLaneTri * 32;                                                      // sorted list [i][lane]
#define LAST(level) lv32[((level) * 2 + 0) * 32 + lane]
#define PAIR(level) lv32[((level) * 2 + 1) * 32 + lane]
#define CHAR(level) lv16[((level) * 2 + 0) * 32 + lane]
#define NEED(level) lv16[((level) * 2 + 1) * 32 + lane]
#define LC(level, j) lc[((level) * ((level) + 1) / 2 - 1 + (j)) * 32 + lane]   /* level 1..15, j <= level */
    const uint32_t kMaxI32 = 0x7fffffffu;
    const uint32_t nb = *nblocks_dev;
    const uint32_t ntasks = (nb + 31) / 32;  // a task: 32 consecutive blocks, one per lane
    for (uint32_t task = blockIdx.x * kLaneWarps + w; task < ntasks; task += gridDim.x * kLaneWarps) {
        const uint32_t b = task * 32 + lane;
        uint32_t n = b < nb ? g_count[b] : 0;
        if (n == kHuffDone || n <= 2 || n >= kLaneList) n = 0;  // nothing to do for this lane (n >= 258 cannot happen here)
        if (n) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(g_sfreq + (size_t)b * kSplitSlots);
            for (uint32_t i = 0; i < (n + 1) / 2; i++) {
                const uint32_t v = src[i];
                sl[(2 * i) * 32 + lane] = (uint16_t)v;
                sl[(2 * i + 1) * 32 + lane] = (uint16_t)(v >> 16);
            }
            uint32_t max_bits = 15;
            if (max_bits > n - 1) max_bits = n - 1;  // :131
            for (uint32_t i = 0; i < kLaneTri; i++) lc[i * 32 + lane] = 0;
            for (uint32_t i = 0; i < 17 * 2; i++) { lv32[i * 32 + lane] = 0; lv16[i * 32 + lane] = 0; }
            const uint32_t f0 = sl[lane], f1 = sl[32 + lane], f2 = sl[64 + lane];
            for (uint32_t level = 1; level <= max_bits; level++) {  // :144-161
                LAST(level) = f1;
                CHAR(level) = (uint16_t)f2;
                PAIR(level) = level == 1 ? kMaxI32 : f0 + f1;
                NEED(level) = 0;
                LC(level, level) = 2;
            }
            NEED(max_bits) = (uint16_t)(2 * n - 4);  // :164
            uint32_t level = max_bits;
            while (true) {  // :168-224
                const uint32_t last = LAST(level), next_char = CHAR(level), next_pair = PAIR(level);
                uint32_t needed = NEED(level);
                // :170 `next_pair == maxInt and next_char == maxInt` cannot hold: the leaf sentinel is 65535 (maxNode, :282)
                uint32_t new_last;
                if (next_char < next_pair) {  // :182 next item is a leaf
                    const uint32_t next = (uint32_t)LC(level, level) + 1;
                    new_last = next_char;
                    LC(level, level) = (uint16_t)next;
                    CHAR(level) = next >= n ? (uint16_t)65535u : sl[next * 32 + lane];  // :188-192
                } else {  // :193 next item is a pair from the level below
                    new_last = next_pair;
                    if (level > 1)
                        for (uint32_t j = 0; j < level; j++) LC(level, j) = LC(level - 1, j);  // :199 (the row of level 0 is all zero)
                    else
                        LC(1, 0) = 0;
                    NEED(level - 1) = 2;
                }
                needed -= 1;
                LAST(level) = new_last;
                NEED(level) = (uint16_t)needed;
                if (needed == 0) {  // :204
                    if (level == max_bits) break;
                    PAIR(level + 1) = last + new_last;
                    level += 1;
                } else {
                    while (NEED(level - 1) > 0) {  // :217
                        level -= 1;
                        if (level == 0) break;
                    }
                }
            }
            uint16_t* out = g_bit_count + (size_t)b * 16;
            for (uint32_t i = 0; i < 16; i++) out[i] = 0;
            uint32_t bits = 1;
            for (uint32_t l = max_bits; l > 0; l--) {  // :235-245
                out[bits] = (uint16_t)(LC(max_bits, l) - LC(max_bits, l - 1));
                bits++;
            }
        }
        __syncwarp();
    }
#undef LAST
#undef PAIR
#undef CHAR
#undef NEED
#undef LC
}

// ------------------------------------------------------------------------------------------
// K5 split, middle pass, data-parallel form: the same per-length counts as bitCounts (huffman_encoder.zig:122-247) from
// the EAGER package-merge, one warp per block.  bitCounts is the lazy ("boundary") evaluation of this construction:
// level 1 is the sorted leaves; level l is the merge of the leaves with the pairs (sums of consecutive items) of level
// l - 1, a pair going first when it ties with a leaf (`next_char_freq < next_pair_freq` takes the leaf, :182); the top
// level takes 2n - 2 items, a level that contributes p pairs makes the level below contribute 2p items, and the number
// of leaves a_l among the items a level contributes gives the counts: bit_count[L - l + 1] = a_l - a_(l-1).  A merge is
// two binary searches per item (rank of a leaf among the pairs, of a pair among the leaves), so a level is 17 parallel
// steps of a warp instead of ~500 dependent ones.  Equality with the lazy form is a CPU test of its own
// (tests/test_eager_package_merge_cpu.py: frequency sets with ties, binding length limits, 15- and 7-bit limits, against the
// sequential restatement of bitCounts) and is checked on the GPU by the bit-exact huffman-only tests.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kEagerWarps = 8;
__global__ void __launch_bounds__(kEagerWarps * 32)
bit_counts_eager_kernel(const uint32_t* __restrict__ nblocks_dev, const uint16_t* __restrict__ g_sfreq, const uint32_t* __restrict__ g_count,
                        uint16_t* __restrict__ g_bit_count) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    EagerShared& S = reinterpret_cast<EagerShared*>(smem_raw)[w];
    const uint32_t b = blockIdx.x * kEagerWarps + w;
    if (b >= *nblocks_dev) return;
    const uint32_t n = g_count[b];
    if (n == kHuffDone || n <= 2 || n > kSplitSlots) return;
    const uint16_t* list = g_sfreq + (size_t)b * kSplitSlots;
    for (uint32_t i = lane; i < n; i += 32) S.leaf[i] = list[i];
    __syncwarp();
    bit_counts_eager_warp(S, n, 15, g_bit_count + (size_t)b * 16);
}

// ------------------------------------------------------------------------------------------
// K5b: block bit offsets.  Huffman blocks are not byte aligned; a stored block pads after its
// 3 header bits (block_writer.zig:283-291), so the offset recurrence is sequential.
// ------------------------------------------------------------------------------------------
// A run of blocks moves the write position x to  has ? align8(x + pre) + post : x + pre  (pre ends with the 3 header
// bits of the run's first stored block, post counts from the byte boundary that block aligns to).  Runs compose, so
// the recurrence is a scan: every thread summarises a contiguous range of blocks, one thread chains the 1024
// summaries, every thread then walks its range again from its start position.
struct RunBits {
    uint64_t pre, post;
    uint32_t has;
};
__device__ __forceinline__ uint64_t run_apply(uint64_t x, const RunBits& r) {
    return r.has ? ((x + r.pre + 7) & ~(uint64_t)7) + r.post : x + r.pre;
}
__device__ __forceinline__ void run_append(RunBits& a, uint32_t type, uint32_t in_len, uint64_t bits) {  // a := a then one block
    if (type == kStored) {
        if (a.has) a.post = ((a.post + 3 + 7) & ~(uint64_t)7) + 32 + 8ull * in_len;
        else { a.pre += 3; a.has = 1; a.post = 32 + 8ull * in_len; }
    } else if (a.has) {
        a.post += bits;
    } else {
        a.pre += bits;
    }
}
constexpr uint32_t kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads)
scan_block_offsets_kernel(BlockDesc* __restrict__ descs, const uint32_t* __restrict__ nblocks_dev,
                                          uint64_t start_bits, uint64_t* __restrict__ total_bits) {
    // total_bits[0] = end of the stream in bits, [1] = number of blocks, [2 + i] = start bit of block
    // pack_part_begin(nb, i + 1) (lets the host overlap the device-to-host copy of finished parts with the packing of
    // later ones), [16..18] = (pre, has, post) of the whole run (block-range sharded streams)
    __shared__ RunBits runs[kScanThreads];
    __shared__ uint64_t starts[kScanThreads];
    const uint32_t nb = *nblocks_dev;
    const uint32_t per = (nb + kScanThreads - 1) / kScanThreads;
    const uint32_t b0 = min(threadIdx.x * per, nb), b1 = min(b0 + per, nb);
    RunBits mine{0, 0, 0};
    for (uint32_t b = b0; b < b1; b++) run_append(mine, descs[b].type, descs[b].in_len, (uint64_t)descs[b].hdr_bits + descs[b].body_bits);
    runs[threadIdx.x] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t x = start_bits;
        RunBits all{0, 0, 0};
        for (uint32_t t = 0; t < kScanThreads; t++) {
            const RunBits r = runs[t];
            starts[t] = x;
            x = run_apply(x, r);
            if (!all.has) { all.pre += r.pre; all.has = r.has; all.post = r.post; }
            else if (r.has) all.post = ((all.post + r.pre + 7) & ~(uint64_t)7) + r.post;
            else all.post += r.pre;
        }
        total_bits[0] = x;
        total_bits[1] = nb;
        total_bits[16] = all.pre;
        total_bits[17] = all.has;
        total_bits[18] = all.has ? all.post : 0;
    }
    __syncthreads();
    uint32_t mark[kPackParts - 1];
#pragma unroll
    for (uint32_t i = 0; i + 1 < kPackParts; i++) mark[i] = pack_part_begin(nb, i + 1);
    uint64_t off = starts[threadIdx.x];
    for (uint32_t b = b0; b < b1; b++) {
        descs[b].bit_offset = off;
#pragma unroll
        for (uint32_t i = 0; i + 1 < kPackParts; i++)
            if (b == mark[i]) total_bits[2 + i] = off;
        if (descs[b].type == kStored) off = ((off + 3 + 7) & ~(uint64_t)7) + 32 + 8ull * descs[b].in_len;
        else off += (uint64_t)descs[b].hdr_bits + descs[b].body_bits;
    }
}

// ------------------------------------------------------------------------------------------
// K6: bit-pack.  One CTA per deflate block.  Each thread owns a contiguous run of tokens (or
// bytes): pass 1 sums code lengths, a block-wide exclusive scan turns them into bit offsets,
// pass 2 emits the codes LSB-first (bit_writer.zig:63-79 semantics) into 32-bit words.  Words a
// thread fully owns are stored; boundary words are merged with atomicOr (output is pre-zeroed).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kPackThreads = 1024;

struct BitSink {
    uint32_t* out;
    uint64_t word;  // next output word index
    uint64_t acc;
    uint32_t nacc;
    bool first;
    __device__ void init(uint32_t* o, uint64_t bitpos) {
        out = o; word = bitpos >> 5; acc = 0; nacc = (uint32_t)(bitpos & 31); first = true;
    }
    __device__ __forceinline__ void put(uint32_t v, uint32_t nb) {  // nb <= 32, nacc < 32 on entry
        acc |= (uint64_t)v << nacc;
        nacc += nb;
        if (nacc >= 32) {
            const uint32_t w = (uint32_t)acc;
            if (first) { atomicOr(out + word, w); first = false; }
            else out[word] = w;
            word++;
            acc >>= 32;
            nacc -= 32;
        }
    }
    __device__ void finish() {
        if (nacc > 0 && (uint32_t)acc != 0) atomicOr(out + word, (uint32_t)acc);
    }
};

__device__ __forceinline__ uint32_t token_bits(uint32_t t, const uint32_t* lit, const uint32_t* dist) {
    if (!(t & kTokMatch)) return lit[t & 255u] >> 16;
    uint32_t lc, leb, lev, dc, deb, dev;
    length_code(t & 255u, lc, leb, lev);
    distance_code((t >> 8) & 0x7fffu, dc, deb, dev);
    return (lit[257 + lc] >> 16) + leb + (dist[dc] >> 16) + deb;
}

template <bool kBytes>  // kBytes: huffman-only / store streams (no token list)
__global__ void __launch_bounds__(kPackThreads)
pack_blocks_kernel(const uint8_t* __restrict__ in, const uint32_t* __restrict__ tokens, const BlockDesc* __restrict__ descs,
                   const uint32_t* __restrict__ nblocks_dev, uint32_t first_block, uint32_t* __restrict__ out) {
    __shared__ uint32_t lit[kNumLit];
    __shared__ uint32_t dist[kNumDist];
    __shared__ uint32_t warp_sums[kPackThreads / 32];
    const uint32_t b = first_block + blockIdx.x;
    if (b >= *nblocks_dev) return;
    const BlockDesc& d = descs[b];
    const uint64_t o = d.bit_offset;
    const uint32_t type = d.type;

    if (type == kStored) {
        // 3 header bits, pad to byte, LEN, NLEN, raw bytes (block_writer.zig:283-291, 385-388)
        if (threadIdx.x == 0) atomicOr(out + (o >> 5), d.hdr[0] << (o & 31));  // 3 bits never straddle a byte
        const uint64_t a = ((o + 3 + 7) >> 3);  // first byte of LEN
        const uint32_t len = d.in_len;
        const uint32_t total = len + 4;
        const uint8_t* src = in + d.in_begin;
        const uint64_t w0 = a >> 2, w1 = (a + total + 3) >> 2;
        auto emit_word = [&](uint64_t w) {
            const int64_t s0 = (int64_t)(w * 4) - (int64_t)a;  // stream index of the word's first byte
            if (s0 >= 4 && s0 + 4 <= (int64_t)total) {
                // interior word: four payload bytes from an arbitrarily aligned source (two aligned loads + funnel shift)
                const uint8_t* p = src + (s0 - 4);
                const uint32_t* pw = reinterpret_cast<const uint32_t*>((uintptr_t)p & ~(uintptr_t)3);
                const uint32_t sh = (uint32_t)((uintptr_t)p & 3) * 8;
                const uint32_t lo = pw[0];
                const uint32_t hi = sh ? pw[1] : 0;  // when aligned the second word may lie past the buffer
                out[w] = __funnelshift_r(lo, hi, sh);
                return;
            }
            uint32_t v = 0;
            bool full = true;
            for (uint32_t j = 0; j < 4; j++) {
                const int64_t si = (int64_t)(w * 4 + j) - (int64_t)a;
                if (si < 0 || si >= (int64_t)total) { full = false; continue; }
                uint32_t byte;
                if (si == 0) byte = len & 255u;
                else if (si == 1) byte = (len >> 8) & 255u;
                else if (si == 2) byte = (~len) & 255u;
                else if (si == 3) byte = ((~len) >> 8) & 255u;
                else byte = src[si - 4];
                v |= byte << (8 * j);
            }
            if (full) out[w] = v;
            else if (v) atomicOr(out + w, v);
        };
        // the bulk in 16-byte stores (the output buffer is 16-byte aligned): five aligned source words give four output words
        const uint64_t g0 = (w0 + 3) >> 2, g1 = w1 >> 2;
        for (uint64_t g = g0 + threadIdx.x; g < g1; g += kPackThreads) {
            const int64_t s0 = (int64_t)(g * 16) - (int64_t)a;
            if (s0 >= 4 && s0 + 16 <= (int64_t)total) {
                const uint8_t* p = src + (s0 - 4);
                const uint32_t* pw = reinterpret_cast<const uint32_t*>((uintptr_t)p & ~(uintptr_t)3);
                const uint32_t sh = (uint32_t)((uintptr_t)p & 3) * 8;
                const uint32_t x0 = pw[0], x1 = pw[1], x2 = pw[2], x3 = pw[3], x4 = sh ? pw[4] : 0;
                reinterpret_cast<uint4*>(out)[g] = make_uint4(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh),
                                                              __funnelshift_r(x2, x3, sh), __funnelshift_r(x3, x4, sh));
            } else {
                for (uint32_t k = 0; k < 4; k++) emit_word(g * 4 + k);
            }
        }
        const uint64_t head_end = g0 * 4 < w1 ? g0 * 4 : w1, tail_begin = g1 > g0 ? g1 * 4 : head_end;
        for (uint64_t w = w0 + threadIdx.x; w < head_end; w += kPackThreads) emit_word(w);
        for (uint64_t w = tail_begin + threadIdx.x; w < w1; w += kPackThreads) emit_word(w);
        return;
    }

    for (uint32_t i = threadIdx.x; i < kNumLit; i += kPackThreads) lit[i] = d.lit_code[i];
    for (uint32_t i = threadIdx.x; i < kNumDist; i += kPackThreads) dist[i] = d.dist_code[i];
    // header words
    const uint32_t hdr_bits = d.hdr_bits;
    for (uint32_t w = threadIdx.x; w * 32 < hdr_bits; w += kPackThreads) {
        const uint32_t v = d.hdr[w];
        if (v) {
            const uint64_t pos = o + (uint64_t)w * 32;
            const uint32_t sh = (uint32_t)(pos & 31);
            atomicOr(out + (pos >> 5), v << sh);
            if (sh && (v >> (32 - sh))) atomicOr(out + (pos >> 5) + 1, v >> (32 - sh));
        }
    }
    __syncthreads();

    const uint32_t items = ((kBytes || tokens == nullptr) ? d.in_len : d.tok_count) + 1;  // + end-of-block
    const uint32_t per = (items + kPackThreads - 1) / kPackThreads;
    const uint32_t i0 = min(threadIdx.x * per, items), i1 = min(i0 + per, items);
    const uint32_t* tk = (!kBytes && tokens) ? tokens + d.tok_begin : nullptr;
    const uint8_t* src = in ? in + d.in_begin : nullptr;
    // huffman-only slices (no token list): a thread walks its run of bytes as aligned words (a byte-wise walk asks
    // the L1 for 32 different sectors with every load); the second pass finds them in the L1
    const bool word_bytes = kBytes && tk == nullptr;
    uint32_t nbytes = 0, mis = 0, nwords = 0;
    const uint32_t* pw = nullptr;
    if (word_bytes) {
        nbytes = min(i1, items - 1) - min(i0, items - 1);  // the last item is the end-of-block symbol, not a byte
        const uint8_t* p = src + i0;
        mis = (uint32_t)((uintptr_t)p & 3);
        pw = reinterpret_cast<const uint32_t*>(p - mis);
        nwords = nbytes ? (nbytes + mis + 3) / 4 : 0;
    }
    uint32_t mybits = 0;
    if (word_bytes) {
        for (uint32_t w = 0; w < nwords; w++) {
            const uint32_t v = pw[w];
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t bi = w * 4 + k - mis;  // wraps for the bytes before the run
                if (bi < nbytes) mybits += lit[(v >> (8 * k)) & 255u] >> 16;
            }
        }
        if (i1 == items && i0 < i1) mybits += lit[kEndBlock] >> 16;
    } else {
        for (uint32_t i = i0; i < i1; i++) {
            if (i == items - 1) mybits += lit[kEndBlock] >> 16;
            else if (tk) mybits += token_bits(tk[i], lit, dist);
            else mybits += lit[src[i]] >> 16;
        }
    }
    // block-wide exclusive scan of mybits
    uint32_t x = mybits;
    for (int s = 1; s < 32; s <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, s);
        if ((threadIdx.x & 31) >= s) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = warp_sums[threadIdx.x];
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, w, s);
            if (threadIdx.x >= s) w += y;
        }
        warp_sums[threadIdx.x] = w;
    }
    __syncthreads();
    const uint32_t excl = ((threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + x - mybits;
    if (i0 >= i1) return;
    BitSink bs;
    bs.init(out, o + hdr_bits + excl);
    if (word_bytes) {
        for (uint32_t w = 0; w < nwords; w++) {
            const uint32_t v = pw[w];
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t bi = w * 4 + k - mis;
                if (bi < nbytes) {
                    const uint32_t c = lit[(v >> (8 * k)) & 255u];
                    bs.put(c & 0xffffu, c >> 16);
                }
            }
        }
        if (i1 == items) bs.put(lit[kEndBlock] & 0xffffu, lit[kEndBlock] >> 16);
        bs.finish();
        return;
    }
    for (uint32_t i = i0; i < i1; i++) {
        if (i == items - 1) {
            bs.put(lit[kEndBlock] & 0xffffu, lit[kEndBlock] >> 16);
        } else if (tk) {
            const uint32_t t = tk[i];
            if (!(t & kTokMatch)) {
                const uint32_t c = lit[t & 255u];
                bs.put(c & 0xffffu, c >> 16);
            } else {  // block_writer.zig:504-516
                uint32_t lc, leb, lev, dc, deb, dev;
                length_code(t & 255u, lc, leb, lev);
                distance_code((t >> 8) & 0x7fffu, dc, deb, dev);
                const uint32_t c = lit[257 + lc];
                bs.put((c & 0xffffu) | (lev << (c >> 16)), (c >> 16) + leb);
                const uint32_t e = dist[dc];
                bs.put((e & 0xffffu) | (dev << (e >> 16)), (e >> 16) + deb);
            }
        } else {
            const uint32_t c = lit[src[i]];
            bs.put(c & 0xffffu, c >> 16);
        }
    }
    bs.finish();
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
cudaError_t plan_level_blocks(const uint32_t* total_tokens, const uint32_t* cut_rp, uint32_t begin, uint32_t n,
                              uint32_t max_blocks, uint32_t final_flush, BlockPlan* plans, uint32_t* nblocks, cudaStream_t st,
                              uint32_t fp0) {
    plan_level_blocks_kernel<<<(max_blocks + 127) / 128, 128, 0, st>>>(total_tokens, cut_rp, begin, n, max_blocks, final_flush,
                                                                       plans, nblocks, fp0);
    return cudaGetLastError();
}
cudaError_t histogram_tokens(const uint32_t* tokens, const BlockPlan* plans, const uint32_t* nblocks_dev,
                             uint32_t max_blocks, uint32_t* lit_freq, uint32_t* dist_freq, cudaStream_t st) {
    histogram_tokens_kernel<<<max_blocks, kHistThreads, 0, st>>>(tokens, plans, nblocks_dev, lit_freq, dist_freq);
    return cudaGetLastError();
}
cudaError_t histogram_bytes(const uint8_t* in, const BlockPlan* plans, uint32_t nblocks, uint32_t* lit_freq,
                            cudaStream_t st) {
    histogram_bytes_kernel<<<nblocks, kHistThreads, 0, st>>>(in, plans, nblocks, lit_freq);
    return cudaGetLastError();
}
size_t build_blocks_split_bytes(uint32_t max_blocks) { return (size_t)max_blocks * (kSplitSlots * 2 * 2 + 4 + 16 * 2) + 64; }
cudaError_t build_blocks(const BlockPlan* plans, const uint32_t* nblocks_dev, uint32_t max_blocks,
                         const uint32_t* lit_freq, const uint32_t* dist_freq, BlockDesc* descs, cudaStream_t st, void* split_scratch) {
    const uint32_t grid = (max_blocks + kBuildWarps - 1) / kBuildWarps;
    if (!split_scratch) {
        build_blocks_kernel<kHuffFull><<<grid, kBuildWarps * 32, 0, st>>>(plans, nblocks_dev, lit_freq, dist_freq, descs, nullptr, nullptr,
                                                                        nullptr, nullptr);
        return cudaGetLastError();
    }
    // split form: sort -> bitCounts of 32 blocks per warp -> the rest (split_scratch holds build_blocks_split_bytes(max_blocks))
    uint8_t* base = reinterpret_cast<uint8_t*>(split_scratch);
    uint16_t* slit = reinterpret_cast<uint16_t*>(base);
    uint16_t* sfreq = slit + (size_t)max_blocks * kSplitSlots;
    uint16_t* bitc = sfreq + (size_t)max_blocks * kSplitSlots;
    uint32_t* count = reinterpret_cast<uint32_t*>(bitc + (size_t)max_blocks * 16);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaFuncSetAttribute(bit_counts_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kLaneWarps * kLaneSmemPerWarp));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    build_blocks_kernel<kHuffSortOnly><<<grid, kBuildWarps * 32, 0, st>>>(plans, nblocks_dev, lit_freq, dist_freq, descs, slit, sfreq, count, bitc);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t ntasks = (max_blocks + 31) / 32;  // persistent: one CTA of kLaneWarps warps per SM, tasks dealt round robin
    const uint32_t lgrid = min((uint32_t)sms, (ntasks + kLaneWarps - 1) / kLaneWarps);
    static const bool lazy_lanes = [] { const char* e = getenv("FB200_BITCOUNTS"); return e && e[0] == 'l'; }();  // FB200_BITCOUNTS=lanes: the lazy form
    if (lazy_lanes) {
        bit_counts_lanes_kernel<<<lgrid, kLaneWarps * 32, kLaneWarps * kLaneSmemPerWarp, st>>>(nblocks_dev, sfreq, count, bitc);
    } else {
        static bool eager_attr[64] = {};
        if (dev < 0 || dev >= 64 || !eager_attr[dev]) {
            cudaFuncSetAttribute(bit_counts_eager_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kEagerWarps * sizeof(EagerShared)));
            if (dev >= 0 && dev < 64) eager_attr[dev] = true;
        }
        bit_counts_eager_kernel<<<(max_blocks + kEagerWarps - 1) / kEagerWarps, kEagerWarps * 32, kEagerWarps * sizeof(EagerShared), st>>>(
            nblocks_dev, sfreq, count, bitc);
    }
    build_blocks_kernel<kHuffFromCounts><<<grid, kBuildWarps * 32, 0, st>>>(plans, nblocks_dev, lit_freq, dist_freq, descs, slit, sfreq, count, bitc);
    return cudaGetLastError();
}
cudaError_t scan_block_offsets(BlockDesc* descs, const uint32_t* nblocks_dev, uint64_t start_bits, uint64_t* total_bits,
                               cudaStream_t st) {
    scan_block_offsets_kernel<<<1, kScanThreads, 0, st>>>(descs, nblocks_dev, start_bits, total_bits);
    return cudaGetLastError();
}
cudaError_t pack_blocks(const uint8_t* in, const uint32_t* tokens, const BlockDesc* descs, const uint32_t* nblocks_dev,
                        uint32_t max_blocks, uint32_t* out_words, cudaStream_t st) {
    if (tokens) pack_blocks_kernel<false><<<max_blocks, kPackThreads, 0, st>>>(in, tokens, descs, nblocks_dev, 0, out_words);
    else pack_blocks_kernel<true><<<max_blocks, kPackThreads, 0, st>>>(in, tokens, descs, nblocks_dev, 0, out_words);
    return cudaGetLastError();
}
cudaError_t pack_blocks_range(const uint8_t* in, const uint32_t* tokens, const BlockDesc* descs, const uint32_t* nblocks_dev,
                              uint32_t first_block, uint32_t count, uint32_t* out_words, cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    if (tokens) pack_blocks_kernel<false><<<count, kPackThreads, 0, st>>>(in, tokens, descs, nblocks_dev, first_block, out_words);
    else pack_blocks_kernel<true><<<count, kPackThreads, 0, st>>>(in, tokens, descs, nblocks_dev, first_block, out_words);
    return cudaGetLastError();
}

}  // namespace fb
