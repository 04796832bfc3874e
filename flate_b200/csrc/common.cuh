// common.cuh -- shared device/host definitions for the B200 DEFLATE engine.
//
// Vocabulary follows the reference (ianic/flate): window, lookup chain, token, block, codegen.
// Citations "file:line" are to the reference repository (src/flate/...).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fb {

// ---- consts.zig:1-49 ----
constexpr uint32_t kTokensPerBlock = 1u << 15;  // consts.zig:6
constexpr uint32_t kMinMatch = 4;               // consts.zig:11
constexpr uint32_t kMaxMatch = 258;             // consts.zig:12
constexpr uint32_t kMaxDist = 32768;            // consts.zig:15
constexpr uint32_t kHist = 32768;
constexpr uint32_t kMinLookahead = kMinMatch + kMaxMatch;  // SlidingWindow.zig:13
constexpr uint32_t kNumLit = 286, kNumDist = 30, kNumCodegen = 19;
constexpr uint32_t kEndBlock = 256;
constexpr uint32_t kMaxStore = 65535;

// Token (ours; Token.zig:19-22 is a packed struct we do not mirror):
//   literal: byte value; match: 0x80000000 | (distance-1) << 8 | (length-3)
constexpr uint32_t kTokMatch = 0x80000000u;
__host__ __device__ inline uint32_t tok_match(uint32_t dist, uint32_t len) {
    return kTokMatch | ((dist - 1) << 8) | (len - 3);
}

// deflate.zig:35-53 LevelArgs
struct LevelArgs {
    uint32_t good, nice, lazy, chain;
};
__host__ __device__ inline bool level_args(int level, LevelArgs& a) {
    switch (level) {
        case 4: a = {4, 16, 4, 16}; return true;
        case 5: a = {8, 32, 16, 32}; return true;
        case 6: a = {8, 128, 16, 128}; return true;
        case 7: a = {8, 128, 32, 256}; return true;
        case 8: a = {32, 258, 128, 1024}; return true;
        case 9: a = {32, 258, 258, 4096}; return true;
    }
    return false;
}

// Slide schedule in closed form (deflate.zig:304-321, SlidingWindow.zig:36-44, Lookup.zig:43-51).
// The window slides by 32768 whenever 64 KiB are buffered, so the lookup base in force when
// position p is searched is a pure function of p and the stream length n.  Candidates q must
// satisfy base < q (window-relative position 0 means "none", deflate.zig:248).
__host__ __device__ inline uint32_t slide_base(uint64_t p, uint64_t n) {
    uint64_t j = (p + kMinLookahead) / kHist;
    j = j > 0 ? j - 1 : 0;
    uint64_t J = n / kHist;
    J = J > 0 ? J - 1 : 0;
    return (uint32_t)((j < J ? j : J) * kHist);
}

// Packed match-search result: len | (dist-1) << 9 ; 0 = none  (len in 4..258)
__host__ __device__ inline uint32_t pack_match(uint32_t len, uint32_t dist) { return len | ((dist - 1) << 9); }
__host__ __device__ inline uint32_t match_len_of(uint32_t r) { return r & 511u; }
__host__ __device__ inline uint32_t match_dist_of(uint32_t r) { return (r >> 9) + 1; }

// Packed lazy-parse step for a clean arrival at p (see deflate.cu lazy_step_kernel):
//   k (deferred literals before the match) | (len-3) << 8 | dist << 16 ; dist == 0 => plain literal
constexpr uint32_t kNxInvalid = 0xFFFFFFFFu;  // entry never evaluated by the sparse parse (distance 0xFFFF cannot occur)
__host__ __device__ inline uint32_t nx_clean(uint32_t nx) { return nx == kNxInvalid ? 0u : nx; }
__host__ __device__ inline uint32_t nx_step(uint32_t nx) { return (nx >> 16) ? (nx & 255u) + ((nx >> 8) & 255u) + 3u : 1u; }

// ---- Token.zig:58-103 code tables, in arithmetic form (RFC 1951 3.2.5) ----
// length-3 -> (code index 0..28, extra bits, extra value)
__host__ __device__ inline void length_code(uint32_t l3, uint32_t& code, uint32_t& eb, uint32_t& ev) {
    if (l3 < 8) {
        code = l3; eb = 0; ev = 0;
    } else if (l3 == 255) {
        code = 28; eb = 0; ev = 0;
    } else {
#ifdef __CUDA_ARCH__
        uint32_t msb = 31 - __clz(l3);
#else
        uint32_t msb = 31 - __builtin_clz(l3);
#endif
        eb = msb - 2;
        code = 4 * (msb - 1) + ((l3 >> eb) & 3);
        ev = l3 & ((1u << eb) - 1);
    }
}
// distance-1 -> (code 0..29, extra bits, extra value)
__host__ __device__ inline void distance_code(uint32_t d1, uint32_t& code, uint32_t& eb, uint32_t& ev) {
    if (d1 < 4) {
        code = d1; eb = 0; ev = 0;
    } else {
#ifdef __CUDA_ARCH__
        uint32_t msb = 31 - __clz(d1);
#else
        uint32_t msb = 31 - __builtin_clz(d1);
#endif
        eb = msb - 1;
        code = 2 * msb + ((d1 >> eb) & 1);
        ev = d1 & ((1u << eb) - 1);
    }
}
__host__ __device__ inline uint32_t length_extra_bits(uint32_t code) {  // code index 0..28
    return (code < 8 || code == 28) ? 0 : (code - 4) / 4;
}
__host__ __device__ inline uint32_t distance_extra_bits(uint32_t code) { return code < 4 ? 0 : (code - 2) / 2; }

// Block descriptor produced by the block-writer kernels.
enum BlockType : uint32_t { kStored = 0, kFixed = 1, kDynamic = 2 };
constexpr uint32_t kHdrWords = 160;  // dynamic header <= 17 + 19*3 + 316*(7+7) bits = 4498 bits

struct BlockDesc {
    uint32_t type;        // BlockType
    uint32_t eof;         // BFINAL
    uint32_t tok_begin;   // first token (level modes) / unused
    uint32_t tok_count;
    uint64_t in_begin;    // stored: source byte range; huffman-only: slice begin
    uint32_t in_len;
    uint32_t hdr_bits;    // bits in hdr[] (block header incl. BFINAL/BTYPE)
    uint64_t body_bits;   // Huffman body incl. EOB (huffman types) / 0
    uint64_t bit_offset;  // absolute start bit in the output stream (filled by the offset scan)
    uint32_t lit_code[kNumLit];   // code | len << 16 (codes bit-reversed: LSB-first ready)
    uint32_t dist_code[kNumDist];
    uint32_t hdr[kHdrWords];
};

#define FB_CUDA_CHECK(x)                                                        \
    do {                                                                        \
        cudaError_t e_ = (x);                                                   \
        if (e_ != cudaSuccess) { fb::set_last_cuda_error(e_, __FILE__, __LINE__); return FB200_ERR_CUDA; } \
    } while (0)

void set_last_cuda_error(cudaError_t e, const char* file, int line);

}  // namespace fb
