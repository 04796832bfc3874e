// inflate_dev.cuh -- device helpers shared by the inflate kernels (inflate.cu: one warp per member;
// inflate_par.cu: one CTA per member with lane-parallel symbol decode): the exact bit cursor of
// bit_reader.zig, canonical decoder construction with the validation order of huffman_decoder.zig:126-153,
// and the CRC-32 / Adler-32 pieces of container.zig:168-206.
#pragma once
#include "../../include/flate_b200.h"
#include "inflate.cuh"
#include "inflate_span.cuh"

namespace fb {

static __constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
static __constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                        67, 83, 99, 115, 131, 163, 195, 227, 258};
static __constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                         1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};

constexpr uint32_t kLitFast = 10, kDistFast = 8;
constexpr uint32_t kInflateWarps = 4;

// direct-table entry: code length (4 bits, 0 = use the canonical fallback) | symbol << 4 (9 bits) |
// extra-bit count << 13 (4 bits) | base value << 17 (length or distance base; 15 bits)
constexpr uint32_t kQueue = 64;        // match descriptors per batch
constexpr uint32_t kBatchSpan = 1024;  // output bytes lane 0 may run ahead of the warp's copies
struct WarpTables {
    uint32_t lit_fast[1 << kLitFast];
    uint32_t dist_fast[1 << kDistFast];
    uint32_t queue[2 * kQueue];          // (position, length << 16 | distance - 1)
    uint16_t lit_count[16], dist_count[16];
    uint16_t lit_sym[kNumLit + 2], dist_sym[kNumDist + 2];
    uint8_t lit_lens[kNumLit + 2], dist_lens[kNumDist + 2];
    uint32_t crc_tab[256];
};

__device__ __forceinline__ uint32_t bfe32(uint32_t v, uint32_t pos, uint32_t len) {  // bit-field extract, len may be 0
    uint32_t r;
    asm("bfe.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(pos), "r"(len));
    return r;
}

struct BitCursor {  // lane 0 only
    const uint8_t* next;
    const uint8_t* end;
    uint64_t buf;
    uint32_t cnt;
    __device__ __forceinline__ void refill() {
        while (cnt <= 32 && next < end) {
            if ((((uintptr_t)next) & 3) == 0 && next + 4 <= end) {
                buf |= (uint64_t)(*reinterpret_cast<const uint32_t*>(next)) << cnt;
                next += 4;
                cnt += 32;
            } else {
                buf |= (uint64_t)(*next) << cnt;
                next += 1;
                cnt += 8;
            }
        }
    }
    __device__ __forceinline__ bool empty() { refill(); return cnt == 0; }             // fill(): EndOfStream
    __device__ __forceinline__ uint32_t peek(uint32_t nb) { return (uint32_t)buf & ((1u << nb) - 1); }
    __device__ __forceinline__ bool shift(uint32_t nb) {                                 // false => EndOfStream
        if (nb > cnt) { refill(); if (nb > cnt) return false; }
        buf >>= nb;
        cnt -= nb;
        return true;
    }
    // read(U): fill + shift
    __device__ __forceinline__ int read(uint32_t nb, uint32_t& v) {
        refill();
        if (cnt == 0) return FB200_END_OF_STREAM;
        v = nb >= 32 ? (uint32_t)buf : ((uint32_t)buf & ((1u << nb) - 1));
        if (nb > cnt) return FB200_END_OF_STREAM;
        buf >>= nb;
        cnt -= nb;
        return FB200_OK;
    }
    __device__ __forceinline__ void align_to_byte() {
        const uint32_t r = cnt & 7;  // whole bytes are buffered, so cnt mod 8 is the stream's bit phase
        buf >>= r;
        cnt -= r;
    }
    __device__ __forceinline__ const uint8_t* byte_pos() const { return next - (cnt >> 3); }  // when aligned
    // the cursor as an absolute bit address (8 * byte address + bit), and back
    __device__ __forceinline__ unsigned long long bit_address() const { return (unsigned long long)(uintptr_t)next * 8ull - cnt; }
    __device__ __forceinline__ void seek(unsigned long long bit_addr) {
        next = reinterpret_cast<const uint8_t*>((uintptr_t)(bit_addr >> 3));
        buf = 0;
        cnt = 0;
        refill();
        const uint32_t r = (uint32_t)(bit_addr & 7);  // r != 0 means the byte exists, so cnt >= 8 here
        buf >>= r;
        cnt -= r;
    }
};

// canonical decode on a zero-padded LSB-first peek (huffman_decoder.zig:156-175 find semantics)
__device__ __forceinline__ int slow_find(const uint16_t* count, const uint16_t* symbol, uint32_t max_bits, uint32_t peek,
                                         uint32_t& sym, uint32_t& nbits) {
    int code = 0, first = 0, index = 0;
    for (uint32_t len = 1; len <= max_bits; len++) {
        code |= (int)(peek & 1);
        peek >>= 1;
        const int cnt = count[len];
        if (code - cnt < first) {
            sym = symbol[index + (code - first)];
            nbits = len;
            return FB200_OK;
        }
        index += cnt;
        first += cnt;
        first <<= 1;
        code <<= 1;
    }
    return FB200_INVALID_CODE;
}

// huffman_decoder.zig:126-153 checkCompletnes + canonical tables.  Whole warp; returns status (uniform).
static __device__ int build_decoder(const uint8_t* lens, uint32_t n, bool is_lit, uint32_t max_code_bits, uint16_t* count,
                             uint16_t* symbol, uint32_t* fast, uint32_t fast_bits) {
    const uint32_t lane = threadIdx.x & 31;
    int status = FB200_OK;
    __shared__ uint16_t offs_all[kInflateWarps][17];
    uint16_t* offs = offs_all[(threadIdx.x >> 5) % kInflateWarps];
    if (lane == 0) {
        if (is_lit && lens[256] == 0) status = FB200_MISSING_END_OF_BLOCK_CODE;  // :127-128
        if (status == FB200_OK) {
            for (uint32_t i = 0; i < 16; i++) count[i] = 0;
            uint32_t mx = 0;
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t l = lens[i];
                if (l == 0) continue;
                if (l > mx) mx = l;
                count[l]++;
            }
            if (mx != 0) {
                int left = 1;
                for (uint32_t len = 1; len <= max_code_bits; len++) {
                    left <<= 1;
                    if ((int)count[len] > left) { status = FB200_OVERSUBSCRIBED_HUFFMAN_TREE; break; }
                    left -= count[len];
                }
                if (status == FB200_OK && left > 0) {
                    if (!(max_code_bits > 7 && mx == count[1])) status = FB200_INCOMPLETE_HUFFMAN_TREE;  // :148-151
                }
            }
            if (status == FB200_OK) {
                offs[1] = 0;
                for (uint32_t len = 1; len < 16; len++) offs[len + 1] = offs[len] + count[len];
                for (uint32_t i = 0; i < n; i++)
                    if (lens[i]) symbol[offs[lens[i]]++] = (uint16_t)i;
            }
        }
    }
    __syncwarp();
    status = __shfl_sync(0xffffffffu, status, 0);
    if (status != FB200_OK) return status;
    if (fast == nullptr) return status;
    for (uint32_t i = lane; i < (1u << fast_bits); i += 32) fast[i] = 0;
    __syncwarp();
    // first canonical code of each length
    uint32_t code = 0, index = 0;
    for (uint32_t len = 1; len <= fast_bits && len <= max_code_bits; len++) {
        const uint32_t cnt = count[len];
        // symbols symbol[index .. index+cnt) have codes code .. code+cnt-1 (MSB-first)
        for (uint32_t k = lane; k < cnt; k += 32) {
            const uint32_t rev = __brev(code + k) >> (32 - len);
            const uint32_t sym = symbol[index + k];
            uint32_t eb = 0, base_v = 0;
            if (is_lit) {
                if (sym >= 257 && sym <= 285) {
                    eb = length_extra_bits(sym - 257);
                    base_v = c_len_base[sym - 257];
                }
            } else if (sym <= 29) {
                eb = distance_extra_bits(sym);
                base_v = c_dist_base[sym];
            }
            const uint32_t entry = len | (sym << 4) | (eb << 13) | (base_v << 17);
            for (uint32_t e = rev; e < (1u << fast_bits); e += (1u << len)) fast[e] = entry;
        }
        code = (code + cnt) << 1;
        index += cnt;
    }
    __syncwarp();
    return status;
}

static __device__ void build_fixed_lens(uint8_t* lit_lens, uint8_t* dist_lens) {
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t i = lane; i < 288; i += 32) lit_lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
    for (uint32_t i = lane; i < 32; i += 32) dist_lens[i] = 5;
    __syncwarp();
}

// ---- checksums over the member's output, warp-parallel ----
__device__ __forceinline__ uint32_t multmodp(uint32_t a, uint32_t b) {  // GF(2)[x] mod the reflected CRC-32 polynomial
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
static __device__ uint32_t x8nmodp(uint64_t nbytes) {  // x^(8 n) mod p
    uint32_t sq = 1u << 30;                     // x^1
    sq = multmodp(sq, sq);                      // x^2
    sq = multmodp(sq, sq);                      // x^4
    sq = multmodp(sq, sq);                      // x^8
    uint32_t p = 1u << 31;                      // x^0
    while (nbytes) {
        if (nbytes & 1) p = multmodp(sq, p);
        sq = multmodp(sq, sq);
        nbytes >>= 1;
    }
    return p;
}
static __device__ uint32_t crc32_chunk(const uint32_t* tab, const uint8_t* p, uint64_t n) {
    uint32_t c = 0xffffffffu;
    uint64_t i = 0;
    while (i < n && (((uintptr_t)(p + i)) & 15)) {
        c = tab[(c ^ p[i]) & 0xff] ^ (c >> 8);
        i++;
    }
    for (; i + 16 <= n; i += 16) {
        const uint4 v = *reinterpret_cast<const uint4*>(p + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t x = w[k];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                c = tab[(c ^ x) & 0xff] ^ (c >> 8);
                x >>= 8;
            }
        }
    }
    for (; i < n; i++) c = tab[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return ~c;
}
static __device__ uint32_t warp_crc32(const uint32_t* tab, const uint8_t* p, uint64_t n) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t per = (n + 31) / 32;
    const uint64_t b = min(per * lane, n), e = min(b + per, n);
    uint32_t term = 0;
    if (e > b) term = multmodp(x8nmodp(n - e), crc32_chunk(tab, p + b, e - b));
    for (int o = 16; o > 0; o >>= 1) term ^= __shfl_xor_sync(0xffffffffu, term, o);
    return term;
}
static __device__ uint32_t warp_adler32(const uint8_t* p, uint64_t n) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t per = (n + 31) / 32;
    const uint64_t b = min(per * lane, n), e = min(b + per, n);
    uint64_t a = 0, s = 0, i;  // a = sum of bytes, s = sum of (e - i) * byte_i (weights relative to the lane's end)
    for (i = b; i < e; i++) {
        a += p[i];
        s += a;
        if ((i & 2047) == 2047) { a %= 65521; s %= 65521; }
    }
    a %= 65521;
    s %= 65521;
    // Adler over the whole: A = 1 + sum a_l ; B = n + sum_l (s_l + a_l * (n - e_l))
    uint64_t A = a, B = (s + a * ((n - e) % 65521)) % 65521;
    for (int o = 16; o > 0; o >>= 1) {
        A += __shfl_xor_sync(0xffffffffu, A, o);
        B += __shfl_xor_sync(0xffffffffu, B, o);
    }
    A = (A + 1) % 65521;
    B = (B + n % 65521) % 65521;
    return (uint32_t)((B << 16) | A);
}

}  // namespace fb
