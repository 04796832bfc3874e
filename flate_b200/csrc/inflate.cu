// inflate.cu -- inflate on sm_100a.  One warp per member (independent gzip/zlib/raw stream).
//
// Lane 0 owns the bit cursor and walks the symbol stream; Huffman tables are rebuilt per deflate
// block by the whole warp into shared memory (10-bit direct table + canonical fallback), and
// back-references are resolved by the whole warp against the flat output buffer in HBM (the
// reference's 64 KiB CircularBuffer, CircularBuffer.zig:44-75, becomes a flat buffer with the
// same InvalidMatch rule).  Error classes and their order follow inflate.zig / huffman_decoder.zig /
// bit_reader.zig exactly; the bit cursor reproduces bit_reader.zig's
// "fill fails only when no bit is left, shift fails when fewer than n bits are left" rules on a
// flat input with zero-padded peeks.
#include "inflate_dev.cuh"

namespace fb {


#define BCAST(x) __shfl_sync(0xffffffffu, (x), 0)

// ---- output window: the reference's CircularBuffer (CircularBuffer.zig) as a per-warp ring in
// shared memory.  Literals and match copies stay on the SM; the ring is drained to HBM in
// 16-byte coalesced stores.  Ring slot of output position p is (p + A) mod kRing with
// A = (address of out) mod 16, so ring and global alignment agree.
constexpr uint32_t kRing = 16384;
constexpr uint32_t kFlushAt = 4096;  // drain when this many bytes are pending

struct OutWindow {
    uint8_t* ring;
    uint8_t* out;
    uint32_t A;
    uint64_t flushed;  // output positions [0, flushed) are in HBM
    __device__ __forceinline__ uint32_t slot(uint64_t p) const { return (uint32_t)(p + A) & (kRing - 1); }
    // drain [flushed, upto); whole warp.  upto is either 16-byte aligned in (p + A) or final.
    __device__ void drain(uint64_t upto) {
        const uint32_t lane = threadIdx.x & 31;
        uint64_t f = flushed;
        // head: bytes until (f + A) is 16-byte aligned
        const uint64_t head_end = min(upto, (f + A + 15) / 16 * 16 - A);
        for (uint64_t p = f + lane; p < head_end; p += 32) out[p] = ring[slot(p)];
        f = head_end;
        const uint64_t nvec = (upto - f) / 16;
        for (uint64_t v = lane; v < nvec; v += 32) {
            const uint64_t p = f + v * 16;
            *reinterpret_cast<uint4*>(out + p) = *reinterpret_cast<const uint4*>(ring + slot(p));
        }
        f += nvec * 16;
        for (uint64_t p = f + lane; p < upto; p += 32) out[p] = ring[slot(p)];
        flushed = upto;
        __syncwarp();
    }
};

__global__ void __launch_bounds__(32)
inflate_members_kernel(int container, const uint8_t* __restrict__ d_in, const MemberDesc* __restrict__ descs, uint32_t k,
                       uint8_t* d_out, MemberResult* __restrict__ results) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t m = blockIdx.x;
    if (m >= k) return;
    WarpTables& T = *reinterpret_cast<WarpTables*>(smem_raw + kRing);
    const MemberDesc md = descs[m];
    uint8_t* out = d_out + md.out_off;
    const uint64_t cap = md.out_cap;
    uint64_t pos = 0;  // bytes produced (uniform across the warp after each broadcast)
    int status = FB200_OK;
    OutWindow W;
    W.ring = smem_raw;
    W.out = out;
    W.A = (uint32_t)((uintptr_t)out & 15);
    W.flushed = 0;

    // shared-window addresses, made opaque so that they live in registers instead of being re-derived (S2R) per token
    uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(smem_raw);
    asm volatile("mov.u32 %0, %0;" : "+r"(ring_addr));
    const uint32_t lit_fast_addr = ring_addr + kRing + (uint32_t)offsetof(WarpTables, lit_fast);
    const uint32_t dist_fast_addr = ring_addr + kRing + (uint32_t)offsetof(WarpTables, dist_fast);
    const uint32_t queue_addr = ring_addr + kRing + (uint32_t)offsetof(WarpTables, queue);
    BitCursor bc;
    bc.next = d_in + md.in_off;
    bc.end = bc.next + md.in_len;
    bc.buf = 0;
    bc.cnt = 0;

    // ---- container header (container.zig:111-152), lane 0 ----
    if (lane == 0 && container != FB200_RAW) {
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
    status = BCAST(status);

    bool fixed_ready = false;
    while (status == FB200_OK) {  // inflate.zig:251-280 step: one deflate block per iteration
        uint32_t bfinal = 0, btype = 0;
        if (lane == 0) {
            status = bc.read(1, bfinal);
            if (!status) status = bc.read(2, btype);
        }
        status = BCAST(status);
        if (status) break;
        bfinal = BCAST(bfinal);
        btype = BCAST(btype);

        if (btype == 0) {  // stored block, inflate.zig:89-102
            uint32_t len = 0;
            const uint8_t* src = nullptr;
            if (lane == 0) {
                bc.align_to_byte();
                uint32_t nlen = 0;
                status = bc.read(16, len);
                if (!status) status = bc.read(16, nlen);
                if (!status && len != ((~nlen) & 0xffffu)) status = FB200_WRONG_STORED_BLOCK_NLEN;
                if (!status) {
                    src = bc.byte_pos();
                    if ((uint64_t)(bc.end - src) < len) status = FB200_END_OF_STREAM;
                    else if (pos + len > cap) status = FB200_NO_SPACE_LEFT;
                }
            }
            status = BCAST(status);
            if (status) break;
            len = BCAST(len);
            src = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)src, 0);
            for (uint32_t done = 0; done < len;) {  // through the ring in pieces so history stays valid
                const uint32_t piece = min(len - done, 4096u);
                for (uint32_t i = lane; i < piece; i += 32) W.ring[W.slot(pos + i)] = src[done + i];
                __syncwarp();
                pos += piece;
                done += piece;
                if (pos - W.flushed >= kFlushAt) W.drain(pos - ((pos + W.A) & 15));
            }
            if (lane == 0) {  // reposition the cursor after the raw bytes
                bc.next = src + len;
                bc.buf = 0;
                bc.cnt = 0;
            }
            __syncwarp();
        } else if (btype == 1 || btype == 2) {
            if (btype == 2) {  // dynamicBlockHeader, inflate.zig:144-185
                fixed_ready = false;
                uint32_t hlit = 0, hdist = 0;
                if (lane == 0) {
                    uint32_t v = 0, hclen = 0;
                    status = bc.read(5, v); hlit = v + 257;
                    if (!status) { status = bc.read(5, v); hdist = v + 1; }
                    if (!status) { status = bc.read(4, v); hclen = v + 4; }
                    if (!status && (hlit > 286 || hdist > 30)) status = FB200_INVALID_DYNAMIC_BLOCK_HEADER;
                    // code-length code lengths go to dist_lens[0..19) temporarily
                    if (!status) {
                        for (uint32_t i = 0; i < 19; i++) T.dist_lens[i] = 0;
                        for (uint32_t i = 0; i < hclen && !status; i++) {
                            status = bc.read(3, v);
                            T.dist_lens[c_cl_order[i]] = (uint8_t)v;
                        }
                    }
                }
                status = BCAST(status);
                if (status) break;
                // CodegenDecoder(19, 7, 7): built into dist_count / dist_sym (no fast table)
                status = build_decoder(T.dist_lens, 19, false, 7, T.dist_count, T.dist_sym, nullptr, 0);
                if (status) break;
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
                            uint32_t sym = 0, nb = 0;
                            status = slow_find(T.dist_count, T.dist_sym, 7, bc.peek(7), sym, nb);
                            if (status) break;
                            if (!bc.shift(nb)) { status = FB200_END_OF_STREAM; break; }
                            // dynamicCodeLength, inflate.zig:189-216
                            if (p >= lens_len) { status = FB200_INVALID_DYNAMIC_BLOCK_HEADER; break; }
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
                if (status) break;
                __syncwarp();
                status = build_decoder(T.lit_lens, kNumLit, true, 15, T.lit_count, T.lit_sym, T.lit_fast, kLitFast);
                if (status) break;
                status = build_decoder(T.dist_lens, kNumDist, false, 15, T.dist_count, T.dist_sym, T.dist_fast, kDistFast);
                if (status) break;
            } else if (!fixed_ready) {
                // fixed block: the reference decodes by arithmetic (bit_reader.zig:205-217); the same symbols
                // come out of the canonical code with lengths 8/9/7/8 over 288 symbols and 32 five-bit
                // distance codes.  286/287 and 30/31 decode and are then rejected (inflate.zig:111,136).
                build_fixed_lens(T.lit_lens, T.dist_lens);
                status = build_decoder(T.lit_lens, 288, true, 15, T.lit_count, T.lit_sym, T.lit_fast, kLitFast);
                if (!status) status = build_decoder(T.dist_lens, 32, false, 15, T.dist_count, T.dist_sym, T.dist_fast, kDistFast);
                if (status) break;
                fixed_ready = true;
            }
            // ---- symbol loop (inflate.zig:220-239 dynamicBlock / :104-124 fixedBlock) ----
            bool done = false;
            while (!done) {
                // ---- batch: lane 0 decodes ahead.  Literals go straight into the ring, matches are queued
                // (their bytes do not influence decoding) and resolved afterwards by the whole warp.  Only
                // regular cases are committed here; anything else (codes outside the direct tables, end of
                // block, short input, invalid symbols, capacity) is left to the exact general path below.
                uint32_t nq = 0;
                const uint64_t pos0 = pos;  // uniform here; lane 0 runs ahead from it
                if (lane == 0) {
                    // Local state, 32-bit arithmetic only: every instruction here is on the member's critical
                    // path (one warp issues a dependent instruction every ~6 cycles).  The cursor is a bit
                    // offset `bo` (< 32) into three consecutive little-endian words w0, w1, w2 of the input;
                    // w2 is loaded 32 bits ahead of its use, so the load latency stays off the chain.
                    const uint64_t stop = min(min(cap, W.flushed + kFlushAt), pos + kBatchSpan);
                    const uint32_t stop32 = (uint32_t)(stop - pos0);          // < 2^13
                    const uint32_t room32 = (uint32_t)min(cap - pos0, (uint64_t)0x7fffffffu);
                    const uint32_t slot0 = (uint32_t)pos0 + W.A;
                    const uint64_t reach = md.hist + pos0;                     // bytes a match may reach back from pos0
                    const unsigned long long q0 = (unsigned long long)(uintptr_t)bc.next * 8ull - bc.cnt;  // cursor as a bit address
                    const uintptr_t wa = (uintptr_t)((q0 >> 3) & ~3ull);
                    const uintptr_t lo_addr = (uintptr_t)(d_in + md.in_off), end_addr = (uintptr_t)bc.end;
                    uint32_t p32 = 0;
                    if (wa >= lo_addr && wa + 12 <= end_addr) {
                        const uint32_t* const wbase = reinterpret_cast<const uint32_t*>(wa);
                        const uint32_t wlim = (uint32_t)(((end_addr & ~(uintptr_t)3) - wa) >> 2);  // whole words from wbase
                        uint32_t w0 = __ldg(wbase), w1 = __ldg(wbase + 1), w2 = __ldg(wbase + 2);
                        uint32_t wi = 3;                                       // next word to load
                        uint32_t bo = (uint32_t)(q0 - (unsigned long long)wa * 8ull);
                        const uint32_t reach32 = (uint32_t)min(reach, (uint64_t)0xffff0000u);
                        // a token takes at most 10 + 5 + 8 + 13 = 36 bits: two word advances, both must be loadable
                        while (p32 < stop32 && nq < kQueue && wi + 2 <= wlim) {
                            const uint32_t win = __funnelshift_r(w0, w1, bo);
                            uint32_t e;
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(lit_fast_addr + ((win & ((1u << kLitFast) - 1)) << 2)));
                            const uint32_t nb = e & 15;
                            if (nb == 0) break;
                            const uint32_t sym = bfe32(e, 4, 9);
                            if (sym < 256) {
                                asm volatile("st.shared.u8 [%0], %1;" ::"r"(ring_addr + ((slot0 + p32) & (kRing - 1))), "r"(sym));
                                bo += nb;
                                p32++;
                                if (bo >= 32) {
                                    bo -= 32;
                                    w0 = w1;
                                    w1 = w2;
                                    w2 = __ldg(wbase + wi);
                                    wi++;
                                }
                                continue;
                            }
                            if (sym - 257u > 28u) break;  // end of block, or 286/287
                            // length + distance; an irregular case restores the cursor from (bo0, wi0)
                            const uint32_t bo0 = bo, wi0 = wi;
                            const uint32_t leb = bfe32(e, 13, 4);
                            const uint32_t length = (e >> 17) + bfe32(win, nb, leb);
                            bo += nb + leb;
                            if (bo >= 32) {
                                bo -= 32;
                                w0 = w1;
                                w1 = w2;
                                w2 = __ldg(wbase + wi);
                                wi++;
                            }
                            const uint32_t dwin = __funnelshift_r(w0, w1, bo);
                            uint32_t de;
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(de) : "r"(dist_fast_addr + ((dwin & ((1u << kDistFast) - 1)) << 2)));
                            const uint32_t dnb = de & 15, deb = bfe32(de, 13, 4);
                            const uint32_t distance = (de >> 17) + bfe32(dwin, dnb, deb);  // dnb + deb <= 21 bits
                            if (dnb == 0 || bfe32(de, 4, 9) > 29 || reach32 + p32 < distance || p32 + length > room32) {
                                bo = bo0;  // rare: put the cursor back to the start of the token
                                wi = wi0;
                                w0 = __ldg(wbase + wi - 3);
                                w1 = __ldg(wbase + wi - 2);
                                w2 = __ldg(wbase + wi - 1);
                                break;
                            }
                            bo += dnb + deb;
                            if (bo >= 32) {
                                bo -= 32;
                                w0 = w1;
                                w1 = w2;
                                w2 = __ldg(wbase + wi);
                                wi++;
                            }
                            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(queue_addr + 8 * nq), "r"(p32), "r"((length << 16) | (distance - 1)));
                            nq++;
                            p32 += length;
                        }
                        if (p32) {
                            // back to the cursor of the exact path, straight from the registers: w0 and w1 are whole
                            // bytes of the stream, so cnt mod 8 stays the stream's bit phase (align_to_byte relies on it)
                            bc.buf = (((uint64_t)w1 << 32) | w0) >> bo;
                            bc.cnt = 64 - bo;
                            bc.next = reinterpret_cast<const uint8_t*>(wbase + wi - 1);
                        }
                    }
                    pos = pos0 + p32;
                }
                __syncwarp();
                nq = BCAST(nq);
                // Resolve the queue: four matches at a time, eight lanes each.  A match may only run together
                // with earlier ones of its group if its source ends before the group's first destination
                // (conservative); otherwise it starts the next group.
                // (32-bit positions relative to the batch start; the ring is addressed through its shared window)
                const uint32_t slot0r = (uint32_t)pos0 + W.A;
                const uint32_t back32 = (uint32_t)min(pos0, (uint64_t)0xffff0000u);  // output bytes before the batch
                for (uint32_t k = 0; k < nq;) {
                    const uint32_t g = lane >> 3, sub = lane & 7;
                    const bool valid = k + g < nq;
                    uint32_t qlo = 0, w1 = 0;
                    if (valid) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(qlo), "=r"(w1) : "r"(queue_addr + 8 * (k + g)));
                    const uint32_t q_len = w1 >> 16, q_dist = (w1 & 0xffffu) + 1;
                    const uint32_t first_lo = __shfl_sync(0xffffffffu, qlo, 0);
                    // source bytes actually read: [qlo - dist, qlo - dist + min(len, dist))
                    const int32_t src_end = (int32_t)qlo - (int32_t)q_dist + (int32_t)min(q_len, q_dist);
                    const bool conflict = valid && g > 0 && src_end > (int32_t)first_lo;
                    const uint32_t cmask = __ballot_sync(0xffffffffu, conflict);
                    const uint32_t take = cmask ? (uint32_t)(__ffs(cmask) - 1) >> 3 : min(4u, nq - k);
                    if (valid && g < take) {
                        if (q_dist <= kRing - kBatchSpan - 512 && q_dist <= back32 + qlo) {
                            const uint32_t to = slot0r + qlo, from = to - q_dist;
                            if (q_dist >= q_len) {
                                for (uint32_t i = sub; i < q_len; i += 8) {
                                    uint32_t b;
                                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(ring_addr + ((from + i) & (kRing - 1))));
                                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(ring_addr + ((to + i) & (kRing - 1))), "r"(b));
                                }
                            } else {
                                for (uint32_t i = sub; i < q_len; i += 8) {
                                    uint32_t b;
                                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(ring_addr + ((from + i % q_dist) & (kRing - 1))));
                                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(ring_addr + ((to + i) & (kRing - 1))), "r"(b));
                                }
                            }
                        } else {
                            // far match, or one that reaches into an earlier member's output: the source left the ring
                            // but was drained to HBM long ago (pending <= kFlushAt + kBatchSpan + 258)
                            const uint64_t qpos = pos0 + qlo;
                            for (uint32_t i = sub; i < q_len; i += 8) {
                                const int64_t sp = (int64_t)qpos - q_dist + (q_dist >= q_len ? i : i % q_dist);
                                uint8_t b;
                                if (sp >= (int64_t)W.flushed) b = W.ring[W.slot((uint64_t)sp)];
                                else b = __ldcg(out + sp);
                                W.ring[W.slot(qpos + i)] = b;
                            }
                        }
                    }
                    __syncwarp();
                    k += take;
                }
                uint32_t ev_len = 0, ev_dist = 0;  // match event (len > 0); otherwise flush / end of block / error
                if (lane == 0) {
                    for (;;) {
                        if (pos - W.flushed >= kFlushAt) break;                    // drain request
                        if (bc.empty()) { status = FB200_END_OF_STREAM; break; }  // fill(15) / fill(7+2)
                        uint32_t sym, nb;
                        const uint32_t e = T.lit_fast[bc.peek(kLitFast)];
                        if (e & 15) {
                            sym = (e >> 4) & 511u;
                            nb = e & 15;
                        } else {
                            status = slow_find(T.lit_count, T.lit_sym, 15, bc.peek(15), sym, nb);
                            if (status) break;
                        }
                        if (!bc.shift(nb)) { status = FB200_END_OF_STREAM; break; }
                        if (sym < 256) {
                            if (pos >= cap) { status = FB200_NO_SPACE_LEFT; break; }
                            W.ring[W.slot(pos)] = (uint8_t)sym;
                            pos++;
                            continue;
                        }
                        if (sym == 256) { done = true; break; }
                        // match: fill(5+15+13), decodeLength, distance symbol, decodeDistance.  Symbols 286/287
                        // only exist in the fixed code, where the reference rejects them before any fill
                        // (inflate.zig:111); in a dynamic block they cannot occur.
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
                        const uint32_t de = T.dist_fast[bc.peek(kDistFast)];
                        if (de & 15) {
                            dsym = (de >> 4) & 511u;
                            dnb = de & 15;
                        } else {
                            status = slow_find(T.dist_count, T.dist_sym, 15, bc.peek(15), dsym, dnb);
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
                        // CircularBuffer.zig:45-50 writeMatch validation
                        if (md.hist + pos < distance) { status = FB200_INVALID_MATCH; break; }
                        if (pos + length > cap) { status = FB200_NO_SPACE_LEFT; break; }
                        ev_len = length;
                        ev_dist = distance;
                        break;
                    }
                }
                __syncwarp();
                status = BCAST(status);
                if (status) break;
                done = BCAST(done);
                pos = __shfl_sync(0xffffffffu, (unsigned long long)pos, 0);
                ev_len = BCAST(ev_len);
                if (ev_len) {
                    ev_dist = BCAST(ev_dist);
                    if (ev_dist <= kRing - 264 && ev_dist <= pos) {
                        // source still in the ring: shared memory to shared memory
                        const uint64_t from = pos - ev_dist;
                        if (ev_dist >= ev_len) {
                            for (uint32_t i = lane; i < ev_len; i += 32) W.ring[W.slot(pos + i)] = W.ring[W.slot(from + i)];
                        } else {
                            for (uint32_t i = lane; i < ev_len; i += 32)
                                W.ring[W.slot(pos + i)] = W.ring[W.slot(from + i % ev_dist)];
                        }
                    } else {
                        // far match, or one that reaches into an earlier member's output: the source left the
                        // ring but was drained to HBM long ago (pending <= kFlushAt + 258 < kRing - 264)
                        for (uint32_t i = lane; i < ev_len; i += 32) {
                            const int64_t sp = (int64_t)pos - ev_dist + (ev_dist >= ev_len ? i : i % ev_dist);
                            uint8_t b;
                            if (sp >= (int64_t)W.flushed) b = W.ring[W.slot((uint64_t)sp)];
                            else b = __ldcg(out + sp);
                            W.ring[W.slot(pos + i)] = b;
                        }
                    }
                    pos += ev_len;
                    __syncwarp();
                }
                if (pos - W.flushed >= kFlushAt) W.drain(pos - ((pos + W.A) & 15));
            }
            if (status) break;
        } else {
            status = FB200_INVALID_BLOCK_TYPE;  // inflate.zig:267
            break;
        }
        if (bfinal) break;
    }
    pos = __shfl_sync(0xffffffffu, (unsigned long long)pos, 0);
    __syncwarp();
    W.drain(pos);  // whatever was produced, also on error (the caller sees out_len and the status)

    // ---- protocol footer (inflate.zig:271-275, container.zig:154-166) ----
    if (status == FB200_OK && container != FB200_RAW) {
        __syncwarp();
        uint32_t sum;
        if (container == FB200_GZIP) {
            for (uint32_t i = lane; i < 256; i += 32) {
                uint32_t c = i;
                for (int b = 0; b < 8; b++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
                T.crc_tab[i] = c;
            }
            __syncwarp();
            sum = warp_crc32(T.crc_tab, out, pos);
        } else {
            sum = warp_adler32(out, pos);
        }
        if (lane == 0) {
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
    if (lane == 0) {
        if (status == FB200_OK) bc.align_to_byte();
        MemberResult r{};
        r.out_len = pos;
        // bytes consumed: everything handed to the cursor minus whole bytes still buffered
        r.consumed = (uint64_t)((bc.next - (bc.cnt >> 3)) - (d_in + md.in_off));
        r.status = (uint32_t)status;
        r.pad = 0;
        results[m] = r;
    }
}

// ---- standalone checksum kernels (used for the gzip/zlib footers of compress) ----
constexpr uint32_t kSumChunk = 4096;
__global__ void __launch_bounds__(256)
crc32_chunks_kernel(const uint8_t* __restrict__ data, uint64_t n, uint32_t* __restrict__ result) {
    __shared__ uint32_t tab[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = i;
        for (int b = 0; b < 8; b++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        tab[i] = c;
    }
    __syncthreads();
    const uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t b = chunk * kSumChunk;
    uint32_t term = 0;
    if (b < n) {
        const uint64_t e = min(b + kSumChunk, n);
        term = multmodp(x8nmodp(n - e), crc32_chunk(tab, data + b, e - b));
    }
    for (int o = 16; o > 0; o >>= 1) term ^= __shfl_xor_sync(0xffffffffu, term, o);
    if ((threadIdx.x & 31) == 0 && term) atomicXor(result, term);
}
__global__ void __launch_bounds__(256)
adler32_chunks_kernel(const uint8_t* __restrict__ data, uint64_t n, unsigned long long* __restrict__ acc) {
    const uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t b = chunk * kSumChunk;
    unsigned long long A = 0, B = 0;
    if (b < n) {
        const uint64_t e = min(b + kSumChunk, n);
        uint64_t a = 0, s = 0;
        for (uint64_t i = b; i < e; i++) {
            a += data[i];
            s += a;
        }
        a %= 65521;
        s %= 65521;
        A = a;
        B = (s + a * ((n - e) % 65521)) % 65521;
    }
    for (int o = 16; o > 0; o >>= 1) {
        A += __shfl_xor_sync(0xffffffffu, A, o);
        B += __shfl_xor_sync(0xffffffffu, B, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc, A);
        atomicAdd(acc + 1, B);
    }
}
__global__ void adler32_finish_kernel(const unsigned long long* __restrict__ acc, uint64_t n, uint32_t* __restrict__ result) {
    const uint64_t A = (acc[0] + 1) % 65521;
    const uint64_t B = (acc[1] + n % 65521) % 65521;
    *result = (uint32_t)((B << 16) | A);
}

cudaError_t inflate_members(int container, const uint8_t* d_in, const MemberDesc* d_desc, uint32_t k, uint8_t* d_out,
                            MemberResult* d_res, cudaStream_t st) {
    if (k == 0) return cudaSuccess;
    static bool attr_set[64] = {};  // per device
    const size_t smem = kRing + sizeof(WarpTables);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaFuncSetAttribute(inflate_members_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    inflate_members_kernel<<<k, 32, smem, st>>>(container, d_in, d_desc, k, d_out, d_res);
    return cudaGetLastError();
}
// ---- member discovery: every position that could start a gzip member ----
__global__ void __launch_bounds__(256)
gzip_candidates_kernel(const uint8_t* __restrict__ data, uint64_t n, uint64_t* __restrict__ list, uint32_t cap, uint32_t* __restrict__ count) {
    // a thread looks at four positions through two aligned words (the buffer is readable a few bytes past n)
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t mis = (uint32_t)((uintptr_t)data & 3);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(data - mis);
    const uint64_t nwords = (n + mis + 3) / 4;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride) {
        const uint32_t w0 = words[w], w1 = w + 1 < nwords ? words[w + 1] : 0u;
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t v = __funnelshift_r(w0, w1, 8 * k);
            if ((v & 0x00ffffffu) == 0x00088b1fu && (v & 0xe0000000u) == 0) {
                const uint64_t at = w * 4 + k;
                if (at >= mis && at - mis + 18 <= n) {  // a member is at least header + empty block + footer
                    const uint32_t slot = atomicAdd(count, 1u);
                    if (slot < cap) list[slot] = at - mis;
                }
            }
        }
    }
}
cudaError_t gzip_candidates_device(const uint8_t* d_data, uint64_t n, uint64_t* d_list, uint32_t cap, uint32_t* d_count, cudaStream_t st) {
    cudaMemsetAsync(d_count, 0, 4, st);
    if (n >= 18) gzip_candidates_kernel<<<148 * 8, 256, 0, st>>>(d_data, n, d_list, cap, d_count);
    return cudaGetLastError();
}

cudaError_t crc32_device(const uint8_t* d_data, uint64_t n, uint32_t* d_result, cudaStream_t st) {
    cudaMemsetAsync(d_result, 0, 4, st);
    const uint64_t chunks = (n + kSumChunk - 1) / kSumChunk;
    if (chunks) crc32_chunks_kernel<<<(uint32_t)((chunks + 255) / 256), 256, 0, st>>>(d_data, n, d_result);
    return cudaGetLastError();
}
cudaError_t adler32_device(const uint8_t* d_data, uint64_t n, uint32_t* d_result, uint64_t* d_scratch2, cudaStream_t st) {
    cudaMemsetAsync(d_scratch2, 0, 16, st);
    const uint64_t chunks = (n + kSumChunk - 1) / kSumChunk;
    if (chunks)
        adler32_chunks_kernel<<<(uint32_t)((chunks + 255) / 256), 256, 0, st>>>(d_data, n, (unsigned long long*)d_scratch2);
    adler32_finish_kernel<<<1, 1, 0, st>>>((const unsigned long long*)d_scratch2, n, d_result);
    return cudaGetLastError();
}

}  // namespace fb
