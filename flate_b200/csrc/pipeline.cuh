// pipeline.cuh -- device workspace layout and kernel launchers shared by the C-ABI layer.
#pragma once
#include "common.cuh"

namespace fb {

// Optional per-phase device timing (CUDA events on the launching stream), used by bench.py to report
// the dominant kernel's duration live.  Off by default: mark() is then a no-op.
enum Phase : int {
    kPhLink = 0, kPhSearch, kPhLazy, kPhChunkExit, kPhResolve, kPhMark, kPhScan, kPhEmit, kPhHist, kPhBuild,
    kPhOffsets, kPhPack, kPhInflate, kPhSparse, kPhCount
};
struct PhaseTimer {
    static constexpr int kMaxMarks = 64;
    bool on = false;
    cudaEvent_t ev[kMaxMarks];
    int phase_of[kMaxMarks];
    int n = 0;
    bool created = false;
    double ms[kPhCount] = {0};
    uint64_t count[kPhCount] = {0};
    void begin(cudaStream_t st) {
        if (!on) return;
        if (!created) {
            for (int i = 0; i < kMaxMarks; i++) cudaEventCreate(&ev[i]);
            created = true;
        }
        n = 0;
        cudaEventRecord(ev[n], st);
        phase_of[n++] = -1;
    }
    void mark(cudaStream_t st, int phase) {  // call after the phase's kernels were launched
        if (!on || n >= kMaxMarks) return;
        cudaEventRecord(ev[n], st);
        phase_of[n++] = phase;
    }
    void destroy() {
        if (created)
            for (int i = 0; i < kMaxMarks; i++) cudaEventDestroy(ev[i]);
        created = false;
    }
    void collect() {  // after the stream was synchronized
        if (!on) return;
        for (int i = 1; i < n; i++) {
            float t = 0;
            if (cudaEventElapsedTime(&t, ev[i - 1], ev[i]) == cudaSuccess && phase_of[i] >= 0) {
                ms[phase_of[i]] += t;
                count[phase_of[i]] += 1;
            }
        }
        n = 0;
    }
};

constexpr uint32_t kChunk = 4096;    // positions per lazy-parse chunk
constexpr uint32_t kEntries = 516;   // possible entry offsets into a chunk (step <= 515)
constexpr uint32_t kGroup = 256;     // chunks per resolution group

// Workspace of the LZ77 phases for an input of n positions (all device pointers).
struct Lz77Buffers {
    uint16_t* link;          // n        distance to previous same-hash position, 0 = none
    uint32_t* r_full;        // n        findMatch result, full budget
    uint32_t* r_quarter;     // n        findMatch result, chain >> 2 budget
    uint32_t* nx;            // n        packed lazy step from a clean arrival
    uint16_t* exits;         // nchunks * kEntries
    uint16_t* gexits;        // ngroups * kEntries
    uint16_t* gentry;        // ngroups
    uint16_t* entry;         // nchunks
    uint16_t* jumps;         // nchunks * kChunk   sub-chunk jump table (chunk_exit -> orbit_mark, sparse strategy)
    uint32_t* bitmap;        // nchunks * kChunk/32   arrivals actually visited
    uint32_t* chunk_tokens;  // nchunks
    uint32_t* tok_offset;    // nchunks
    uint32_t* total_tokens;  // 1
    uint32_t* tokens;        // up to n
    uint32_t* cut_rp;        // nblocks: reference `rp` when block b's last token was added
};

// Tokenizes stream positions [begin, n); earlier bytes are history (begin > 0 after a sync flush).
// d_skip/nskip: history positions the reference never inserted into its hash chains.  r_full,
// r_quarter, nx, tokens and cut_rp are relative to `begin`.
cudaError_t lz77_tokenize(const Lz77Buffers& b, const uint8_t* d_in, uint32_t begin, uint32_t n, const uint32_t* d_skip,
                          uint32_t nskip, const LevelArgs& lv, cudaStream_t st, PhaseTimer* pt = nullptr);

// the two halves of lz77_tokenize, for callers that overlap the input copy with the search
cudaError_t lz77_search_range(const Lz77Buffers& b, const uint8_t* d_in, uint32_t seg_begin, uint32_t from, uint32_t range_end,
                              uint32_t n, const uint32_t* d_skip, uint32_t nskip, const LevelArgs& lv, cudaStream_t st,
                              PhaseTimer* pt = nullptr, uint32_t link_from = 0xffffffffu);
cudaError_t lz77_parse(const Lz77Buffers& b, const uint8_t* d_in, uint32_t begin, uint32_t n, const LevelArgs& lv,
                       cudaStream_t st, PhaseTimer* pt = nullptr);

// position-sharded single stream: stage 1 (every rank) and stage 2 (one rank, complete nx in b.nx)
cudaError_t lz77_shard_search(const Lz77Buffers& b, const uint8_t* d_in, uint32_t from, uint32_t to, uint32_t n,
                              const LevelArgs& lv, uint32_t* nx_out, cudaStream_t st, PhaseTimer* pt = nullptr);
cudaError_t lz77_parse_from_nx(const Lz77Buffers& b, const uint8_t* d_in, uint32_t n, const LevelArgs& lv, cudaStream_t st,
                               PhaseTimer* pt = nullptr, uint32_t* flags = nullptr, uint32_t tok_carry = 0);

// sparse parse (whole streams, begin = 0): links, then the speculative sparse kernel writes nx directly;
// lz77_parse_from_nx finishes.  *flags != 0 afterwards means the speculation did not cover the true
// orbit: bit 0 = some chunk_fail[c] is set (repair: lz77_sparse_dense_chunks on the chunks after the
// failed ones, repeated while the repaired chunks fail themselves, then parse again), bit 1 = the token
// emitter met an unevaluated entry (redo the stream with lz77_tokenize, dense tables).
cudaError_t lz77_link_range(const Lz77Buffers& b, const uint8_t* d_in, uint32_t link_from, uint32_t range_end, uint32_t n,
                            cudaStream_t st, PhaseTimer* pt = nullptr, const uint32_t* d_skip = nullptr, uint32_t nskip = 0);
uint32_t lz77_sparse_chunk();      // positions per sparse-parse CTA
uint32_t lz77_sparse_lookahead();  // positions past a chunk's end that the chunk's CTA reads (overlap + lazy halo + compare slack)
uint32_t lz77_sparse_link_ahead(); // positions past a chunk range's end that must be linked and resident (whole load epochs)
// run_counter: one device word for the rolling kernel's run hand-out (nullptr: chunk kernel only)
cudaError_t lz77_sparse_range(const Lz77Buffers& b, const uint8_t* d_in, uint32_t first_chunk, uint32_t end_chunk, uint32_t n,
                              const LevelArgs& lv, uint32_t* chunk_fail, uint32_t* flags, cudaStream_t st, PhaseTimer* pt = nullptr,
                              uint32_t begin = 0, uint32_t* run_counter = nullptr);
// repair after a failed coverage check: evaluate every position of the listed chunks
cudaError_t lz77_sparse_dense_chunks(const Lz77Buffers& b, const uint8_t* d_in, const uint32_t* chunk_list, uint32_t count, uint32_t n,
                                     const LevelArgs& lv, uint32_t* chunk_fail, uint32_t* flags, cudaStream_t st,
                                     PhaseTimer* pt = nullptr, uint32_t begin = 0);

// ---- block writer ----
enum WriteKind : uint32_t { kWrite = 0, kDynamicBlock = 1, kHuffmanBlock = 2 };  // block_writer.zig:307,395,524

struct BlockPlan {           // per-block inputs of the build kernel (device array)
    uint32_t tok_begin, tok_count;
    uint64_t in_begin;       // candidate stored-input range / huffman slice
    uint32_t in_len;
    uint32_t has_input;      // `input != null` (SlidingWindow.zig:119-123)
    uint32_t eof;
    uint32_t kind;           // WriteKind; 3 = stored block verbatim (store mode / sync marker)
};

// plans for the level modes are derived on the device from the token count and cut_rp
// final_flush: 0 = sync flush (open block closed, marker appended), 1 = finish, 2 = a part in the middle of a stream
// (complete blocks only; the open one is carried).  fp0: position of the last cut before these tokens (the first
// block's stored-input candidate starts there); pass `begin` when the tokens start a segment.
cudaError_t plan_level_blocks(const uint32_t* total_tokens, const uint32_t* cut_rp, uint32_t begin, uint32_t n, uint32_t max_blocks,
                              uint32_t final_flush, BlockPlan* plans, uint32_t* nblocks, cudaStream_t st, uint32_t fp0);
cudaError_t histogram_tokens(const uint32_t* tokens, const BlockPlan* plans, const uint32_t* nblocks_dev,
                             uint32_t max_blocks, uint32_t* lit_freq, uint32_t* dist_freq, cudaStream_t st);
cudaError_t histogram_bytes(const uint8_t* in, const BlockPlan* plans, uint32_t nblocks, uint32_t* lit_freq,
                            cudaStream_t st);
// split_scratch (may be null): build_blocks_split_bytes(max_blocks) bytes of device memory; with it the code construction
// runs in its split form (sort / bitCounts of 32 blocks per warp / the rest), which pays off for many blocks with large
// alphabets (huffman-only streams)
size_t build_blocks_split_bytes(uint32_t max_blocks);
cudaError_t build_blocks(const BlockPlan* plans, const uint32_t* nblocks_dev, uint32_t max_blocks,
                         const uint32_t* lit_freq, const uint32_t* dist_freq, BlockDesc* descs, cudaStream_t st,
                         void* split_scratch = nullptr);
// sequential offset scan; start_bits = container header bits. Writes total bits to *total_bits.
cudaError_t scan_block_offsets(BlockDesc* descs, const uint32_t* nblocks_dev, uint64_t start_bits, uint64_t* total_bits,
                               cudaStream_t st);
cudaError_t pack_blocks(const uint8_t* in, const uint32_t* tokens, const BlockDesc* descs, const uint32_t* nblocks_dev,
                        uint32_t max_blocks, uint32_t* out_words, cudaStream_t st);
constexpr uint32_t kPackParts = 4;  // scan_block_offsets reports kPackParts-1 interior part boundaries after total_bits
// first block of part i out of nb blocks: a small first part (10, 30, 30, 30 %), so that the copy back -- which takes
// several times longer than the packing -- starts as early as possible
__host__ __device__ inline uint32_t pack_part_begin(uint32_t nb, uint32_t i) {
    const uint32_t tenths = i == 0 ? 0u : i == 1 ? 1u : i == 2 ? 4u : i == 3 ? 7u : 10u;
    return (uint32_t)(((uint64_t)nb * tenths) / 10);
}
cudaError_t pack_blocks_range(const uint8_t* in, const uint32_t* tokens, const BlockDesc* descs, const uint32_t* nblocks_dev,
                              uint32_t first_block, uint32_t count, uint32_t* out_words, cudaStream_t st);

}  // namespace fb
