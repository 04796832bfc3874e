// capi.cu -- the extern "C" boundary (include/flate_b200.h) over the CUDA pipelines.
// No CPU fallback anywhere: without a CUDA device every entry point fails with FB200_NO_DEVICE.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <mutex>
#include <new>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/flate_b200.h"
#include "common.cuh"
#include "inflate.cuh"
#include "pipeline.cuh"

namespace fb {
static char g_cuda_err[256] = "";  // process-wide: worker threads (parts of a stream, devices of a pool) report through it too
void set_last_cuda_error(cudaError_t e, const char* file, int line) {
    snprintf(g_cuda_err, sizeof g_cuda_err, "%s (%s:%d)", cudaGetErrorString(e), file, line);
}

__global__ void plan_simple_blocks_kernel(uint64_t begin, uint64_t end, uint32_t nblocks, uint32_t kind, uint32_t final_flush,
                                          BlockPlan* __restrict__ plans, uint32_t* __restrict__ nblocks_dev) {
    // SimpleCompressor (deflate.zig:449-529): the segment is cut into 65535-byte slices, the last
    // (possibly empty) one closes the segment; a sync flush appends an empty stored block (:474-478).
    // final_flush: 0 = sync flush, 1 = finish, 2 = a shard in the middle of a block-range sharded stream
    // (whole slices only, neither BFINAL nor a marker)
    const uint32_t total = nblocks + (final_flush ? 0 : 1);
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *nblocks_dev = total;
    if (b >= total) return;
    BlockPlan pl;
    pl.tok_begin = 0;
    pl.tok_count = 0;
    pl.has_input = 1;
    if (b < nblocks) {
        pl.in_begin = begin + (uint64_t)b * kMaxStore;  // deflate.zig:456: 65535-byte slices
        const uint64_t rem = end - pl.in_begin;
        pl.in_len = (uint32_t)(rem < kMaxStore ? rem : kMaxStore);
        pl.eof = (b + 1 == nblocks) && final_flush == 1;
        pl.kind = kind;
    } else {
        pl.in_begin = 0;
        pl.in_len = 0;
        pl.eof = 0;
        pl.kind = 3;
    }
    plans[b] = pl;
}

// zero [0, ceil(total_bits/32)+2) words of the output, total_bits read on the device
__global__ void zero_output_kernel(uint32_t* __restrict__ out, const uint64_t* __restrict__ total_bits, uint64_t cap_words) {
    uint64_t words = ((*total_bits + 31) >> 5) + 2;
    if (words > cap_words) words = cap_words;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint4* o4 = reinterpret_cast<uint4*>(out);
    const uint64_t n4 = words / 4;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) o4[i] = make_uint4(0, 0, 0, 0);
    for (uint64_t i = n4 * 4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride) out[i] = 0;
}
}  // namespace fb

using namespace fb;

namespace {
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // the same from the device's stream-ordered memory pool (kept warm: a streaming compressor is created and
    // destroyed per stream, and cudaMalloc / cudaFree of its windows would cost more than compressing a short stream)
    cudaError_t ensure_pool(size_t n, cudaStream_t st) {
        if (n <= cap) return cudaSuccess;
        release_pool(st);
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&p), n * sizeof(T), st);
        if (e == cudaSuccess) cap = n;
        else p = nullptr;
        return e;
    }
    void release_pool(cudaStream_t st) {
        if (p) cudaFreeAsync(p, st);
        p = nullptr;
        cap = 0;
    }
};
// Pinned host buffers are expensive to create (the pages are locked one by one): a few are kept for the next stream.
struct PinnedCache {
    std::mutex mu;
    struct Entry { uint8_t* p; size_t cap; } slot[4] = {};
    uint8_t* get(size_t want, size_t* cap) {
        {
            std::lock_guard<std::mutex> g(mu);
            for (auto& e : slot)
                if (e.p && e.cap >= want) {
                    uint8_t* p = e.p;
                    *cap = e.cap;
                    e.p = nullptr;
                    return p;
                }
        }
        uint8_t* p = nullptr;
        if (cudaMallocHost(&p, want) != cudaSuccess) return nullptr;
        *cap = want;
        return p;
    }
    void put(uint8_t* p, size_t cap) {
        if (!p) return;
        {
            std::lock_guard<std::mutex> g(mu);
            for (auto& e : slot)
                if (!e.p) {
                    e.p = p;
                    e.cap = cap;
                    return;
                }
        }
        cudaFreeHost(p);
    }
};
inline PinnedCache& pinned_cache() {
    static PinnedCache* c = new PinnedCache();  // never destroyed: the CUDA runtime may be gone before static destructors run
    return *c;
}
inline void keep_pool_warm(int device) {
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return;
    done[device] = true;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) != cudaSuccess) return;
    uint64_t keep = ~0ull;  // freed blocks stay in the pool instead of going back to the driver at every synchronisation
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
}
}  // namespace

struct fb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host-to-device slabs, overlapped with the search of earlier slabs
    cudaStream_t copy_stream2 = nullptr;  // device-to-host copies of finished inflate batches
    cudaEvent_t slab_ev[16] = {};
    uint64_t launches = 0;
    PhaseTimer timer;
    int parse_mode = 0;              // 0 = sparse parse with dense fallback, 1 = dense tables always
    uint64_t sparse_fallbacks = 0;   // streams redone with the dense tables
    uint64_t sparse_repairs = 0;     // streams whose sparse parse was completed by evaluating some chunks densely
    // LZ77 workspace
    DevBuf<uint16_t> link, exits, gexits, gentry, entry, jumps;
    DevBuf<uint32_t> r_full, r_quarter, nx, bitmap, chunk_tokens, tok_offset, tokens, cut_rp;
    DevBuf<uint32_t> chunk_fail, chunk_list;  // sparse parse: coverage check per chunk, chunks to repair
    // block writer workspace
    DevBuf<BlockPlan> plans;
    DevBuf<BlockDesc> descs;
    DevBuf<uint32_t> lit_freq, dist_freq;
    DevBuf<uint8_t> huff_split;  // sorted lists and per-length counts between the passes of the split code construction
    // staging
    DevBuf<uint8_t> d_in, d_out;
    // device scalars: u32[0] total_tokens, u32[1] nblocks; u64 at byte 8: total_bits, nblocks, kPackParts-1 part
    // offsets (written by scan_block_offsets); u32 at byte 64: container checksum; u64[2] at byte 96: Adler scratch
    uint32_t* d_scalars = nullptr;
    uint64_t* h_scalars = nullptr;  // pinned: [0] total_bits [1] nblocks [2..] part offsets; [8] checksum
    // inflate workspace
    DevBuf<uint64_t> m_desc;        // member descriptors / results
    DevBuf<uint8_t> m_scratch;      // work counter + per-CTA match queues of the member-parallel inflate kernel
    int sm_count = 148;
    // One call at a time works on a context's buffers: every entry point that uses them takes this lock, and so does
    // the worker thread of a streaming compressor while a part of its stream is being compressed.
    std::recursive_mutex mu;
    // block-range sharded simple stream: state between the plan and the pack stage
    const uint8_t* shard_in = nullptr;
    uint32_t shard_blocks = 0;
    std::vector<uint64_t> h_members;
};

static inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" {

int fb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
const char* fb200_last_cuda_error(void) { return g_cuda_err; }
const char* fb200_strerror(int code) {
    static const char* names[] = {"Ok", "EndOfStream", "InvalidCode", "InvalidMatch", "InvalidBlockType",
                                  "WrongStoredBlockNlen", "InvalidDynamicBlockHeader", "OversubscribedHuffmanTree",
                                  "IncompleteHuffmanTree", "MissingEndOfBlockCode", "BadGzipHeader", "BadZlibHeader",
                                  "WrongGzipChecksum", "WrongGzipSize", "WrongZlibChecksum", "UnfinishedBits",
                                  "InvalidState", "NoSpaceLeft", "InvalidArgument", "CudaError", "NoDevice", "RetryDense"};
    if (code < 0 || code > 21) return "Unknown";
    return names[code];
}
int fb200_ctx_set_parse_mode(fb200_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 1) return FB200_INVALID_ARGUMENT;
    ctx->parse_mode = mode;
    return FB200_OK;
}
uint64_t fb200_sparse_fallbacks(const fb200_ctx* ctx) { return ctx ? ctx->sparse_fallbacks : 0; }
uint64_t fb200_sparse_repairs(const fb200_ctx* ctx) { return ctx ? ctx->sparse_repairs : 0; }
uint64_t fb200_kernel_launches(const fb200_ctx* ctx) { return ctx ? ctx->launches : 0; }
int fb200_profile_enable(fb200_ctx* ctx, int on) {
    if (!ctx) return FB200_INVALID_ARGUMENT;
    ctx->timer.on = on != 0;
    for (int i = 0; i < kPhCount; i++) ctx->timer.ms[i] = 0, ctx->timer.count[i] = 0;
    return FB200_OK;
}
int fb200_profile_phases(void) { return kPhCount; }
const char* fb200_profile_phase_name(int i) {
    static const char* names[] = {"hash_link", "match_search", "lazy_step(fused)", "lazy+chunk_exit", "resolve_entries", "orbit_mark",
                                  "scan_tokens", "emit_tokens", "plan+histogram", "build_blocks", "offsets+zero", "pack_blocks",
                                  "inflate_members", "sparse_parse"};
    return (i >= 0 && i < kPhCount) ? names[i] : "?";
}
int fb200_profile_read(const fb200_ctx* ctx, double* ms, uint64_t* count, int n) {
    if (!ctx || !ms || !count) return FB200_INVALID_ARGUMENT;
    for (int i = 0; i < n && i < kPhCount; i++) ms[i] = ctx->timer.ms[i], count[i] = ctx->timer.count[i];
    return FB200_OK;
}

int fb200_ctx_create(int device, fb200_ctx** out) {
    if (!out) return FB200_INVALID_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return FB200_NO_DEVICE;
    if (device < 0 || device >= ndev) return FB200_INVALID_ARGUMENT;
    FB_CUDA_CHECK(cudaSetDevice(device));
    fb200_ctx* c = new (std::nothrow) fb200_ctx();
    if (!c) return FB200_INVALID_ARGUMENT;
    c->device = device;
    if (cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c->sm_count <= 0) c->sm_count = 148;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream2, cudaStreamNonBlocking);
    for (auto& ev : c->slab_ev)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_scalars, 256);
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_scalars, 128);
    if (e != cudaSuccess) {  // release whatever was created
        fb::set_last_cuda_error(e, __FILE__, __LINE__);
        fb200_ctx_destroy(c);
        return FB200_ERR_CUDA;
    }
    *out = c;
    return FB200_OK;
}

void fb200_ctx_destroy(fb200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->link.release(); c->exits.release(); c->gexits.release(); c->gentry.release(); c->entry.release();
    c->r_full.release(); c->r_quarter.release(); c->nx.release(); c->bitmap.release(); c->chunk_tokens.release();
    c->tok_offset.release(); c->tokens.release(); c->cut_rp.release(); c->plans.release(); c->descs.release();
    c->lit_freq.release(); c->dist_freq.release(); c->d_in.release(); c->d_out.release(); c->m_desc.release();
    c->jumps.release(); c->chunk_fail.release(); c->chunk_list.release(); c->m_scratch.release(); c->huff_split.release();
    c->timer.destroy();
    if (c->d_scalars) cudaFree(c->d_scalars);
    if (c->h_scalars) cudaFreeHost(c->h_scalars);
    for (auto& e : c->slab_ev)
        if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->copy_stream2) cudaStreamDestroy(c->copy_stream2);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

size_t fb200_compress_bound(size_t n, int mode) {
    (void)mode;
    // stored blocks cost 5 bytes each; an unstorable Huffman block of incompressible bytes < 9/8 n + header;
    // the last term covers the container header and footer of any container (gzip: 10 + 8 bytes)
    return n + (n >> 3) + (n / 32768 + 2) * 640 + 64 + 32;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
static int ensure_lz77(fb200_ctx* c, size_t n) {
    const size_t nchunks = (n + kChunk - 1) / kChunk + 1;
    const size_t ngroups = (nchunks + kGroup - 1) / kGroup + 1;
    FB_CUDA_CHECK(c->link.ensure(n + 64));
    FB_CUDA_CHECK(c->r_full.ensure(n + 64));
    FB_CUDA_CHECK(c->r_quarter.ensure(n + 64));
    FB_CUDA_CHECK(c->nx.ensure(n + 64));
    FB_CUDA_CHECK(c->exits.ensure(nchunks * kEntries));
    FB_CUDA_CHECK(c->gexits.ensure(ngroups * kEntries));
    FB_CUDA_CHECK(c->gentry.ensure(ngroups));
    FB_CUDA_CHECK(c->entry.ensure(nchunks));
    FB_CUDA_CHECK(c->jumps.ensure(nchunks * kChunk));
    FB_CUDA_CHECK(c->bitmap.ensure(nchunks * (kChunk / 32)));
    FB_CUDA_CHECK(c->chunk_tokens.ensure(nchunks));
    FB_CUDA_CHECK(c->tok_offset.ensure(nchunks));
    FB_CUDA_CHECK(c->tokens.ensure(n + 64));
    FB_CUDA_CHECK(c->cut_rp.ensure(n / kTokensPerBlock + 4));
    FB_CUDA_CHECK(c->chunk_fail.ensure(n / lz77_sparse_chunk() + 2));
    FB_CUDA_CHECK(c->chunk_list.ensure(n / lz77_sparse_chunk() + 2));
    return FB200_OK;
}
static int ensure_blocks(fb200_ctx* c, size_t max_blocks) {
    FB_CUDA_CHECK(c->plans.ensure(max_blocks));
    FB_CUDA_CHECK(c->descs.ensure(max_blocks));
    FB_CUDA_CHECK(c->lit_freq.ensure(max_blocks * kNumLit));
    FB_CUDA_CHECK(c->dist_freq.ensure(max_blocks * kNumDist));
    return FB200_OK;
}
static Lz77Buffers lz77_view(fb200_ctx* c) {
    Lz77Buffers b;
    b.link = c->link.p; b.r_full = c->r_full.p; b.r_quarter = c->r_quarter.p; b.nx = c->nx.p;
    b.exits = c->exits.p; b.gexits = c->gexits.p; b.gentry = c->gentry.p; b.entry = c->entry.p; b.jumps = c->jumps.p;
    b.bitmap = c->bitmap.p; b.chunk_tokens = c->chunk_tokens.p; b.tok_offset = c->tok_offset.p;
    b.total_tokens = c->d_scalars; b.tokens = c->tokens.p; b.cut_rp = c->cut_rp.p;
    return b;
}

static const uint8_t kGzipHeader[10] = {0x1f, 0x8b, 0x08, 0, 0, 0, 0, 0, 0, 0x03};  // container.zig:64
static const uint8_t kZlibHeader[2] = {0x78, 0x9c};                                  // container.zig:78
static inline size_t header_size(int container) { return container == FB200_GZIP ? 10 : container == FB200_ZLIB ? 2 : 0; }
static inline size_t footer_size(int container) { return container == FB200_GZIP ? 8 : container == FB200_ZLIB ? 4 : 0; }

// Device flags of the sparse parse (u32 index 20 of d_scalars): bit 0 = an orbit left its span unjoined,
// bit 1 = the token emitter met an entry that was never evaluated.  Non-zero => redo with dense tables.
constexpr int kSparseFlagIdx = 20;
constexpr int kRunCounterIdx = 21;  // run hand-out of the rolling sparse parse
constexpr size_t kSpHaloPlus = 256 + 272;  // lz77_sparse_lookahead() = overlap + lazy halo + compare look-ahead
static int sparse_begin(fb200_ctx* c, const Lz77Buffers& b, size_t count, cudaStream_t st) {
    FB_CUDA_CHECK(cudaMemsetAsync(b.nx, 0xFF, count * sizeof(uint32_t), st));
    FB_CUDA_CHECK(cudaMemsetAsync(c->d_scalars + kSparseFlagIdx, 0, sizeof(uint32_t), st));
    return FB200_OK;
}
// segment [begin, n) of a stream that is already on the device (begin > 0: earlier bytes are history, their links
// are in place; d_skip/nskip: history positions the reference never inserted into its chains)
static int sparse_verdict(fb200_ctx* c, const Lz77Buffers& b, const uint8_t* d_in, size_t begin, size_t n, const LevelArgs& lv,
                          cudaStream_t st, bool* dense);
static int sparse_tokenize(fb200_ctx* c, const Lz77Buffers& b, const uint8_t* d_in, size_t begin, size_t n, const uint32_t* d_skip,
                           uint32_t nskip, const LevelArgs& lv, cudaStream_t st, bool* dense) {
    *dense = false;
    if (n == begin) {
        FB_CUDA_CHECK(lz77_parse_from_nx(b, d_in, 0, lv, st, &c->timer));
        return FB200_OK;
    }
    int rc = sparse_begin(c, b, n - begin, st);
    if (rc) return rc;
    const uint32_t T = lz77_sparse_chunk();
    FB_CUDA_CHECK(lz77_link_range(b, d_in, (uint32_t)begin, (uint32_t)n, (uint32_t)n, st, &c->timer, d_skip, nskip));
    FB_CUDA_CHECK(lz77_sparse_range(b, d_in, (uint32_t)(begin / T), (uint32_t)((n + T - 1) / T), (uint32_t)n, lv, c->chunk_fail.p,
                                    c->d_scalars + kSparseFlagIdx, st, &c->timer, (uint32_t)begin, c->d_scalars + kRunCounterIdx));
    c->launches += 2;
    if ((rc = sparse_verdict(c, b, d_in, begin, n, lv, st, dense)) || *dense) return rc;
    FB_CUDA_CHECK(lz77_parse_from_nx(b, d_in + begin, (uint32_t)(n - begin), lv, st, &c->timer, c->d_scalars + kSparseFlagIdx));
    c->launches += 7;
    return FB200_OK;
}

// The coverage check of some chunks failed (stream synchronized, flag bit 0 set, bit 1 may be a consequence):
// evaluate every position of the chunks that follow a failed chunk; a repaired chunk that fails itself
// hands the problem to its successor.  Afterwards nx is closed under the lazy step again and the caller
// parses once more.  Returns FB200_OK and *ok = false when it gives up (caller redoes the stream densely).
static int sparse_repair(fb200_ctx* c, const Lz77Buffers& b, const uint8_t* d_in, size_t begin, size_t n, const LevelArgs& lv,
                         cudaStream_t st, bool* ok) {
    *ok = false;
    const uint32_t T = lz77_sparse_chunk();
    const uint32_t nch = (uint32_t)((n + T - 1) / T);
    std::vector<uint32_t> fail(nch), list;
    std::vector<uint8_t> dense(nch, 0);
    FB_CUDA_CHECK(cudaMemcpyAsync(fail.data(), c->chunk_fail.p, nch * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    for (uint32_t i = (uint32_t)(begin / T); i + 1 < nch; i++)  // chunks before the segment were not evaluated now
        if (fail[i]) list.push_back(i + 1);
    size_t total = 0;
    for (int round = 0; round < 6 && !list.empty(); round++) {
        total += list.size();
        if (total > nch / 2 + 1) return FB200_OK;  // mostly periodic data: the dense tables are the better tool
        for (uint32_t ch : list) dense[ch] = 1;
        FB_CUDA_CHECK(cudaMemcpyAsync(c->chunk_list.p, list.data(), list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        FB_CUDA_CHECK(lz77_sparse_dense_chunks(b, d_in, c->chunk_list.p, (uint32_t)list.size(), (uint32_t)n, lv, c->chunk_fail.p,
                                               c->d_scalars + kSparseFlagIdx, st, &c->timer, (uint32_t)begin));
        c->launches += 1;
        FB_CUDA_CHECK(cudaMemcpyAsync(fail.data(), c->chunk_fail.p, nch * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        FB_CUDA_CHECK(cudaStreamSynchronize(st));
        std::vector<uint32_t> next;
        for (uint32_t ch : list)
            if (fail[ch] && ch + 1 < nch && !dense[ch + 1]) next.push_back(ch + 1);
        list.swap(next);
    }
    *ok = list.empty();
    return FB200_OK;
}

// The verdict of the coverage check is known as soon as the sparse parse has run (flag bit 0): repair at once, so that
// the parse and the block writer only ever run over a table that is closed under the lazy step.  *dense = true: the
// repair gave up (mostly periodic data), the caller goes to the dense match tables.  Costs the happy path one
// synchronisation (the kernels before and after it are long).
static int sparse_verdict(fb200_ctx* c, const Lz77Buffers& b, const uint8_t* d_in, size_t begin, size_t n, const LevelArgs& lv,
                          cudaStream_t st, bool* dense) {
    *dense = false;
    c->h_scalars[12] = 0;
    FB_CUDA_CHECK(cudaMemcpyAsync(c->h_scalars + 12, c->d_scalars + kSparseFlagIdx, 4, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    if (((uint32_t)c->h_scalars[12] & 1u) == 0) return FB200_OK;
    bool ok = false;
    int rc = sparse_repair(c, b, d_in, begin, n, lv, st, &ok);
    if (rc) return rc;
    if (ok) {
        c->sparse_repairs++;
        FB_CUDA_CHECK(cudaMemsetAsync(c->d_scalars + kSparseFlagIdx, 0, sizeof(uint32_t), st));
    } else {
        c->sparse_fallbacks++;
        *dense = true;
    }
    return FB200_OK;
}

// State a streaming compressor carries from one part of its stream to the next (level modes), all positions relative
// to the device window that d_in points at.  A part starts at a clean arrival of the lazy parse (so its parse starts
// at offset 0), ends on its own 4096-position chunk grid at parse_end, and writes complete 32768-token blocks only:
// the tokens of the block left open stay at the front of the context's token buffer, the output continues at the
// bit where the previous part stopped.
struct PartCarry {
    // in
    uint32_t* nx = nullptr;      // the stream's lazy-step table, entry of position p at nx[p] (kept across parts)
    uint32_t carry_tok = 0;      // tokens of the open block at the front of the token buffer
    bool has_fp = false;         // fp0 valid (else the tokens start a segment: fp = begin)
    uint32_t fp0 = 0;            // position of the last block cut
    uint32_t bit_phase = 0;      // the output starts at this bit of its first byte (bits below it stay zero)
    size_t link_from = SIZE_MAX; // links are in place below this position (SIZE_MAX: from `begin`)
    size_t chunks_from = SIZE_MAX;  // sparse-parse chunks below this one are evaluated (SIZE_MAX: from begin's chunk)
    size_t parse_end = 0;        // 0: flush / finish over [begin, n).  Else a part: tokens of the arrivals in
    size_t chunk_end = 0;        //    [begin, parse_end), sparse chunks up to chunk_end, links up to link_to
    size_t link_to = 0;
    // out
    uint32_t exit = 0;           // first arrival at or past parse_end, relative to it
    uint32_t leftover = 0;       // tokens of the open block (already moved to the front of the token buffer)
    bool cut = false;            // a block was completed ...
    uint32_t last_rp = 0;        // ... and this is the position of the last cut
    uint64_t total_bits = 0;     // end of the written bits, counted from bit 0 of the output's first byte
};

// Runs the deflate body of stream positions [begin, n) on the device.  One-shot calls pass begin = 0
// and get the container header in front; the streaming compressor passes the flush point and gets
// just the blocks of the segment (plus the sync marker when !final_flush).  Returns the end of the
// written bytes (from d_out start).  The checksum of d_in[0..n) is left in c->h_scalars[1].
static int deflate_body_device(fb200_ctx* c, int container, int mode, const uint8_t* d_in, size_t begin, size_t n,
                               const uint32_t* d_skip, uint32_t nskip, uint8_t* d_out, size_t cap, size_t* end_bytes,
                               bool final_flush, bool with_header, cudaStream_t st, const uint8_t* h_src = nullptr,
                               uint8_t* h_dst = nullptr, size_t h_cap = 0, const uint32_t* d_nx_given = nullptr,
                               int redo = 0, PartCarry* carry = nullptr) {
    // carry != nullptr: streaming compressor (see PartCarry); redo == 3 then means "evaluate every position of the
    // chunks concerned" (the coverage check of the sparse parse failed on the first attempt).
    // redo: 1 = links and a repaired nx table are in place, only parse again; 2 = dense match tables
    // d_nx_given != nullptr: the lazy-step table of the whole stream was produced elsewhere (position-sharded
    // search on several GPUs); only the parse and the block writer run here.
    // h_dst != nullptr: the packed bytes are also copied to host memory at h_dst, part by part, while later
    // blocks are still being packed (the caller must not copy them again).
    // h_src != nullptr: the bytes [begin, n) still live in (pinned or pageable) host memory at h_src and are
    // copied to d_in + begin here, in slabs, so that hash links and match search of slab k run while slab k+1
    // is still crossing PCIe.
    if (container < 0 || container > 2 || begin > n) return FB200_INVALID_ARGUMENT;
    if (((uintptr_t)d_out & 15) != 0) return FB200_INVALID_ARGUMENT;
    const size_t hdr = with_header ? header_size(container) : 0;
    if (cap < fb200_compress_bound(n - begin, mode)) return FB200_NO_SPACE_LEFT;  // the bound includes the container overhead
    uint32_t* nblocks_dev = c->d_scalars + 1;
    uint64_t* total_bits_dev = reinterpret_cast<uint64_t*>(c->d_scalars + 2);
    uint32_t max_blocks;
    const uint32_t* tokens = nullptr;
    LevelArgs lv;
    bool sparse = false;
    const bool part = carry && carry->parse_end != 0;
    const uint32_t fmode = part ? 2u : final_flush ? 1u : 0u;  // plan_level_blocks / plan_simple_blocks_kernel
    uint32_t carry_groups = 0;
    if (level_args(mode, lv)) {
        if (n > (1ull << 31)) return FB200_INVALID_ARGUMENT;  // single-stream position space is 32-bit
        int rc = ensure_lz77(c, n);
        if (rc) return rc;
        max_blocks = (uint32_t)((n - begin + (carry ? carry->carry_tok : 0)) / kTokensPerBlock + 3);
        if ((rc = ensure_blocks(c, max_blocks))) return rc;
        Lz77Buffers b = lz77_view(c);
        c->timer.begin(st);
        constexpr size_t kLag = 8192;  // the search of a slab lags one hash tile behind its copy
        // slab-wise input copy: a quarter of the input per slab (4..64 MiB), a small first slab to start early
        size_t kSlab = ((n - begin) / 4 + (1u << 20) - 1) >> 20 << 20;
        kSlab = kSlab < (4u << 20) ? (4u << 20) : kSlab > (64u << 20) ? (64u << 20) : kSlab;
        size_t kFirst = kSlab / 4;
        if (const char* e = getenv("FB200_SLAB")) {  // development knob: FB200_SLAB="first_MiB,slab_MiB"
            int a = 0, b2 = 0;
            if (sscanf(e, "%d,%d", &a, &b2) == 2 && a > 0 && b2 > 0) {
                kFirst = (size_t)a << 20;
                kSlab = (size_t)b2 << 20;
            }
        }
        if (carry) {
            const uint32_t T = lz77_sparse_chunk();
            const size_t pend = part ? carry->parse_end : n;
            const size_t c_from = carry->chunks_from == SIZE_MAX ? begin / T : carry->chunks_from;
            const size_t c_to = part ? carry->chunk_end : (n + T - 1) / T;
            uint32_t* flags = c->d_scalars + kSparseFlagIdx;
            sparse = true;
            b.nx = carry->nx + begin;  // the kernels index the table from the segment start
            FB_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(uint32_t), st));
            if (pend == begin) {
                // an empty segment: only the carried tokens, if any
            } else if (redo != 3) {
                // entries below the previous run's overlap are in place; what this run can write starts out "not evaluated"
                const size_t ms_from = carry->chunks_from == SIZE_MAX ? begin : c_from * T + 1024;
                const size_t ms_to = part ? (c_to * T + 1024 < n ? c_to * T + 1024 : n) : n;
                if (ms_to > ms_from) FB_CUDA_CHECK(cudaMemsetAsync(carry->nx + ms_from, 0xFF, (ms_to - ms_from) * sizeof(uint32_t), st));
                const size_t l_from = carry->link_from == SIZE_MAX ? begin : carry->link_from;
                const size_t l_to = part ? carry->link_to : n;
                if (l_to > l_from) FB_CUDA_CHECK(lz77_link_range(b, d_in, (uint32_t)l_from, (uint32_t)l_to, (uint32_t)n, st, &c->timer, d_skip, nskip));
                if (c_to > c_from)
                    FB_CUDA_CHECK(lz77_sparse_range(b, d_in, (uint32_t)c_from, (uint32_t)c_to, (uint32_t)n, lv, c->chunk_fail.p, flags, st, &c->timer,
                                                    (uint32_t)begin, c->d_scalars + kRunCounterIdx));
                c->launches += 2;
            } else {
                // every position of the chunks the parse walks: nothing is left to the speculation
                std::vector<uint32_t> list;
                for (size_t ch = begin / T; ch < c_to; ch++) list.push_back((uint32_t)ch);
                if (!list.empty()) {
                    FB_CUDA_CHECK(c->chunk_list.ensure(list.size()));
                    FB_CUDA_CHECK(cudaMemcpyAsync(c->chunk_list.p, list.data(), list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
                    FB_CUDA_CHECK(cudaStreamSynchronize(st));  // `list` is pageable and about to go out of scope
                    FB_CUDA_CHECK(lz77_sparse_dense_chunks(b, d_in, c->chunk_list.p, (uint32_t)list.size(), (uint32_t)n, lv, c->chunk_fail.p, flags, st,
                                                           &c->timer, (uint32_t)begin));
                    FB_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(uint32_t), st));  // hand-over failures between dense chunks mean nothing
                }
                c->launches += 1;
            }
            FB_CUDA_CHECK(lz77_parse_from_nx(b, d_in + begin, (uint32_t)(pend - begin), lv, st, &c->timer, flags, carry->carry_tok));
            const uint32_t nchunks = (uint32_t)((pend - begin + kChunk - 1) / kChunk);
            carry_groups = (nchunks + kGroup - 1) / kGroup;
            c->launches += pend > begin ? 7 : 0;
        } else if (redo == 1) {
            sparse = true;
            FB_CUDA_CHECK(cudaMemsetAsync(c->d_scalars + kSparseFlagIdx, 0, sizeof(uint32_t), st));
            FB_CUDA_CHECK(lz77_parse_from_nx(b, d_in + begin, (uint32_t)(n - begin), lv, st, &c->timer, c->d_scalars + kSparseFlagIdx));
            c->launches += 7;
        } else if (d_nx_given) {
            b.nx = const_cast<uint32_t*>(d_nx_given);
            FB_CUDA_CHECK(cudaMemsetAsync(c->d_scalars + kSparseFlagIdx, 0, sizeof(uint32_t), st));
            FB_CUDA_CHECK(lz77_parse_from_nx(b, d_in, (uint32_t)n, lv, st, &c->timer, c->d_scalars + kSparseFlagIdx));
            c->launches += n ? 7 : 0;
        } else if ((sparse = (n > begin && redo == 0 && c->parse_mode == 0)) && !(h_src && begin == 0 && n > kSlab + kLag)) {
            if (h_src) FB_CUDA_CHECK(cudaMemcpyAsync(const_cast<uint8_t*>(d_in) + begin, h_src, n - begin, cudaMemcpyHostToDevice, st));
            bool dense = false;
            if ((rc = sparse_tokenize(c, b, d_in, begin, n, d_skip, nskip, lv, st, &dense))) return rc;
            if (dense)  // mostly periodic data: the dense tables are the better tool (nothing was parsed or packed yet)
                return deflate_body_device(c, container, mode, d_in, begin, n, d_skip, nskip, d_out, cap, end_bytes, final_flush, with_header,
                                           st, nullptr, h_dst, h_cap, nullptr, 2);
        } else if (sparse) {
            // slab-overlapped copy: links follow the copy front one hash tile behind, the sparse parse follows the
            // links by its look-ahead
            if ((rc = sparse_begin(c, b, n, st))) return rc;
            const uint32_t T = lz77_sparse_chunk(), ahead = lz77_sparse_link_ahead();
            size_t copied = 0, linked = 0;
            uint32_t chunks_done = 0;
            int k = 0;
            while (copied < n) {
                const size_t slab_end = copied == 0 ? (kFirst < n ? kFirst : n) : (copied + kSlab < n ? copied + kSlab : n);
                cudaEvent_t ev = c->slab_ev[k % 16];
                if (k >= 16) FB_CUDA_CHECK(cudaEventSynchronize(ev));
                FB_CUDA_CHECK(cudaMemcpyAsync(const_cast<uint8_t*>(d_in) + copied, h_src + copied, slab_end - copied,
                                              cudaMemcpyHostToDevice, c->copy_stream));
                FB_CUDA_CHECK(cudaEventRecord(ev, c->copy_stream));
                FB_CUDA_CHECK(cudaStreamWaitEvent(st, ev, 0));
                const size_t range_end = slab_end == n ? n : slab_end - kLag;
                FB_CUDA_CHECK(lz77_link_range(b, d_in, (uint32_t)linked, (uint32_t)range_end, (uint32_t)n, st, &c->timer));
                const uint32_t chunks_end = range_end == n ? (uint32_t)((n + T - 1) / T)
                                                           : (uint32_t)(range_end > ahead ? (range_end - ahead) / T : 0);
                if (chunks_end > chunks_done) {
                    FB_CUDA_CHECK(lz77_sparse_range(b, d_in, chunks_done, chunks_end, (uint32_t)n, lv, c->chunk_fail.p, c->d_scalars + kSparseFlagIdx, st, &c->timer, 0,
                                                    c->d_scalars + kRunCounterIdx));
                    chunks_done = chunks_end;
                }
                c->launches += 2;
                linked = range_end;
                copied = slab_end;
                k++;
            }
            bool dense = false;
            if ((rc = sparse_verdict(c, b, d_in, 0, n, lv, st, &dense))) return rc;
            if (dense)
                return deflate_body_device(c, container, mode, d_in, begin, n, d_skip, nskip, d_out, cap, end_bytes, final_flush, with_header,
                                           st, nullptr, h_dst, h_cap, nullptr, 2);
            FB_CUDA_CHECK(lz77_parse_from_nx(b, d_in, (uint32_t)n, lv, st, &c->timer, c->d_scalars + kSparseFlagIdx));
            c->launches += 7;
        } else if (h_src && n - begin > kSlab + kLag) {
            const size_t first_end = (begin / kFirst + 1) * kFirst;  // a small first slab shortens the initial wait
            size_t copied = begin, searched = begin;
            int k = 0;
            while (copied < n) {
                const size_t slab_end = copied == begin ? (first_end < n ? first_end : n) : (copied + kSlab < n ? copied + kSlab : n);
                cudaEvent_t ev = c->slab_ev[k % 16];
                if (k >= 16) FB_CUDA_CHECK(cudaEventSynchronize(ev));
                FB_CUDA_CHECK(cudaMemcpyAsync(const_cast<uint8_t*>(d_in) + copied, h_src + (copied - begin), slab_end - copied,
                                              cudaMemcpyHostToDevice, c->copy_stream));
                FB_CUDA_CHECK(cudaEventRecord(ev, c->copy_stream));
                FB_CUDA_CHECK(cudaStreamWaitEvent(st, ev, 0));
                const size_t range_end = slab_end == n ? n : slab_end - kLag;
                FB_CUDA_CHECK(lz77_search_range(b, d_in, (uint32_t)begin, (uint32_t)searched, (uint32_t)range_end, (uint32_t)n, d_skip,
                                                nskip, lv, st, &c->timer));
                c->launches += 2;
                searched = range_end;
                copied = slab_end;
                k++;
            }
            FB_CUDA_CHECK(lz77_parse(b, d_in, (uint32_t)begin, (uint32_t)n, lv, st, &c->timer));
            c->launches += 7;
        } else {
            if (h_src && n > begin) FB_CUDA_CHECK(cudaMemcpyAsync(const_cast<uint8_t*>(d_in) + begin, h_src, n - begin, cudaMemcpyHostToDevice, st));
            FB_CUDA_CHECK(lz77_tokenize(b, d_in, (uint32_t)begin, (uint32_t)n, d_skip, nskip, lv, st, &c->timer));
            c->launches += n > begin ? 9 : 0;
        }
        FB_CUDA_CHECK(plan_level_blocks(b.total_tokens, b.cut_rp, (uint32_t)begin, (uint32_t)n, max_blocks, fmode, c->plans.p, nblocks_dev, st,
                                        carry && carry->has_fp ? carry->fp0 : (uint32_t)begin));
        FB_CUDA_CHECK(histogram_tokens(b.tokens, c->plans.p, nblocks_dev, max_blocks, c->lit_freq.p, c->dist_freq.p, st));
        c->launches += 2;
        c->timer.mark(st, kPhHist);
        tokens = b.tokens;
    } else if (mode == FB200_MODE_HUFFMAN || mode == FB200_MODE_STORE) {
        // a part of a stream is a whole number of slices; flush and finish close the segment with its last, shorter
        // (possibly empty) slice
        const uint64_t nb64 = part ? (n - begin) / kMaxStore : (n - begin) / kMaxStore + 1;
        if (nb64 > 0x7ffffff0ull || (part && (nb64 == 0 || (n - begin) % kMaxStore))) return FB200_INVALID_ARGUMENT;
        const uint32_t nslices = (uint32_t)nb64;
        max_blocks = nslices + (fmode ? 0 : 1);
        int rc = ensure_blocks(c, max_blocks);
        if (rc) return rc;
        if (h_src && n > begin) FB_CUDA_CHECK(cudaMemcpyAsync(const_cast<uint8_t*>(d_in) + begin, h_src, n - begin, cudaMemcpyHostToDevice, st));
        c->timer.begin(st);
        plan_simple_blocks_kernel<<<(max_blocks + 255) / 256, 256, 0, st>>>(begin, n, nslices,
                                                                            mode == FB200_MODE_HUFFMAN ? kHuffmanBlock : 3u,
                                                                            fmode, c->plans.p, nblocks_dev);
        FB_CUDA_CHECK(cudaGetLastError());
        c->launches += 1;
        if (mode == FB200_MODE_HUFFMAN) {
            FB_CUDA_CHECK(histogram_bytes(d_in, c->plans.p, max_blocks, c->lit_freq.p, st));
            c->launches += 1;
            c->timer.mark(st, kPhHist);
        }
    } else {
        return FB200_INVALID_ARGUMENT;
    }
    void* split = nullptr;
    if (mode == FB200_MODE_HUFFMAN && max_blocks >= 64) {
        FB_CUDA_CHECK(c->huff_split.ensure(build_blocks_split_bytes(max_blocks)));
        split = c->huff_split.p;
    }
    FB_CUDA_CHECK(build_blocks(c->plans.p, nblocks_dev, max_blocks, c->lit_freq.p, c->dist_freq.p, c->descs.p, st, split));
    c->launches += split ? 2 : 0;
    c->timer.mark(st, kPhBuild);
    FB_CUDA_CHECK(scan_block_offsets(c->descs.p, nblocks_dev, hdr * 8 + (carry ? carry->bit_phase : 0), total_bits_dev, st));
    zero_output_kernel<<<148 * 4, 256, 0, st>>>(reinterpret_cast<uint32_t*>(d_out), total_bits_dev, cap / 4);
    FB_CUDA_CHECK(cudaGetLastError());
    if (hdr && container == FB200_GZIP) FB_CUDA_CHECK(cudaMemcpyAsync(d_out, kGzipHeader, 10, cudaMemcpyHostToDevice, st));
    if (hdr && container == FB200_ZLIB) FB_CUDA_CHECK(cudaMemcpyAsync(d_out, kZlibHeader, 2, cudaMemcpyHostToDevice, st));
    c->timer.mark(st, kPhOffsets);
    bool copied_out = false;
    if (h_dst) {
        // one small read-back tells the host the block count, the total size and the part boundaries
        FB_CUDA_CHECK(cudaMemcpyAsync(c->h_scalars, total_bits_dev, 8 * (2 + kPackParts - 1), cudaMemcpyDeviceToHost, st));
        c->h_scalars[9] = 0;
        if (sparse || d_nx_given) FB_CUDA_CHECK(cudaMemcpyAsync(c->h_scalars + 9, c->d_scalars + kSparseFlagIdx, 4, cudaMemcpyDeviceToHost, st));
        FB_CUDA_CHECK(cudaStreamSynchronize(st));
        const uint64_t total_bits = c->h_scalars[0];
        const uint32_t nb = (uint32_t)c->h_scalars[1];
        const size_t out_bytes = (size_t)((total_bits + 7) >> 3);
        // a failed speculation is known by now (both flag bits are set before the block writer runs): nothing of
        // this attempt is packed or copied back, the repair / dense redo below produces the stream
        const bool speculation_failed = (uint32_t)c->h_scalars[9] != 0;
        if (!speculation_failed && out_bytes + footer_size(container) > h_cap) return FB200_NO_SPACE_LEFT;
        size_t byte_lo = 0;
        for (uint32_t i = 0; i < kPackParts && !speculation_failed; i++) {
            const uint32_t b_lo = pack_part_begin(nb, i), b_hi = pack_part_begin(nb, i + 1);
            FB_CUDA_CHECK(pack_blocks_range(d_in, tokens, c->descs.p, nblocks_dev, b_lo, b_hi - b_lo, reinterpret_cast<uint32_t*>(d_out), st));
            // bytes below the next part's first bit are final once this part is packed (a shared byte goes with the next part)
            const size_t byte_hi = i + 1 == kPackParts ? out_bytes : (size_t)(c->h_scalars[2 + i] >> 3);
            FB_CUDA_CHECK(cudaEventRecord(c->slab_ev[i], st));
            FB_CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->slab_ev[i], 0));
            if (byte_hi > byte_lo)
                FB_CUDA_CHECK(cudaMemcpyAsync(h_dst + byte_lo, d_out + byte_lo, byte_hi - byte_lo, cudaMemcpyDeviceToHost, c->copy_stream));
            byte_lo = byte_hi > byte_lo ? byte_hi : byte_lo;
        }
        copied_out = true;
    } else {
        FB_CUDA_CHECK(pack_blocks(d_in, tokens, c->descs.p, nblocks_dev, max_blocks, reinterpret_cast<uint32_t*>(d_out), st));
    }
    c->timer.mark(st, kPhPack);
    c->launches += 4;
    // container checksum of the plain bytes on the device (container.zig:168-206), SURVEY.md §8(f) rank 1
    uint32_t* sum_dev = c->d_scalars + 16;
    if (!final_flush || carry) {
        // the footer is only written by finish(); a streaming compressor sums its stream part by part itself
    } else if (container == FB200_GZIP) {
        FB_CUDA_CHECK(crc32_device(d_in, n, sum_dev, st));
        c->launches += n ? 1 : 0;
    } else if (container == FB200_ZLIB) {
        FB_CUDA_CHECK(adler32_device(d_in, n, sum_dev, reinterpret_cast<uint64_t*>(c->d_scalars + 24), st));
        c->launches += n ? 2 : 1;
    }
    if (container != FB200_RAW && final_flush && !carry) FB_CUDA_CHECK(cudaMemcpyAsync(c->h_scalars + 8, sum_dev, 4, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaMemcpyAsync(c->h_scalars, total_bits_dev, 8, cudaMemcpyDeviceToHost, st));
    if (sparse || d_nx_given) FB_CUDA_CHECK(cudaMemcpyAsync(c->h_scalars + 9, c->d_scalars + kSparseFlagIdx, 4, cudaMemcpyDeviceToHost, st));
    if (carry) {  // token count | block count, and where the orbit left the part
        FB_CUDA_CHECK(cudaMemcpyAsync(c->h_scalars + 10, c->d_scalars, 8, cudaMemcpyDeviceToHost, st));
        c->h_scalars[11] = 0;
        if (tokens) FB_CUDA_CHECK(cudaMemcpyAsync(c->h_scalars + 11, c->gentry.p + carry_groups, 2, cudaMemcpyDeviceToHost, st));
    }
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    if (copied_out) FB_CUDA_CHECK(cudaStreamSynchronize(c->copy_stream));
    c->timer.collect();
    if (carry && sparse && (uint32_t)c->h_scalars[9] != 0) {
        if (redo == 3) return FB200_RETRY_DENSE;  // cannot happen: every entry the parse can meet was evaluated
        c->sparse_repairs++;
        return deflate_body_device(c, container, mode, d_in, begin, n, d_skip, nskip, d_out, cap, end_bytes, final_flush, with_header, st,
                                   nullptr, nullptr, 0, nullptr, 3, carry);
    }
    if (carry) {
        carry->total_bits = c->h_scalars[0];
        const uint32_t ntok = (uint32_t)c->h_scalars[10], nb = (uint32_t)(c->h_scalars[10] >> 32);
        carry->exit = (uint32_t)(c->h_scalars[11] & 0xffffu);
        carry->leftover = 0;
        carry->cut = false;
        if (part && tokens) {
            carry->leftover = ntok - nb * kTokensPerBlock;
            if (nb) {
                carry->cut = true;
                FB_CUDA_CHECK(cudaMemcpyAsync(&carry->last_rp, c->cut_rp.p + nb - 1, 4, cudaMemcpyDeviceToHost, st));
                // the open block's tokens move to the front (no overlap: fewer than one block's worth, from at least one block up)
                if (carry->leftover)
                    FB_CUDA_CHECK(cudaMemcpyAsync(c->tokens.p, c->tokens.p + (size_t)nb * kTokensPerBlock, (size_t)carry->leftover * 4,
                                                  cudaMemcpyDeviceToDevice, st));
                FB_CUDA_CHECK(cudaStreamSynchronize(st));
                carry->last_rp += (uint32_t)begin;  // cut_rp is relative to the segment start
            }
        }
        *end_bytes = (size_t)((carry->total_bits + 7) >> 3);
        return FB200_OK;
    }
    if (d_nx_given && (uint32_t)c->h_scalars[9] != 0) return FB200_RETRY_DENSE;  // the given table does not cover the orbit
    if (sparse && (uint32_t)c->h_scalars[9] != 0) {
        // The speculation did not cover the true orbit (periodic data, where parses started at different
        // positions never fall into step).  First try to repair the chunks concerned; if that is not enough
        // redo the stream with the dense match tables.
        bool ok = false;
        if (redo == 0) {
            Lz77Buffers b = lz77_view(c);
            int rc = sparse_repair(c, b, d_in, begin, n, lv, st, &ok);
            if (rc) return rc;
        }
        if (ok) c->sparse_repairs++;
        else c->sparse_fallbacks++;
        return deflate_body_device(c, container, mode, d_in, begin, n, d_skip, nskip, d_out, cap, end_bytes, final_flush, with_header,
                                   st, nullptr, h_dst, h_cap, nullptr, ok ? 1 : 2);
    }
    *end_bytes = (size_t)((c->h_scalars[0] + 7) >> 3);
    return FB200_OK;
}

// container footer (container.zig:85-109) from the device-computed checksum
static size_t make_footer(int container, uint32_t sum, size_t n, uint8_t* f) {
    if (container == FB200_GZIP) {
        const uint32_t c = sum, sz = (uint32_t)n;
        for (int i = 0; i < 4; i++) f[i] = (uint8_t)(c >> (8 * i)), f[4 + i] = (uint8_t)(sz >> (8 * i));
        return 8;
    }
    if (container == FB200_ZLIB) {
        const uint32_t c = sum;
        for (int i = 0; i < 4; i++) f[i] = (uint8_t)(c >> (24 - 8 * i));
        return 4;
    }
    return 0;
}

extern "C" {

int fb200_compress_device(fb200_ctx* c, int container, int mode, const void* d_in, size_t n, void* d_out, size_t cap,
                          size_t* out_len, void* stream) {
    if (!c || !out_len || (!d_in && n)) return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    size_t end = 0;
    int rc = deflate_body_device(c, container, mode, (const uint8_t*)d_in, 0, n, nullptr, 0, (uint8_t*)d_out, cap, &end, true, true, st);
    if (rc) return rc;
    uint8_t footer[8];
    const size_t flen = make_footer(container, (uint32_t)c->h_scalars[8], n, footer);
    if (flen) FB_CUDA_CHECK(cudaMemcpyAsync((uint8_t*)d_out + end, footer, flen, cudaMemcpyHostToDevice, st));
    if (flen) FB_CUDA_CHECK(cudaStreamSynchronize(st));
    *out_len = end + flen;
    return FB200_OK;
}

int fb200_compress(fb200_ctx* c, int container, int mode, const uint8_t* in, size_t n, uint8_t* out, size_t cap,
                   size_t* out_len) {
    if (!c || !out_len || (!in && n) || !out) return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    const size_t bound = fb200_compress_bound(n, mode) + 32;
    FB_CUDA_CHECK(c->d_in.ensure(n + 512));
    FB_CUDA_CHECK(c->d_out.ensure(round_up(bound, 16)));
    cudaStream_t st = c->stream;
    size_t end = 0;
    int rc = deflate_body_device(c, container, mode, c->d_in.p, 0, n, nullptr, 0, c->d_out.p, c->d_out.cap, &end, true, true, st, in,
                                 out, cap);
    if (rc) return rc;
    uint8_t footer[8];
    const size_t flen = make_footer(container, (uint32_t)c->h_scalars[8], n, footer);
    if (end + flen > cap) return FB200_NO_SPACE_LEFT;
    memcpy(out + end, footer, flen);
    *out_len = end + flen;
    return FB200_OK;
}

// ---- position-sharded single stream ----
size_t fb200_shard_overlap(void) { return lz77_sparse_lookahead() - kSpHaloPlus; }
size_t fb200_shard_align(void) { return lz77_sparse_chunk(); }
int fb200_deflate_shard_search(fb200_ctx* c, int level, const void* d_in, size_t n, size_t from, size_t to, void* d_nx,
                               void* stream) {
    LevelArgs lv;
    if (!c || !level_args(level, lv) || !d_nx || n > (1ull << 31) || from > to || to > n || (from % 8192) != 0)
        return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    if (to == from) return FB200_OK;
    const size_t T = lz77_sparse_chunk(), W = fb200_shard_overlap();
    Lz77Buffers b = lz77_view(c);
    if (c->parse_mode == 0 && from % T == 0 && (to % T == 0 || to == n)) {
        // sparse parse of the range's chunks; entries of [to, to + W) belong to the overlap of the last chunk
        FB_CUDA_CHECK(c->link.ensure(n + 64));
        FB_CUDA_CHECK(c->chunk_fail.ensure(n / T + 2));
        b = lz77_view(c);
        b.nx = reinterpret_cast<uint32_t*>(d_nx);
        c->timer.begin(st);
        const size_t nx_end = to + W < n ? to + W : n;
        FB_CUDA_CHECK(cudaMemsetAsync(b.nx + from, 0xFF, (nx_end - from) * sizeof(uint32_t), st));
        FB_CUDA_CHECK(cudaMemsetAsync(c->d_scalars + kSparseFlagIdx, 0, sizeof(uint32_t), st));
        const size_t need = to + lz77_sparse_link_ahead();
        const size_t link_end = need >= n ? n : (need + 8191) / 8192 * 8192 < n ? (need + 8191) / 8192 * 8192 : n;
        FB_CUDA_CHECK(lz77_link_range(b, (const uint8_t*)d_in, (uint32_t)(from >= kHist ? from - kHist : 0), (uint32_t)link_end,
                                      (uint32_t)n, st, &c->timer));
        FB_CUDA_CHECK(lz77_sparse_range(b, (const uint8_t*)d_in, (uint32_t)(from / T), (uint32_t)((to + T - 1) / T), (uint32_t)n, lv,
                                        c->chunk_fail.p, c->d_scalars + kSparseFlagIdx, st, &c->timer, 0, c->d_scalars + kRunCounterIdx));
        c->launches += 2;
        uint32_t bad = 0;
        FB_CUDA_CHECK(cudaMemcpyAsync(&bad, c->d_scalars + kSparseFlagIdx, 4, cudaMemcpyDeviceToHost, st));
        FB_CUDA_CHECK(cudaStreamSynchronize(st));
        c->timer.collect();
        return bad ? FB200_RETRY_DENSE : FB200_OK;
    }
    const size_t span = to - from + 256 + 8192 + 64;
    FB_CUDA_CHECK(c->link.ensure(n + 64));
    FB_CUDA_CHECK(c->r_full.ensure(span));
    FB_CUDA_CHECK(c->r_quarter.ensure(span));
    b = lz77_view(c);
    c->timer.begin(st);
    FB_CUDA_CHECK(lz77_shard_search(b, (const uint8_t*)d_in, (uint32_t)from, (uint32_t)to, (uint32_t)n, lv,
                                    reinterpret_cast<uint32_t*>(d_nx) + from, st, &c->timer));
    c->launches += 3;
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    c->timer.collect();
    return FB200_OK;
}
int fb200_deflate_shard_finish(fb200_ctx* c, int container, int level, const void* d_in, size_t n, const void* d_nx,
                               void* d_out, size_t cap, size_t* out_len, void* stream) {
    LevelArgs lv;
    if (!c || !level_args(level, lv) || !out_len || (!d_nx && n) || n > (1ull << 31)) return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    size_t end = 0;
    static const uint32_t dummy_nx = 0;
    int rc = deflate_body_device(c, container, level, (const uint8_t*)d_in, 0, n, nullptr, 0, (uint8_t*)d_out, cap, &end, true, true,
                                 st, nullptr, nullptr, 0, n ? (const uint32_t*)d_nx : &dummy_nx);
    if (rc) return rc;
    uint8_t footer[8];
    const size_t flen = make_footer(container, (uint32_t)c->h_scalars[8], n, footer);
    if (flen) FB_CUDA_CHECK(cudaMemcpyAsync((uint8_t*)d_out + end, footer, flen, cudaMemcpyHostToDevice, st));
    if (flen) FB_CUDA_CHECK(cudaStreamSynchronize(st));
    *out_len = end + flen;
    return FB200_OK;
}

// ---- huffman-only / store stream sharded by 65535-byte block ranges ----
int fb200_simple_shard_plan(fb200_ctx* c, int container, int mode, const void* d_in, size_t shard_bytes, int is_last,
                            uint64_t* pre_bits, int* has_stored, uint64_t* post_bits, uint32_t* checksum, void* stream) {
    if (!c || (mode != FB200_MODE_HUFFMAN && mode != FB200_MODE_STORE) || container < 0 || container > 2 || (!d_in && shard_bytes) ||
        !pre_bits || !has_stored || !post_bits)
        return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    if (!is_last && (shard_bytes == 0 || shard_bytes % kMaxStore != 0)) return FB200_INVALID_ARGUMENT;  // whole slices only
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const uint64_t nb64 = shard_bytes / kMaxStore + (is_last ? 1 : 0);  // deflate.zig:456: the last slice may be empty
    if (nb64 == 0 || nb64 > 0x7ffffff0ull) return FB200_INVALID_ARGUMENT;
    const uint32_t nslices = (uint32_t)nb64;
    int rc = ensure_blocks(c, nslices);
    if (rc) return rc;
    uint32_t* nblocks_dev = c->d_scalars + 1;
    uint64_t* total_bits_dev = reinterpret_cast<uint64_t*>(c->d_scalars + 2);
    c->timer.begin(st);
    plan_simple_blocks_kernel<<<(nslices + 255) / 256, 256, 0, st>>>(0, shard_bytes, nslices, mode == FB200_MODE_HUFFMAN ? kHuffmanBlock : 3u,
                                                                     is_last ? 1 : 2, c->plans.p, nblocks_dev);
    FB_CUDA_CHECK(cudaGetLastError());
    c->launches += 1;
    if (mode == FB200_MODE_HUFFMAN) {
        FB_CUDA_CHECK(histogram_bytes((const uint8_t*)d_in, c->plans.p, nslices, c->lit_freq.p, st));
        c->launches += 1;
        c->timer.mark(st, kPhHist);
    }
    void* split = nullptr;
    if (mode == FB200_MODE_HUFFMAN && nslices >= 64) {
        FB_CUDA_CHECK(c->huff_split.ensure(build_blocks_split_bytes(nslices)));
        split = c->huff_split.p;
    }
    FB_CUDA_CHECK(build_blocks(c->plans.p, nblocks_dev, nslices, c->lit_freq.p, c->dist_freq.p, c->descs.p, st, split));
    c->timer.mark(st, kPhBuild);
    FB_CUDA_CHECK(scan_block_offsets(c->descs.p, nblocks_dev, 0, total_bits_dev, st));
    c->launches += 2;
    uint32_t* sum_dev = c->d_scalars + 16;
    if (container == FB200_GZIP) {
        FB_CUDA_CHECK(crc32_device((const uint8_t*)d_in, shard_bytes, sum_dev, st));
        c->launches += shard_bytes ? 1 : 0;
    } else if (container == FB200_ZLIB) {
        FB_CUDA_CHECK(adler32_device((const uint8_t*)d_in, shard_bytes, sum_dev, reinterpret_cast<uint64_t*>(c->d_scalars + 24), st));
        c->launches += shard_bytes ? 2 : 1;
    }
    uint64_t summary[3] = {0, 0, 0};
    uint32_t sum = 0;
    FB_CUDA_CHECK(cudaMemcpyAsync(summary, total_bits_dev + 16, sizeof summary, cudaMemcpyDeviceToHost, st));
    if (container != FB200_RAW) FB_CUDA_CHECK(cudaMemcpyAsync(&sum, sum_dev, 4, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    c->timer.collect();
    *pre_bits = summary[0];
    *has_stored = (int)summary[1];
    *post_bits = summary[2];
    if (checksum) *checksum = sum;
    c->shard_in = (const uint8_t*)d_in;
    c->shard_blocks = nslices;
    return FB200_OK;
}

int fb200_simple_shard_pack(fb200_ctx* c, uint64_t start_bit, void* d_out, size_t cap, uint64_t* byte_lo, size_t* nbytes,
                            uint64_t* end_bit, void* stream) {
    if (!c || !d_out || !byte_lo || !nbytes || !end_bit || c->shard_blocks == 0 || ((uintptr_t)d_out & 15) != 0) return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const uint64_t lo = (start_bit >> 3) & ~15ull;  // stream byte held by d_out[0]; 16-byte steps keep word alignment
    const uint64_t local_start = start_bit - 8 * lo;
    uint32_t* nblocks_dev = c->d_scalars + 1;
    uint64_t* total_bits_dev = reinterpret_cast<uint64_t*>(c->d_scalars + 2);
    c->timer.begin(st);
    FB_CUDA_CHECK(scan_block_offsets(c->descs.p, nblocks_dev, local_start, total_bits_dev, st));
    zero_output_kernel<<<148 * 4, 256, 0, st>>>(reinterpret_cast<uint32_t*>(d_out), total_bits_dev, cap / 4);
    FB_CUDA_CHECK(cudaGetLastError());
    c->timer.mark(st, kPhOffsets);
    // the size is known from the plan stage only up to the alignment of the first stored block: check the capacity
    // against the scan's result before packing
    uint64_t total_bits = 0;
    FB_CUDA_CHECK(cudaMemcpyAsync(&total_bits, total_bits_dev, 8, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    if (((total_bits + 31) >> 5) * 4 + 8 > cap) return FB200_NO_SPACE_LEFT;
    FB_CUDA_CHECK(pack_blocks(c->shard_in, nullptr, c->descs.p, nblocks_dev, c->shard_blocks, reinterpret_cast<uint32_t*>(d_out), st));
    c->timer.mark(st, kPhPack);
    c->launches += 3;
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    c->timer.collect();
    *byte_lo = lo;
    *nbytes = (size_t)((total_bits + 7) >> 3);
    *end_bit = 8 * lo + total_bits;
    return FB200_OK;
}

// zlib's crc32_combine / adler32_combine: checksum of A||B from the checksums of A and B and the length of B
static uint32_t gf2_times(const uint32_t* mat, uint32_t vec) {
    uint32_t sum = 0;
    for (int i = 0; vec; vec >>= 1, i++)
        if (vec & 1) sum ^= mat[i];
    return sum;
}
static void gf2_square(uint32_t* sq, const uint32_t* mat) {
    for (int n = 0; n < 32; n++) sq[n] = gf2_times(mat, mat[n]);
}
uint32_t fb200_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
    if (len2 == 0) return crc1;
    uint32_t even[32], odd[32];
    odd[0] = 0xEDB88320u;  // one zero bit
    uint32_t row = 1;
    for (int n = 1; n < 32; n++) { odd[n] = row; row <<= 1; }
    gf2_square(even, odd);  // two zero bits
    gf2_square(odd, even);  // four
    do {
        gf2_square(even, odd);
        if (len2 & 1) crc1 = gf2_times(even, crc1);
        len2 >>= 1;
        if (len2 == 0) break;
        gf2_square(odd, even);
        if (len2 & 1) crc1 = gf2_times(odd, crc1);
        len2 >>= 1;
    } while (len2 != 0);
    return crc1 ^ crc2;
}
uint32_t fb200_adler32_combine(uint32_t adler1, uint32_t adler2, uint64_t len2) {
    const uint32_t BASE = 65521;
    const uint32_t rem = (uint32_t)(len2 % BASE);
    uint32_t sum1 = adler1 & 0xffff;
    uint32_t sum2 = (uint32_t)(((uint64_t)rem * sum1) % BASE);
    sum1 += (adler2 & 0xffff) + BASE - 1;
    sum2 += ((adler1 >> 16) & 0xffff) + ((adler2 >> 16) & 0xffff) + BASE - rem;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum2 >= (BASE << 1)) sum2 -= (BASE << 1);
    if (sum2 >= BASE) sum2 -= BASE;
    return sum1 | (sum2 << 16);
}

// ---- test seams ----
int fb200_debug_tokens(fb200_ctx* c, int level, const uint8_t* in, size_t n, uint32_t* tokens, size_t cap, size_t* ntok) {
    LevelArgs lv;
    if (!c || !level_args(level, lv) || !ntok || n > (1ull << 31)) return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    FB_CUDA_CHECK(c->d_in.ensure(n + 512));
    int rc = ensure_lz77(c, n);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    if (n) FB_CUDA_CHECK(cudaMemcpyAsync(c->d_in.p, in, n, cudaMemcpyHostToDevice, st));
    Lz77Buffers b = lz77_view(c);
    uint32_t total = 0, bad = 0;
    if (c->parse_mode == 0 && n) {
        bool dense = false;  // the repair inside already gave up
        if ((rc = sparse_tokenize(c, b, c->d_in.p, 0, n, nullptr, 0, lv, st, &dense))) return rc;
        if (!dense) FB_CUDA_CHECK(cudaMemcpyAsync(&bad, c->d_scalars + kSparseFlagIdx, 4, cudaMemcpyDeviceToHost, st));
        FB_CUDA_CHECK(cudaStreamSynchronize(st));
        if (dense) bad = 1;
        else if (bad) {
            bool ok = false;
            if ((rc = sparse_repair(c, b, c->d_in.p, 0, n, lv, st, &ok))) return rc;
            if (ok) {
                FB_CUDA_CHECK(cudaMemsetAsync(c->d_scalars + kSparseFlagIdx, 0, sizeof(uint32_t), st));
                FB_CUDA_CHECK(lz77_parse_from_nx(b, c->d_in.p, (uint32_t)n, lv, st, &c->timer, c->d_scalars + kSparseFlagIdx));
                FB_CUDA_CHECK(cudaMemcpyAsync(&bad, c->d_scalars + kSparseFlagIdx, 4, cudaMemcpyDeviceToHost, st));
                FB_CUDA_CHECK(cudaStreamSynchronize(st));
                if (!bad) c->sparse_repairs++;
            }
            if (bad) c->sparse_fallbacks++;
        }
    }
    if (c->parse_mode != 0 || n == 0 || bad) {
        FB_CUDA_CHECK(lz77_tokenize(b, c->d_in.p, 0, (uint32_t)n, nullptr, 0, lv, st));
        c->launches += n ? 9 : 0;
    }
    FB_CUDA_CHECK(cudaMemcpyAsync(&total, b.total_tokens, 4, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    *ntok = total;
    if (total > cap) return FB200_NO_SPACE_LEFT;
    if (total) FB_CUDA_CHECK(cudaMemcpy(tokens, b.tokens, (size_t)total * 4, cudaMemcpyDeviceToHost));
    return FB200_OK;
}

int fb200_debug_match_tables(fb200_ctx* c, int level, const uint8_t* in, size_t n, uint32_t* r_full, uint32_t* r_quarter) {
    LevelArgs lv;
    if (!c || !level_args(level, lv) || n > (1ull << 31) || n == 0) return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    FB_CUDA_CHECK(c->d_in.ensure(n + 512));
    int rc = ensure_lz77(c, n);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    FB_CUDA_CHECK(cudaMemcpyAsync(c->d_in.p, in, n, cudaMemcpyHostToDevice, st));
    Lz77Buffers b = lz77_view(c);
    FB_CUDA_CHECK(lz77_tokenize(b, c->d_in.p, 0, (uint32_t)n, nullptr, 0, lv, st));
    c->launches += 9;
    FB_CUDA_CHECK(cudaMemcpyAsync(r_full, b.r_full, n * 4, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaMemcpyAsync(r_quarter, b.r_quarter, n * 4, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    return FB200_OK;
}

int fb200_debug_block_write(fb200_ctx* c, int kind, const uint32_t* tokens, size_t ntok, int eof, const uint8_t* input,
                            size_t input_len, int has_input, uint8_t* out, size_t cap, size_t* out_len) {
    if (!c || kind < 0 || kind > 2 || !out_len || ntok > (1u << 20)) return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    if (kind == 2 && !has_input) return FB200_INVALID_ARGUMENT;
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    FB_CUDA_CHECK(c->d_in.ensure(input_len + 512));
    FB_CUDA_CHECK(c->tokens.ensure(ntok + 64));
    int rc = ensure_blocks(c, 2);
    if (rc) return rc;
    const size_t bound = round_up(ntok * 8 + input_len + input_len / 8 + 8192, 16);
    FB_CUDA_CHECK(c->d_out.ensure(bound));
    if (has_input && input_len) FB_CUDA_CHECK(cudaMemcpyAsync(c->d_in.p, input, input_len, cudaMemcpyHostToDevice, st));
    if (ntok) FB_CUDA_CHECK(cudaMemcpyAsync(c->tokens.p, tokens, ntok * 4, cudaMemcpyHostToDevice, st));
    BlockPlan pl;
    pl.tok_begin = 0;
    pl.tok_count = (uint32_t)ntok;
    pl.in_begin = 0;
    pl.in_len = has_input ? (uint32_t)input_len : 0;
    pl.has_input = has_input ? 1 : 0;
    pl.eof = eof ? 1 : 0;
    pl.kind = (uint32_t)kind;
    if (has_input && input_len > kMaxStore) pl.in_len = (uint32_t)input_len;  // not storable; never dereferenced
    uint32_t one = 1;
    uint32_t* nblocks_dev = c->d_scalars + 1;
    uint64_t* total_bits_dev = reinterpret_cast<uint64_t*>(c->d_scalars + 2);
    FB_CUDA_CHECK(cudaMemcpyAsync(c->plans.p, &pl, sizeof pl, cudaMemcpyHostToDevice, st));
    FB_CUDA_CHECK(cudaMemcpyAsync(nblocks_dev, &one, 4, cudaMemcpyHostToDevice, st));
    if (kind == 2) FB_CUDA_CHECK(histogram_bytes(c->d_in.p, c->plans.p, 1, c->lit_freq.p, st));
    else FB_CUDA_CHECK(histogram_tokens(c->tokens.p, c->plans.p, nblocks_dev, 1, c->lit_freq.p, c->dist_freq.p, st));
    FB_CUDA_CHECK(build_blocks(c->plans.p, nblocks_dev, 1, c->lit_freq.p, c->dist_freq.p, c->descs.p, st));
    FB_CUDA_CHECK(scan_block_offsets(c->descs.p, nblocks_dev, 0, total_bits_dev, st));
    zero_output_kernel<<<64, 256, 0, st>>>(reinterpret_cast<uint32_t*>(c->d_out.p), total_bits_dev, c->d_out.cap / 4);
    FB_CUDA_CHECK(cudaGetLastError());
    FB_CUDA_CHECK(pack_blocks(c->d_in.p, kind == 2 ? nullptr : c->tokens.p, c->descs.p, nblocks_dev, 1,
                              reinterpret_cast<uint32_t*>(c->d_out.p), st));
    c->launches += 5;
    FB_CUDA_CHECK(cudaMemcpyAsync(c->h_scalars, total_bits_dev, 8, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    const size_t end = (size_t)((c->h_scalars[0] + 7) >> 3);
    if (end > cap) return FB200_NO_SPACE_LEFT;
    if (end) FB_CUDA_CHECK(cudaMemcpy(out, c->d_out.p, end, cudaMemcpyDeviceToHost));
    *out_len = end;
    return FB200_OK;
}

}  // extern "C"

// =============================================================================================
// inflate entry points
// =============================================================================================
static int run_members(fb200_ctx* c, int container, const uint8_t* d_in, const MemberDesc* h_desc, size_t k, uint8_t* d_out,
                       MemberResult* h_res, cudaStream_t st) {
    const size_t desc_words = k * (sizeof(MemberDesc) / 8), res_words = k * (sizeof(MemberResult) / 8);
    FB_CUDA_CHECK(c->m_desc.ensure(desc_words + res_words));
    MemberDesc* d_desc = reinterpret_cast<MemberDesc*>(c->m_desc.p);
    MemberResult* d_res = reinterpret_cast<MemberResult*>(c->m_desc.p + desc_words);
    FB_CUDA_CHECK(cudaMemcpyAsync(d_desc, h_desc, k * sizeof(MemberDesc), cudaMemcpyHostToDevice, st));
    c->timer.begin(st);
    static const bool warp_kernel = [] { const char* e = getenv("FB200_INFLATE"); return e && strcmp(e, "warp") == 0; }();
    if (warp_kernel) FB_CUDA_CHECK(inflate_members(container, d_in, d_desc, (uint32_t)k, d_out, d_res, st));
    else {
        FB_CUDA_CHECK(c->m_scratch.ensure(inflate_par_scratch_bytes((uint32_t)k, c->sm_count)));
        bool pieces = false;  // a streaming decompressor's piece of a member: the kernel variant that can stop and resume
        for (size_t i = 0; i < k; i++) pieces = pieces || h_desc[i].flags != 0 || h_desc[i].start_bit != 0;
        FB_CUDA_CHECK(inflate_members_par(container, d_in, d_desc, (uint32_t)k, d_out, d_res, c->m_scratch.p, c->sm_count, st, pieces));
    }
    c->timer.mark(st, kPhInflate);
    c->launches += 1;
    FB_CUDA_CHECK(cudaMemcpyAsync(h_res, d_res, k * sizeof(MemberResult), cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    c->timer.collect();
    static const bool stats = getenv("FB200_INFLATE_STATS") != nullptr;  // development: rounds / retries / exact-path calls
    if (stats)
        for (size_t i = 0; i < k && i < 4; i++)
            fprintf(stderr, "inflate member %zu: status %u out %llu rounds %u retries %u exact %u\n", i, h_res[i].status,
                    (unsigned long long)h_res[i].out_len, h_res[i].pad & 0xfff, (h_res[i].pad >> 12) & 0x3ff, h_res[i].pad >> 22);
    return FB200_OK;
}

extern "C" {

int fb200_decompress_members_device(fb200_ctx* c, int container, const void* d_in, const uint64_t* in_off,
                                    const uint64_t* in_len, size_t k, void* d_out, const uint64_t* out_off,
                                    const uint64_t* out_cap, uint64_t* out_len, uint64_t* consumed, int* status,
                                    void* stream) {
    if (!c || container < 0 || container > 2 || (k && (!in_off || !in_len || !out_off || !out_cap))) return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    if (k > 0x7fffffffu) return FB200_INVALID_ARGUMENT;
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    std::vector<MemberDesc> desc(k);
    std::vector<MemberResult> res(k);
    for (size_t i = 0; i < k; i++) desc[i] = MemberDesc{in_off[i], in_len[i], out_off[i], out_cap[i], 0};
    int rc = run_members(c, container, (const uint8_t*)d_in, desc.data(), k, (uint8_t*)d_out, res.data(), st);
    if (rc) return rc;
    int first = FB200_OK;
    for (size_t i = 0; i < k; i++) {
        if (out_len) out_len[i] = res[i].out_len;
        if (consumed) consumed[i] = res[i].consumed;
        if (status) status[i] = (int)res[i].status;
        if (first == FB200_OK && res[i].status) first = (int)res[i].status;
    }
    return first;
}

int fb200_decompress_members(fb200_ctx* c, int container, const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len,
                             size_t k, uint8_t* out, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                             uint64_t* consumed, int* status) {
    if (!c || container < 0 || container > 2 || k > 0x7fffffffu || (k && (!in || !out || !in_off || !in_len || !out_off || !out_cap || !out_len)))
        return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    for (size_t i = 0; i < k; i++) {  // nothing below may leave these unwritten
        out_len[i] = 0;
        if (consumed) consumed[i] = 0;
        if (status) status[i] = FB200_OK;
    }
    if (k == 0) return FB200_OK;
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    uint64_t in_end = 0, out_end = 0;
    bool ordered = true;  // members one after the other in both buffers: the copies can be pipelined batch by batch
    for (size_t i = 0; i < k; i++) {
        if (in_off[i] + in_len[i] > in_end) in_end = in_off[i] + in_len[i];
        if (out_off[i] + out_cap[i] > out_end) out_end = out_off[i] + out_cap[i];
        if (i && (in_off[i] < in_off[i - 1] + in_len[i - 1] || out_off[i] < out_off[i - 1] + out_cap[i - 1])) ordered = false;
    }
    FB_CUDA_CHECK(c->d_in.ensure(in_end + 512));
    FB_CUDA_CHECK(c->d_out.ensure(out_end + 512));
    cudaStream_t st = c->stream;
    const size_t desc_words = k * (sizeof(MemberDesc) / 8), res_words = k * (sizeof(MemberResult) / 8);
    FB_CUDA_CHECK(c->m_desc.ensure(desc_words + res_words));
    FB_CUDA_CHECK(c->m_scratch.ensure(inflate_par_scratch_bytes((uint32_t)k, c->sm_count)));
    MemberDesc* d_desc = reinterpret_cast<MemberDesc*>(c->m_desc.p);
    MemberResult* d_res = reinterpret_cast<MemberResult*>(c->m_desc.p + desc_words);
    std::vector<MemberDesc> desc(k);
    std::vector<MemberResult> res(k);
    for (size_t i = 0; i < k; i++) desc[i] = MemberDesc{in_off[i], in_len[i], out_off[i], out_cap[i], 0};
    FB_CUDA_CHECK(cudaMemcpyAsync(d_desc, desc.data(), k * sizeof(MemberDesc), cudaMemcpyHostToDevice, st));
    // Batches of members: the input of batch b+1 crosses PCIe while batch b is inflated and the output of batch b-1
    // goes back (three streams).  A batch's output copy covers its members' whole capacity ranges: out_len is only
    // known on the host at the end, and bytes past out_len inside out_cap belong to the caller's buffer anyway.
    const size_t nbatch = !ordered ? 1 : k >= 64 ? 8 : k >= 8 ? 4 : 1;
    static const bool warp_kernel = [] { const char* e = getenv("FB200_INFLATE"); return e && strcmp(e, "warp") == 0; }();
    c->timer.begin(st);
    for (size_t b = 0; b < nbatch; b++) {
        const size_t m0 = k * b / nbatch, m1 = k * (b + 1) / nbatch;
        if (m1 == m0) continue;
        uint64_t lo = in_off[m0], hi = in_off[m1 - 1] + in_len[m1 - 1], olo = out_off[m0], ohi = out_off[m1 - 1] + out_cap[m1 - 1];
        if (!ordered) { lo = 0; hi = in_end; olo = 0; ohi = out_end; }
        cudaEvent_t ev_in = c->slab_ev[(2 * b) % 16], ev_k = c->slab_ev[(2 * b + 1) % 16];
        if (hi > lo) FB_CUDA_CHECK(cudaMemcpyAsync(c->d_in.p + lo, in + lo, hi - lo, cudaMemcpyHostToDevice, c->copy_stream));
        FB_CUDA_CHECK(cudaEventRecord(ev_in, c->copy_stream));
        FB_CUDA_CHECK(cudaStreamWaitEvent(st, ev_in, 0));
        if (warp_kernel) FB_CUDA_CHECK(inflate_members(container, c->d_in.p, d_desc + m0, (uint32_t)(m1 - m0), c->d_out.p, d_res + m0, st));
        else FB_CUDA_CHECK(inflate_members_par(container, c->d_in.p, d_desc + m0, (uint32_t)(m1 - m0), c->d_out.p, d_res + m0, c->m_scratch.p, c->sm_count, st));
        c->launches += 1;
        FB_CUDA_CHECK(cudaEventRecord(ev_k, st));
        FB_CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream2, ev_k, 0));
        if (ohi > olo) FB_CUDA_CHECK(cudaMemcpyAsync(out + olo, c->d_out.p + olo, ohi - olo, cudaMemcpyDeviceToHost, c->copy_stream2));
    }
    c->timer.mark(st, kPhInflate);
    FB_CUDA_CHECK(cudaMemcpyAsync(res.data(), d_res, k * sizeof(MemberResult), cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    FB_CUDA_CHECK(cudaStreamSynchronize(c->copy_stream2));
    c->timer.collect();
    int first = FB200_OK;
    for (size_t i = 0; i < k; i++) {
        out_len[i] = res[i].out_len;
        if (consumed) consumed[i] = res[i].consumed;
        if (status) status[i] = (int)res[i].status;
        if (first == FB200_OK && res[i].status) first = (int)res[i].status;
    }
    return first;
}

int fb200_decompress(fb200_ctx* c, int container, const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len,
                     size_t* consumed) {
    if (!c || (!in && n) || (!out && cap)) return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    uint64_t in_off = 0, in_len = n, out_off = 0, out_cap = cap, olen = 0, used = 0;
    int status = 0;
    uint8_t dummy_in = 0, dummy_out = 0;
    int rc = fb200_decompress_members(c, container, in ? in : &dummy_in, &in_off, &in_len, 1, out ? out : &dummy_out, &out_off,
                                      &out_cap, &olen, &used, &status);
    if (out_len) *out_len = (size_t)olen;
    if (consumed) *consumed = (size_t)used;
    return rc;
}

// =============================================================================================
// streaming compressor: Compressor / SimpleCompressor (deflate.zig:121-373, 449-529).
//
// write() copies the caller's bytes straight into a device window (no host copy of the stream is kept).  Once a
// part's worth of bytes has arrived, the part of the stream whose look-ahead is complete is compressed and its
// finished blocks go to the writer (deflate.zig:363-371 emits blocks as they fill; the bytes do not depend on how
// the input was cut into write() calls).  What is carried from part to part: the clean arrival where the parse
// goes on, the tokens of the open block, the position of the last block cut and the bits of the last, partial output
// byte (PartCarry).  The window keeps 64 KiB of history before the next parse position and slides by multiples of
// 32768, so that the reference's slide schedule (a function of position) reads the same in window coordinates.
// A part runs on a worker thread while the caller goes on writing into the window behind it; the writer is only
// ever called on the caller's thread.  flush() / finish() close the segment with the existing segment pipeline.
// =============================================================================================
struct fb200_deflate {
    fb200_ctx* ctx;
    int device = 0;                 // the context may be destroyed before us: never dereference it in destroy()
    int container, mode;
    bool level_mode = false;
    fb200_write_fn writer;
    void* user;
    // device window, two sets (a slide copies into the other one); index 0 is stream position wbase
    DevBuf<uint8_t> win[2];
    DevBuf<uint16_t> link[2];
    DevBuf<uint32_t> nx[2];
    DevBuf<uint32_t> tokbuf;        // tokens of the open block between parts (the context's token buffer is per call)
    int cur = 0;
    size_t cap = 0;                 // positions a window holds now (grows with the stream up to cap_full)
    size_t cap_full = 0;            // history + two parts + slack
    size_t part = 0;                // bytes per part
    uint64_t wbase = 0;             // stream position of window index 0 (multiple of 32768)
    size_t filled = 0;              // bytes in the window
    size_t seg_begin = 0;           // where the parse goes on: a clean arrival (or the segment start)
    bool seg_fresh = true;          // nothing of the current segment has been linked / evaluated yet
    size_t linked = 0, chunks_done = 0;
    uint32_t carry_tok = 0, fp0 = 0, bit_phase = 0;
    bool has_fp = true;             // fp0 = position of the last block cut, or the start of the segment
    uint8_t carry_byte = 0;
    std::vector<uint64_t> skip;     // stream positions never inserted into the chains (3 before every flush point)
    DevBuf<uint32_t> d_skip;
    // container checksum of everything written, summed part by part on the device and combined here
    uint32_t sum = 0;
    size_t sum_done = 0;            // window bytes below this are in `sum`
    uint64_t total_in = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_ev = nullptr;
    uint8_t* h_outs[2] = {nullptr, nullptr};  // pinned: a part's output on its way to the writer (one being filled, one being emitted)
    size_t h_out_caps[2] = {0, 0};
    int h_cur = 0;                  // the buffer the next stream_run fills
    uint8_t* h_out = nullptr;       // = h_outs[h_cur] during a stream_run
    const uint8_t* pend_ptr = nullptr;  // a finished part's bytes, not yet handed to the writer
    size_t pend_len = 0;
    int pend_buf = 0;
    DevBuf<uint8_t> d_outs[2];      // a part's packed bytes on the device: copied back while the next part is being compressed
    cudaStream_t out_stream = nullptr;
    cudaEvent_t body_ev = nullptr, out_ev[2] = {nullptr, nullptr};
    uint8_t carry_in[2] = {0, 0};   // bits of the previous part that share the first byte of the buffer's bytes
    size_t last_trigger = 0;        // value of `filled` when the last part was started (or the segment began)
    // the part in flight (worker thread).  The worker only reads the stream state; what it found is applied by the
    // caller's thread when it joins (stream_apply), so that write() never races with it.
    std::thread worker;
    bool job_running = false;
    int job_rc = 0;
    size_t job_out = 0;             // bytes of h_out to hand to the writer
    struct Result {
        bool part = false;
        uint32_t bit_phase = 0, carry_tok = 0, fp0 = 0;
        uint8_t carry_byte = 0;
        bool cut = false;
        size_t seg_begin = 0, linked = 0, chunks_done = 0;
    } res;
    bool finished = false;
    int err = 0;
};

namespace {
constexpr size_t kStreamHist = 65536;  // 32 KiB of match history and 32 KiB before it to rebuild that history's chains

int stream_alloc(fb200_deflate* d) {
    keep_pool_warm(d->device);
    FB_CUDA_CHECK(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
    FB_CUDA_CHECK(cudaStreamCreateWithFlags(&d->out_stream, cudaStreamNonBlocking));
    FB_CUDA_CHECK(cudaEventCreateWithFlags(&d->copy_ev, cudaEventDisableTiming));
    FB_CUDA_CHECK(cudaEventCreateWithFlags(&d->body_ev, cudaEventDisableTiming));
    for (auto& e : d->out_ev) FB_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    FB_CUDA_CHECK(d->win[0].ensure_pool(d->cap + 1024, d->copy_stream));
    if (d->level_mode) {
        FB_CUDA_CHECK(d->link[0].ensure_pool(d->cap + 1024, d->copy_stream));
        FB_CUDA_CHECK(d->nx[0].ensure_pool(d->cap + 1024, d->copy_stream));
        FB_CUDA_CHECK(d->tokbuf.ensure_pool(kTokensPerBlock, d->copy_stream));
    }
    FB_CUDA_CHECK(cudaStreamSynchronize(d->copy_stream));
    return FB200_OK;
}

// Moves the window contents [from, filled) to the start of the other buffer set, which gets room for `newcap` positions.
int stream_move(fb200_deflate* d, size_t from, size_t newcap, cudaStream_t st) {
    const int o = d->cur ^ 1;
    const size_t keep = d->filled - from;
    FB_CUDA_CHECK(cudaStreamSynchronize(d->copy_stream));  // the caller's bytes have landed; nothing else writes the window
    FB_CUDA_CHECK(d->win[o].ensure_pool(newcap + 1024, st));
    if (keep) FB_CUDA_CHECK(cudaMemcpyAsync(d->win[o].p, d->win[d->cur].p + from, keep, cudaMemcpyDeviceToDevice, st));
    if (d->level_mode) {
        FB_CUDA_CHECK(d->link[o].ensure_pool(newcap + 1024, st));
        FB_CUDA_CHECK(d->nx[o].ensure_pool(newcap + 1024, st));
        if (keep) {
            FB_CUDA_CHECK(cudaMemcpyAsync(d->link[o].p, d->link[d->cur].p + from, keep * sizeof(uint16_t), cudaMemcpyDeviceToDevice, st));
            FB_CUDA_CHECK(cudaMemcpyAsync(d->nx[o].p, d->nx[d->cur].p + from, keep * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        }
    }
    if (newcap != d->cap) {  // growing: the old, smaller set goes back to the pool
        d->win[d->cur].release_pool(st);
        d->link[d->cur].release_pool(st);
        d->nx[d->cur].release_pool(st);
    }
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    d->cur = o;
    d->cap = newcap;
    return FB200_OK;
}

// container checksum of the window bytes [sum_done, upto), folded into d->sum (stream `st` is synchronized)
int stream_sum(fb200_deflate* d, size_t upto, cudaStream_t st) {
    if (d->container == FB200_RAW || upto <= d->sum_done) {
        d->sum_done = upto > d->sum_done ? upto : d->sum_done;
        return FB200_OK;
    }
    fb200_ctx* c = d->ctx;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    FB_CUDA_CHECK(cudaStreamSynchronize(d->copy_stream));  // the bytes being summed have landed (a write() may be in the middle of its copies)
    const size_t len = upto - d->sum_done;
    uint32_t* sum_dev = c->d_scalars + 16;
    uint32_t got = 0;
    if (d->container == FB200_GZIP) FB_CUDA_CHECK(crc32_device(d->win[d->cur].p + d->sum_done, len, sum_dev, st));
    else FB_CUDA_CHECK(adler32_device(d->win[d->cur].p + d->sum_done, len, sum_dev, reinterpret_cast<uint64_t*>(c->d_scalars + 24), st));
    FB_CUDA_CHECK(cudaMemcpyAsync(&got, sum_dev, 4, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    d->sum = d->container == FB200_GZIP ? fb200_crc32_combine(d->sum, got, len) : fb200_adler32_combine(d->sum, got, len);
    d->sum_done = upto;
    return FB200_OK;
}

// Slides the window so that it starts 64 KiB (rounded down to 32768) before the next parse position.
int stream_slide(fb200_deflate* d, cudaStream_t st) {
    if (d->seg_begin < kStreamHist + 32768) return FB200_OK;
    const size_t shift = (d->seg_begin - kStreamHist) / 32768 * 32768;
    int rc = stream_sum(d, d->filled, st);  // bytes about to leave the window are summed first
    if (rc) return rc;
    if ((rc = stream_move(d, shift, d->cap, st))) return rc;
    d->wbase += shift;
    d->filled -= shift;
    d->seg_begin -= shift;
    d->sum_done -= shift;
    d->last_trigger = d->last_trigger > shift ? d->last_trigger - shift : 0;
    if (!d->seg_fresh) {
        d->linked -= shift;
        d->chunks_done -= shift / 32768;
    }
    if (d->has_fp) d->fp0 = d->fp0 > shift ? d->fp0 - (uint32_t)shift : 0;  // a cut that far back cannot be stored anyway (> 65535 bytes)
    return FB200_OK;
}

// Uploads the skip list in window coordinates; returns the count.
int stream_skip(fb200_deflate* d, cudaStream_t st, uint32_t* count) {
    std::vector<uint32_t> rel;
    for (uint64_t p : d->skip)
        if (p >= d->wbase) rel.push_back((uint32_t)(p - d->wbase));
    *count = (uint32_t)rel.size();
    if (rel.empty()) return FB200_OK;
    FB_CUDA_CHECK(d->d_skip.ensure(rel.size()));
    FB_CUDA_CHECK(cudaMemcpyAsync(d->d_skip.p, rel.data(), rel.size() * 4, cudaMemcpyHostToDevice, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    return FB200_OK;
}

// One call of the segment pipeline over window positions [seg_begin, n): a part when parse_end != 0, else the
// closing flush / finish.  Leaves the output bytes in d->h_out (job_out of them are final) and updates the carry.
int stream_run(fb200_deflate* d, size_t n, size_t parse_end, size_t chunk_end, size_t link_to, bool final_flush) {
    fb200_ctx* c = d->ctx;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    FB_CUDA_CHECK(cudaStreamWaitEvent(st, d->copy_ev, 0));  // the window bytes below n have landed
    uint32_t nskip = 0;
    int rc = d->level_mode ? stream_skip(d, st, &nskip) : FB200_OK;
    if (rc) return rc;
    const size_t begin = d->seg_begin;
    const size_t bound = round_up(fb200_compress_bound(n - begin, d->mode) + 64, 16);
    const int hb = d->h_cur;
    FB_CUDA_CHECK(d->d_outs[hb].ensure_pool(bound + bound / 8, st));
    if (bound + 16 > d->h_out_caps[hb]) {
        pinned_cache().put(d->h_outs[hb], d->h_out_caps[hb]);
        d->h_out_caps[hb] = 0;
        d->h_outs[hb] = pinned_cache().get(bound + bound / 8 + 16, &d->h_out_caps[hb]);
        if (!d->h_outs[hb]) {
            fb::set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__);
            return FB200_ERR_CUDA;
        }
    }
    d->h_out = d->h_outs[hb];
    PartCarry carry;
    carry.bit_phase = d->bit_phase;
    carry.parse_end = parse_end;
    carry.chunk_end = chunk_end;
    carry.link_to = link_to;
    if (d->level_mode) {
        if ((rc = ensure_lz77(c, n))) return rc;  // before the open block's tokens go to the front of the token buffer
        carry.nx = d->nx[d->cur].p;
        carry.carry_tok = d->carry_tok;
        carry.has_fp = d->has_fp;
        carry.fp0 = d->fp0;
        if (!d->seg_fresh) {
            carry.link_from = d->linked;
            carry.chunks_from = d->chunks_done;
        }
        if (d->carry_tok) FB_CUDA_CHECK(cudaMemcpyAsync(c->tokens.p, d->tokbuf.p, (size_t)d->carry_tok * 4, cudaMemcpyDeviceToDevice, st));
        std::swap(c->link, d->link[d->cur]);  // the pipeline uses the context's link buffer: lend it ours for the call
    }
    size_t out_end = 0;
    static const bool trace = getenv("FB200_STREAM_TRACE") != nullptr;  // development: host-side times of every part
    const auto t0 = std::chrono::steady_clock::now();
    rc = deflate_body_device(c, d->container, d->mode, d->win[d->cur].p, begin, n, d->d_skip.p, nskip, d->d_outs[hb].p, d->d_outs[hb].cap,
                             &out_end, final_flush, false, st, nullptr, nullptr, 0, nullptr, 0, &carry);
    const auto t1 = std::chrono::steady_clock::now();
    if (d->level_mode) std::swap(c->link, d->link[d->cur]);
    if (rc) return rc;
    // output: whole bytes go to the writer, the bits of the last partial byte wait for the next part
    const size_t nfull = parse_end ? (size_t)(carry.total_bits >> 3) : out_end;
    const size_t ncopy = (size_t)((carry.total_bits + 7) >> 3);
    // the packed bytes go back on their own stream (the pipeline has synchronized `st`, so they are complete): a part
    // returns as soon as its carry is known, and the copy runs while the next part is being compressed
    if (ncopy) FB_CUDA_CHECK(cudaMemcpyAsync(d->h_out, d->d_outs[hb].p, ncopy, cudaMemcpyDeviceToHost, d->out_stream));
    FB_CUDA_CHECK(cudaEventRecord(d->out_ev[hb], d->out_stream));
    if (d->level_mode && carry.leftover) FB_CUDA_CHECK(cudaMemcpyAsync(d->tokbuf.p, c->tokens.p, (size_t)carry.leftover * 4, cudaMemcpyDeviceToDevice, st));
    const uint32_t phase = parse_end ? (uint32_t)(carry.total_bits & 7) : 0;
    uint8_t* last_pinned = reinterpret_cast<uint8_t*>(c->h_scalars + 13);  // the byte the next part's bits share
    *last_pinned = 0;
    if (phase) FB_CUDA_CHECK(cudaMemcpyAsync(last_pinned, d->d_outs[hb].p + nfull, 1, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    const uint8_t last = *last_pinned;
    if (!parse_end) FB_CUDA_CHECK(cudaEventSynchronize(d->out_ev[hb]));  // flush / finish hand their bytes out at once
    if (trace) {
        const auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "stream part: begin %zu n %zu parse_end %zu: body %.3f ms, carry %.3f ms, %zu bytes on their way back\n", begin, n, parse_end,
                std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count(), ncopy);
    }
    d->carry_in[hb] = d->carry_byte;
    if (!parse_end) {
        if (ncopy == 0) d->h_out[0] = 0;
        d->h_out[0] |= d->carry_byte;
    }
    d->job_out = nfull;
    fb200_deflate::Result& r = d->res;
    r.part = parse_end != 0;
    r.bit_phase = phase;
    r.carry_byte = phase ? (uint8_t)(last | (nfull == 0 ? d->carry_byte : 0)) : 0;
    r.carry_tok = carry.leftover;
    r.cut = carry.cut;
    r.fp0 = carry.last_rp;
    r.seg_begin = (parse_end && d->level_mode) ? parse_end + carry.exit : n;
    r.linked = link_to;
    r.chunks_done = chunk_end;
    return FB200_OK;
}

// the stream state after a finished stream_run (caller's thread)
void stream_apply(fb200_deflate* d) {
    const fb200_deflate::Result& r = d->res;
    d->bit_phase = r.bit_phase;
    d->carry_byte = r.carry_byte;
    d->carry_tok = r.carry_tok;
    d->seg_begin = r.seg_begin;
    if (r.part) {
        if (r.cut) {
            d->has_fp = true;
            d->fp0 = r.fp0;
        }
        if (d->level_mode) {
            d->linked = r.linked;
            d->chunks_done = r.chunks_done;
            d->seg_fresh = false;
        }
    } else {
        d->has_fp = true;  // the next block's stored-input candidate starts at the flush point (SlidingWindow.zig:113)
        d->fp0 = (uint32_t)r.seg_begin;
        d->seg_fresh = true;
    }
}

// waits for the part in flight and takes its results (caller's thread); its bytes wait for stream_emit
int stream_join(fb200_deflate* d) {
    if (!d->job_running) return FB200_OK;
    static const bool trace = getenv("FB200_STREAM_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    d->worker.join();
    if (trace) fprintf(stderr, "stream join: waited %.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    d->job_running = false;
    if (d->job_rc) return d->job_rc;
    stream_apply(d);
    d->pend_ptr = d->h_out;  // handed to the writer by stream_emit, after the next part has been started
    d->pend_len = d->job_out;
    d->pend_buf = d->h_cur;
    d->job_out = 0;
    d->h_cur ^= 1;
    return FB200_OK;
}
int stream_emit(fb200_deflate* d) {
    const size_t n = d->pend_len;
    d->pend_len = 0;
    if (n == 0) return FB200_OK;
    FB_CUDA_CHECK(cudaEventSynchronize(d->out_ev[d->pend_buf]));  // the part's bytes have arrived in pinned memory
    d->h_outs[d->pend_buf][0] |= d->carry_in[d->pend_buf];
    if (d->writer(d->user, d->pend_ptr, n)) return FB200_NO_SPACE_LEFT;
    return FB200_OK;
}

// Starts a part if a part's worth of input is waiting.  force: the window is full, a part must make room.
int stream_maybe_part(fb200_deflate* d, bool force) {
    int rc = stream_join(d);
    if (rc) return rc;
    const size_t pending = d->filled - d->seg_begin;
    if (pending < d->part / 2 && !force) return FB200_OK;
    size_t n, parse_end, chunk_end = 0, link_to = 0;
    if (d->level_mode) {
        if (d->filled < 2 * 32768) return FB200_OK;
        chunk_end = d->filled / 32768 - 1;  // the chunks whose look-ahead (and slide schedule) no later byte can change
        const size_t limit = chunk_end * 32768;
        if (limit <= d->seg_begin + kChunk || (!d->seg_fresh && chunk_end <= d->chunks_done)) return FB200_OK;
        parse_end = d->seg_begin + (limit - d->seg_begin) / kChunk * kChunk;
        link_to = limit + 8192;
        n = d->filled;
    } else {
        const size_t slices = pending / kMaxStore;
        if (slices == 0) return FB200_OK;
        n = d->seg_begin + slices * kMaxStore;
        parse_end = n;
    }
    FB_CUDA_CHECK(cudaEventRecord(d->copy_ev, d->copy_stream));
    d->last_trigger = d->filled;
    d->job_running = true;
    d->job_rc = 0;
    d->worker = std::thread([d, n, parse_end, chunk_end, link_to] { d->job_rc = stream_run(d, n, parse_end, chunk_end, link_to, false); });
    return FB200_OK;
}
}  // namespace

int fb200_deflate_create(fb200_ctx* ctx, int container, int mode, fb200_write_fn writer, void* user, fb200_deflate** out) {
    LevelArgs lv;
    if (!ctx || !out || !writer || container < 0 || container > 2) return FB200_INVALID_ARGUMENT;
    if (mode != FB200_MODE_STORE && mode != FB200_MODE_HUFFMAN && !level_args(mode, lv)) return FB200_INVALID_ARGUMENT;
    *out = nullptr;
    fb200_deflate* d = new (std::nothrow) fb200_deflate();
    if (!d) return FB200_INVALID_ARGUMENT;
    d->ctx = ctx;
    d->device = ctx->device;
    d->container = container;
    d->mode = mode;
    d->level_mode = mode >= 4;
    d->writer = writer;
    d->user = user;
    d->sum = container == FB200_ZLIB ? 1u : 0u;  // Adler-32 / CRC-32 of nothing
    // bytes per part: FB200_STREAM_PART (KiB) for tests and tuning; 64 MiB parts run at 0.8 of the one-shot rate (the
    // fixed cost of a part, its synchronisations and short kernels, is about 0.8 ms)
    size_t part_kib = 65536;
    if (const char* e = getenv("FB200_STREAM_PART")) {
        const long v = atol(e);
        if (v >= 128 && v <= (1 << 20)) part_kib = (size_t)v;
    }
    d->part = part_kib << 10;
    d->cap_full = kStreamHist + 2 * d->part + (1u << 20);
    d->cap = d->cap_full < kStreamHist + (3u << 20) ? d->cap_full : kStreamHist + (3u << 20);  // short streams stay small
    if (cudaSetDevice(ctx->device) != cudaSuccess || stream_alloc(d) != FB200_OK) {
        fb200_deflate_destroy(d);
        return FB200_ERR_CUDA;
    }
    // the reference writes the container header in init (deflate.zig:144, :470)
    int rc = 0;
    if (container == FB200_GZIP) rc = writer(user, kGzipHeader, 10);
    else if (container == FB200_ZLIB) rc = writer(user, kZlibHeader, 2);
    if (rc) {
        fb200_deflate_destroy(d);
        return FB200_NO_SPACE_LEFT;
    }
    *out = d;
    return FB200_OK;
}

int fb200_deflate_write(fb200_deflate* d, const uint8_t* data, size_t n) {
    if (!d || (!data && n)) return FB200_INVALID_ARGUMENT;
    if (d->err) return d->err;
    if (d->finished) return d->err = FB200_INVALID_STATE;
    if (cudaSetDevice(d->device) != cudaSuccess) return d->err = FB200_ERR_CUDA;
    while (n) {
        if (d->filled == d->cap && d->cap < d->cap_full) {  // a stream that keeps coming gets a bigger window
            int rc = stream_join(d);
            if (!rc) rc = stream_emit(d);
            const size_t bigger = d->cap * 4 < d->cap_full ? d->cap * 4 : d->cap_full;
            if (!rc) rc = stream_move(d, 0, bigger, d->ctx->stream);
            if (rc) return d->err = rc;
        } else if (d->filled == d->cap) {  // make room: finish the part in flight, run another if need be, slide
            int rc = stream_maybe_part(d, true);
            if (!rc) rc = stream_emit(d);
            if (!rc) rc = stream_join(d);
            if (!rc) rc = stream_emit(d);
            if (!rc) rc = stream_slide(d, d->ctx->stream);
            if (rc) return d->err = rc;
            if (d->filled == d->cap) return d->err = FB200_NO_SPACE_LEFT;  // cannot happen: a part always frees more than it keeps
        }
        const size_t take = n < d->cap - d->filled ? n : d->cap - d->filled;
        if (cudaMemcpyAsync(d->win[d->cur].p + d->filled, data, take, cudaMemcpyHostToDevice, d->copy_stream) != cudaSuccess) {
            fb::set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__);
            return d->err = FB200_ERR_CUDA;
        }
        d->filled += take;
        d->total_in += take;
        data += take;
        n -= take;
        if (d->filled - d->last_trigger >= d->part) {  // a part's worth since the last one: take its results, start the next
            static const bool trace = getenv("FB200_STREAM_TRACE") != nullptr;
            const auto t0 = std::chrono::steady_clock::now();
            int rc = stream_join(d);
            const auto t1 = std::chrono::steady_clock::now();
            if (!rc) rc = stream_slide(d, d->ctx->stream);  // between parts: nothing reads the window
            const auto t2 = std::chrono::steady_clock::now();
            if (!rc) rc = stream_maybe_part(d, false);
            if (!rc) rc = stream_emit(d);  // the finished part's bytes go out while the next one is being compressed
            if (trace) {
                const auto t3 = std::chrono::steady_clock::now();
                static std::chrono::steady_clock::time_point last = t0;
                fprintf(stderr, "stream trigger: since last %.3f ms | join+writer %.3f ms, slide %.3f ms, launch %.3f ms\n",
                        std::chrono::duration<double, std::milli>(t0 - last).count(), std::chrono::duration<double, std::milli>(t1 - t0).count(),
                        std::chrono::duration<double, std::milli>(t2 - t1).count(), std::chrono::duration<double, std::milli>(t3 - t2).count());
                last = t3;
            }
            if (rc) return d->err = rc;
        }
    }
    // the caller may reuse `data` at once: the copy has to be out of it
    if (cudaStreamSynchronize(d->copy_stream) != cudaSuccess) return d->err = FB200_ERR_CUDA;
    return FB200_OK;
}

// Closes the segment: non-final + sync marker for flush (deflate.zig:335-337, :268-288), final + footer for finish
// (:344-347).  How the input was cut into write() calls is not observable in the reference's output (the slide
// schedule depends on stream position only).
static int deflate_close_segment(fb200_deflate* d, bool final_flush) {
    if (cudaSetDevice(d->device) != cudaSuccess) return FB200_ERR_CUDA;
    int rc = stream_join(d);
    if (!rc) rc = stream_emit(d);
    if (rc) return rc;
    fb200_ctx* c = d->ctx;
    FB_CUDA_CHECK(cudaStreamSynchronize(d->copy_stream));
    FB_CUDA_CHECK(cudaEventRecord(d->copy_ev, d->copy_stream));
    const size_t end = d->filled;
    if ((rc = stream_run(d, end, 0, 0, 0, final_flush))) return rc;
    stream_apply(d);
    d->last_trigger = d->filled;
    size_t total = d->job_out;
    d->job_out = 0;
    if (final_flush) {
        if ((rc = stream_sum(d, d->filled, c->stream))) return rc;
        total += make_footer(d->container, d->sum, (size_t)d->total_in, d->h_out + total);
    }
    if (total && d->writer(d->user, d->h_out, total)) return FB200_NO_SPACE_LEFT;
    // positions with fewer than 4 bytes before the flush point were never hashed (Lookup.zig:24) and
    // stay that way for later segments
    if (!final_flush && d->level_mode) {
        const uint64_t aend = d->wbase + end;
        for (uint64_t p = aend >= 3 ? aend - 3 : 0; p < aend; p++)
            if (d->skip.empty() || d->skip.back() < p) d->skip.push_back(p);
    }
    // only flush points within the last 32 KiB + one tile can matter again
    while (!d->skip.empty() && d->skip.front() + 2 * kHist < d->wbase + end) d->skip.erase(d->skip.begin());
    return FB200_OK;
}
int fb200_deflate_flush(fb200_deflate* d) {
    if (!d) return FB200_INVALID_ARGUMENT;
    if (d->err) return d->err;
    if (d->finished) return d->err = FB200_INVALID_STATE;
    int rc = deflate_close_segment(d, false);
    if (rc) d->err = rc;
    return rc;
}
int fb200_deflate_finish(fb200_deflate* d) {
    if (!d) return FB200_INVALID_ARGUMENT;
    if (d->err) return d->err;
    if (d->finished) return d->err = FB200_INVALID_STATE;
    int rc = deflate_close_segment(d, true);
    if (rc) return d->err = rc;
    d->finished = true;
    return FB200_OK;
}
void fb200_deflate_set_writer(fb200_deflate* d, fb200_write_fn writer, void* user) {
    if (!d) return;
    d->writer = writer;
    d->user = user;
}
void fb200_deflate_destroy(fb200_deflate* d) {
    if (!d) return;
    if (d->job_running) d->worker.join();
    if (cudaSetDevice(d->device) != cudaSuccess) (void)cudaGetLastError();
    if (d->out_stream) cudaStreamSynchronize(d->out_stream);  // a part's bytes may still be on their way back
    for (int i = 0; i < 2; i++) {
        d->win[i].release_pool(d->copy_stream);
        d->link[i].release_pool(d->copy_stream);
        d->nx[i].release_pool(d->copy_stream);
        pinned_cache().put(d->h_outs[i], d->h_out_caps[i]);
    }
    d->tokbuf.release_pool(d->copy_stream);
    for (auto& b : d->d_outs) b.release_pool(d->copy_stream);
    if (d->out_stream) cudaStreamSynchronize(d->out_stream);
    for (auto& e : d->out_ev)
        if (e) cudaEventDestroy(e);
    if (d->body_ev) cudaEventDestroy(d->body_ev);
    if (d->out_stream) cudaStreamDestroy(d->out_stream);
    d->d_skip.release();
    if (d->copy_stream) cudaStreamSynchronize(d->copy_stream);
    if (d->copy_ev) cudaEventDestroy(d->copy_ev);
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    delete d;
}

// =============================================================================================
// streaming decompressor: Decompressor (inflate.zig:43-355).  One member per decompress()/next()
// sequence; reset() continues with the next member of the same reader, keeping the history
// (CircularBuffer.wp is not reset, so later members may reach into earlier output).
//
// The reader is pulled on demand, a chunk at a time, and a member is decoded piece by piece: every call of the
// kernel takes the input read so far and decodes whole deflate blocks until the input, or the room for output, runs
// out inside a block; it then stops at the start of that block (kMemberPartial) and the next call resumes there
// with more input, the last 32 KiB of output as history.  Output is handed out as soon as a piece is decoded, so
// memory stays bounded by the piece size whatever the member's size.  Reading stops once the member's footer has
// been seen; what was read past it stays available (next member after reset(), or fb200_inflate_unused).
// =============================================================================================
struct fb200_inflate {
    fb200_ctx* ctx;
    int container;
    fb200_read_fn reader;
    void* user;
    std::vector<uint8_t> in;     // input read and not yet consumed; decoding goes on at bit `start_bit` of in[0]
    uint32_t start_bit = 0;
    bool reader_eof = false;
    bool started = false;        // the current member's container header has been consumed
    std::vector<uint8_t> hist;   // last <= 32768 bytes of output (earlier members included)
    uint64_t member_out = 0;     // plain bytes of the current member so far
    uint32_t sum = 0;            // their CRC-32 / Adler-32
    std::vector<uint8_t> out;    // the decoded piece being handed out
    size_t out_pos = 0;
    size_t read_chunk = 65536;   // grows to 8 MiB while a member keeps asking for more
    size_t out_cap = 32u << 20;  // room for one piece's output
    enum { kDecoding, kEnd } state = kDecoding;
    int err = 0;
};

int fb200_inflate_create(fb200_ctx* ctx, int container, fb200_read_fn reader, void* user, fb200_inflate** out) {
    if (!ctx || !out || container < 0 || container > 2) return FB200_INVALID_ARGUMENT;
    fb200_inflate* s = new (std::nothrow) fb200_inflate();
    if (!s) return FB200_INVALID_ARGUMENT;
    s->ctx = ctx;
    s->container = container;
    s->reader = reader;
    s->user = user;
    s->sum = container == FB200_ZLIB ? 1u : 0u;
    *out = s;
    return FB200_OK;
}
// reads until `want` bytes are buffered or the reader is exhausted
static void inflate_fill(fb200_inflate* s, size_t want) {
    while (!s->reader_eof && s->in.size() < want) {
        if (!s->reader) { s->reader_eof = true; break; }
        const size_t have = s->in.size(), ask = want - have < 65536 ? 65536 : want - have;
        s->in.resize(have + ask);
        const size_t got = s->reader(s->user, s->in.data() + have, ask);
        s->in.resize(have + (got <= ask ? got : ask));
        if (got == 0) s->reader_eof = true;
    }
}
// container footer after the final block (inflate.zig:271-275, container.zig:154-166): same order of checks
static int inflate_footer(fb200_inflate* s) {
    const size_t at = s->start_bit ? 1 : 0;  // the footer starts at the next byte boundary
    const size_t need = s->container == FB200_GZIP ? 8 : s->container == FB200_ZLIB ? 4 : 0;
    inflate_fill(s, at + need);
    const size_t have = s->in.size() > at ? s->in.size() - at : 0;
    const uint8_t* f = s->in.data() + at;
    if (need) {
        if (have < 4) return FB200_END_OF_STREAM;
        if (s->container == FB200_GZIP) {
            const uint32_t crc = (uint32_t)f[0] | ((uint32_t)f[1] << 8) | ((uint32_t)f[2] << 16) | ((uint32_t)f[3] << 24);
            if (crc != s->sum) return FB200_WRONG_GZIP_CHECKSUM;
            if (have < 8) return FB200_END_OF_STREAM;
            const uint32_t sz = (uint32_t)f[4] | ((uint32_t)f[5] << 8) | ((uint32_t)f[6] << 16) | ((uint32_t)f[7] << 24);
            if (sz != (uint32_t)s->member_out) return FB200_WRONG_GZIP_SIZE;
        } else {
            const uint32_t ad = ((uint32_t)f[0] << 24) | ((uint32_t)f[1] << 16) | ((uint32_t)f[2] << 8) | (uint32_t)f[3];
            if (ad != s->sum) return FB200_WRONG_ZLIB_CHECKSUM;
        }
    }
    const size_t used = (at + need) < s->in.size() ? at + need : s->in.size();
    s->in.erase(s->in.begin(), s->in.begin() + used);
    s->start_bit = 0;
    return FB200_OK;
}
// Decodes the next piece of the current member into s->out (possibly nothing, when the member ends).
static int inflate_decode_piece(fb200_inflate* s) {
    fb200_ctx* c = s->ctx;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    s->out.clear();
    s->out_pos = 0;
    if (s->in.size() < s->read_chunk / 2) inflate_fill(s, s->read_chunk);
    for (;;) {
        const size_t n = s->in.size(), hist = s->hist.size();
        FB_CUDA_CHECK(c->d_in.ensure(n + 512));
        FB_CUDA_CHECK(c->d_out.ensure(hist + s->out_cap + 512));
        if (n) FB_CUDA_CHECK(cudaMemcpyAsync(c->d_in.p, s->in.data(), n, cudaMemcpyHostToDevice, st));
        if (hist) FB_CUDA_CHECK(cudaMemcpyAsync(c->d_out.p, s->hist.data(), hist, cudaMemcpyHostToDevice, st));
        // the reference only checks `wp < distance` (CircularBuffer.zig:45): all of the history kept is reachable
        MemberDesc md{0, n, hist, s->out_cap, hist, s->start_bit,
                      kMemberNoFooter | (s->started ? kMemberResume : 0u) | (s->reader_eof ? 0u : kMemberPartial)};
        MemberResult res;
        int rc = run_members(c, s->container, c->d_in.p, &md, 1, c->d_out.p, &res, st);
        if (rc) return rc;
        if (res.status) return (int)res.status;  // the input is complete (or the error is real): the reference's error class
        const bool moved = res.resume_bits > s->start_bit;
        if (res.out_len) {
            s->out.resize(res.out_len);
            FB_CUDA_CHECK(cudaMemcpy(s->out.data(), c->d_out.p + hist, res.out_len, cudaMemcpyDeviceToHost));
            s->member_out += res.out_len;
            if (s->container == FB200_GZIP) s->sum = fb200_crc32_combine(s->sum, res.sum, res.out_len);
            else if (s->container == FB200_ZLIB) s->sum = fb200_adler32_combine(s->sum, res.sum, res.out_len);
            s->hist.insert(s->hist.end(), s->out.begin(), s->out.end());
            if (s->hist.size() > kMaxDist) s->hist.erase(s->hist.begin(), s->hist.end() - kMaxDist);
        }
        if (moved) {
            s->started = true;
            const size_t whole = (size_t)(res.resume_bits >> 3);
            s->in.erase(s->in.begin(), s->in.begin() + (whole < s->in.size() ? whole : s->in.size()));
            s->start_bit = (uint32_t)(res.resume_bits & 7);
        }
        if (res.info & 1u) {  // the final block is done
            rc = inflate_footer(s);
            if (rc) return rc;
            s->state = fb200_inflate::kEnd;
            return FB200_OK;
        }
        if (!s->out.empty()) return FB200_OK;
        const uint32_t why = (res.info >> 8) & 0xffu;
        if (!moved && why == FB200_NO_SPACE_LEFT) {  // one block larger than a piece's room
            if (s->out_cap >= ((size_t)1 << 34)) return FB200_NO_SPACE_LEFT;
            s->out_cap *= 4;
            continue;
        }
        // a block does not fit in what has been read (or nothing but headers went by): more input
        const size_t before = s->in.size();
        inflate_fill(s, before + s->read_chunk);
        if (s->read_chunk < (8u << 20)) s->read_chunk *= 2;
        if (s->in.size() == before) s->reader_eof = true;  // nothing more: the next pass decides what the truncated stream means
    }
}
int fb200_inflate_get(fb200_inflate* s, size_t limit, const uint8_t** data, size_t* len) {
    if (!s || !data || !len) return FB200_INVALID_ARGUMENT;
    *data = nullptr;
    *len = 0;
    if (s->err) return s->err;
    while (s->out_pos == s->out.size()) {
        if (s->state == fb200_inflate::kEnd) return FB200_OK;
        int rc = inflate_decode_piece(s);
        if (rc) return s->err = rc;
    }
    // the reference hands out at most one 64 KiB ring's worth per call (inflate.zig:321-325)
    size_t take = s->out.size() - s->out_pos;
    if (take > 65536) take = 65536;
    if (limit && take > limit) take = limit;
    *data = s->out.data() + s->out_pos;
    *len = take;
    s->out_pos += take;
    return FB200_OK;
}
int fb200_inflate_next(fb200_inflate* s, const uint8_t** data, size_t* len) { return fb200_inflate_get(s, 0, data, len); }
int fb200_inflate_read(fb200_inflate* s, uint8_t* buf, size_t cap, size_t* n) {
    if (!n) return FB200_INVALID_ARGUMENT;
    *n = 0;
    if (cap == 0) return FB200_OK;
    const uint8_t* p;
    size_t len;
    int rc = fb200_inflate_get(s, cap, &p, &len);
    if (rc) return rc;
    if (len) memcpy(buf, p, len);
    *n = len;
    return FB200_OK;
}
int fb200_inflate_reset(fb200_inflate* s) {
    if (!s) return FB200_INVALID_ARGUMENT;
    // inflate.zig:301-309: only legal once the current member has been fully delivered
    if (s->err || s->state != fb200_inflate::kEnd || s->out_pos != s->out.size()) return FB200_INVALID_STATE;
    s->out.clear();
    s->out_pos = 0;
    s->member_out = 0;
    s->sum = s->container == FB200_ZLIB ? 1u : 0u;
    s->started = false;
    s->state = fb200_inflate::kDecoding;
    return FB200_OK;
}
int fb200_inflate_unused(fb200_inflate* s, const uint8_t** data, size_t* len) {
    if (!s || !data || !len) return FB200_INVALID_ARGUMENT;
    *data = s->in.empty() ? nullptr : s->in.data();
    *len = s->in.size();
    return FB200_OK;
}
void fb200_inflate_set_reader(fb200_inflate* s, fb200_read_fn reader, void* user) {
    if (!s) return;
    s->reader = reader;
    s->user = user;
    s->reader_eof = false;
    if (s->state == fb200_inflate::kEnd && s->out_pos == s->out.size()) {  // inflate.zig:283-288
        s->out.clear();
        s->out_pos = 0;
        s->member_out = 0;
        s->sum = s->container == FB200_ZLIB ? 1u : 0u;
        s->started = false;
        s->state = fb200_inflate::kDecoding;
    }
}
// =============================================================================================
// multi-member gzip file without an index: speculative cuts at header look-alikes, validated piece by piece
// =============================================================================================
int fb200_decompress_gzip_file(fb200_ctx* c, const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len, size_t* consumed,
                               size_t* members) {
    if (!c || (!in && n) || (!out && cap) || !out_len) return FB200_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);
    *out_len = 0;
    if (consumed) *consumed = 0;
    if (members) *members = 0;
    FB_CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    FB_CUDA_CHECK(c->d_in.ensure(n + 512));
    if (n) FB_CUDA_CHECK(cudaMemcpyAsync(c->d_in.p, in, n, cudaMemcpyHostToDevice, st));
    // 1. candidates
    uint32_t list_cap = (uint32_t)(n / 4096 + 1024);
    std::vector<uint64_t> cuts;
    for (;;) {
        FB_CUDA_CHECK(c->m_desc.ensure(list_cap + 8));
        uint32_t* d_count = c->d_scalars + 22;
        FB_CUDA_CHECK(gzip_candidates_device(c->d_in.p, n, c->m_desc.p, list_cap, d_count, st));
        uint32_t count = 0;
        FB_CUDA_CHECK(cudaMemcpyAsync(&count, d_count, 4, cudaMemcpyDeviceToHost, st));
        FB_CUDA_CHECK(cudaStreamSynchronize(st));
        c->launches += 1;
        if (count > list_cap) {  // a file made of header look-alikes: take them all
            list_cap = count + 1024;
            continue;
        }
        cuts.resize(count);
        if (count) FB_CUDA_CHECK(cudaMemcpy(cuts.data(), c->m_desc.p, (size_t)count * 8, cudaMemcpyDeviceToHost));
        break;
    }
    std::sort(cuts.begin(), cuts.end());
    if (cuts.empty() || cuts[0] != 0) cuts.insert(cuts.begin(), 0);  // the file's first member starts at 0 whatever it looks like
    // 2. decode the pieces; settle them in stream order
    auto le32 = [&](uint64_t at) { return (uint64_t)in[at] | ((uint64_t)in[at + 1] << 8) | ((uint64_t)in[at + 2] << 16) | ((uint64_t)in[at + 3] << 24); };
    size_t settled = 0;        // pieces [0, settled) are members
    uint64_t settled_out = 0;  // their plain bytes
    uint64_t end_of_members = 0;
    while (settled < cuts.size()) {
        const size_t k = cuts.size() - settled;
        std::vector<uint64_t> io(k), il(k), oo(k), oc(k), ol(k), used(k);
        std::vector<int> stv(k);
        uint64_t total = settled_out, claimed = settled_out;
        for (size_t i = 0; i < k; i++) {
            const bool last = settled + i + 1 == cuts.size();
            const uint64_t b = cuts[settled + i], e = last ? n : cuts[settled + i + 1];
            io[i] = b;
            il[i] = e - b;
            const uint64_t isize = e - b >= 18 ? le32(e - 4) : 0;  // ISIZE of the member that ends where the next piece starts
            claimed += isize;
            // the last piece may be followed by bytes that belong to no member: its size is not known, it gets the room left
            oc[i] = last ? (cap > total ? cap - total : 0) : isize;
            oo[i] = total;
            total += oc[i];
        }
        if (claimed > cap || total > cap) {
            *out_len = (size_t)(claimed > total ? claimed : total);
            return FB200_NO_SPACE_LEFT;
        }
        FB_CUDA_CHECK(c->d_out.ensure(total + 512));
        int rc = fb200_decompress_members_device(c, FB200_GZIP, c->d_in.p, io.data(), il.data(), k, c->d_out.p, oo.data(), oc.data(), ol.data(),
                                                 used.data(), stv.data(), nullptr);
        if (rc == FB200_ERR_CUDA || rc == FB200_INVALID_ARGUMENT) return rc;
        size_t good = 0;
        while (good < k && stv[good] == FB200_OK && (used[good] == il[good] || settled + good + 1 == cuts.size())) good++;
        // the good prefix is final: its bytes go home now (the pieces of a round are contiguous in the output)
        if (good) {
            const uint64_t lo = oo[0], hi = oo[good - 1] + ol[good - 1];
            if (hi > lo) FB_CUDA_CHECK(cudaMemcpy(out + lo, c->d_out.p + lo, hi - lo, cudaMemcpyDeviceToHost));
            settled_out = hi;
            end_of_members = io[good - 1] + used[good - 1];
        }
        settled += good;
        if (good == k) break;
        // The first piece that is not a member.  What the sequential decoder does here is decode from this position with
        // the rest of the file behind it: do exactly that.  If the member is sound, the cuts inside it were look-alikes;
        // if it is not, this is the file's first error.
        const uint64_t b = cuts[settled];
        uint64_t solo_in = n - b, solo_out = settled_out, solo_cap = cap - settled_out, solo_len = 0, solo_used = 0;
        int solo_st = 0;
        FB_CUDA_CHECK(c->d_out.ensure(cap + 512));
        rc = fb200_decompress_members_device(c, FB200_GZIP, c->d_in.p, &b, &solo_in, 1, c->d_out.p, &solo_out, &solo_cap, &solo_len, &solo_used,
                                             &solo_st, nullptr);
        if (rc == FB200_ERR_CUDA || rc == FB200_INVALID_ARGUMENT) return rc;
        if (solo_st != FB200_OK) {
            *out_len = (size_t)settled_out;
            if (consumed) *consumed = (size_t)end_of_members;
            if (members) *members = settled;
            return solo_st;
        }
        if (solo_len) FB_CUDA_CHECK(cudaMemcpy(out + solo_out, c->d_out.p + solo_out, solo_len, cudaMemcpyDeviceToHost));
        settled_out += solo_len;
        end_of_members = b + solo_used;
        settled += 1;
        size_t drop = settled;
        while (drop < cuts.size() && cuts[drop] < end_of_members) drop++;
        cuts.erase(cuts.begin() + settled, cuts.begin() + drop);
        if (settled < cuts.size() && cuts[settled] != end_of_members) cuts.resize(settled);  // what follows the member is not a member
    }
    *out_len = (size_t)settled_out;
    if (consumed) *consumed = (size_t)end_of_members;
    if (members) *members = settled;
    return FB200_OK;
}

// =============================================================================================
// several GPUs behind one call: one context and one host thread per device
// =============================================================================================
struct fb200_pool {
    std::vector<fb200_ctx*> ctxs;
};
int fb200_pool_create(uint64_t device_mask, fb200_pool** out) {
    if (!out) return FB200_INVALID_ARGUMENT;
    *out = nullptr;
    const int ndev = fb200_device_count();
    if (ndev == 0) return FB200_NO_DEVICE;
    fb200_pool* p = new (std::nothrow) fb200_pool();
    if (!p) return FB200_INVALID_ARGUMENT;
    for (int dev = 0; dev < ndev && dev < 64; dev++) {
        if (device_mask && !((device_mask >> dev) & 1)) continue;
        fb200_ctx* c = nullptr;
        const int rc = fb200_ctx_create(dev, &c);
        if (rc) {
            fb200_pool_destroy(p);
            return rc;
        }
        p->ctxs.push_back(c);
    }
    if (p->ctxs.empty()) {
        fb200_pool_destroy(p);
        return FB200_INVALID_ARGUMENT;
    }
    *out = p;
    return FB200_OK;
}
int fb200_pool_devices(const fb200_pool* p) { return p ? (int)p->ctxs.size() : 0; }
void fb200_pool_destroy(fb200_pool* p) {
    if (!p) return;
    for (fb200_ctx* c : p->ctxs) fb200_ctx_destroy(c);
    delete p;
}
int fb200_compress_batch(fb200_pool* p, int container, int mode, size_t k, const uint8_t* const* in, const size_t* in_len,
                         uint8_t* const* out, const size_t* out_cap, size_t* out_len, int* status) {
    if (!p || (k && (!in || !in_len || !out || !out_cap || !out_len))) return FB200_INVALID_ARGUMENT;
    const size_t ndev = p->ctxs.size();
    // largest first onto the least loaded device (the time of a stream grows with its size)
    std::vector<size_t> order(k);
    for (size_t i = 0; i < k; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return in_len[a] != in_len[b] ? in_len[a] > in_len[b] : a < b; });
    std::vector<std::vector<size_t>> share(ndev);
    std::vector<uint64_t> load(ndev, 0);
    for (size_t i : order) {
        size_t best = 0;
        for (size_t d = 1; d < ndev; d++)
            if (load[d] < load[best]) best = d;
        share[best].push_back(i);
        load[best] += in_len[i] + 65536;
    }
    std::vector<int> st(k, FB200_OK);
    std::vector<std::thread> workers;
    for (size_t d = 0; d < ndev; d++) {
        if (share[d].empty()) continue;
        workers.emplace_back([&, d] {
            for (size_t i : share[d]) {
                out_len[i] = 0;
                st[i] = fb200_compress(p->ctxs[d], container, mode, in[i], in_len[i], out[i], out_cap[i], &out_len[i]);
            }
        });
    }
    for (auto& w : workers) w.join();
    int first = FB200_OK;
    for (size_t i = 0; i < k; i++) {
        if (status) status[i] = st[i];
        if (first == FB200_OK && st[i]) first = st[i];
    }
    return first;
}
int fb200_decompress_members_batch(fb200_pool* p, int container, const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len,
                                   size_t k, uint8_t* out, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                                   uint64_t* consumed, int* status) {
    if (!p || (k && (!in || !out || !in_off || !in_len || !out_off || !out_cap || !out_len))) return FB200_INVALID_ARGUMENT;
    const size_t ndev = p->ctxs.size();
    // contiguous ranges of members with about equal compressed bytes
    uint64_t total = 0;
    for (size_t i = 0; i < k; i++) total += in_len[i] + 4096;
    std::vector<size_t> cut(ndev + 1, k);
    cut[0] = 0;
    uint64_t acc = 0;
    size_t d = 1;
    for (size_t i = 0; i < k && d < ndev; i++) {
        acc += in_len[i] + 4096;
        if (acc * ndev >= total * d) cut[d++] = i + 1;
    }
    std::vector<int> rcs(ndev, FB200_OK);
    std::vector<std::thread> workers;
    for (size_t dv = 0; dv < ndev; dv++) {
        const size_t lo = cut[dv], hi = cut[dv + 1];
        if (hi <= lo) continue;
        workers.emplace_back([&, dv, lo, hi] {
            // the members of a range are addressed inside the whole buffers: the device copies only what the range spans
            uint64_t in_lo = ~0ull, out_lo = ~0ull;
            for (size_t i = lo; i < hi; i++) {
                in_lo = in_off[i] < in_lo ? in_off[i] : in_lo;
                out_lo = out_off[i] < out_lo ? out_off[i] : out_lo;
            }
            std::vector<uint64_t> io(hi - lo), oo(hi - lo);
            for (size_t i = lo; i < hi; i++) {
                io[i - lo] = in_off[i] - in_lo;
                oo[i - lo] = out_off[i] - out_lo;
            }
            rcs[dv] = fb200_decompress_members(p->ctxs[dv], container, in + in_lo, io.data(), in_len + lo, hi - lo, out + out_lo, oo.data(),
                                               out_cap + lo, out_len + lo, consumed ? consumed + lo : nullptr, status ? status + lo : nullptr);
        });
    }
    for (auto& w : workers) w.join();
    for (size_t dv = 0; dv < ndev; dv++)
        if (rcs[dv]) return rcs[dv];
    return FB200_OK;
}

void fb200_inflate_rebind(fb200_inflate* s, fb200_read_fn reader, void* user) {
    if (!s) return;
    s->reader = reader;  // same reader object at a new address: the read state is left alone (unlike set_reader)
    s->user = user;
}
void fb200_inflate_destroy(fb200_inflate* s) { delete s; }

}  // extern "C"
