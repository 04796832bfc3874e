/*
 * flate_b200.h -- C ABI of the B200-native DEFLATE engine (libflate_b200.so).
 *
 * This is the drop-in boundary for the hot path of ianic/flate: LZ77 match search + Huffman
 * encode/decode behind the gzip/zlib/raw compress()/decompress()/Compressor/Decompressor surface.
 * Plain pointers and sizes only; every call returns a status code that maps 1:1 to the Zig error
 * names of the reference (see fb200_strerror).  INTEGRATION.md shows the Zig `extern` block and
 * the shim that re-creates src/{flate,gzip,zlib}.zig on top of these entry points.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   fb200_compress / fb200_compress_device   <- deflate.zig:56-60  compress(container, reader, writer, options)
 *                                               deflate.zig:402-406 huffman.compress, :421-425 store.compress
 *   fb200_deflate_*                          <- deflate.zig:63-74,138,304,335,344,351,363 Compressor.{init,compress,
 *                                               write,flush,finish,setWriter}; :449-529 SimpleCompressor
 *   fb200_decompress / _device / _members    <- inflate.zig:14-17 decompress(container, reader, writer)
 *   fb200_inflate_*                          <- inflate.zig:20-22,80,283-353 Decompressor.{init,decompress,next,
 *                                               get,read,reset,setReader}
 *   fb200_debug_tokens                       <- the BlockWriterType seam used by the reference's tests
 *                                               (deflate.zig:118-121, TestTokenWriter :578-608)
 *   fb200_debug_block_write                  <- block_writer.zig:307 write, :395 dynamicBlock, :524 huffmanBlock
 *   fb200_debug_match_tables                 <- deflate.zig:233-266 findMatch (per position, both budgets)
 */
#ifndef FLATE_B200_H
#define FLATE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes; 1..16 are the reference's error set (inflate.zig:72-78, huffman_decoder.zig:35-40,
 * container.zig:45-51, bit_writer.zig:35, inflate.zig:302-304) */
enum {
    FB200_OK = 0,
    FB200_END_OF_STREAM = 1,
    FB200_INVALID_CODE = 2,
    FB200_INVALID_MATCH = 3,
    FB200_INVALID_BLOCK_TYPE = 4,
    FB200_WRONG_STORED_BLOCK_NLEN = 5,
    FB200_INVALID_DYNAMIC_BLOCK_HEADER = 6,
    FB200_OVERSUBSCRIBED_HUFFMAN_TREE = 7,
    FB200_INCOMPLETE_HUFFMAN_TREE = 8,
    FB200_MISSING_END_OF_BLOCK_CODE = 9,
    FB200_BAD_GZIP_HEADER = 10,
    FB200_BAD_ZLIB_HEADER = 11,
    FB200_WRONG_GZIP_CHECKSUM = 12,
    FB200_WRONG_GZIP_SIZE = 13,
    FB200_WRONG_ZLIB_CHECKSUM = 14,
    FB200_UNFINISHED_BITS = 15,
    FB200_INVALID_STATE = 16,
    /* ours */
    FB200_NO_SPACE_LEFT = 17, /* caller's output buffer too small (the writer's error in the reference) */
    FB200_INVALID_ARGUMENT = 18,
    FB200_ERR_CUDA = 19,      /* CUDA runtime failure; fb200_last_cuda_error() has the text */
    FB200_NO_DEVICE = 20,     /* no CUDA device: there is NO CPU fallback */
    FB200_RETRY_DENSE = 21    /* sharded search only: see fb200_deflate_shard_search */
};

/* container.zig:17-21 */
enum { FB200_RAW = 0, FB200_GZIP = 1, FB200_ZLIB = 2 };
/* mode: 0 = store (deflate.zig:420), 1 = huffman-only (deflate.zig:401), 4..9 = Level (deflate.zig:23-32);
 * FB200_LEVEL_DEFAULT = 6, fast = 4, best = 9 */
enum { FB200_MODE_STORE = 0, FB200_MODE_HUFFMAN = 1, FB200_LEVEL_FAST = 4, FB200_LEVEL_DEFAULT = 6, FB200_LEVEL_BEST = 9 };

/* token layout of fb200_debug_* : literal = byte; match = 0x80000000 | (distance-1) << 8 | (length-3) */
#define FB200_TOKEN_MATCH 0x80000000u

typedef struct fb200_ctx fb200_ctx;

/* One context per GPU (and per host thread using it).  Owns a stream, the device workspace and
 * pinned staging buffers; the workspace grows to the largest input seen. */
int fb200_ctx_create(int device, fb200_ctx** ctx);
void fb200_ctx_destroy(fb200_ctx* ctx);
int fb200_device_count(void);
const char* fb200_strerror(int code);
const char* fb200_last_cuda_error(void);
/* number of this library's kernels launched by ctx since creation (bench.py's gpu_launches) */
uint64_t fb200_kernel_launches(const fb200_ctx* ctx);
/* Parse strategy of the level modes.  0 (default): f(p), the lazy-parse step of deflate.zig:160-193, is
 * evaluated only on the orbits of seeds placed every few positions (speculative sparse parse), with an exact
 * coverage check and an automatic redo with mode 1 when the check fails; 1: match tables for every position
 * (findMatch deflate.zig:233-266 under both budgets).  Both give the reference's bytes; mode 1 exists for
 * tests and for inputs known to be periodic.  fb200_sparse_fallbacks counts the automatic redos. */
int fb200_ctx_set_parse_mode(fb200_ctx* ctx, int mode);
uint64_t fb200_sparse_fallbacks(const fb200_ctx* ctx);
/* streams whose coverage check failed for some chunks and was repaired by evaluating those chunks densely */
uint64_t fb200_sparse_repairs(const fb200_ctx* ctx);
/* optional per-phase device timing with CUDA events on the launching stream (for bench.py's roofline) */
int fb200_profile_enable(fb200_ctx* ctx, int on); /* also clears the accumulated times */
int fb200_profile_phases(void);
const char* fb200_profile_phase_name(int phase);
int fb200_profile_read(const fb200_ctx* ctx, double* ms, uint64_t* count, int n);

/* ---- one-shot, host buffers (H2D and D2H inside) ---- */
/* Upper bound of the compressed size of n bytes in any mode, container header and footer included. */
size_t fb200_compress_bound(size_t n, int mode);
int fb200_compress(fb200_ctx* ctx, int container, int mode, const uint8_t* in, size_t n, uint8_t* out, size_t cap,
                   size_t* out_len);
/* One member (like the reference's decompress()).  *consumed = bytes of `in` used, so callers can
 * loop over concatenated members the way the reference loops reset() (inflate.zig:301, :544-563). */
int fb200_decompress(fb200_ctx* ctx, int container, const uint8_t* in, size_t n, uint8_t* out, size_t cap,
                     size_t* out_len, size_t* consumed);

/* ---- device-resident variants: d_in/d_out are device pointers on ctx's GPU; `stream` is a
 * cudaStream_t (NULL = the context's own stream).  d_out must hold fb200_compress_bound() bytes and be 16-byte
 * aligned (any cudaMalloc pointer is; FB200_INVALID_ARGUMENT otherwise: the packer writes whole words). ---- */
int fb200_compress_device(fb200_ctx* ctx, int container, int mode, const void* d_in, size_t n, void* d_out, size_t cap,
                          size_t* out_len, void* stream);
/* k independent members (e.g. a multi-member gzip file with a member index): member i occupies
 * d_in[in_off[i] .. in_off[i]+in_len[i]) and is inflated to d_out[out_off[i] .. +out_cap[i]).
 * Per member: status[i], out_len[i], consumed[i] (host arrays).  Returns the first non-OK status. */
int fb200_decompress_members_device(fb200_ctx* ctx, int container, const void* d_in, const uint64_t* in_off,
                                    const uint64_t* in_len, size_t k, void* d_out, const uint64_t* out_off,
                                    const uint64_t* out_cap, uint64_t* out_len, uint64_t* consumed, int* status,
                                    void* stream);
/* same with host buffers */
int fb200_decompress_members(fb200_ctx* ctx, int container, const uint8_t* in, const uint64_t* in_off,
                             const uint64_t* in_len, size_t k, uint8_t* out, const uint64_t* out_off,
                             const uint64_t* out_cap, uint64_t* out_len, uint64_t* consumed, int* status);

/* ---- one stream sharded by position over several GPUs ----
 * The lazy-parse step from a position does not depend on the parse, so any range of positions can be
 * evaluated given a 32 KiB halo of input before it and a little after it.  Stage 1 runs on every rank for its
 * own range [from, to) and writes packed lazy-parse steps (uint32 per position, 0xFFFFFFFF = not evaluated)
 * into d_nx, which is indexed by stream position and holds n entries.
 *   - `from` and `to` multiples of fb200_shard_align() (`to` may also be n): the sparse parse runs; it writes
 *     d_nx[from .. min(n, to + fb200_shard_overlap())).  The entries past `to` continue this range's orbits into
 *     the next rank's range until they join orbits the next rank evaluates itself.
 *     d_in must be readable on [max(0, from - 65536), min(n, to + fb200_shard_overlap() + 8720)): matches reach 32 KiB
 *     back, and the hash chains of that history are rebuilt from another 32 KiB before it.
 *   - otherwise (`from` a multiple of 8192), or after fb200_ctx_set_parse_mode(ctx, 1): dense tables, writes
 *     d_nx[from .. to); d_in readable on [max(0, from - 65536), min(n, to + 8464)).
 * The ranks exchange their tables (NCCL all-gather of the [from, to) parts; for a position covered by two ranks
 * any evaluated entry is the right one) and stage 2 runs on one rank over the joined table: lazy-parse orbit,
 * block cut, Huffman construction, bit-pack.  The result is byte-identical to fb200_compress_device on the
 * same stream.  Either stage returns FB200_RETRY_DENSE when the sparse parse cannot vouch for its coverage
 * (periodic data): every rank then repeats stage 1 in parse mode 1.  Replaces deflate.zig:304-347 for one
 * large stream. */
size_t fb200_shard_align(void);
size_t fb200_shard_overlap(void);
int fb200_deflate_shard_search(fb200_ctx* ctx, int level, const void* d_in, size_t n, size_t from, size_t to, void* d_nx,
                               void* stream);
int fb200_deflate_shard_finish(fb200_ctx* ctx, int container, int level, const void* d_in, size_t n, const void* d_nx,
                               void* d_out, size_t cap, size_t* out_len, void* stream);

/* ---- huffman-only / store stream sharded by 65535-byte block ranges over several GPUs ----
 * The blocks of SimpleCompressor (deflate.zig:449-529) are the 65535-byte slices of the stream and do not depend on
 * each other; only their bit offsets do.  Every rank takes a contiguous range of whole slices (is_last = 0; shard_bytes
 * a multiple of 65535) and the last rank the rest, which ends with the stream's final, possibly empty, slice.
 *   stage 1, fb200_simple_shard_plan: histograms, codes and block sizes of the shard's slices; returns the shard's
 *     size summary: placed at stream bit x the shard ends at bit
 *         has_stored ? ((x + pre_bits + 7) & ~7) + post_bits : x + pre_bits
 *     (a stored block re-aligns to a byte boundary, block_writer.zig:283-291), and with container != raw the
 *     CRC-32 / Adler-32 of the shard's plain bytes (join them with fb200_crc32_combine / fb200_adler32_combine).
 *   exchange: an exclusive scan of the summaries over the ranks (the container header's bits first) gives every
 *     shard's start bit.
 *   stage 2, fb200_simple_shard_pack: packs the shard at its start bit into d_out (16-byte aligned, device), whose
 *     byte 0 is stream byte *byte_lo = (start_bit / 8) & ~15; bits before start_bit are zero, so the stream is the OR
 *     of the shards' bytes: a shard overlaps its predecessor in at most 17 bytes.  cap >= shard_bytes + shard_bytes / 8
 *     + 1024 always suffices.
 * Byte-identical with fb200_compress_device on the whole stream.  Replaces deflate.zig:449-529 for one large
 * huffman-only / store stream. */
int fb200_simple_shard_plan(fb200_ctx* ctx, int container, int mode, const void* d_in, size_t shard_bytes, int is_last,
                            uint64_t* pre_bits, int* has_stored, uint64_t* post_bits, uint32_t* checksum, void* stream);
int fb200_simple_shard_pack(fb200_ctx* ctx, uint64_t start_bit, void* d_out, size_t cap, uint64_t* byte_lo, size_t* nbytes,
                            uint64_t* end_bit, void* stream);
uint32_t fb200_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);
uint32_t fb200_adler32_combine(uint32_t adler1, uint32_t adler2, uint64_t len2);

/* ---- streaming compressor (Compressor / SimpleCompressor) ---- */
typedef struct fb200_deflate fb200_deflate;
typedef int (*fb200_write_fn)(void* user, const uint8_t* data, size_t len); /* the `writer`; non-zero = error */
int fb200_deflate_create(fb200_ctx* ctx, int container, int mode, fb200_write_fn writer, void* user, fb200_deflate** d);
int fb200_deflate_write(fb200_deflate* d, const uint8_t* data, size_t n);
int fb200_deflate_flush(fb200_deflate* d);
int fb200_deflate_finish(fb200_deflate* d);
void fb200_deflate_set_writer(fb200_deflate* d, fb200_write_fn writer, void* user);
void fb200_deflate_destroy(fb200_deflate* d);

/* ---- streaming decompressor (Decompressor) ---- */
typedef struct fb200_inflate fb200_inflate;
typedef size_t (*fb200_read_fn)(void* user, uint8_t* buf, size_t cap); /* the `reader`; 0 = end of input */
int fb200_inflate_create(fb200_ctx* ctx, int container, fb200_read_fn reader, void* user, fb200_inflate** s);
/* borrowed slice valid until the next call; *len == 0 at end of the member (inflate.zig:313-336) */
int fb200_inflate_next(fb200_inflate* s, const uint8_t** data, size_t* len);
int fb200_inflate_get(fb200_inflate* s, size_t limit, const uint8_t** data, size_t* len);
int fb200_inflate_read(fb200_inflate* s, uint8_t* buf, size_t cap, size_t* n);
int fb200_inflate_reset(fb200_inflate* s);
void fb200_inflate_set_reader(fb200_inflate* s, fb200_read_fn reader, void* user);
/* The reader is pulled a chunk at a time, so bytes past the end of the member may have been read: they are kept for
 * the next member (reset) and can be looked at here (valid until the next call on s).  The reference's bit reader
 * holds at most 8 such bytes (bit_reader.zig:18-44). */
int fb200_inflate_unused(fb200_inflate* s, const uint8_t** data, size_t* len);
/* the caller's reader object moved (a by-value host struct, like the reference's Inflate): new callback context,
 * no change of state (fb200_inflate_set_reader restarts the header parse after an end of member, inflate.zig:283-288) */
void fb200_inflate_rebind(fb200_inflate* s, fb200_read_fn reader, void* user);
void fb200_inflate_destroy(fb200_inflate* s);

/* ---- a multi-member gzip file without a member index ----
 * The reference decodes concatenated members one after the other (reset() per member, inflate.zig:301-309): where a
 * member ends is only known once it has been decoded.  Here every position that looks like a member header is found on
 * the device (1f 8b 08, reserved flag bits zero), the file is cut there, every piece is inflated as a member in one
 * launch (its size is the ISIZE field in front of the next cut), and a piece only counts if it decodes without error,
 * its footer checks out and it ends exactly where the next piece starts.  The first piece that does not is the victim
 * of a header look-alike inside compressed data: the cut after it is dropped and the pieces from there on are decoded
 * again.  By induction the accepted pieces are exactly the members the sequential loop finds, so the output (and the
 * first error, if the file is corrupt) is the sequential one.  Bytes after the last member are left alone: *consumed
 * tells where the members end.  out must hold the sum of the members' sizes; when it does not, FB200_NO_SPACE_LEFT is
 * returned and *out_len is the size the file claims (the sum of the ISIZE fields seen).  zlib and raw streams carry no
 * marker to look for and have to go through the sequential Decompressor. */
int fb200_decompress_gzip_file(fb200_ctx* ctx, const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len,
                               size_t* consumed, size_t* members);

/* ---- several GPUs of one node behind one call (SURVEY.md section 8b "batch form", 8e-i) ----
 * A pool owns one context and one host thread per device of the mask.  The k inputs of a batch are independent
 * streams (the reference's compress() called k times, src/gzip.zig:12): they are dealt to the devices, largest first
 * onto the least loaded one, and every device works through its share with fb200_compress (host buffers, copies
 * overlapped inside).  Outputs land in the caller's host buffers, so there is nothing to gather afterwards.
 * status[i] / out_len[i] per item; the call returns the first non-OK status.  The members form splits the member
 * index of fb200_decompress_members into one contiguous range per device. */
typedef struct fb200_pool fb200_pool;
int fb200_pool_create(uint64_t device_mask, fb200_pool** pool); /* bit d = CUDA device d; 0 = every device present */
int fb200_pool_devices(const fb200_pool* pool);
void fb200_pool_destroy(fb200_pool* pool);
int fb200_compress_batch(fb200_pool* pool, int container, int mode, size_t k, const uint8_t* const* in, const size_t* in_len,
                         uint8_t* const* out, const size_t* out_cap, size_t* out_len, int* status);
int fb200_decompress_members_batch(fb200_pool* pool, int container, const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len,
                                   size_t k, uint8_t* out, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                                   uint64_t* consumed, int* status);

/* ---- test seams (all run on the GPU) ---- */
int fb200_debug_tokens(fb200_ctx* ctx, int level, const uint8_t* in, size_t n, uint32_t* tokens, size_t cap,
                       size_t* ntok);
int fb200_debug_match_tables(fb200_ctx* ctx, int level, const uint8_t* in, size_t n, uint32_t* r_full,
                             uint32_t* r_quarter);
/* kind: 0 = write, 1 = dynamicBlock, 2 = huffmanBlock */
int fb200_debug_block_write(fb200_ctx* ctx, int kind, const uint32_t* tokens, size_t ntok, int eof,
                            const uint8_t* input, size_t input_len, int has_input, uint8_t* out, size_t cap,
                            size_t* out_len);

#ifdef __cplusplus
}
#endif
#endif
