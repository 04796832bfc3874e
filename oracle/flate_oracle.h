/*
 * flate_oracle.h -- CPU restatement of the ianic/flate hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle: a plain, single-threaded C restatement of the reference's
 * *sequential* algorithm (64 KiB sliding window, u16 head/chain tables, lazy matching,
 * Go-lineage length-limited Huffman construction, block writer, inflate state machine).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (flate_b200/csrc) never links, includes or calls anything here.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks it against every golden vector the
 * reference's own tests hold for this path (token lists, 36 token counts, 43 byte-exact block
 * encodings, 32 container sizes, Huffman KATs, 40 fuzz error classes, header/footer KATs).
 *
 * Each function cites the reference file:line it follows (paths relative to the reference
 * repository root, i.e. src/flate/...).
 */
#ifndef FLATE_ORACLE_H
#define FLATE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes: same numbering as include/flate_b200.h (1:1 with the Zig error names) */
enum {
    FO_OK = 0,
    FO_END_OF_STREAM = 1,
    FO_INVALID_CODE = 2,
    FO_INVALID_MATCH = 3,
    FO_INVALID_BLOCK_TYPE = 4,
    FO_WRONG_STORED_BLOCK_NLEN = 5,
    FO_INVALID_DYNAMIC_BLOCK_HEADER = 6,
    FO_OVERSUBSCRIBED_HUFFMAN_TREE = 7,
    FO_INCOMPLETE_HUFFMAN_TREE = 8,
    FO_MISSING_END_OF_BLOCK_CODE = 9,
    FO_BAD_GZIP_HEADER = 10,
    FO_BAD_ZLIB_HEADER = 11,
    FO_WRONG_GZIP_CHECKSUM = 12,
    FO_WRONG_GZIP_SIZE = 13,
    FO_WRONG_ZLIB_CHECKSUM = 14,
    FO_UNFINISHED_BITS = 15,
    FO_INVALID_STATE = 16,
    FO_NO_SPACE_LEFT = 17,
    FO_INVALID_ARGUMENT = 18
};

enum { FO_RAW = 0, FO_GZIP = 1, FO_ZLIB = 2 };
/* mode: 0 = store, 1 = huffman-only, 4..9 = deflate levels (deflate.zig:23-53) */
enum { FO_MODE_STORE = 0, FO_MODE_HUFFMAN = 1 };

/* Token layout (ours, not the reference's packed struct): literal = byte value;
 * match = 0x80000000 | (distance-1) << 8 | (length-3). */
#define FO_TOK_MATCH 0x80000000u
static inline uint32_t fo_tok_match(uint32_t dist, uint32_t len) { return FO_TOK_MATCH | ((dist - 1) << 8) | (len - 3); }

/* ---- one-shot ---- */
int fo_compress(int container, int mode, const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len);
size_t fo_compress_bound(size_t n);
/* one member; *consumed = bytes of `in` belonging to the member (header..footer) */
int fo_decompress(int container, const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len,
                  size_t* consumed);
/* like fo_decompress but the decoder may reference `hist_len` bytes of earlier output located
 * directly before `out` (Inflate.reset keeps CircularBuffer.wp, CircularBuffer.zig:44-49). */
int fo_decompress_hist(int container, const uint8_t* in, size_t n, uint8_t* out, size_t hist_len, size_t cap,
                       size_t* out_len, size_t* consumed);

/* ---- streaming compressor (Deflate / SimpleCompressor, deflate.zig:121-373, 449-529) ---- */
typedef struct fo_deflate fo_deflate;
fo_deflate* fo_deflate_create(int container, int mode);
int fo_deflate_write(fo_deflate* d, const uint8_t* data, size_t n);
int fo_deflate_flush(fo_deflate* d);
int fo_deflate_finish(fo_deflate* d);
/* borrow everything written so far; `take` resets the sink (== setWriter to a fresh writer) */
const uint8_t* fo_deflate_output(fo_deflate* d, size_t* len);
void fo_deflate_take(fo_deflate* d);
void fo_deflate_destroy(fo_deflate* d);

/* ---- test seams ---- */
/* Token list produced by compress()+flush() through a recording block writer
 * (TestTokenWriter, deflate.zig:578-608).  Returns token count via *ntok. */
int fo_tokenize(int level, const uint8_t* in, size_t n, uint32_t* tokens, size_t cap, size_t* ntok);
/* findMatch result tables for every position under the closed-form slide schedule (SURVEY §7
 * facts 1-3): packed len | (dist-1) << 9, 0 = none.  r_full uses budget chain, r_quarter
 * chain >> 2.  Used to check the GPU match-search kernel in isolation. */
int fo_match_tables(int level, const uint8_t* in, size_t n, uint32_t* r_full, uint32_t* r_quarter);
/* BlockWriter.{write,dynamicBlock,huffmanBlock} + flush (block_writer.zig:307,395,524).
 * kind: 0 = write, 1 = dynamicBlock, 2 = huffmanBlock.  input may be NULL (has_input = 0). */
int fo_block_write(int kind, const uint32_t* tokens, size_t ntok, int eof, const uint8_t* input, size_t input_len,
                   int has_input, uint8_t* out, size_t cap, size_t* out_len);
/* HuffmanEncoder.generate (huffman_encoder.zig:62-95); codes are bit-reversed as stored there. */
void fo_huffman_generate(const uint16_t* freq, int n, int max_bits, uint16_t* codes, uint16_t* lens);
uint32_t fo_crc32(uint32_t crc, const uint8_t* p, size_t n);
uint32_t fo_adler32(uint32_t adler, const uint8_t* p, size_t n);
const char* fo_strerror(int code);

#ifdef __cplusplus
}
#endif
#endif
