/*
 * flate_oracle.c -- CPU restatement of the ianic/flate hot path.  TEST INFRASTRUCTURE ONLY:
 * see flate_oracle.h.  Never linked into, included by, or called from the shipped CUDA path.
 *
 * The restatement keeps the reference's *sequential streaming* shape on purpose (64 KiB window,
 * u16 head/chain with saturating slide, 32768-token list, 64-bit bit accumulator), so that it is
 * an independent check of the data-parallel re-derivation used on the GPU.
 */
#include "flate_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ consts.zig:1-49 */
enum {
    TOKENS_PER_BLOCK = 1 << 15, /* consts.zig:6 */
    BASE_LENGTH = 3,
    MIN_MATCH = 4,   /* consts.zig:11 */
    MAX_MATCH = 258, /* consts.zig:12 */
    MAX_DIST = 32768,
    HIST_LEN = 32768,
    WIN_LEN = 65536,
    MIN_LOOKAHEAD = MIN_MATCH + MAX_MATCH, /* SlidingWindow.zig:13 */
    MAX_RP = WIN_LEN - MIN_LOOKAHEAD,
    HASH_SHIFT = 17, /* consts.zig:22-26 */
    NUM_LIT = 286,
    NUM_DIST = 30,
    NUM_CODEGEN = 19,
    END_BLOCK = 256,
    MAX_STORE = 65535
};
static const uint8_t codegen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

/* ------------------------------------------------------------------ growable sink */
typedef struct {
    uint8_t* p;
    size_t len, cap;
    int oom;
} Sink;

static void sink_write(Sink* s, const uint8_t* b, size_t n) {
    if (s->len + n > s->cap) {
        size_t nc = s->cap ? s->cap * 2 : 4096;
        while (nc < s->len + n) nc *= 2;
        uint8_t* np = (uint8_t*)realloc(s->p, nc);
        if (!np) {
            s->oom = 1;
            return;
        }
        s->p = np;
        s->cap = nc;
    }
    if (n) memcpy(s->p + s->len, b, n);
    s->len += n;
}

/* ------------------------------------------------------------------ checksums (Zig std.hash.Crc32 / Adler32) */
static uint32_t crc_table[8][256];
static int crc_ready = 0;
static void crc_init(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crc_table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++) crc_table[t][i] = (crc_table[t - 1][i] >> 8) ^ crc_table[0][crc_table[t - 1][i] & 0xff];
    crc_ready = 1;
}
uint32_t fo_crc32(uint32_t crc, const uint8_t* p, size_t n) {
    if (!crc_ready) crc_init();
    uint32_t c = ~crc;
    while (n >= 8) {
        uint32_t a = c ^ ((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24);
        uint32_t b = (uint32_t)p[4] | (uint32_t)p[5] << 8 | (uint32_t)p[6] << 16 | (uint32_t)p[7] << 24;
        c = crc_table[7][a & 0xff] ^ crc_table[6][(a >> 8) & 0xff] ^ crc_table[5][(a >> 16) & 0xff] ^
            crc_table[4][a >> 24] ^ crc_table[3][b & 0xff] ^ crc_table[2][(b >> 8) & 0xff] ^
            crc_table[1][(b >> 16) & 0xff] ^ crc_table[0][b >> 24];
        p += 8;
        n -= 8;
    }
    while (n--) c = crc_table[0][(c ^ *p++) & 0xff] ^ (c >> 8);
    return ~c;
}
uint32_t fo_adler32(uint32_t adler, const uint8_t* p, size_t n) {
    uint32_t a = adler & 0xffff, b = adler >> 16;
    while (n) {
        size_t k = n < 5552 ? n : 5552;
        n -= k;
        while (k--) {
            a += *p++;
            b += a;
        }
        a %= 65521;
        b %= 65521;
    }
    return (b << 16) | a;
}

/* container.zig:168-206 Hasher */
typedef struct {
    int container;
    uint32_t state;
    uint64_t bytes;
} Hasher;
static void hasher_init(Hasher* h, int container) {
    h->container = container;
    h->state = (container == FO_ZLIB) ? 1u : 0u;
    h->bytes = 0;
}
static void hasher_update(Hasher* h, const uint8_t* p, size_t n) {
    if (h->container == FO_GZIP) h->state = fo_crc32(h->state, p, n);
    else if (h->container == FO_ZLIB) h->state = fo_adler32(h->state, p, n);
    else return;
    h->bytes += n;
}

/* ------------------------------------------------------------------ bit_writer.zig:46-97 */
typedef struct {
    Sink* w;
    uint64_t bits;
    uint32_t nbits;
} BitWriter;

static void bw_write_bits(BitWriter* b, uint32_t v, uint32_t nb) { /* bit_writer.zig:63-79 */
    b->bits |= (uint64_t)v << b->nbits;
    b->nbits += nb;
    if (b->nbits < 48) return;
    uint8_t six[6];
    for (int i = 0; i < 6; i++) six[i] = (uint8_t)(b->bits >> (8 * i));
    sink_write(b->w, six, 6);
    b->bits >>= 48;
    b->nbits -= 48;
}
static void bw_flush(BitWriter* b) { /* bit_writer.zig:46-61: pads the last byte with zero bits */
    while (b->nbits != 0) {
        uint8_t c = (uint8_t)b->bits;
        sink_write(b->w, &c, 1);
        b->bits >>= 8;
        b->nbits = b->nbits > 8 ? b->nbits - 8 : 0;
    }
    b->bits = 0;
}
static int bw_write_bytes(BitWriter* b, const uint8_t* p, size_t n) { /* bit_writer.zig:81-97 */
    if (b->nbits & 7) return FO_UNFINISHED_BITS;
    while (b->nbits != 0) {
        uint8_t c = (uint8_t)b->bits;
        sink_write(b->w, &c, 1);
        b->bits >>= 8;
        b->nbits -= 8;
    }
    sink_write(b->w, p, n);
    return FO_OK;
}

/* ------------------------------------------------------------------ Token.zig:58-103 (RFC 1951 3.2.5 code tables) */
static uint16_t len_base[29], dist_base[30];
static uint8_t len_extra[29], dist_extra[30];
static uint8_t len_code_of[256];   /* len-3 -> code index 0..28 */
static uint8_t dist_code_small[256]; /* Token.zig:203-220 index for (dist-1) < 256 */
static int tables_ready = 0;

static void tables_init(void) {
    /* length codes 257..284: groups of 4 share an extra-bit count; 285 = 258 with 0 extra */
    int base = 3;
    for (int c = 0; c < 28; c++) {
        int eb = c < 8 ? 0 : (c - 4) / 4;
        len_base[c] = (uint16_t)base;
        len_extra[c] = (uint8_t)eb;
        base += 1 << eb;
    }
    len_base[28] = 258;
    len_extra[28] = 0;
    for (int l = 3; l <= 258; l++) {
        int c = 28;
        if (l < 258) {
            c = 0;
            while (c + 1 < 28 && len_base[c + 1] <= l) c++;
        }
        len_code_of[l - 3] = (uint8_t)c;
    }
    base = 1;
    for (int c = 0; c < 30; c++) {
        int eb = c < 4 ? 0 : (c - 2) / 2;
        dist_base[c] = (uint16_t)base;
        dist_extra[c] = (uint8_t)eb;
        base += 1 << eb;
    }
    for (int d = 0; d < 256; d++) {
        int c = 0;
        while (c + 1 < 30 && (int)dist_base[c + 1] - 1 <= d) c++;
        dist_code_small[d] = (uint8_t)c;
    }
    tables_ready = 1;
}
static inline int tok_is_match(uint32_t t) { return (t & FO_TOK_MATCH) != 0; }
static inline uint32_t tok_len3(uint32_t t) { return t & 0xff; }           /* length - 3 or literal */
static inline uint32_t tok_dist1(uint32_t t) { return (t >> 8) & 0x7fff; } /* distance - 1 */
static inline uint32_t tok_length(uint32_t t) { return tok_len3(t) + BASE_LENGTH; }
/* Token.zig:70-82 distanceCode: two >>7 steps over a 256-entry index */
static inline uint32_t tok_dist_code(uint32_t t) {
    uint32_t d = tok_dist1(t);
    if (d < 256) return dist_code_small[d];
    d >>= 7;
    if (d < 256) return dist_code_small[d] + 14u;
    d >>= 7;
    return dist_code_small[d] + 28u;
}

/* ------------------------------------------------------------------ huffman_encoder.zig */
typedef struct {
    uint16_t code, len;
} HuffCode;
typedef struct {
    uint16_t literal, freq;
} LitNode;

static int by_freq(const void* a, const void* b) { /* huffman_encoder.zig:355-361 */
    const LitNode *x = (const LitNode*)a, *y = (const LitNode*)b;
    if (x->freq == y->freq) return (int)x->literal - (int)y->literal;
    return (int)x->freq - (int)y->freq;
}
static int by_literal(const void* a, const void* b) { /* huffman_encoder.zig:350-353 */
    return (int)((const LitNode*)a)->literal - (int)((const LitNode*)b)->literal;
}
static uint16_t bit_reverse16(uint16_t v, unsigned n) {
    uint16_t r = 0;
    for (unsigned i = 0; i < n; i++)
        if (v & (1u << i)) r |= (uint16_t)(1u << (n - 1 - i));
    return r;
}

#define MAXI32 0x7fffffffu
typedef struct {
    uint32_t level, last_freq, next_char_freq, next_pair_freq, needed;
} LevelInfo;

/* huffman_encoder.zig:122-247 bitCounts.  `list` sorted by (freq, literal); n >= 3.
 * Returns the (possibly reduced) max_bits; bit_count[1..max_bits] filled. */
static uint32_t huff_bit_counts(const LitNode* list, uint32_t n, uint32_t max_bits, uint32_t* bit_count) {
    if (max_bits > n - 1) max_bits = n - 1; /* :131 */
    LevelInfo levels[17];
    uint32_t leaf_counts[17][16];
    memset(levels, 0, sizeof levels);
    memset(leaf_counts, 0, sizeof leaf_counts);
    for (uint32_t level = 1; level <= max_bits; level++) { /* :144-161 */
        levels[level].level = level;
        levels[level].last_freq = list[1].freq;
        levels[level].next_char_freq = list[2].freq;
        levels[level].next_pair_freq = (uint32_t)list[0].freq + (uint32_t)list[1].freq;
        levels[level].needed = 0;
        leaf_counts[level][level] = 2;
        if (level == 1) levels[level].next_pair_freq = MAXI32;
    }
    levels[max_bits].needed = 2 * n - 4; /* :164 */
    uint32_t level = max_bits;
    for (;;) { /* :168-224 */
        LevelInfo* l = &levels[level];
        if (l->next_pair_freq == MAXI32 && l->next_char_freq == MAXI32) { /* :170 (unreachable: leaf sentinel is 65535) */
            l->needed = 0;
            levels[level + 1].next_pair_freq = MAXI32;
            level += 1;
            continue;
        }
        uint32_t prev_freq = l->last_freq;
        if (l->next_char_freq < l->next_pair_freq) { /* :182 leaf */
            uint32_t next = leaf_counts[level][level] + 1;
            l->last_freq = l->next_char_freq;
            leaf_counts[level][level] = next;
            l->next_char_freq = (next >= n) ? 65535u /* maxNode().freq :282-287 */ : list[next].freq;
        } else { /* :193 pair from the level below */
            l->last_freq = l->next_pair_freq;
            memcpy(leaf_counts[level], leaf_counts[level - 1], level * sizeof(uint32_t));
            levels[l->level - 1].needed = 2;
        }
        l->needed -= 1;
        if (l->needed == 0) { /* :204 */
            if (l->level == max_bits) break;
            levels[l->level + 1].next_pair_freq = prev_freq + l->last_freq;
            level += 1;
        } else {
            while (levels[level - 1].needed > 0) { /* :217 */
                level -= 1;
                if (level == 0) break;
            }
        }
    }
    uint32_t bits = 1;
    const uint32_t* counts = leaf_counts[max_bits];
    for (uint32_t lv = max_bits; lv > 0; lv--) { /* :235-245 */
        bit_count[bits] = counts[lv] - counts[lv - 1];
        bits++;
    }
    return max_bits;
}

/* huffman_encoder.zig:62-95 generate + :251-278 assignEncodingAndSize */
static void huff_generate(HuffCode* codes, const uint16_t* freq, int n, uint32_t max_bits) {
    LitNode list[NUM_LIT + 1];
    uint32_t count = 0;
    for (int i = 0; i < n; i++) {
        if (freq[i] != 0) {
            list[count].literal = (uint16_t)i;
            list[count].freq = freq[i];
            count++;
        } else {
            codes[i].len = 0; /* code field keeps its previous value, as in the reference */
        }
    }
    if (count <= 2) { /* :79-87 */
        for (uint32_t i = 0; i < count; i++) {
            codes[list[i].literal].code = (uint16_t)i;
            codes[list[i].literal].len = 1;
        }
        return;
    }
    qsort(list, count, sizeof(LitNode), by_freq); /* total order => any sort agrees with std.mem.sort */
    uint32_t bit_count[17];
    memset(bit_count, 0, sizeof bit_count);
    uint32_t mb = huff_bit_counts(list, count, max_bits, bit_count);
    uint16_t code = 0;
    uint32_t remaining = count;
    for (uint32_t nb = 0; nb <= mb; nb++) { /* :255-277 */
        code = (uint16_t)(code << 1);
        if (nb == 0 || bit_count[nb] == 0) continue;
        uint32_t bits = bit_count[nb];
        LitNode* chunk = list + (remaining - bits);
        qsort(chunk, bits, sizeof(LitNode), by_literal);
        for (uint32_t k = 0; k < bits; k++) {
            codes[chunk[k].literal].code = bit_reverse16(code, nb);
            codes[chunk[k].literal].len = (uint16_t)nb;
            code++;
        }
        remaining -= bits;
    }
}
static uint32_t huff_bit_length(const HuffCode* codes, const uint16_t* freq, int n) { /* :97-105 */
    uint32_t total = 0;
    for (int i = 0; i < n; i++)
        if (freq[i] != 0) total += (uint32_t)freq[i] * codes[i].len;
    return total;
}
void fo_huffman_generate(const uint16_t* freq, int n, int max_bits, uint16_t* codes, uint16_t* lens) {
    HuffCode hc[NUM_LIT];
    memset(hc, 0, sizeof hc);
    huff_generate(hc, freq, n, (uint32_t)max_bits);
    for (int i = 0; i < n; i++) {
        codes[i] = hc[i].code;
        lens[i] = hc[i].len;
    }
}

/* ------------------------------------------------------------------ block_writer.zig */
typedef struct {
    BitWriter bw;
    uint16_t codegen_freq[NUM_CODEGEN];
    uint16_t literal_freq[NUM_LIT];
    uint16_t distance_freq[NUM_DIST];
    uint8_t codegen[NUM_LIT + NUM_DIST + 1];
    HuffCode literal_enc[NUM_LIT], distance_enc[NUM_DIST], codegen_enc[NUM_CODEGEN];
    HuffCode fixed_lit[NUM_LIT], fixed_dist[NUM_DIST], huff_dist[NUM_DIST];
} BlockWriter;

static void blw_init(BlockWriter* b, Sink* s) { /* block_writer.zig:38-45 */
    if (!tables_ready) tables_init();
    memset(b, 0, sizeof *b);
    b->bw.w = s;
    for (int ch = 0; ch < NUM_LIT; ch++) { /* huffman_encoder.zig:298-330 */
        uint16_t bits, size;
        if (ch <= 143) bits = (uint16_t)(ch + 48), size = 8;
        else if (ch <= 255) bits = (uint16_t)(ch + 400 - 144), size = 9;
        else if (ch <= 279) bits = (uint16_t)(ch - 256), size = 7;
        else bits = (uint16_t)(ch + 192 - 280), size = 8;
        b->fixed_lit[ch].code = bit_reverse16(bits, size);
        b->fixed_lit[ch].len = size;
    }
    for (int ch = 0; ch < NUM_DIST; ch++) { /* :332-338 */
        b->fixed_dist[ch].code = bit_reverse16((uint16_t)ch, 5);
        b->fixed_dist[ch].len = 5;
    }
    uint16_t df[NUM_DIST] = {0}; /* :340-348 huffmanDistanceEncoder */
    df[0] = 1;
    huff_generate(b->huff_dist, df, NUM_DIST, 15);
}
static void blw_write_code(BlockWriter* b, HuffCode c) { bw_write_bits(&b->bw, c.code, c.len); }

/* block_writer.zig:78-171 */
static void blw_generate_codegen(BlockWriter* b, uint32_t num_literals, uint32_t num_distances, const HuffCode* lit_enc,
                                 const HuffCode* dist_enc) {
    memset(b->codegen_freq, 0, sizeof b->codegen_freq);
    uint8_t* codegen = b->codegen;
    for (uint32_t i = 0; i < num_literals; i++) codegen[i] = (uint8_t)lit_enc[i].len;
    for (uint32_t i = 0; i < num_distances; i++) codegen[num_literals + i] = (uint8_t)dist_enc[i].len;
    codegen[num_literals + num_distances] = 255;

    uint8_t size = codegen[0];
    int32_t count = 1;
    uint32_t out_index = 0;
    for (uint32_t in_index = 1; size != 255; in_index++) {
        uint8_t next_size = codegen[in_index];
        if (next_size == size) {
            count++;
            continue;
        }
        if (size != 0) {
            codegen[out_index++] = size;
            b->codegen_freq[size]++;
            count--;
            while (count >= 3) {
                int32_t n = count < 6 ? count : 6;
                codegen[out_index++] = 16;
                codegen[out_index++] = (uint8_t)(n - 3);
                b->codegen_freq[16]++;
                count -= n;
            }
        } else {
            while (count >= 11) {
                int32_t n = count < 138 ? count : 138;
                codegen[out_index++] = 18;
                codegen[out_index++] = (uint8_t)(n - 11);
                b->codegen_freq[18]++;
                count -= n;
            }
            if (count >= 3) {
                codegen[out_index++] = 17;
                codegen[out_index++] = (uint8_t)(count - 3);
                b->codegen_freq[17]++;
                count = 0;
            }
        }
        count--;
        for (; count >= 0; count--) {
            codegen[out_index++] = size;
            b->codegen_freq[size]++;
        }
        size = next_size;
        count = 1;
    }
    codegen[out_index] = 255;
}

/* block_writer.zig:179-203 */
static uint32_t blw_dynamic_size(BlockWriter* b, const HuffCode* lit_enc, const HuffCode* dist_enc, uint32_t extra_bits,
                                 uint32_t* num_codegens_out) {
    uint32_t num_codegens = NUM_CODEGEN;
    while (num_codegens > 4 && b->codegen_freq[codegen_order[num_codegens - 1]] == 0) num_codegens--;
    uint32_t header = 3 + 5 + 5 + 4 + 3 * num_codegens + huff_bit_length(b->codegen_enc, b->codegen_freq, NUM_CODEGEN) +
                      (uint32_t)b->codegen_freq[16] * 2 + (uint32_t)b->codegen_freq[17] * 3 +
                      (uint32_t)b->codegen_freq[18] * 7;
    *num_codegens_out = num_codegens;
    return header + huff_bit_length(lit_enc, b->literal_freq, NUM_LIT) +
           huff_bit_length(dist_enc, b->distance_freq, NUM_DIST) + extra_bits;
}
static uint32_t blw_fixed_size(BlockWriter* b, uint32_t extra_bits) { /* :206-211 */
    return 3 + huff_bit_length(b->fixed_lit, b->literal_freq, NUM_LIT) +
           huff_bit_length(b->fixed_dist, b->distance_freq, NUM_DIST) + extra_bits;
}
static void blw_dynamic_header(BlockWriter* b, uint32_t num_literals, uint32_t num_distances, uint32_t num_codegens,
                               int eof) { /* :237-281 */
    bw_write_bits(&b->bw, eof ? 5 : 4, 3);
    bw_write_bits(&b->bw, num_literals - 257, 5);
    bw_write_bits(&b->bw, num_distances - 1, 5);
    bw_write_bits(&b->bw, num_codegens - 4, 4);
    for (uint32_t i = 0; i < num_codegens; i++) bw_write_bits(&b->bw, b->codegen_enc[codegen_order[i]].len, 3);
    uint32_t i = 0;
    for (;;) {
        uint32_t cw = b->codegen[i++];
        if (cw == 255) break;
        blw_write_code(b, b->codegen_enc[cw]);
        if (cw == 16) bw_write_bits(&b->bw, b->codegen[i++], 2);
        else if (cw == 17) bw_write_bits(&b->bw, b->codegen[i++], 3);
        else if (cw == 18) bw_write_bits(&b->bw, b->codegen[i++], 7);
    }
}
static int blw_stored_block(BlockWriter* b, const uint8_t* input, size_t len, int eof) { /* :283-291, 385-388 */
    bw_write_bits(&b->bw, eof ? 1 : 0, 3);
    bw_flush(&b->bw);
    bw_write_bits(&b->bw, (uint32_t)len & 0xffff, 16);
    bw_write_bits(&b->bw, (~(uint32_t)len) & 0xffff, 16);
    return bw_write_bytes(&b->bw, input, len);
}
/* block_writer.zig:444-488 */
static void blw_index_tokens(BlockWriter* b, const uint32_t* tokens, size_t ntok, uint32_t* num_literals,
                             uint32_t* num_distances) {
    memset(b->literal_freq, 0, sizeof b->literal_freq);
    memset(b->distance_freq, 0, sizeof b->distance_freq);
    for (size_t i = 0; i < ntok; i++) {
        uint32_t t = tokens[i];
        if (!tok_is_match(t)) {
            b->literal_freq[t & 0xff]++;
            continue;
        }
        b->literal_freq[257 + len_code_of[tok_len3(t)]]++;
        b->distance_freq[tok_dist_code(t)]++;
    }
    b->literal_freq[END_BLOCK]++;
    uint32_t nl = NUM_LIT;
    while (b->literal_freq[nl - 1] == 0) nl--;
    uint32_t nd = NUM_DIST;
    while (nd > 0 && b->distance_freq[nd - 1] == 0) nd--;
    if (nd == 0) {
        b->distance_freq[0] = 1;
        nd = 1;
    }
    huff_generate(b->literal_enc, b->literal_freq, NUM_LIT, 15);
    huff_generate(b->distance_enc, b->distance_freq, NUM_DIST, 15);
    *num_literals = nl;
    *num_distances = nd;
}
/* block_writer.zig:492-520 */
static void blw_write_tokens(BlockWriter* b, const uint32_t* tokens, size_t ntok, const HuffCode* le, const HuffCode* oe) {
    for (size_t i = 0; i < ntok; i++) {
        uint32_t t = tokens[i];
        if (!tok_is_match(t)) {
            blw_write_code(b, le[t & 0xff]);
            continue;
        }
        uint32_t lc = len_code_of[tok_len3(t)];
        blw_write_code(b, le[257 + lc]);
        if (len_extra[lc]) bw_write_bits(&b->bw, tok_length(t) - len_base[lc], len_extra[lc]);
        uint32_t dc = tok_dist_code(t);
        blw_write_code(b, oe[dc]);
        if (dist_extra[dc]) bw_write_bits(&b->bw, tok_dist1(t) + 1 - dist_base[dc], dist_extra[dc]);
    }
    blw_write_code(b, le[END_BLOCK]);
}
/* block_writer.zig:307-383 */
static int blw_write(BlockWriter* b, const uint32_t* tokens, size_t ntok, int eof, const uint8_t* input, size_t input_len,
                     int has_input) {
    uint32_t num_literals, num_distances;
    blw_index_tokens(b, tokens, ntok, &num_literals, &num_distances);
    uint32_t extra_bits = 0;
    int storable = has_input && input_len <= MAX_STORE; /* :221-229 */
    uint32_t stored_size = storable ? (uint32_t)((input_len + 5) * 8) : 0;
    if (storable) { /* :317-334 */
        for (uint32_t lc = 257 + 8; lc < num_literals; lc++)
            extra_bits += (uint32_t)b->literal_freq[lc] * len_extra[lc - 257];
        for (uint32_t dc = 4; dc < num_distances; dc++) extra_bits += (uint32_t)b->distance_freq[dc] * dist_extra[dc];
    }
    const HuffCode* lit = b->fixed_lit;
    const HuffCode* dist = b->fixed_dist;
    uint32_t size = blw_fixed_size(b, extra_bits);
    uint32_t num_codegens = 0;
    blw_generate_codegen(b, num_literals, num_distances, b->literal_enc, b->distance_enc);
    huff_generate(b->codegen_enc, b->codegen_freq, NUM_CODEGEN, 7);
    uint32_t dyn_size = blw_dynamic_size(b, b->literal_enc, b->distance_enc, extra_bits, &num_codegens);
    if (dyn_size < size) {
        size = dyn_size;
        lit = b->literal_enc;
        dist = b->distance_enc;
    }
    if (storable && stored_size < size) return blw_stored_block(b, input, input_len, eof);
    if (lit == b->fixed_lit) bw_write_bits(&b->bw, eof ? 3 : 2, 3); /* :293-300 */
    else blw_dynamic_header(b, num_literals, num_distances, num_codegens, eof);
    blw_write_tokens(b, tokens, ntok, lit, dist);
    return FO_OK;
}
/* block_writer.zig:395-433 */
static int blw_dynamic_block(BlockWriter* b, const uint32_t* tokens, size_t ntok, int eof, const uint8_t* input,
                             size_t input_len, int has_input) {
    uint32_t num_literals, num_distances, num_codegens;
    blw_index_tokens(b, tokens, ntok, &num_literals, &num_distances);
    blw_generate_codegen(b, num_literals, num_distances, b->literal_enc, b->distance_enc);
    huff_generate(b->codegen_enc, b->codegen_freq, NUM_CODEGEN, 7);
    uint32_t size = blw_dynamic_size(b, b->literal_enc, b->distance_enc, 0, &num_codegens);
    int storable = has_input && input_len <= MAX_STORE;
    uint32_t ssize = storable ? (uint32_t)((input_len + 5) * 8) : 0;
    if (storable && ssize < (size + (size >> 4))) return blw_stored_block(b, input, input_len, eof);
    blw_dynamic_header(b, num_literals, num_distances, num_codegens, eof);
    blw_write_tokens(b, tokens, ntok, b->literal_enc, b->distance_enc);
    return FO_OK;
}
/* block_writer.zig:524-585 */
static int blw_huffman_block(BlockWriter* b, const uint8_t* input, size_t len, int eof) {
    memset(b->literal_freq, 0, sizeof b->literal_freq);
    for (size_t i = 0; i < len; i++) b->literal_freq[input[i]]++;
    b->literal_freq[END_BLOCK] = 1;
    const uint32_t num_literals = END_BLOCK + 1, num_distances = 1;
    b->distance_freq[0] = 1; /* other entries are whatever they were; huff_dist lens there are 0 */
    huff_generate(b->literal_enc, b->literal_freq, NUM_LIT, 15);
    uint32_t num_codegens = 0;
    blw_generate_codegen(b, num_literals, num_distances, b->literal_enc, b->huff_dist);
    huff_generate(b->codegen_enc, b->codegen_freq, NUM_CODEGEN, 7);
    uint32_t size = blw_dynamic_size(b, b->literal_enc, b->huff_dist, 0, &num_codegens);
    int storable = len <= MAX_STORE;
    uint32_t ssize = storable ? (uint32_t)((len + 5) * 8) : 0;
    if (storable && ssize < (size + (size >> 4))) return blw_stored_block(b, input, len, eof);
    blw_dynamic_header(b, num_literals, num_distances, num_codegens, eof);
    for (size_t i = 0; i < len; i++) blw_write_code(b, b->literal_enc[input[i]]);
    blw_write_code(b, b->literal_enc[END_BLOCK]);
    return FO_OK;
}

int fo_block_write(int kind, const uint32_t* tokens, size_t ntok, int eof, const uint8_t* input, size_t input_len,
                   int has_input, uint8_t* out, size_t cap, size_t* out_len) {
    Sink s = {0};
    BlockWriter* b = (BlockWriter*)malloc(sizeof *b);
    blw_init(b, &s);
    int rc;
    if (kind == 0) rc = blw_write(b, tokens, ntok, eof, input, input_len, has_input);
    else if (kind == 1) rc = blw_dynamic_block(b, tokens, ntok, eof, input, input_len, has_input);
    else rc = blw_huffman_block(b, input, input_len, eof);
    bw_flush(&b->bw);
    free(b);
    if (rc == FO_OK) {
        if (s.len > cap) rc = FO_NO_SPACE_LEFT;
        else {
            memcpy(out, s.p, s.len);
            *out_len = s.len;
        }
    }
    free(s.p);
    return rc;
}

/* ------------------------------------------------------------------ container.zig:53-109 */
static void container_write_header(int container, Sink* s) {
    static const uint8_t gz[10] = {0x1f, 0x8b, 0x08, 0, 0, 0, 0, 0, 0, 0x03};
    static const uint8_t zl[2] = {0x78, 0x9c};
    if (container == FO_GZIP) sink_write(s, gz, 10);
    else if (container == FO_ZLIB) sink_write(s, zl, 2);
}
static void container_write_footer(int container, Hasher* h, Sink* s) {
    uint8_t b[8];
    if (container == FO_GZIP) {
        uint32_t c = h->state, n = (uint32_t)h->bytes;
        for (int i = 0; i < 4; i++) b[i] = (uint8_t)(c >> (8 * i)), b[4 + i] = (uint8_t)(n >> (8 * i));
        sink_write(s, b, 8);
    } else if (container == FO_ZLIB) {
        uint32_t c = h->state;
        for (int i = 0; i < 4; i++) b[i] = (uint8_t)(c >> (24 - 8 * i));
        sink_write(s, b, 4);
    }
}

/* ------------------------------------------------------------------ deflate.zig:35-53 LevelArgs */
typedef struct {
    uint16_t good, nice, lazy, chain;
} LevelArgs;
static int level_args(int level, LevelArgs* a) {
    switch (level) {
        case 4: *a = (LevelArgs){4, 16, 4, 16}; return 1;
        case 5: *a = (LevelArgs){8, 32, 16, 32}; return 1;
        case 6: *a = (LevelArgs){8, 128, 16, 128}; return 1;
        case 7: *a = (LevelArgs){8, 128, 32, 256}; return 1;
        case 8: *a = (LevelArgs){32, 258, 128, 1024}; return 1;
        case 9: *a = (LevelArgs){32, 258, 258, 4096}; return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ the streaming compressor */
typedef int (*TokenSinkFn)(void* ctx, const uint32_t* tokens, size_t ntok, int eof, const uint8_t* input, size_t input_len,
                           int has_input);

struct fo_deflate {
    int container, mode;
    LevelArgs level;
    Sink out;
    Hasher hasher;
    BlockWriter blw;
    /* Lookup.zig:16-18 */
    uint16_t head[1 << 15];
    uint16_t chain[WIN_LEN];
    /* SlidingWindow.zig:18-21 */
    uint8_t win[WIN_LEN];
    size_t wp, rp;
    ptrdiff_t fp;
    /* deflate.zig:376-396 */
    uint32_t tokens[TOKENS_PER_BLOCK];
    size_t ntok;
    int has_prev_match, has_prev_literal;
    uint32_t prev_match;
    uint8_t prev_literal;
    /* SimpleCompressor buffer, deflate.zig:456-457 */
    uint8_t sbuf[MAX_STORE];
    size_t swp;
    /* test seam (BlockWriterType, deflate.zig:118-121) */
    TokenSinkFn token_sink;
    void* token_ctx;
    int err;
};

static inline uint32_t hash4(const uint8_t* b) { /* Lookup.zig:75-84 */
    uint32_t v = (uint32_t)b[3] | (uint32_t)b[2] << 8 | (uint32_t)b[1] << 16 | (uint32_t)b[0] << 24;
    return (v * 0x9E3779B1u) >> HASH_SHIFT;
}
static inline uint16_t lookup_set(fo_deflate* d, uint32_t h, uint16_t pos) { /* Lookup.zig:35-40 */
    uint16_t p = d->head[h];
    d->head[h] = pos;
    d->chain[pos] = p;
    return p;
}
static inline uint16_t lookup_add(fo_deflate* d, const uint8_t* data, size_t len, uint16_t pos) { /* Lookup.zig:23-27 */
    if (len < 4) return 0;
    return lookup_set(d, hash4(data), pos);
}
static void lookup_bulk_add(fo_deflate* d, const uint8_t* data, size_t dlen, uint16_t len, uint16_t pos) { /* :55-72 */
    if (len == 0 || dlen < MIN_MATCH) return;
    lookup_set(d, hash4(data), pos);
    size_t end = (size_t)len + 3 < dlen ? (size_t)len + 3 : dlen;
    uint16_t i = pos;
    for (size_t j = 4; j < end; j++) {
        i++;
        lookup_set(d, hash4(data + j - 3), i);
    }
}
static void lookup_slide(fo_deflate* d, uint16_t n) { /* Lookup.zig:43-51 */
    for (size_t i = 0; i < (1 << 15); i++) d->head[i] = d->head[i] > n ? (uint16_t)(d->head[i] - n) : 0;
    for (size_t i = 0; i < n; i++) d->chain[i] = d->chain[i + n] > n ? (uint16_t)(d->chain[i + n] - n) : 0;
}
/* SlidingWindow.zig:81-104 */
static uint16_t win_match(fo_deflate* d, uint16_t prev_pos, uint16_t curr_pos, uint16_t min_len) {
    size_t max_len = d->wp - curr_pos;
    if (max_len > MAX_MATCH) max_len = MAX_MATCH;
    const uint8_t* prev_lh = d->win + prev_pos;
    const uint8_t* curr_lh = d->win + curr_pos;
    size_t i = min_len;
    if (i > 0) {
        if (max_len <= i) return 0;
        for (;;) {
            if (prev_lh[i] != curr_lh[i]) return 0;
            if (i == 0) break;
            i--;
        }
        i = min_len;
    }
    while (i < max_len && prev_lh[i] == curr_lh[i]) i++;
    return i >= MIN_MATCH ? (uint16_t)i : 0;
}

static int flush_tokens(fo_deflate* d, int flush_opt);
static int add_token(fo_deflate* d, uint32_t t) { /* deflate.zig:227-230 */
    d->tokens[d->ntok++] = t;
    if (d->ntok == TOKENS_PER_BLOCK) return flush_tokens(d, 0);
    return FO_OK;
}
static int add_prev_literal(fo_deflate* d) { /* :214-216 */
    if (d->has_prev_literal) return add_token(d, d->prev_literal);
    return FO_OK;
}
static int add_match(fo_deflate* d, uint32_t m, uint16_t* len) { /* :220-225 */
    int rc = add_token(d, m);
    d->has_prev_literal = 0;
    d->has_prev_match = 0;
    *len = (uint16_t)tok_length(m);
    return rc;
}
/* deflate.zig:233-266 */
static int find_match(fo_deflate* d, uint16_t pos, size_t lh_len, uint16_t min_len, uint32_t* match) {
    uint16_t len = min_len;
    uint16_t prev_pos = lookup_add(d, d->win + pos, lh_len, pos);
    int found = 0;
    size_t chain = d->level.chain;
    if (len >= d->level.good) chain >>= 2;
    for (; prev_pos > 0 && chain > 0; chain--) {
        uint32_t distance = (uint32_t)pos - prev_pos;
        if (distance > MAX_DIST) break;
        uint16_t new_len = win_match(d, prev_pos, pos, len);
        if (new_len > len) {
            *match = fo_tok_match(distance, new_len);
            found = 1;
            if (new_len >= d->level.nice) return 1;
            len = new_len;
        }
        prev_pos = d->chain[prev_pos];
    }
    return found;
}
/* deflate.zig:268-288; flush_opt: 0 none, 1 flush, 2 final */
static int flush_tokens(fo_deflate* d, int flush_opt) {
    const uint8_t* input = NULL;
    size_t input_len = 0;
    int has_input = 0;
    if (d->fp >= 0) { /* SlidingWindow.zig:119-123 tokensBuffer */
        input = d->win + d->fp;
        input_len = d->rp - (size_t)d->fp;
        has_input = 1;
    }
    int rc;
    if (d->token_sink) rc = d->token_sink(d->token_ctx, d->tokens, d->ntok, flush_opt == 2, input, input_len, has_input);
    else rc = blw_write(&d->blw, d->tokens, d->ntok, flush_opt == 2, input, input_len, has_input);
    if (rc) return rc;
    if (flush_opt == 1 && !d->token_sink) {
        rc = blw_stored_block(&d->blw, (const uint8_t*)"", 0, 0);
        if (rc) return rc;
    }
    if (flush_opt != 0) bw_flush(&d->blw.bw);
    d->ntok = 0;
    d->fp = (ptrdiff_t)d->rp; /* SlidingWindow.zig:113-115 */
    return FO_OK;
}
/* deflate.zig:154-205 */
static int tokenize(fo_deflate* d, int flush_opt) {
    int should_flush = flush_opt != 0;
    int rc;
    for (;;) {
        size_t lh_len = d->wp - d->rp; /* SlidingWindow.zig:56-60 */
        size_t min = should_flush ? 0 : MIN_LOOKAHEAD;
        if (!(lh_len > min)) break;
        uint16_t step = 1;
        uint16_t pos = (uint16_t)d->rp;
        uint8_t literal = d->win[pos];
        uint16_t min_len = d->has_prev_match ? (uint16_t)tok_length(d->prev_match) : 0;
        uint32_t match;
        if (find_match(d, pos, lh_len, min_len, &match)) {
            if ((rc = add_prev_literal(d))) return rc;
            if (tok_length(match) >= d->level.lazy) {
                if ((rc = add_match(d, match, &step))) return rc;
            } else {
                d->prev_literal = literal;
                d->has_prev_literal = 1;
                d->prev_match = match;
                d->has_prev_match = 1;
            }
        } else {
            if (d->has_prev_match) {
                if ((rc = add_match(d, d->prev_match, &step))) return rc;
                step -= 1;
            } else {
                if ((rc = add_prev_literal(d))) return rc;
                d->prev_literal = literal;
                d->has_prev_literal = 1;
            }
        }
        /* windowAdvance, deflate.zig:207-211 */
        lookup_bulk_add(d, d->win + pos + 1, lh_len - 1, (uint16_t)(step - 1), (uint16_t)(pos + 1));
        d->rp += step;
    }
    if (should_flush) {
        if ((rc = add_prev_literal(d))) return rc;
        d->has_prev_literal = 0;
        if ((rc = flush_tokens(d, flush_opt))) return rc;
    }
    return FO_OK;
}
static void deflate_slide(fo_deflate* d) { /* deflate.zig:291-294, SlidingWindow.zig:36-44 */
    size_t n = d->wp - HIST_LEN;
    memmove(d->win, d->win + HIST_LEN, n);
    d->rp -= HIST_LEN;
    d->wp -= HIST_LEN;
    d->fp -= HIST_LEN;
    lookup_slide(d, (uint16_t)n);
}

fo_deflate* fo_deflate_create(int container, int mode) {
    if (container < 0 || container > 2) return NULL;
    LevelArgs la = {0};
    if (mode != FO_MODE_STORE && mode != FO_MODE_HUFFMAN && !level_args(mode, &la)) return NULL;
    fo_deflate* d = (fo_deflate*)calloc(1, sizeof *d);
    if (!d) return NULL;
    d->container = container;
    d->mode = mode;
    d->level = la;
    hasher_init(&d->hasher, container);
    blw_init(&d->blw, &d->out);
    container_write_header(container, &d->out); /* deflate.zig:144, 470 */
    return d;
}
static int simple_flush_buffer(fo_deflate* d, int final) { /* deflate.zig:486-493 */
    int rc = d->mode == FO_MODE_HUFFMAN ? blw_huffman_block(&d->blw, d->sbuf, d->swp, final)
                                        : blw_stored_block(&d->blw, d->sbuf, d->swp, final);
    d->swp = 0;
    return rc;
}
int fo_deflate_write(fo_deflate* d, const uint8_t* data, size_t n) {
    if (d->err) return d->err;
    int rc = FO_OK;
    if (d->mode < 4) { /* SimpleCompressor.compress, deflate.zig:498-511 */
        for (;;) {
            size_t room = MAX_STORE - d->swp;
            if (room == 0) {
                if ((rc = simple_flush_buffer(d, 0))) break;
                continue;
            }
            size_t k = n < room ? n : room;
            memcpy(d->sbuf + d->swp, data, k);
            hasher_update(&d->hasher, data, k);
            d->swp += k;
            data += k;
            n -= k;
            if (k < room) break;
        }
    } else { /* Deflate.compress, deflate.zig:304-321 */
        for (;;) {
            size_t room = WIN_LEN - d->wp;
            if (room == 0) {
                if ((rc = tokenize(d, 0))) break;
                deflate_slide(d);
                continue;
            }
            size_t k = n < room ? n : room;
            memcpy(d->win + d->wp, data, k);
            hasher_update(&d->hasher, data, k);
            d->wp += k;
            data += k;
            n -= k;
            if ((rc = tokenize(d, 0))) break;
            if (k < room) break;
        }
    }
    if (rc) d->err = rc;
    return rc;
}
int fo_deflate_flush(fo_deflate* d) {
    if (d->err) return d->err;
    int rc;
    if (d->mode < 4) { /* deflate.zig:474-478 */
        rc = simple_flush_buffer(d, 0);
        if (!rc) rc = blw_stored_block(&d->blw, (const uint8_t*)"", 0, 0);
        if (!rc) bw_flush(&d->blw.bw);
    } else {
        rc = tokenize(d, 1); /* deflate.zig:335-337 */
    }
    if (rc) d->err = rc;
    return rc;
}
int fo_deflate_finish(fo_deflate* d) {
    if (d->err) return d->err;
    int rc;
    if (d->mode < 4) { /* deflate.zig:480-484 */
        rc = simple_flush_buffer(d, 1);
        if (!rc) bw_flush(&d->blw.bw);
    } else {
        rc = tokenize(d, 2); /* deflate.zig:344-347 */
    }
    if (!rc && !d->token_sink) container_write_footer(d->container, &d->hasher, &d->out);
    if (rc) d->err = rc;
    return rc;
}
const uint8_t* fo_deflate_output(fo_deflate* d, size_t* len) {
    *len = d->out.len;
    return d->out.p;
}
void fo_deflate_take(fo_deflate* d) { d->out.len = 0; }
void fo_deflate_destroy(fo_deflate* d) {
    if (!d) return;
    free(d->out.p);
    free(d);
}

size_t fo_compress_bound(size_t n) { return n + (n / 32768 + 2) * 16 + 64; }

int fo_compress(int container, int mode, const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
    fo_deflate* d = fo_deflate_create(container, mode);
    if (!d) return FO_INVALID_ARGUMENT;
    int rc = fo_deflate_write(d, in, n);
    if (!rc) rc = fo_deflate_finish(d);
    if (!rc) {
        if (d->out.oom || d->out.len > cap) rc = FO_NO_SPACE_LEFT;
        else {
            memcpy(out, d->out.p, d->out.len);
            *out_len = d->out.len;
        }
    }
    fo_deflate_destroy(d);
    return rc;
}

/* recording block writer: TestTokenWriter (deflate.zig:578-608) */
typedef struct {
    uint32_t* tokens;
    size_t cap, n;
} TokRec;
static int tokrec_sink(void* ctx, const uint32_t* tokens, size_t ntok, int eof, const uint8_t* input, size_t input_len,
                       int has_input) {
    (void)eof;
    (void)input;
    (void)input_len;
    (void)has_input;
    TokRec* r = (TokRec*)ctx;
    for (size_t i = 0; i < ntok; i++) {
        if (r->n < r->cap) r->tokens[r->n] = tokens[i];
        r->n++;
    }
    return FO_OK;
}
int fo_tokenize(int level, const uint8_t* in, size_t n, uint32_t* tokens, size_t cap, size_t* ntok) {
    fo_deflate* d = fo_deflate_create(FO_RAW, level);
    if (!d || level < 4) {
        fo_deflate_destroy(d);
        return FO_INVALID_ARGUMENT;
    }
    TokRec r = {tokens, cap, 0};
    d->token_sink = tokrec_sink;
    d->token_ctx = &r;
    int rc = fo_deflate_write(d, in, n);
    if (!rc) rc = fo_deflate_flush(d);
    fo_deflate_destroy(d);
    *ntok = r.n;
    if (!rc && r.n > cap) rc = FO_NO_SPACE_LEFT;
    return rc;
}

/* Parse-independent match tables.  Runs the reference's own
 * head/chain/window machinery, visiting EVERY position with min_len = 0, once per budget. */
int fo_match_tables(int level, const uint8_t* in, size_t n, uint32_t* r_full, uint32_t* r_quarter) {
    for (int pass = 0; pass < 2; pass++) {
        uint32_t* r = pass == 0 ? r_full : r_quarter;
        fo_deflate* d = fo_deflate_create(FO_RAW, level);
        if (!d || level < 4) {
            fo_deflate_destroy(d);
            return FO_INVALID_ARGUMENT;
        }
        if (pass == 1) d->level.chain >>= 2;
        size_t fed = 0, base = 0;
        for (;;) {
            size_t room = WIN_LEN - d->wp;
            if (room == 0) {
                deflate_slide(d);
                base += HIST_LEN;
                continue;
            }
            size_t k = n - fed < room ? n - fed : room;
            memcpy(d->win + d->wp, in + fed, k);
            d->wp += k;
            fed += k;
            int last = k < room;
            for (;;) {
                size_t lh_len = d->wp - d->rp;
                if (!(lh_len > (last ? 0u : (size_t)MIN_LOOKAHEAD))) break;
                uint32_t m = 0;
                uint32_t packed = 0;
                if (find_match(d, (uint16_t)d->rp, lh_len, 0, &m)) packed = tok_length(m) | (tok_dist1(m) << 9);
                r[base + d->rp] = packed;
                d->rp += 1;
            }
            if (last) break;
        }
        fo_deflate_destroy(d);
    }
    return FO_OK;
}

/* ================================================================== inflate */
typedef struct {
    const uint8_t* p;
    size_t n;      /* total bytes */
    uint64_t bitpos; /* bits consumed */
} BitReader;

static inline uint64_t br_remaining(const BitReader* r) { return (uint64_t)r->n * 8 - r->bitpos; }
/* zero-padded peek of up to 32 bits, LSB first (bit_reader.zig:46-68 leaves zero bits past EOF) */
static inline uint32_t br_peek(const BitReader* r, unsigned nb) {
    uint64_t v = 0;
    size_t byte = (size_t)(r->bitpos >> 3);
    unsigned off = (unsigned)(r->bitpos & 7);
    if (byte + 8 <= r->n) memcpy(&v, r->p + byte, 8); /* little-endian host */
    else
        for (unsigned i = 0; i < 6 && byte + i < r->n; i++) v |= (uint64_t)r->p[byte + i] << (8 * i);
    v >>= off;
    return (uint32_t)(v & ((nb >= 32) ? 0xffffffffu : ((1u << nb) - 1)));
}
/* bit_reader.zig:159-163 shift: EndOfStream if fewer than n bits are left */
static inline int br_shift(BitReader* r, unsigned nb) {
    if (nb > br_remaining(r)) return FO_END_OF_STREAM;
    r->bitpos += nb;
    return FO_OK;
}
/* `fill(nice)` only fails when no bit at all is left (bit_reader.zig:65-66) */
static inline int br_fill_check(const BitReader* r) { return br_remaining(r) == 0 ? FO_END_OF_STREAM : FO_OK; }
/* read(U): fill + shift (bit_reader.zig:104-109) */
static inline int br_read(BitReader* r, unsigned nb, uint32_t* v) {
    int rc;
    if ((rc = br_fill_check(r))) return rc;
    *v = br_peek(r, nb);
    return br_shift(r, nb);
}
static inline int br_read_buffered(BitReader* r, unsigned nb, uint32_t* v) { /* flag.buffered: no fill */
    *v = br_peek(r, nb);
    return br_shift(r, nb);
}
static inline void br_align(BitReader* r) { r->bitpos = (r->bitpos + 7) & ~(uint64_t)7; } /* :172-176 */
static inline uint32_t rev_bits(uint32_t v, unsigned n) {
    uint32_t o = 0;
    for (unsigned i = 0; i < n; i++) o |= ((v >> i) & 1u) << (n - 1 - i);
    return o;
}

static inline uint32_t rev_bits(uint32_t v, unsigned n);
/* huffman_decoder.zig: canonical code over (code_bits, alphabet index); the 9-bit table + linked
 * lists (:64-118) are not observable, only code assignment order, completeness rules and the
 * InvalidCode-on-miss behaviour of find (:156-175). */
typedef struct {
    uint16_t count[16];
    uint16_t symbol[NUM_LIT];
    int max_bits;
    /* direct table over the first 9 stream bits (the reference uses lookup_bits = 9 for the literal and
     * distance decoders, huffman_decoder.zig:31-33): symbol << 4 | code length, 0 = longer code / no code */
    uint16_t fast[512];
} HuffDec;

/* huffman_decoder.zig:126-153 checkCompletnes; alphabet 286 => lit, max_code_bits 15 or 7 */
static int hd_generate(HuffDec* h, const uint8_t* lens, int n, int alphabet, int max_code_bits) {
    if (alphabet == 286 && lens[256] == 0) return FO_MISSING_END_OF_BLOCK_CODE;
    memset(h->count, 0, sizeof h->count);
    h->max_bits = max_code_bits;
    int max = 0;
    for (int i = 0; i < n; i++) {
        if (lens[i] == 0) continue;
        if (lens[i] > max) max = lens[i];
        h->count[lens[i]]++;
    }
    if (max != 0) {
        int left = 1;
        for (int len = 1; len <= max_code_bits; len++) {
            left <<= 1;
            if (h->count[len] > left) return FO_OVERSUBSCRIBED_HUFFMAN_TREE;
            left -= h->count[len];
        }
        if (left > 0) {
            if (!(max_code_bits > 7 && max == h->count[1])) return FO_INCOMPLETE_HUFFMAN_TREE; /* count[0] is 0 */
        }
    }
    uint16_t offs[17];
    offs[1] = 0;
    for (int len = 1; len < 16; len++) offs[len + 1] = (uint16_t)(offs[len] + h->count[len]);
    for (int i = 0; i < n; i++)
        if (lens[i]) h->symbol[offs[lens[i]]++] = (uint16_t)i;
    memset(h->fast, 0, sizeof h->fast);
    {
        unsigned code = 0, index = 0;
        for (int len = 1; len <= 9 && len <= max_code_bits; len++) {
            for (unsigned k = 0; k < h->count[len]; k++) {
                unsigned rev = rev_bits(code + k, (unsigned)len);
                for (unsigned e = rev; e < 512; e += 1u << len) h->fast[e] = (uint16_t)((h->symbol[index + k] << 4) | len);
            }
            code = (code + h->count[len]) << 1;
            index += h->count[len];
        }
    }
    return FO_OK;
}
/* find on the zero-padded peek; returns symbol index and code length, or InvalidCode */
static int hd_find(const HuffDec* h, uint32_t peek_lsb_first, int* sym, int* nbits) {
    const uint16_t e = h->fast[peek_lsb_first & 511];
    if (e) {
        *sym = e >> 4;
        *nbits = e & 15;
        return FO_OK;
    }
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= h->max_bits; len++) {
        code |= (int)(peek_lsb_first & 1);
        peek_lsb_first >>= 1;
        int count = h->count[len];
        if (code - count < first) {
            *sym = h->symbol[index + (code - first)];
            *nbits = len;
            return FO_OK;
        }
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return FO_INVALID_CODE;
}

typedef struct {
    BitReader br;
    uint8_t* out;    /* out[-hist .. 0) is earlier history (may be empty) */
    size_t hist;     /* bytes of earlier output reachable before out */
    size_t pos, cap; /* bytes produced / capacity */
    HuffDec lit, dst;
    int no_space;
} Inflate;

static inline int out_literal(Inflate* s, uint8_t c) {
    if (s->pos >= s->cap) return FO_NO_SPACE_LEFT;
    s->out[s->pos++] = c;
    return FO_OK;
}
/* CircularBuffer.zig:44-75 writeMatch */
static int out_match(Inflate* s, uint32_t length, uint32_t distance) {
    if (s->hist + s->pos < distance || length < BASE_LENGTH || length > MAX_MATCH || distance < 1 || distance > MAX_DIST)
        return FO_INVALID_MATCH;
    if (s->pos + length > s->cap) return FO_NO_SPACE_LEFT;
    uint8_t* to = s->out + s->pos;
    const uint8_t* from = to - distance;
    for (uint32_t i = 0; i < length; i++) to[i] = from[i];
    s->pos += length;
    return FO_OK;
}
static int decode_length(Inflate* s, uint32_t code, uint32_t* length) { /* inflate.zig:126-133 */
    if (code > 28) return FO_INVALID_CODE;
    *length = len_base[code];
    if (len_extra[code]) {
        uint32_t e;
        int rc = br_read_buffered(&s->br, len_extra[code], &e);
        if (rc) return rc;
        *length += e;
    }
    return FO_OK;
}
static int decode_distance(Inflate* s, uint32_t code, uint32_t* distance) { /* inflate.zig:135-142 */
    if (code > 29) return FO_INVALID_CODE;
    *distance = dist_base[code];
    if (dist_extra[code]) {
        uint32_t e;
        int rc = br_read_buffered(&s->br, dist_extra[code], &e);
        if (rc) return rc;
        *distance += e;
    }
    return FO_OK;
}
static int inf_stored(Inflate* s) { /* inflate.zig:89-102 */
    br_align(&s->br);
    uint32_t len, nlen;
    int rc;
    if ((rc = br_read(&s->br, 16, &len))) return rc;
    if ((rc = br_read(&s->br, 16, &nlen))) return rc;
    if (len != ((~nlen) & 0xffff)) return FO_WRONG_STORED_BLOCK_NLEN;
    size_t byte = (size_t)(s->br.bitpos >> 3);
    if (byte + len > s->br.n) return FO_END_OF_STREAM;
    if (s->pos + len > s->cap) return FO_NO_SPACE_LEFT;
    memcpy(s->out + s->pos, s->br.p + byte, len);
    s->pos += len;
    s->br.bitpos += (uint64_t)len * 8;
    return FO_OK;
}
static int read_fixed_code(BitReader* r, uint32_t* code) { /* bit_reader.zig:205-217 */
    int rc;
    uint32_t v, e;
    if ((rc = br_fill_check(r))) return rc;
    if ((rc = br_read_buffered(r, 7, &v))) return rc;
    uint32_t code7 = rev_bits(v, 7);
    if (code7 <= 0x17) {
        *code = code7 + 256;
    } else if (code7 <= 0x5f) {
        if ((rc = br_read_buffered(r, 1, &e))) return rc;
        *code = (code7 << 1) + e - 0x30;
    } else if (code7 <= 0x63) {
        if ((rc = br_read_buffered(r, 1, &e))) return rc;
        *code = ((code7 - 0x60) << 1) + e + 280;
    } else {
        if ((rc = br_read_buffered(r, 2, &e))) return rc;
        *code = ((code7 - 0x64) << 2) + rev_bits(e, 2) + 144;
    }
    return FO_OK;
}
static int inf_fixed(Inflate* s) { /* inflate.zig:104-124 */
    int rc;
    for (;;) {
        uint32_t code;
        if ((rc = read_fixed_code(&s->br, &code))) return rc;
        if (code < 256) {
            if ((rc = out_literal(s, (uint8_t)code))) return rc;
        } else if (code == 256) {
            return FO_OK;
        } else if (code <= 285) {
            uint32_t length, distance, d5;
            if ((rc = br_fill_check(&s->br))) return rc; /* fill(5+5+13) */
            if ((rc = decode_length(s, code - 257, &length))) return rc;
            if ((rc = br_read_buffered(&s->br, 5, &d5))) return rc;
            if ((rc = decode_distance(s, rev_bits(d5, 5), &distance))) return rc;
            if ((rc = out_match(s, length, distance))) return rc;
        } else {
            return FO_INVALID_CODE;
        }
    }
}
static int decode_symbol(Inflate* s, const HuffDec* h, int* sym) { /* inflate.zig:241-245 */
    int nbits;
    int rc = hd_find(h, br_peek(&s->br, (unsigned)h->max_bits), sym, &nbits);
    if (rc) return rc;
    return br_shift(&s->br, (unsigned)nbits);
}
static int dynamic_code_length(Inflate* s, int code, uint8_t* lens, size_t lens_len, size_t pos, size_t* adv) {
    /* inflate.zig:189-216 */
    if (pos >= lens_len) return FO_INVALID_DYNAMIC_BLOCK_HEADER;
    uint32_t v;
    int rc;
    if (code <= 15) {
        lens[pos] = (uint8_t)code;
        *adv = 1;
        return FO_OK;
    }
    if (code == 16) {
        if ((rc = br_read(&s->br, 2, &v))) return rc;
        size_t n = v + 3;
        if (pos == 0 || pos + n > lens_len) return FO_INVALID_DYNAMIC_BLOCK_HEADER;
        for (size_t i = 0; i < n; i++) lens[pos + i] = lens[pos + i - 1];
        *adv = n;
        return FO_OK;
    }
    if (code == 17) {
        if ((rc = br_read(&s->br, 3, &v))) return rc;
        *adv = v + 3;
        return FO_OK;
    }
    if (code == 18) {
        if ((rc = br_read(&s->br, 7, &v))) return rc;
        *adv = v + 11;
        return FO_OK;
    }
    return FO_INVALID_DYNAMIC_BLOCK_HEADER;
}
static int inf_dynamic_header(Inflate* s) { /* inflate.zig:144-185 */
    uint32_t v;
    int rc;
    if ((rc = br_read(&s->br, 5, &v))) return rc;
    uint32_t hlit = v + 257;
    if ((rc = br_read(&s->br, 5, &v))) return rc;
    uint32_t hdist = v + 1;
    if ((rc = br_read(&s->br, 4, &v))) return rc;
    uint32_t hclen = v + 4;
    if (hlit > 286 || hdist > 30) return FO_INVALID_DYNAMIC_BLOCK_HEADER;
    uint8_t cl_lens[19] = {0};
    for (uint32_t i = 0; i < hclen; i++) {
        if ((rc = br_read(&s->br, 3, &v))) return rc;
        cl_lens[codegen_order[i]] = (uint8_t)v;
    }
    HuffDec cl;
    if ((rc = hd_generate(&cl, cl_lens, 19, 19, 7))) return rc;
    uint8_t lit_lens[NUM_LIT] = {0};
    size_t pos = 0, adv;
    int sym;
    while (pos < hlit) {
        if ((rc = br_fill_check(&s->br))) return rc; /* peekF(u7, reverse): fill(7) */
        if ((rc = decode_symbol(s, &cl, &sym))) return rc;
        if ((rc = dynamic_code_length(s, sym, lit_lens, NUM_LIT, pos, &adv))) return rc;
        pos += adv;
    }
    if (pos > hlit) return FO_INVALID_DYNAMIC_BLOCK_HEADER;
    uint8_t dst_lens[NUM_DIST] = {0};
    pos = 0;
    while (pos < hdist) {
        if ((rc = br_fill_check(&s->br))) return rc;
        if ((rc = decode_symbol(s, &cl, &sym))) return rc;
        if ((rc = dynamic_code_length(s, sym, dst_lens, NUM_DIST, pos, &adv))) return rc;
        pos += adv;
    }
    if (pos > hdist) return FO_INVALID_DYNAMIC_BLOCK_HEADER;
    if ((rc = hd_generate(&s->lit, lit_lens, NUM_LIT, 286, 15))) return rc;
    return hd_generate(&s->dst, dst_lens, NUM_DIST, 30, 15);
}
static int inf_dynamic(Inflate* s) { /* inflate.zig:220-239 */
    int rc, sym;
    for (;;) {
        if ((rc = br_fill_check(&s->br))) return rc; /* fill(15) */
        if ((rc = decode_symbol(s, &s->lit, &sym))) return rc;
        if (sym < 256) {
            if ((rc = out_literal(s, (uint8_t)sym))) return rc;
        } else if (sym == 256) {
            return FO_OK;
        } else {
            uint32_t length, distance;
            int dsym;
            if ((rc = br_fill_check(&s->br))) return rc; /* fill(5+15+13) */
            if ((rc = decode_length(s, (uint32_t)sym - 257, &length))) return rc;
            if ((rc = decode_symbol(s, &s->dst, &dsym))) return rc;
            if ((rc = decode_distance(s, (uint32_t)dsym, &distance))) return rc;
            if ((rc = out_match(s, length, distance))) return rc;
        }
    }
}
/* container.zig:111-152 */
static int parse_header(int container, BitReader* r) {
    uint32_t v;
    int rc;
    if (container == FO_GZIP) {
        uint32_t magic1, magic2, method, flags;
        if ((rc = br_read(r, 8, &magic1))) return rc;
        if ((rc = br_read(r, 8, &magic2))) return rc;
        if ((rc = br_read(r, 8, &method))) return rc;
        if ((rc = br_read(r, 8, &flags))) return rc;
        for (int i = 0; i < 6; i++)
            if ((rc = br_read(r, 8, &v))) return rc;
        if (magic1 != 0x1f || magic2 != 0x8b || method != 0x08) return FO_BAD_GZIP_HEADER;
        if (flags & 0x04) {
            uint32_t xlen;
            if ((rc = br_read(r, 16, &xlen))) return rc;
            for (uint32_t i = 0; i < xlen; i++)
                if ((rc = br_read(r, 8, &v))) return rc;
        }
        if (flags & 0x08) do {
                if ((rc = br_read(r, 8, &v))) return rc;
            } while (v != 0);
        if (flags & 0x10) do {
                if ((rc = br_read(r, 8, &v))) return rc;
            } while (v != 0);
        if (flags & 0x02) {
            if ((rc = br_read(r, 8, &v))) return rc;
            if ((rc = br_read(r, 8, &v))) return rc;
        }
    } else if (container == FO_ZLIB) {
        uint32_t cm, cinfo;
        if ((rc = br_read(r, 4, &cm))) return rc;
        if ((rc = br_read(r, 4, &cinfo))) return rc;
        if ((rc = br_read(r, 8, &v))) return rc;
        if (cm != 8 || cinfo > 7) return FO_BAD_ZLIB_HEADER;
    }
    return FO_OK;
}
static int parse_footer(int container, BitReader* r, const uint8_t* out, size_t n) {
    uint32_t v;
    int rc;
    if (container == FO_GZIP) {
        if ((rc = br_read(r, 32, &v))) return rc;
        if (v != fo_crc32(0, out, n)) return FO_WRONG_GZIP_CHECKSUM;
        if ((rc = br_read(r, 32, &v))) return rc;
        if (v != (uint32_t)n) return FO_WRONG_GZIP_SIZE;
    } else if (container == FO_ZLIB) {
        if ((rc = br_read(r, 32, &v))) return rc;
        uint32_t a = fo_adler32(1, out, n);
        uint32_t be = (a >> 24) | ((a >> 8) & 0xff00) | ((a << 8) & 0xff0000) | (a << 24);
        if (v != be) return FO_WRONG_ZLIB_CHECKSUM;
    }
    return FO_OK;
}

int fo_decompress_hist(int container, const uint8_t* in, size_t n, uint8_t* out, size_t hist_len, size_t cap,
                       size_t* out_len, size_t* consumed) {
    if (!tables_ready) tables_init();
    Inflate* s = (Inflate*)calloc(1, sizeof *s);
    s->br.p = in;
    s->br.n = n;
    s->out = out;
    s->hist = hist_len;
    s->cap = cap;
    int rc = parse_header(container, &s->br);
    while (!rc) { /* inflate.zig:251-280 step */
        uint32_t bfinal, btype;
        if ((rc = br_read(&s->br, 1, &bfinal))) break;
        if ((rc = br_read(&s->br, 2, &btype))) break;
        if (btype == 2) {
            if ((rc = inf_dynamic_header(s))) break;
            rc = inf_dynamic(s);
        } else if (btype == 0) rc = inf_stored(s);
        else if (btype == 1) rc = inf_fixed(s);
        else rc = FO_INVALID_BLOCK_TYPE;
        if (rc || bfinal) break;
    }
    if (!rc) {
        br_align(&s->br);
        rc = parse_footer(container, &s->br, out, s->pos);
    }
    if (out_len) *out_len = s->pos;
    if (consumed) *consumed = (size_t)((s->br.bitpos + 7) >> 3);
    free(s);
    return rc;
}
int fo_decompress(int container, const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len, size_t* consumed) {
    return fo_decompress_hist(container, in, n, out, 0, cap, out_len, consumed);
}

const char* fo_strerror(int code) {
    static const char* names[] = {"Ok", "EndOfStream", "InvalidCode", "InvalidMatch", "InvalidBlockType",
                                  "WrongStoredBlockNlen", "InvalidDynamicBlockHeader", "OversubscribedHuffmanTree",
                                  "IncompleteHuffmanTree", "MissingEndOfBlockCode", "BadGzipHeader", "BadZlibHeader",
                                  "WrongGzipChecksum", "WrongGzipSize", "WrongZlibChecksum", "UnfinishedBits",
                                  "InvalidState", "NoSpaceLeft", "InvalidArgument"};
    if (code < 0 || code > 18) return "Unknown";
    return names[code];
}
