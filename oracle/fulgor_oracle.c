/*
 * fulgor_oracle.c -- TEST INFRASTRUCTURE ONLY (see fulgor_oracle.h). Plain C11, gcc.
 *
 * CPU restatement of the reference's pseudoalignment hot path. Citations are to files under the
 * reference checkout; "sshash/" = external/sshash/, "pthash/" = external/sshash/external/pthash/,
 * "bits/" = external/sshash/external/pthash/external/bits/.
 * Parity: pinned against the reference itself, see the header of fulgor_oracle.h.
 */
#define _GNU_SOURCE
#include "fulgor_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

static _Thread_local char g_err[256];
const char* fo_last_error(void) { return g_err; }

/* ------------------------------------------------------------------ unaligned little-endian */
static inline uint64_t ld64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t ld32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint16_t ld16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

/* ------------------------------------------------------------------ bits/ primitives */
typedef struct { uint64_t num_bits, nwords; const uint8_t* data; } bitvec;       /* bits/include/bit_vector.hpp:343-351 */
typedef struct { uint64_t size, width, mask, nwords; const uint8_t* data; } cvec; /* bits/include/compact_vector.hpp:290-301 */
typedef struct {                                                                  /* bits/include/darray.hpp:146-158 */
    uint64_t num_positions, n_block, n_sub, n_ovf;
    const uint8_t *block_inv, *sub_inv, *ovf;
} darray;
typedef struct { uint64_t back; bitvec high; darray d1, d0; cvec low; } efseq;     /* bits/include/elias_fano.hpp:303-317 */

static inline uint64_t bv_word(const bitvec* b, uint64_t i) { return ld64(b->data + 8 * i); }

/* bits/include/bit_vector.hpp:185-192 */
static inline uint64_t bv_get_word64(const bitvec* b, uint64_t pos) {
    uint64_t block = pos >> 6, shift = pos & 63;
    uint64_t word = bv_word(b, block) >> shift;
    if (shift && block + 1 < b->nwords) word |= bv_word(b, block + 1) << (64 - shift);
    return word;
}

/* bits/include/compact_vector.hpp:239-247 (operator[]; access() at 252-257 returns the same value) */
static inline uint64_t cv_get(const cvec* c, uint64_t i) {
    uint64_t pos = i * c->width, block = pos >> 6, shift = pos & 63;
    uint64_t lo = ld64(c->data + 8 * block) >> shift;
    if (shift + c->width <= 64) return lo & c->mask;
    return (lo | (ld64(c->data + 8 * (block + 1)) << (64 - shift))) & c->mask;
}

/* bits/include/util.hpp:67-105 (select_in_word), portable form */
static inline uint64_t select_in_word(uint64_t w, uint64_t r) {
    for (uint64_t i = 0; i < r; ++i) w &= w - 1;
    return (uint64_t)__builtin_ctzll(w);
}

/* bits/include/darray.hpp:94-119 (select over the ones of B) */
static uint64_t darray_select(const darray* d, const bitvec* B, uint64_t i) {
    uint64_t block = i / 1024;
    int64_t block_pos = (int64_t)ld64(d->block_inv + 8 * block);
    if (block_pos < 0) {
        uint64_t overflow_pos = (uint64_t)(-block_pos - 1);
        return ld64(d->ovf + 8 * (overflow_pos + (i & 1023)));
    }
    uint64_t subblock = i / 32;
    uint64_t start_pos = (uint64_t)block_pos + ld16(d->sub_inv + 2 * subblock);
    uint64_t reminder = i & 31;
    if (!reminder) return start_pos;
    uint64_t word_idx = start_pos >> 6, word_shift = start_pos & 63;
    uint64_t word = bv_word(B, word_idx) & (UINT64_MAX << word_shift);
    for (;;) {
        uint64_t popcnt = (uint64_t)__builtin_popcountll(word);
        if (reminder < popcnt) break;
        reminder -= popcnt;
        word = bv_word(B, ++word_idx);
    }
    return (word_idx << 6) + select_in_word(word, reminder);
}

/* bits/include/elias_fano.hpp:159-163 */
static inline uint64_t ef_size(const efseq* e) { return e->low.size; }
static uint64_t ef_access(const efseq* e, uint64_t i) {
    uint64_t hi = darray_select(&e->d1, &e->high, i) - i;
    return (hi << e->low.width) | (e->low.width ? cv_get(&e->low, i) : 0);
}

/* bits/include/rank9.hpp:92-103,135-146 */
static uint64_t rank9_rank1(const uint8_t* pairs, uint64_t npairs_words, const bitvec* B, uint64_t i) {
    if (i == B->num_bits) return ld64(pairs + 8 * (npairs_words - 2));
    uint64_t sub_block = i >> 6, block = sub_block / 8, left = sub_block % 8;
    uint64_t r = ld64(pairs + 8 * (block * 2));
    r += ld64(pairs + 8 * (block * 2 + 1)) >> ((7 - left) * 9) & 0x1FF;
    uint64_t sub_left = i % 64;
    if (sub_left) r += (uint64_t)__builtin_popcountll(bv_word(B, sub_block) << (64 - sub_left));
    return r;
}

/* LSB-first bit cursor: bits/include/bit_vector.hpp:194-318 (iterator::take / skip_zeros) */
typedef struct { const bitvec* b; uint64_t pos; } bitcur;
static inline uint64_t cur_take(bitcur* c, uint64_t l) {
    if (l == 0) return 0;
    uint64_t w = bv_get_word64(c->b, c->pos);
    c->pos += l;
    return l == 64 ? w : (w & ((1ULL << l) - 1));
}
static inline uint64_t cur_unary(bitcur* c) { /* number of zeros before the next one; consumes the one */
    uint64_t zeros = 0;
    for (;;) {
        uint64_t w = bv_get_word64(c->b, c->pos);
        /* get_word64 pads with zeros past the end; a valid stream always has a one within reach */
        if (w) {
            uint64_t l = (uint64_t)__builtin_ctzll(w);
            c->pos += l + 1;
            return zeros + l;
        }
        c->pos += 64;
        zeros += 64;
    }
}
/* bits/include/integer_codes.hpp:54-71 */
static inline uint64_t cur_gamma(bitcur* c) { uint64_t b = cur_unary(c); return (cur_take(c, b) | (1ULL << b)) - 1; }
static inline uint64_t cur_delta(bitcur* c) { uint64_t b = cur_gamma(c); return (cur_take(c, b) | (1ULL << b)) - 1; }

/* ------------------------------------------------------------------ pthash */
typedef struct {            /* pthash/include/single_phf.hpp:140-150 */
    uint64_t seed, num_keys, table_size;
    u128 M_128;
    uint64_t M_64;
    uint64_t num_dense, num_sparse; /* pthash/include/utils/bucketers.hpp:197-206 */
    u128 M_dense, M_sparse;
    cvec front_ranks, front_dict, back_ranks, back_dict; /* pthash/include/utils/encoders.hpp:207-215,406-414 */
    efseq free_slots;
} single_phf;
typedef struct {            /* pthash/include/partitioned_phf.hpp:203-210,23-43 */
    uint64_t seed, num_keys, table_size, num_partitions_bucketer;
    uint64_t nparts;
    uint64_t* offsets;
    single_phf* parts;
} part_phf;

/* pthash/include/utils/hasher.hpp:53-117, len = 8 */
static inline uint64_t murmur2_64(uint64_t key, uint64_t seed) {
    const uint64_t m = 0xc6a4a7935bd1e995ULL;
    const int r = 47;
    uint64_t h = seed ^ (8 * m);
    uint64_t k = key;
    k *= m; k ^= k >> r; k *= m;
    h ^= k; h *= m;
    h ^= h >> r; h *= m; h ^= h >> r;
    return h;
}
/* pthash/external/fastmod/fastmod.h:159-162 */
static inline uint64_t fastmod_u64(uint64_t a, u128 M, uint64_t d) {
    u128 lowbits = M * a;
    /* mul128_u64(lowbits, d) = ((lowbits * d) >> 128) */
    u128 bottom_half = (lowbits & UINT64_MAX) * d;
    bottom_half >>= 64;
    u128 top_half = (lowbits >> 64) * d;
    u128 both = bottom_half + top_half;
    both >>= 64;
    return (uint64_t)both;
}
/* pthash/include/utils/bucketers.hpp:163-168; T uses the FLOAT constant a = 0.6f (utils/util.hpp:26) */
static inline uint64_t skew_bucket(const single_phf* f, uint64_t hash) {
    static const float a = 0.6f;
    const uint64_t T = (uint64_t)(a * (double)UINT64_MAX);
    return hash < T ? fastmod_u64(hash, f->M_dense, f->num_dense)
                    : f->num_dense + fastmod_u64(hash, f->M_sparse, f->num_sparse);
}
/* pthash/include/utils/encoders.hpp:391-394,192-195 */
static inline uint64_t pilot_access(const single_phf* f, uint64_t i) {
    if (i < f->front_ranks.size) return cv_get(&f->front_dict, cv_get(&f->front_ranks, i));
    return cv_get(&f->back_dict, cv_get(&f->back_ranks, i - f->front_ranks.size));
}
/* pthash/include/single_phf.hpp:79-101 (xor displacement, minimal) */
static uint64_t single_position(const single_phf* f, uint64_t first, uint64_t second) {
    uint64_t bucket = skew_bucket(f, first);
    uint64_t pilot = pilot_access(f, bucket);
    uint64_t hashed_pilot = murmur2_64(pilot, f->seed);
    uint64_t p = fastmod_u64(second ^ hashed_pilot, f->M_128, f->table_size);
    if (p < f->num_keys) return p;
    return ef_access(&f->free_slots, p - f->num_keys);
}
/* pthash/include/partitioned_phf.hpp:150-159; hash128::mix = first ^ second (utils/hasher.hpp:161-163);
   range_bucketer::bucket (utils/bucketers.hpp:216-218) */
static uint64_t part_position(const part_phf* f, uint64_t first, uint64_t second) {
    uint64_t mix = first ^ second;
    uint64_t b = ((mix >> 32) * f->num_partitions_bucketer) >> 32;
    return f->offsets[b] + single_position(&f->parts[b], first, second);
}
/* murmurhash2_128 on a uint64 key (pthash/include/utils/hasher.hpp:203-207); the k<=31 k-mer hasher
   kmers_pthash_hasher_128 (sshash/include/hash_util.hpp:25-41) reduces to the same two calls */
static uint64_t part_lookup(const part_phf* f, uint64_t key) {
    return part_position(f, murmur2_64(key, f->seed), murmur2_64(key, ~f->seed));
}

/* ------------------------------------------------------------------ file parser (essentials visitor layout) */
typedef struct { const uint8_t* p; const uint8_t* end; int bad; } rd;
static const uint8_t* rd_take(rd* r, uint64_t n) {
    if (r->bad || (uint64_t)(r->end - r->p) < n) { r->bad = 1; return r->p; }
    const uint8_t* q = r->p; r->p += n; return q;
}
static uint64_t rd_u64(rd* r) { const uint8_t* q = rd_take(r, 8); return r->bad ? 0 : ld64(q); }
static uint32_t rd_u32(rd* r) { const uint8_t* q = rd_take(r, 4); return r->bad ? 0 : ld32(q); }
static uint16_t rd_u16(rd* r) { const uint8_t* q = rd_take(r, 2); return r->bad ? 0 : ld16(q); }
static uint8_t rd_u8(rd* r) { const uint8_t* q = rd_take(r, 1); return r->bad ? 0 : *q; }
static u128 rd_u128(rd* r) { uint64_t lo = rd_u64(r), hi = rd_u64(r); return ((u128)hi << 64) | lo; }
/* std::vector<POD>: u64 n, then n*elem bytes (essentials.hpp:287-306,109-115) */
static const uint8_t* rd_vec(rd* r, uint64_t elem, uint64_t* n) {
    *n = rd_u64(r);
    if (*n > (uint64_t)(r->end - r->p) / (elem ? elem : 1)) { r->bad = 1; *n = 0; return r->p; }
    return rd_take(r, *n * elem);
}
static void rd_bitvec(rd* r, bitvec* b) { b->num_bits = rd_u64(r); b->data = rd_vec(r, 8, &b->nwords); }
static void rd_cvec(rd* r, cvec* c) {
    c->size = rd_u64(r); c->width = rd_u64(r); c->mask = rd_u64(r); c->data = rd_vec(r, 8, &c->nwords);
}
static void rd_darray(rd* r, darray* d) {
    d->num_positions = rd_u64(r);
    d->block_inv = rd_vec(r, 8, &d->n_block);
    d->sub_inv = rd_vec(r, 2, &d->n_sub);
    d->ovf = rd_vec(r, 8, &d->n_ovf);
}
static void rd_ef(rd* r, efseq* e) {
    e->back = rd_u64(r); rd_bitvec(r, &e->high); rd_darray(r, &e->d1); rd_darray(r, &e->d0); rd_cvec(r, &e->low);
}
static void rd_single(rd* r, single_phf* f) {
    f->seed = rd_u64(r); f->num_keys = rd_u64(r); f->table_size = rd_u64(r);
    f->M_128 = rd_u128(r); f->M_64 = rd_u64(r);
    f->num_dense = rd_u64(r); f->num_sparse = rd_u64(r); f->M_dense = rd_u128(r); f->M_sparse = rd_u128(r);
    rd_cvec(r, &f->front_ranks); rd_cvec(r, &f->front_dict);
    rd_cvec(r, &f->back_ranks); rd_cvec(r, &f->back_dict);
    rd_ef(r, &f->free_slots);
}
static void rd_part(rd* r, part_phf* f) {
    f->seed = rd_u64(r); f->num_keys = rd_u64(r); f->table_size = rd_u64(r);
    f->num_partitions_bucketer = rd_u64(r);
    (void)rd_u128(r); /* range_bucketer::m_M_num_buckets, unused (utils/bucketers.hpp:244-248) */
    f->nparts = rd_u64(r);
    if (f->nparts > (uint64_t)(r->end - r->p)) { r->bad = 1; f->nparts = 0; }
    f->offsets = (uint64_t*)calloc(f->nparts ? f->nparts : 1, sizeof(uint64_t));
    f->parts = (single_phf*)calloc(f->nparts ? f->nparts : 1, sizeof(single_phf));
    for (uint64_t i = 0; i < f->nparts && !r->bad; ++i) {
        f->offsets[i] = rd_u64(r);
        rd_single(r, &f->parts[i]);
    }
}

typedef struct { /* include/color_sets/hybrid.hpp:339-345 */
    uint32_t num_colors, sparse_thr, very_dense_thr;
    efseq offsets;
    bitvec sets;
} hybrid;
static void rd_hybrid(rd* r, hybrid* h) {
    h->num_colors = rd_u32(r); h->sparse_thr = rd_u32(r); h->very_dense_thr = rd_u32(r);
    rd_ef(r, &h->offsets); rd_bitvec(r, &h->sets);
}

typedef struct { /* include/color_sets/differential.hpp:322-339 */
    uint32_t num_colors;
    efseq representative_offsets, color_set_offsets;
    bitvec sets, clusters;
    const uint8_t* rank9;
    uint64_t rank9_words;
} differential;
static void rd_differential(rd* r, differential* d) {
    d->num_colors = rd_u32(r);
    rd_ef(r, &d->representative_offsets); rd_ef(r, &d->color_set_offsets);
    rd_bitvec(r, &d->sets); rd_bitvec(r, &d->clusters);
    d->rank9 = rd_vec(r, 8, &d->rank9_words);
}

struct fo_index {
    uint8_t* buf;
    uint64_t size;
    int type; /* 0 hybrid, 1 meta, 2 differential, 3 meta-differential */
    /* sshash::dictionary (sshash/include/dictionary.hpp:141-154) */
    uint64_t num_kmers, k, m, magic;
    part_phf minimizers;
    efseq pieces, nskb;  /* sshash/include/buckets.hpp:330-336 */
    cvec offsets;
    bitvec strings;
    uint16_t skew_min_log2, skew_max_log2; /* sshash/include/skew_index.hpp:84-91 */
    uint32_t skew_log2_max;
    uint64_t n_skew;
    part_phf* skew_mphfs;
    cvec* skew_positions;
    /* include/index.hpp:94-102 */
    bitvec u2c;
    const uint8_t* rank9;
    uint64_t rank9_words;
    hybrid hyb;
    /* include/color_sets/meta.hpp:275-281 */
    uint32_t meta_num_colors;
    cvec meta_sets;
    efseq meta_offsets;
    uint64_t n_partial;
    hybrid* partial;
    uint64_t n_endpoints;
    const uint8_t* endpoints; /* {u32 min_color, u32 num_color_sets_before} */
    differential dif;         /* .dfur */
    /* include/color_sets/meta_differential.hpp:306-331 (.mdfur); meta_num_colors and n_partial are shared with .mfur */
    uint32_t md_num_partition_sets;
    efseq md_partition_sets_offsets, md_relative_colors_offsets;
    uint64_t md_n_endpoints;
    const uint8_t* md_endpoints; /* {u64 min_color, u64 num_color_sets} per partition */
    differential* md_partial;
    bitvec md_relative_colors, md_partition_sets, md_partition_sets_partitions;
    const uint8_t* md_rank9;
    uint64_t md_rank9_words;
    uint64_t* md_sets_before;   /* prefix sums of num_color_sets (what forward_iterator::read_partition_id accumulates, :181-183) */
};

static int ends_with(const char* s, const char* suf) {
    size_t a = strlen(s), b = strlen(suf);
    return a >= b && !strcmp(s + a - b, suf);
}

fo_index* fo_open(const char* path) {
    int type;
    if (ends_with(path, ".mdfur")) type = 3;
    else if (ends_with(path, ".dfur")) type = 2;
    else if (ends_with(path, ".mfur")) type = 1;
    else if (ends_with(path, ".fur")) type = 0;
    else { snprintf(g_err, sizeof g_err, "unsupported index suffix: %s", path); return NULL; }
    FILE* f = fopen(path, "rb");
    if (!f) { snprintf(g_err, sizeof g_err, "cannot open %s", path); return NULL; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    fo_index* x = (fo_index*)calloc(1, sizeof *x);
    x->buf = (uint8_t*)malloc((size_t)sz + 16);
    x->size = (uint64_t)sz;
    x->type = type;
    if (fread(x->buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); fo_close(x); snprintf(g_err, sizeof g_err, "short read"); return NULL; }
    fclose(f);
    memset(x->buf + sz, 0, 16);
    rd r = {x->buf, x->buf + sz, 0};
    uint8_t major = rd_u8(&r); rd_u8(&r); rd_u8(&r);            /* include/util.hpp:31-35,91-95 */
    if (major != 4) { snprintf(g_err, sizeof g_err, "MAJOR index version mismatch (%u)", major); fo_close(x); return NULL; }
    uint8_t smajor = rd_u8(&r); rd_u8(&r); rd_u8(&r);           /* sshash/include/util.hpp:149-153 */
    if (smajor != 4) { snprintf(g_err, sizeof g_err, "SSHash MAJOR version mismatch (%u)", smajor); fo_close(x); return NULL; }
    x->num_kmers = rd_u64(&r);
    x->k = rd_u16(&r);
    x->m = rd_u16(&r);
    uint8_t canonical = rd_u8(&r);
    x->magic = rd_u64(&r);                                      /* sshash/include/hash_util.hpp:99-107 */
    rd_part(&r, &x->minimizers);
    rd_ef(&r, &x->pieces); rd_ef(&r, &x->nskb); rd_cvec(&r, &x->offsets); rd_bitvec(&r, &x->strings);
    x->skew_min_log2 = rd_u16(&r); x->skew_max_log2 = rd_u16(&r); x->skew_log2_max = rd_u32(&r);
    x->n_skew = rd_u64(&r);
    if (x->n_skew > 64) r.bad = 1;
    if (!r.bad) {
        x->skew_mphfs = (part_phf*)calloc(x->n_skew ? x->n_skew : 1, sizeof(part_phf));
        for (uint64_t i = 0; i < x->n_skew && !r.bad; ++i) rd_part(&r, &x->skew_mphfs[i]);
        uint64_t npos = rd_u64(&r);
        if (npos != x->n_skew) r.bad = 1;
        x->skew_positions = (cvec*)calloc(x->n_skew ? x->n_skew : 1, sizeof(cvec));
        for (uint64_t i = 0; i < x->n_skew && !r.bad; ++i) rd_cvec(&r, &x->skew_positions[i]);
    }
    { cvec w1, w3; efseq w2; rd_cvec(&r, &w1); rd_ef(&r, &w2); rd_cvec(&r, &w3); } /* weights, empty in Fulgor */
    rd_bitvec(&r, &x->u2c);
    x->rank9 = rd_vec(&r, 8, &x->rank9_words);
    if (type == 0) {
        rd_hybrid(&r, &x->hyb);
    } else if (type == 2) {
        rd_differential(&r, &x->dif);
    } else if (type == 3) {
        x->meta_num_colors = rd_u32(&r);
        x->md_num_partition_sets = rd_u32(&r);
        rd_ef(&r, &x->md_partition_sets_offsets); rd_ef(&r, &x->md_relative_colors_offsets);
        x->md_endpoints = rd_vec(&r, 16, &x->md_n_endpoints);
        x->n_partial = rd_u64(&r);
        if (x->n_partial > (uint64_t)(r.end - r.p) || x->n_partial != x->md_n_endpoints) r.bad = 1;
        if (!r.bad) {
            x->md_partial = (differential*)calloc(x->n_partial ? x->n_partial : 1, sizeof(differential));
            for (uint64_t i = 0; i < x->n_partial && !r.bad; ++i) rd_differential(&r, &x->md_partial[i]);
            x->md_sets_before = (uint64_t*)calloc(x->n_partial + 1, sizeof(uint64_t));
            for (uint64_t i = 0; i < x->n_partial; ++i) x->md_sets_before[i + 1] = x->md_sets_before[i] + ld64(x->md_endpoints + 16 * i + 8);
        }
        rd_bitvec(&r, &x->md_relative_colors); rd_bitvec(&r, &x->md_partition_sets); rd_bitvec(&r, &x->md_partition_sets_partitions);
        x->md_rank9 = rd_vec(&r, 8, &x->md_rank9_words);
    } else {
        x->meta_num_colors = rd_u32(&r);
        rd_cvec(&r, &x->meta_sets);
        rd_ef(&r, &x->meta_offsets);
        x->n_partial = rd_u64(&r);
        if (x->n_partial > (uint64_t)(r.end - r.p)) r.bad = 1;
        if (!r.bad) {
            x->partial = (hybrid*)calloc(x->n_partial ? x->n_partial : 1, sizeof(hybrid));
            for (uint64_t i = 0; i < x->n_partial && !r.bad; ++i) rd_hybrid(&r, &x->partial[i]);
        }
        x->endpoints = rd_vec(&r, 8, &x->n_endpoints);
    }
    { uint64_t n1, n2; rd_vec(&r, 4, &n1); rd_vec(&r, 1, &n2); } /* include/filenames.hpp:37-41 */
    if (r.bad || r.p != r.end || !canonical || x->k > 31 || x->m > x->k) {
        snprintf(g_err, sizeof g_err, "malformed index file (bad=%d, consumed %ld of %ld, canonical=%u, k=%lu)",
                 r.bad, (long)(r.p - x->buf), sz, canonical, (unsigned long)x->k);
        fo_close(x);
        return NULL;
    }
    return x;
}

static void free_part(part_phf* f) { free(f->offsets); free(f->parts); }
void fo_close(fo_index* x) {
    if (!x) return;
    free_part(&x->minimizers);
    if (x->skew_mphfs) for (uint64_t i = 0; i < x->n_skew; ++i) free_part(&x->skew_mphfs[i]);
    free(x->skew_mphfs); free(x->skew_positions); free(x->partial); free(x->md_partial); free(x->md_sets_before); free(x->buf); free(x);
}

static uint32_t index_num_colors(const fo_index* x) {
    return x->type == 0 ? x->hyb.num_colors : x->type == 2 ? x->dif.num_colors : x->meta_num_colors;
}
static uint64_t index_num_color_sets(const fo_index* x) {
    if (x->type == 2) return ef_size(&x->dif.color_set_offsets);            /* differential.hpp:297 */
    if (x->type == 3) return ef_size(&x->md_relative_colors_offsets) - 1;   /* meta_differential.hpp:296 */
    return (x->type == 0 ? ef_size(&x->hyb.offsets) : ef_size(&x->meta_offsets)) - 1;
}
void fo_info(const fo_index* x, uint64_t* out) {
    out[0] = x->k; out[1] = x->m; out[2] = x->num_kmers;
    out[3] = ef_size(&x->pieces) - 1; /* num_unitigs = m_k2u.num_contigs(), include/index.hpp:67 (the differential builders' u2c has one bit more) */
   
    out[4] = index_num_colors(x); out[5] = index_num_color_sets(x); out[6] = (uint64_t)x->type;
}

/* ------------------------------------------------------------------ k-mers (sshash/include/kmer.hpp) */
static inline int char_valid(char c) { /* kmer.hpp:214-224,258-260: exactly ACGTacgt */
    switch (c) { case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return 1; default: return 0; }
}
static inline uint64_t char_code(char c) { return ((uint64_t)(unsigned char)c >> 1) & 3; } /* kmer.hpp:199 */
static inline uint64_t mask2(uint64_t n) { return n >= 32 ? UINT64_MAX : ((1ULL << (2 * n)) - 1); }
/* sshash/include/util.hpp:165-171: base j of the string at bits [2j,2j+1] */
static uint64_t string_to_kmer(const char* s, uint64_t k) {
    uint64_t x = 0;
    for (uint64_t i = k; i-- > 0;) x = (x << 2) | char_code(s[i]);
    return x;
}
/* kmer.hpp:146-170 */
static uint64_t revcomp(uint64_t x, uint64_t k) {
    uint64_t c = x ^ 0xaaaaaaaaaaaaaaaaULL;
    uint64_t res = __builtin_bswap64(c);
    const uint64_t c1 = 0x0f0f0f0f0f0f0f0fULL, c2 = 0x3333333333333333ULL;
    res = ((res & c1) << 4) | ((res & (c1 << 4)) >> 4);
    res = ((res & c2) << 2) | ((res & (c2 << 2)) >> 2);
    return res >> (64 - 2 * k);
}
/* sshash/include/hash_util.hpp:97 */
static inline uint64_t mixer(const fo_index* x, uint64_t v) { return (v * 0x517cc1b727220a95ULL) ^ x->magic; }
/* sshash/include/util.hpp:220-239 */
static uint64_t compute_minimizer(const fo_index* x, uint64_t kmer) {
    uint64_t min_hash = UINT64_MAX, minimizer = UINT64_MAX;
    for (uint64_t i = 0; i != x->k - x->m + 1; ++i) {
        uint64_t mmer = kmer & mask2(x->m);
        uint64_t hash = mixer(x, mmer);
        if (hash < min_hash) { min_hash = hash; minimizer = mmer; }
        kmer >>= 2;
    }
    return minimizer;
}
/* sshash/include/util.hpp:206-215 */
static inline uint64_t read_kmer_at(const fo_index* x, uint64_t base_offset) {
    return bv_get_word64(&x->strings, 2 * base_offset) & mask2(x->k);
}

/* ------------------------------------------------------------------ dictionary lookup */
typedef struct { /* sshash/include/util.hpp:32-55 */
    uint64_t kmer_id, kmer_id_in_contig, contig_id, contig_size;
    int64_t orientation;
    int minimizer_found;
} lookup_result;
static lookup_result lr_default(void) {
    lookup_result r = {FO_INVALID, FO_INVALID, FO_INVALID, FO_INVALID, 1, 1};
    return r;
}

/* sshash/include/buckets.hpp:13-40 (offset_to_id): the unitig u with pieces[u] <= offset < pieces[u+1].
   The reference finds it with elias_fano::locate (bits/include/elias_fano.hpp:262-275); a binary search
   over access() returns the same pair. */
static lookup_result offset_to_id(const fo_index* x, uint64_t offset, uint64_t* contig_end) {
    uint64_t lo = 0, hi = ef_size(&x->pieces) - 1; /* invariant: pieces[lo] <= offset < pieces[hi] */
    while (hi - lo > 1) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (ef_access(&x->pieces, mid) <= offset) lo = mid; else hi = mid;
    }
    uint64_t contig_begin = ef_access(&x->pieces, lo);
    *contig_end = ef_access(&x->pieces, lo + 1);
    lookup_result r = lr_default();
    r.kmer_id = offset - lo * (x->k - 1);
    r.kmer_id_in_contig = offset - contig_begin;
    r.contig_id = lo;
    r.contig_size = (*contig_end - contig_begin) - x->k + 1;
    return r;
}

/* sshash/include/buckets.hpp:133-160 */
static lookup_result lookup_canonical_in_super_kmer(const fo_index* x, uint64_t super_kmer_id, uint64_t kmer, uint64_t kmer_rc) {
    uint64_t offset = cv_get(&x->offsets, super_kmer_id), contig_end;
    lookup_result res = offset_to_id(x, offset, &contig_end);
    uint64_t window = x->k - x->m + 1, lim = contig_end - offset - x->k + 1;
    if (lim < window) window = lim;
    for (uint64_t w = 0; w != window; ++w) {
        uint64_t read = read_kmer_at(x, offset + w);
        if (read == kmer) { res.kmer_id += w; res.kmer_id_in_contig += w; res.orientation = 1; return res; }
        if (read == kmer_rc) { res.kmer_id += w; res.kmer_id_in_contig += w; res.orientation = -1; return res; }
    }
    return lr_default();
}

/* sshash/include/buckets.hpp:162-209 */
static lookup_result buckets_lookup_canonical(const fo_index* x, uint64_t begin, uint64_t end, uint64_t kmer, uint64_t kmer_rc, uint64_t minimizer) {
    {
        uint64_t offset = cv_get(&x->offsets, begin);
        uint64_t read = read_kmer_at(x, offset);
        uint64_t a = compute_minimizer(x, read), b = compute_minimizer(x, revcomp(read, x->k));
        if ((a < b ? a : b) != minimizer) { lookup_result r = lr_default(); r.minimizer_found = 0; return r; }
    }
    for (uint64_t s = begin; s != end; ++s) {
        lookup_result r = lookup_canonical_in_super_kmer(x, s, kmer, kmer_rc);
        if (r.kmer_id != FO_INVALID) return r;
    }
    return lr_default();
}

static inline uint64_t ceil_log2_u32(uint64_t v) { /* bits/include/util.hpp ceil_log2_uint32 */
    return v <= 1 ? 0 : 64 - (uint64_t)__builtin_clzll(v - 1);
}

/* sshash/src/dictionary.cpp:47-77 */
static lookup_result lookup_uint_canonical(const fo_index* x, uint64_t kmer, uint64_t kmer_rc, uint64_t minimizer) {
    uint64_t bucket_id = part_lookup(&x->minimizers, minimizer); /* sshash/include/minimizers.hpp:36-39 */
    /* sshash/include/buckets.hpp:62-67 */
    uint64_t begin = ef_access(&x->nskb, bucket_id) + bucket_id;
    uint64_t end = ef_access(&x->nskb, bucket_id + 1) + bucket_id + 1;
    if (x->n_skew != 0) {
        uint64_t n = end - begin, log2n = ceil_log2_u32(n);
        if (log2n > x->skew_min_log2) {
            /* sshash/include/skew_index.hpp:40-52 */
            uint64_t canon = kmer < kmer_rc ? kmer : kmer_rc;
            uint64_t pid = log2n - (x->skew_min_log2 + 1u);
            if (log2n == x->skew_log2_max || log2n > x->skew_max_log2) pid = x->n_skew - 1;
            uint64_t pos = cv_get(&x->skew_positions[pid], part_lookup(&x->skew_mphfs[pid], canon));
            if (pos < n) {
                lookup_result r = lookup_canonical_in_super_kmer(x, begin + pos, kmer, kmer_rc);
                if (r.kmer_id != FO_INVALID) return r;
            }
            return lr_default();
        }
    }
    return buckets_lookup_canonical(x, begin, end, kmer, kmer_rc, minimizer);
}

/* sshash/src/dictionary.cpp:36-45 via lookup_advanced (canonical) */
uint64_t fo_lookup_kmer(const fo_index* x, const char* s) {
    for (uint64_t i = 0; i < x->k; ++i) if (!char_valid(s[i])) return FO_INVALID;
    uint64_t kmer = string_to_kmer(s, x->k), rc = revcomp(kmer, x->k);
    uint64_t a = compute_minimizer(x, kmer), b = compute_minimizer(x, rc);
    lookup_result r = lookup_uint_canonical(x, kmer, rc, a < b ? a : b);
    return r.kmer_id == FO_INVALID ? FO_INVALID : r.contig_id;
}

/* ------------------------------------------------------------------ streaming query (sshash/include/streaming_query.hpp:10-191) */
typedef struct {
    const fo_index* x;
    int start;
    uint64_t kmer, kmer_rc, curr_min, prev_min;
    lookup_result res;
    uint64_t str_pos;   /* base offset in `strings` of the k-mer last matched (kmer_iterator state) */
    uint64_t remaining; /* m_remaining_contig_bases */
} sq;
static void sq_reset(sq* q) { q->start = 1; q->remaining = 0; q->res = lr_default(); }
static void sq_init(sq* q, const fo_index* x) { q->x = x; q->kmer = q->kmer_rc = FO_INVALID; q->curr_min = q->prev_min = FO_INVALID; sq_reset(q); }

static void sq_seed(sq* q) { /* :144-190 */
    const fo_index* x = q->x;
    q->remaining = 0;
    if (q->curr_min == q->prev_min && q->res.minimizer_found == 0) return; /* :150-157 */
    q->res = lookup_uint_canonical(x, q->kmer, q->kmer_rc, q->curr_min);
    if (q->res.kmer_id == FO_INVALID) return;
    q->str_pos = q->res.kmer_id + q->res.contig_id * (x->k - 1);
    q->remaining = (q->res.contig_size - 1) - q->res.kmer_id_in_contig;
    if (q->res.orientation < 0) q->remaining = q->res.kmer_id_in_contig;
}

static lookup_result sq_lookup(sq* q, const char* s) { /* :50-109 */
    const fo_index* x = q->x;
    const uint64_t k = x->k;
    int valid;
    if (q->start) { valid = 1; for (uint64_t i = 0; i < k; ++i) if (!char_valid(s[i])) { valid = 0; break; } }
    else valid = char_valid(s[k - 1]);
    if (!valid) { sq_reset(q); return q->res; }
    if (!q->start) {
        q->kmer = (q->kmer >> 2) | (char_code(s[k - 1]) << (2 * (k - 1)));
        q->kmer_rc = ((q->kmer_rc << 2) | (char_code(s[k - 1]) ^ 2)) & mask2(k);
    } else {
        q->kmer = string_to_kmer(s, k);
        q->kmer_rc = revcomp(q->kmer, k);
    }
    /* minimizer_enumerator::next is asserted equal to compute_minimizer (minimizer_enumerator.hpp:47) */
    uint64_t a = compute_minimizer(x, q->kmer), b = compute_minimizer(x, q->kmer_rc);
    q->curr_min = a < b ? a : b;
    if (q->remaining == 0) {
        sq_seed(q);
    } else {
        /* kmer_iterator next()/next_reverse() (kmer_iterator.hpp:28-55): the next k-mer of the unitig in
           travel direction */
        uint64_t pos = q->res.orientation > 0 ? q->str_pos + 1 : q->str_pos - 1;
        uint64_t expected = read_kmer_at(x, pos);
        if (expected == q->kmer || expected == q->kmer_rc) {
            q->str_pos = pos;
            q->res.kmer_id += (uint64_t)q->res.orientation;
            q->res.kmer_id_in_contig += (uint64_t)q->res.orientation;
            q->remaining -= 1;
        } else {
            sq_seed(q);
        }
    }
    q->prev_min = q->curr_min;
    q->start = 0;
    return q->res;
}

void fo_lookup_read(const fo_index* x, const char* seq, uint64_t len, uint64_t* contig_ids) {
    if (len < x->k) return;
    sq q; sq_init(&q, x);
    for (uint64_t i = 0; i != len - x->k + 1; ++i) {
        lookup_result r = sq_lookup(&q, seq + i);
        contig_ids[i] = r.kmer_id != FO_INVALID ? r.contig_id : FO_INVALID;
    }
}

/* include/index.hpp:37 */
uint64_t fo_u2c(const fo_index* x, uint64_t unitig_id) { return rank9_rank1(x->rank9, x->rank9_words, &x->u2c, unitig_id); }

/* ------------------------------------------------------------------ color sets */
enum { ENC_DELTA = 0, ENC_BITMAP = 1, ENC_COMP = 2 };
typedef struct {
    int enc;
    uint32_t size;   /* number of colors in the set */
    uint32_t n;      /* number of values in vals: the set (delta/bitmap) or its complement (comp) */
    uint32_t* vals;
} decoded;

/* include/color_sets/hybrid.hpp:162-188 (rewind) + :191-236 (next/next_comp); layout :37-95 */
static void hybrid_decode(const hybrid* h, uint64_t id, decoded* d) {
    bitcur c = {&h->sets, ef_access(&h->offsets, id)};
    uint32_t size = (uint32_t)cur_delta(&c);
    d->size = size;
    if (size < h->sparse_thr) {
        d->enc = ENC_DELTA; d->n = size;
        d->vals = (uint32_t*)malloc(sizeof(uint32_t) * (size ? size : 1));
        uint32_t v = 0;
        for (uint32_t i = 0; i < size; ++i) {
            v = i == 0 ? (uint32_t)cur_delta(&c) : v + (uint32_t)cur_delta(&c) + 1;
            d->vals[i] = v;
        }
    } else if (size < h->very_dense_thr) {
        d->enc = ENC_BITMAP; d->n = size;
        d->vals = (uint32_t*)malloc(sizeof(uint32_t) * (size ? size : 1));
        uint32_t j = 0;
        for (uint32_t col = 0; col < h->num_colors; ++col)
            if (bv_get_word64(&h->sets, c.pos + col) & 1) d->vals[j++] = col;
        d->n = j;
    } else {
        d->enc = ENC_COMP; d->n = h->num_colors - size;
        d->vals = (uint32_t*)malloc(sizeof(uint32_t) * (d->n ? d->n : 1));
        uint32_t v = 0;
        for (uint32_t i = 0; i < d->n; ++i) {
            v = i == 0 ? (uint32_t)cur_delta(&c) : v + (uint32_t)cur_delta(&c) + 1;
            d->vals[i] = v;
        }
    }
}
/* materialise the colors of a decoded hybrid set (what forward_iterator::value()/next() enumerate) */
static uint32_t decoded_expand(const decoded* d, uint32_t num_colors, uint32_t add, uint32_t* out) {
    uint32_t n = 0;
    if (d->enc != ENC_COMP) { for (uint32_t i = 0; i < d->n; ++i) out[n++] = d->vals[i] + add; return n; }
    uint32_t j = 0;
    for (uint32_t col = 0; col < num_colors; ++col) {
        if (j < d->n && d->vals[j] == col) { ++j; continue; }
        out[n++] = col + add;
    }
    return n;
}

/* A differential set opened like differential::color_set + forward_iterator::init (differential.hpp:289-295, :256-278):
   the difference list (header: its length, then the size of the decoded set) and the representative of the set's cluster
   (cluster = rank1(m_clusters, id)), both as ascending values. */
typedef struct {
    uint64_t representative_begin;
    uint32_t size;          /* colors in the decoded set */
    uint32_t nd, nr;
    uint32_t *dv, *rv;      /* difference list, representative */
} diffset;
static void gap_list(bitcur* c, uint32_t n, uint32_t* out) {
    uint32_t v = 0;
    for (uint32_t i = 0; i < n; ++i) { v = i == 0 ? (uint32_t)cur_delta(c) : v + (uint32_t)cur_delta(c) + 1; out[i] = v; }
}
static void diff_open(const differential* d, uint64_t id, diffset* s) {
    bitcur cd = {&d->sets, ef_access(&d->color_set_offsets, id)};
    s->representative_begin = ef_access(&d->representative_offsets, rank9_rank1(d->rank9, d->rank9_words, &d->clusters, id));
    bitcur cr = {&d->sets, s->representative_begin};
    s->nd = (uint32_t)cur_delta(&cd);
    s->nr = (uint32_t)cur_delta(&cr);
    s->size = (uint32_t)cur_delta(&cd);
    s->dv = (uint32_t*)malloc(sizeof(uint32_t) * (s->nd ? s->nd : 1));
    s->rv = (uint32_t*)malloc(sizeof(uint32_t) * (s->nr ? s->nr : 1));
    gap_list(&cd, s->nd, s->dv);
    gap_list(&cr, s->nr, s->rv);
}
static void diff_close(diffset* s) { free(s->dv); free(s->rv); }
/* the colors forward_iterator::next / update_curr_val enumerate (:203-217, :280-288): values in exactly one of the lists */
static uint32_t diff_expand(const diffset* s, uint32_t add, uint32_t* out) {
    uint32_t i = 0, j = 0, n = 0;
    while (i < s->nd || j < s->nr) {
        if (j == s->nr || (i < s->nd && s->dv[i] < s->rv[j])) out[n++] = s->dv[i++] + add;
        else if (i == s->nd || s->rv[j] < s->dv[i]) out[n++] = s->rv[j++] + add;
        else { ++i; ++j; }
    }
    return n;
}

/* include/color_sets/meta.hpp:227-235 */
static uint32_t meta_partition_of(const fo_index* x, uint32_t meta_color, uint32_t partition_id) {
    while (partition_id + 1 < x->n_endpoints && meta_color >= ld32(x->endpoints + 8 * (partition_id + 1) + 4)) ++partition_id;
    return partition_id;
}
static uint32_t ep_min_color(const fo_index* x, uint32_t p) { return ld32(x->endpoints + 8 * p); }
static uint32_t ep_sets_before(const fo_index* x, uint32_t p) { return ld32(x->endpoints + 8 * p + 4); }

typedef struct { uint32_t n; uint32_t* mc; uint32_t* part; } meta_list; /* meta colors of one set + their partitions */
static uint32_t md_min_color(const fo_index* x, uint32_t p) { return (uint32_t)ld64(x->md_endpoints + 16 * p); }
static uint32_t md_sets_before(const fo_index* x, uint32_t p) { return (uint32_t)x->md_sets_before[p]; }
static uint64_t msbll(uint64_t v) { return 63 - (uint64_t)__builtin_clzll(v); } /* bits/include/util.hpp msbll */
/* meta_differential::color_set + forward_iterator::init / read_partition_id (meta_differential.hpp:286-293, :126-193): the set's
   group of partitions is the delta-coded partition set number rank1(m_partition_sets_partitions, id) (gaps between partition
   ids), its relative set ids are fields of msb(num_color_sets of the partition) + 1 bits; meta color = sets before + relative */
static void md_list_load(const fo_index* x, uint64_t id, meta_list* l) {
    uint64_t group = rank9_rank1(x->md_rank9, x->md_rank9_words, &x->md_partition_sets_partitions, id);
    bitcur cp = {&x->md_partition_sets, ef_access(&x->md_partition_sets_offsets, group)};
    bitcur cr = {&x->md_relative_colors, ef_access(&x->md_relative_colors_offsets, id)};
    l->n = (uint32_t)cur_delta(&cp);
    l->mc = (uint32_t*)malloc(sizeof(uint32_t) * (l->n ? l->n : 1));
    l->part = (uint32_t*)malloc(sizeof(uint32_t) * (l->n ? l->n : 1));
    uint64_t pid = 0;
    for (uint32_t i = 0; i < l->n; ++i) {
        pid += cur_delta(&cp);
        uint64_t rel = cur_take(&cr, msbll(ld64(x->md_endpoints + 16 * pid + 8)) + 1);
        l->part[i] = (uint32_t)pid;
        l->mc[i] = md_sets_before(x, (uint32_t)pid) + (uint32_t)rel;
    }
}
static void meta_list_load(const fo_index* x, uint64_t id, meta_list* l) {
    if (x->type == 3) { md_list_load(x, id, l); return; }
    uint64_t b = ef_access(&x->meta_offsets, id);
    l->n = (uint32_t)cv_get(&x->meta_sets, b);
    l->mc = (uint32_t*)malloc(sizeof(uint32_t) * (l->n ? l->n : 1));
    l->part = (uint32_t*)malloc(sizeof(uint32_t) * (l->n ? l->n : 1));
    uint32_t p = 0;
    for (uint32_t i = 0; i < l->n; ++i) {
        l->mc[i] = (uint32_t)cv_get(&x->meta_sets, b + 1 + i);
        p = meta_partition_of(x, l->mc[i], p);
        l->part[i] = p;
    }
}
static int meta_find(const meta_list* l, uint32_t part) { for (uint32_t i = 0; i < l->n; ++i) if (l->part[i] == part) return (int)i; return -1; }

int64_t fo_color_set(const fo_index* x, uint64_t id, uint32_t* out, uint64_t cap) {
    uint32_t C = index_num_colors(x);
    uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * (C ? C : 1));
    uint32_t n = 0;
    if (x->type == 0) {
        decoded d; hybrid_decode(&x->hyb, id, &d);
        n = decoded_expand(&d, C, 0, tmp);
        free(d.vals);
    } else if (x->type == 2) {
        diffset s; diff_open(&x->dif, id, &s);
        n = diff_expand(&s, 0, tmp);
        diff_close(&s);
    } else if (x->type == 3) { /* meta_differential.hpp:93-268: the partial (differential) sets, shifted by min_color */
        meta_list l; meta_list_load(x, id, &l);
        for (uint32_t i = 0; i < l.n; ++i) {
            diffset s; diff_open(&x->md_partial[l.part[i]], l.mc[i] - md_sets_before(x, l.part[i]), &s);
            n += diff_expand(&s, md_min_color(x, l.part[i]), tmp + n);
            diff_close(&s);
        }
        free(l.mc); free(l.part);
    } else { /* include/color_sets/meta.hpp:93-236: concatenation of the partial sets, shifted by min_color */
        uint64_t b = ef_access(&x->meta_offsets, id);
        uint32_t sz = (uint32_t)cv_get(&x->meta_sets, b), p = 0;
        for (uint32_t i = 0; i < sz; ++i) {
            uint32_t mc = (uint32_t)cv_get(&x->meta_sets, b + 1 + i);
            p = meta_partition_of(x, mc, p);
            decoded d; hybrid_decode(&x->partial[p], mc - ep_sets_before(x, p), &d);
            n += decoded_expand(&d, x->partial[p].num_colors, ep_min_color(x, p), tmp + n);
            free(d.vals);
        }
    }
    for (uint32_t i = 0; i < n && i < cap; ++i) out[i] = tmp[i];
    free(tmp);
    return (int64_t)n;
}

/* ------------------------------------------------------------------ stage 1: src/ps_full_intersection.cpp:335-374 */
static int cmp_u64(const void* a, const void* b) { uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }
static int cmp_u32(const void* a, const void* b) { uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b; return x < y ? -1 : x > y; }

uint64_t fo_fetch_color_set_ids(const fo_index* x, const char* seq, uint64_t len, uint32_t* out, uint64_t* num_positive) {
    if (num_positive) *num_positive = 0;
    if (len < x->k) return 0;
    uint64_t nk = len - x->k + 1, nu = 0, npos = 0, prev = FO_INVALID;
    uint64_t* unitigs = (uint64_t*)malloc(sizeof(uint64_t) * nk);
    sq q; sq_init(&q, x);
    for (uint64_t i = 0; i != nk; ++i) {
        lookup_result r = sq_lookup(&q, seq + i);
        if (r.kmer_id != FO_INVALID) {
            ++npos;
            if (r.contig_id != prev) { unitigs[nu++] = r.contig_id; prev = r.contig_id; }
        }
    }
    qsort(unitigs, nu, sizeof(uint64_t), cmp_u64);
    uint64_t n = 0;
    for (uint64_t i = 0; i < nu; ++i) {
        if (i && unitigs[i] == unitigs[i - 1]) continue;
        out[n++] = (uint32_t)fo_u2c(x, unitigs[i]);
    }
    free(unitigs);
    qsort(out, n, sizeof(uint32_t), cmp_u32);
    uint64_t w = 0;
    for (uint64_t i = 0; i < n; ++i) if (!i || out[i] != out[i - 1]) out[w++] = out[i];
    if (num_positive) *num_positive = npos;
    return w;
}

/* ------------------------------------------------------------------ stage 2, hybrid: src/ps_full_intersection.cpp:33-127 */
static int cmp_decoded_size(const void* a, const void* b) { uint32_t x = ((const decoded*)a)->size, y = ((const decoded*)b)->size; return x < y ? -1 : x > y; }

/* leap-frog over ascending arrays (the next_geq loop at :105-126 / :7-30), with an optional exclusion bitmap */
static uint32_t leapfrog(uint32_t** lists, const uint32_t* lens, uint32_t n, const uint8_t* excluded, uint32_t* out) {
    uint32_t cnt = 0;
    uint32_t* pos = (uint32_t*)calloc(n, sizeof(uint32_t));
    for (uint32_t a = 0; a < lens[0]; ++a) {
        uint32_t cand = lists[0][a];
        int ok = 1;
        for (uint32_t i = 1; i < n && ok; ++i) {
            while (pos[i] < lens[i] && lists[i][pos[i]] < cand) ++pos[i];
            if (pos[i] == lens[i] || lists[i][pos[i]] != cand) ok = 0;
        }
        if (ok && !(excluded && excluded[cand])) out[cnt++] = cand;
    }
    free(pos);
    return cnt;
}

static uint64_t hybrid_intersect(const hybrid* h, decoded* its, uint64_t n, uint32_t add, uint32_t* out) {
    if (n == 0) return 0;
    const uint32_t C = h->num_colors;
    qsort(its, n, sizeof(decoded), cmp_decoded_size);                 /* :41-42 */
    uint64_t num_sparse = 0;
    while (num_sparse != n && its[num_sparse].enc != ENC_COMP) ++num_sparse; /* :45-49 */
    uint8_t* excluded = (uint8_t*)calloc(C ? C : 1, 1);
    for (uint64_t i = num_sparse; i < n; ++i)                          /* :51-73 / :93-101: union of the complements */
        for (uint32_t j = 0; j < its[i].n; ++j) excluded[its[i].vals[j]] = 1;
    uint64_t cnt = 0;
    if (num_sparse == 0) {                                             /* :75-91 */
        for (uint32_t c = 0; c < C; ++c) if (!excluded[c]) out[cnt++] = c + add;
    } else {                                                           /* :105-126 */
        uint32_t** lists = (uint32_t**)malloc(sizeof(uint32_t*) * num_sparse);
        uint32_t* lens = (uint32_t*)malloc(sizeof(uint32_t) * num_sparse);
        for (uint64_t i = 0; i < num_sparse; ++i) { lists[i] = its[i].vals; lens[i] = its[i].n; }
        cnt = leapfrog(lists, lens, (uint32_t)num_sparse, excluded, out);
        for (uint64_t i = 0; i < cnt; ++i) out[i] += add;
        free(lists); free(lens);
    }
    free(excluded);
    return cnt;
}

/* ------------------------------------------------------------------ stage 2, meta: src/ps_full_intersection.cpp:243-332 */
/* ------------------------------------------------------------------ stage 2, differential: src/ps_full_intersection.cpp:130-240 */
static int cmp_diffset_rep(const void* a, const void* b) {
    uint64_t x = ((const diffset*)a)->representative_begin, y = ((const diffset*)b)->representative_begin;
    return x < y ? -1 : x > y;
}
/* Sets that share a representative are intersected through it: counts[c] = how many difference lists contain c; c is in all
   sets of the group iff (c not in the representative and all lists flip it in) or (c in the representative and no list
   flips it out) (:166-195). A group of one is decoded (:155-163). The groups are then intersected smallest first (:200-239). */
static uint64_t diff_intersect(const differential* d, diffset* its, uint64_t n, uint32_t lower_bound, uint32_t* out) {
    if (n == 0) return 0;
    const uint32_t C = d->num_colors;
    qsort(its, n, sizeof(diffset), cmp_diffset_rep);                    /* :136-138 */
    uint32_t** lists = (uint32_t**)malloc(sizeof(uint32_t*) * n);
    uint32_t* lens = (uint32_t*)malloc(sizeof(uint32_t) * n);
    uint32_t* counts = (uint32_t*)calloc(C ? C : 1, sizeof(uint32_t));
    uint32_t ng = 0;
    for (uint64_t a = 0; a < n;) {
        uint64_t b = a + 1;
        while (b < n && its[b].representative_begin == its[a].representative_begin) ++b;
        uint32_t* l = (uint32_t*)malloc(sizeof(uint32_t) * (C ? C : 1));
        uint32_t len = 0;
        if (b - a == 1) {
            len = diff_expand(&its[a], 0, l);
        } else {
            const uint32_t gsize = (uint32_t)(b - a);
            for (uint64_t i = a; i < b; ++i) for (uint32_t j = 0; j < its[i].nd; ++j) ++counts[its[i].dv[j]];
            const diffset* last = &its[b - 1];
            uint32_t rp = 0;
            for (uint32_t color = 0; color < C; ++color) {
                while (rp < last->nr && last->rv[rp] < color) ++rp;
                const int in_rep = rp < last->nr && last->rv[rp] == color;
                if ((counts[color] == gsize && !in_rep) || (counts[color] == 0 && in_rep)) l[len++] = color;
            }
            memset(counts, 0, sizeof(uint32_t) * C);
        }
        lists[ng] = l; lens[ng] = len; ++ng;
        a = b;
    }
    for (uint32_t a = 0; a < ng; ++a) for (uint32_t b = a + 1; b < ng; ++b) if (lens[b] < lens[a]) { /* :197-198 */
        uint32_t t = lens[a]; lens[a] = lens[b]; lens[b] = t; uint32_t* tp = lists[a]; lists[a] = lists[b]; lists[b] = tp;
    }
    uint64_t cnt = 0;
    int any_empty = 0;
    for (uint32_t g = 0; g < ng; ++g) if (lens[g] == 0) any_empty = 1;  /* :201-204 */
    if (!any_empty) {
        cnt = leapfrog(lists, lens, ng, NULL, out);
        for (uint64_t i = 0; i < cnt; ++i) out[i] += lower_bound;
    }
    for (uint32_t g = 0; g < ng; ++g) free(lists[g]);
    free(lists); free(lens); free(counts);
    return cnt;
}

static uint64_t meta_intersect(const fo_index* x, const uint32_t* cids, uint64_t n, uint32_t* out) {
    if (n == 0) return 0;
    meta_list* ls = (meta_list*)malloc(sizeof(meta_list) * n);
    for (uint64_t i = 0; i < n; ++i) meta_list_load(x, cids[i], &ls[i]);
    uint64_t cnt = 0;
    const uint32_t P = x->type == 3 ? (uint32_t)x->n_partial : (uint32_t)(x->n_endpoints - 1);
    for (uint32_t p = 0; p < P; ++p) {
        /* step 1 (:258-281): partition must be present in every set */
        int all = 1;
        for (uint64_t i = 0; i < n && all; ++i) if (meta_find(&ls[i], p) < 0) all = 0;
        if (!all) continue;
        /* step 2 (:283-330): distinct meta colors in this partition */
        uint32_t* mcs = (uint32_t*)malloc(sizeof(uint32_t) * n);
        uint32_t nm = 0;
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t mc = ls[i].mc[meta_find(&ls[i], p)];
            int dup = 0;
            for (uint32_t j = 0; j < nm; ++j) if (mcs[j] == mc) dup = 1;
            if (!dup) mcs[nm++] = mc;
        }
        if (x->type == 3) { /* meta_intersect<Iterator, true>: the partial sets are differential */
            const differential* d = &x->md_partial[p];
            diffset* ds = (diffset*)malloc(sizeof(diffset) * nm);
            for (uint32_t j = 0; j < nm; ++j) diff_open(d, mcs[j] - md_sets_before(x, p), &ds[j]);
            if (nm == 1) cnt += diff_expand(&ds[0], md_min_color(x, p), out + cnt);               /* :301-306 */
            else cnt += diff_intersect(d, ds, nm, md_min_color(x, p), out + cnt);                 /* :322-329 */
            for (uint32_t j = 0; j < nm; ++j) diff_close(&ds[j]);
            free(ds); free(mcs);
            continue;
        }
        const hybrid* h = &x->partial[p];
        const uint32_t base = ep_min_color(x, p), sets_before = ep_sets_before(x, p);
        if (nm == 1) { /* same_meta_color: copy the partial set once (:301-306) */
            decoded d; hybrid_decode(h, mcs[0] - sets_before, &d);
            cnt += decoded_expand(&d, h->num_colors, base, out + cnt);
            free(d.vals);
        } else {       /* next_geq_intersect over the materialised partial sets (:307-329, :7-30) */
            uint32_t** lists = (uint32_t**)malloc(sizeof(uint32_t*) * nm);
            uint32_t* lens = (uint32_t*)malloc(sizeof(uint32_t) * nm);
            for (uint32_t j = 0; j < nm; ++j) {
                decoded d; hybrid_decode(h, mcs[j] - sets_before, &d);
                lists[j] = (uint32_t*)malloc(sizeof(uint32_t) * (h->num_colors ? h->num_colors : 1));
                lens[j] = decoded_expand(&d, h->num_colors, base, lists[j]);
                free(d.vals);
            }
            /* smallest first, as the reference sorts by partial_set_size (:309-313) */
            for (uint32_t a = 0; a < nm; ++a) for (uint32_t b = a + 1; b < nm; ++b) if (lens[b] < lens[a]) {
                uint32_t t = lens[a]; lens[a] = lens[b]; lens[b] = t; uint32_t* tp = lists[a]; lists[a] = lists[b]; lists[b] = tp;
            }
            cnt += leapfrog(lists, lens, nm, NULL, out + cnt);
            for (uint32_t j = 0; j < nm; ++j) free(lists[j]);
            free(lists); free(lens);
        }
        free(mcs);
    }
    for (uint64_t i = 0; i < n; ++i) { free(ls[i].mc); free(ls[i].part); }
    free(ls);
    return cnt;
}

/* src/ps_full_intersection.cpp:377-400 */
uint64_t fo_full_intersection(const fo_index* x, const uint32_t* cids, uint64_t n, uint32_t* out) {
    if (x->type == 1 || x->type == 3) return meta_intersect(x, cids, n, out);
    if (x->type == 2) {
        diffset* ds = (diffset*)malloc(sizeof(diffset) * (n ? n : 1));
        for (uint64_t i = 0; i < n; ++i) diff_open(&x->dif, cids[i], &ds[i]);
        uint64_t c = diff_intersect(&x->dif, ds, n, 0, out);
        for (uint64_t i = 0; i < n; ++i) diff_close(&ds[i]);
        free(ds);
        return c;
    }
    decoded* its = (decoded*)malloc(sizeof(decoded) * (n ? n : 1));
    for (uint64_t i = 0; i < n; ++i) hybrid_decode(&x->hyb, cids[i], &its[i]);
    uint64_t cnt = hybrid_intersect(&x->hyb, its, n, 0, out);
    for (uint64_t i = 0; i < n; ++i) free(its[i].vals);
    free(its);
    return cnt;
}

/* ------------------------------------------------------------------ threshold union: src/ps_threshold_union.cpp */
typedef struct { uint64_t item; uint32_t score; } scored;
static int cmp_scored(const void* a, const void* b) { uint64_t x = ((const scored*)a)->item, y = ((const scored*)b)->item; return x < y ? -1 : x > y; }

/* merge (:17-40): int32 scores, complement sets subtract from min_score and from the missing colors */
static uint64_t tu_merge_hybrid(const hybrid* h, const scored* sets, uint64_t n, int64_t min_score, uint32_t* out) {
    if (n == 0) return 0;
    const uint32_t C = h->num_colors;
    int32_t* scores = (int32_t*)calloc(C ? C : 1, sizeof(int32_t));
    for (uint64_t i = 0; i < n; ++i) {
        decoded d; hybrid_decode(h, sets[i].item, &d);
        if (d.enc == ENC_COMP) {
            min_score -= sets[i].score;
            for (uint32_t j = 0; j < d.n; ++j) scores[d.vals[j]] -= (int32_t)sets[i].score;
        } else {
            for (uint32_t j = 0; j < d.n; ++j) scores[d.vals[j]] += (int32_t)sets[i].score;
        }
        free(d.vals);
    }
    uint64_t cnt = 0;
    for (uint32_t c = 0; c < C; ++c) if ((int64_t)scores[c] >= min_score) out[cnt++] = c;
    free(scores);
    return cnt;
}

/* merge_meta (:43-120): partitions whose summed score reaches min_score, then per-color uint32 scores */
static uint64_t tu_merge_meta(const fo_index* x, const scored* sets, uint64_t n, uint64_t min_score, uint32_t* out) {
    if (n == 0) return 0;
    const uint32_t C = x->meta_num_colors, P = (uint32_t)(x->n_endpoints - 1);
    meta_list* ls = (meta_list*)malloc(sizeof(meta_list) * n);
    for (uint64_t i = 0; i < n; ++i) meta_list_load(x, sets[i].item, &ls[i]);
    uint32_t* scores = (uint32_t*)calloc(C ? C : 1, sizeof(uint32_t));
    uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * (C ? C : 1));
    for (uint32_t p = 0; p < P; ++p) {
        uint32_t pscore = 0; int present = 0;
        for (uint64_t i = 0; i < n; ++i) if (meta_find(&ls[i], p) >= 0) { pscore += sets[i].score; present = 1; }
        if (!present || (uint64_t)pscore < min_score) continue; /* :57-73 */
        const hybrid* h = &x->partial[p];
        for (uint64_t i = 0; i < n; ++i) {                      /* :80-115: add each set's score to its partial set's colors */
            int j = meta_find(&ls[i], p);
            if (j < 0) continue;
            decoded d; hybrid_decode(h, ls[i].mc[j] - ep_sets_before(x, p), &d);
            uint32_t m = decoded_expand(&d, h->num_colors, ep_min_color(x, p), tmp);
            for (uint32_t t = 0; t < m; ++t) scores[tmp[t]] += sets[i].score;
            free(d.vals);
        }
    }
    uint64_t cnt = 0;
    for (uint32_t c = 0; c < C; ++c) if ((uint64_t)scores[c] >= min_score) out[cnt++] = c; /* :117-119 */
    for (uint64_t i = 0; i < n; ++i) { free(ls[i].mc); free(ls[i].part); }
    free(ls); free(scores); free(tmp);
    return cnt;
}

/* merge_diff (:123-186) for the sets of ONE differential container, adding into scores[add + color]: sets sharing a
   representative are summed through it -- partition_scores[c] = sum of the scores of the lists that flip c, and a color of the
   representative gets (group score - partition_scores[c]), any other color partition_scores[c] (:160-176); a group of one is
   decoded (:148-155). */
typedef struct { diffset s; uint32_t score; } scored_diffset;
static int cmp_scored_diffset_rep(const void* a, const void* b) {
    uint64_t x = ((const scored_diffset*)a)->s.representative_begin, y = ((const scored_diffset*)b)->s.representative_begin;
    return x < y ? -1 : x > y;
}
static void tu_add_diff_sets(const differential* d, scored_diffset* its, uint64_t n, uint32_t add, uint32_t* scores) {
    const uint32_t C = d->num_colors;
    qsort(its, n, sizeof(scored_diffset), cmp_scored_diffset_rep);
    uint32_t* partition_scores = (uint32_t*)calloc(C ? C : 1, sizeof(uint32_t));
    uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * (C ? C : 1));
    for (uint64_t a = 0; a < n;) {
        uint64_t b = a + 1;
        while (b < n && its[b].s.representative_begin == its[a].s.representative_begin) ++b;
        if (b - a == 1) {
            uint32_t m = diff_expand(&its[a].s, 0, tmp);
            for (uint32_t t = 0; t < m; ++t) scores[add + tmp[t]] += its[a].score;
        } else {
            uint32_t score = 0;
            for (uint64_t i = a; i < b; ++i) {
                score += its[i].score;
                for (uint32_t j = 0; j < its[i].s.nd; ++j) partition_scores[its[i].s.dv[j]] += its[i].score;
            }
            const diffset* last = &its[b - 1].s;
            uint32_t rp = 0;
            for (uint32_t color = 0; color < C; ++color) {
                if (rp < last->nr && last->rv[rp] == color) { scores[add + color] += score - partition_scores[color]; ++rp; }
                else scores[add + color] += partition_scores[color];
            }
            memset(partition_scores, 0, sizeof(uint32_t) * C);
        }
        a = b;
    }
    free(partition_scores); free(tmp);
}
static uint64_t tu_merge_diff(const fo_index* x, const scored* sets, uint64_t n, uint64_t min_score, uint32_t* out) {
    if (n == 0) return 0;
    const uint32_t C = x->dif.num_colors;
    uint32_t* scores = (uint32_t*)calloc(C ? C : 1, sizeof(uint32_t));
    scored_diffset* its = (scored_diffset*)malloc(sizeof(scored_diffset) * n);
    for (uint64_t i = 0; i < n; ++i) { diff_open(&x->dif, sets[i].item, &its[i].s); its[i].score = sets[i].score; }
    tu_add_diff_sets(&x->dif, its, n, 0, scores);
    uint64_t cnt = 0;
    for (uint32_t c = 0; c < C; ++c) if ((uint64_t)scores[c] >= min_score) out[cnt++] = c;  /* :183-185 */
    for (uint64_t i = 0; i < n; ++i) diff_close(&its[i].s);
    free(its); free(scores);
    return cnt;
}
/* merge_metadiff (:188-318): partitions whose summed score reaches min_score (:200-223); inside such a partition, sets with the
   same meta color count once with their scores summed (:272-278), then merge_diff's grouping by representative (:280-315) */
static uint64_t tu_merge_metadiff(const fo_index* x, const scored* sets, uint64_t n, uint64_t min_score, uint32_t* out) {
    if (n == 0) return 0;
    const uint32_t C = x->meta_num_colors, P = (uint32_t)x->n_partial;
    meta_list* ls = (meta_list*)malloc(sizeof(meta_list) * n);
    for (uint64_t i = 0; i < n; ++i) meta_list_load(x, sets[i].item, &ls[i]);
    uint32_t* scores = (uint32_t*)calloc(C ? C : 1, sizeof(uint32_t));
    scored* mcs = (scored*)malloc(sizeof(scored) * n);
    scored_diffset* its = (scored_diffset*)malloc(sizeof(scored_diffset) * n);
    for (uint32_t p = 0; p < P; ++p) {
        uint32_t pscore = 0; uint64_t nm = 0;
        for (uint64_t i = 0; i < n; ++i) {
            int j = meta_find(&ls[i], p);
            if (j < 0) continue;
            pscore += sets[i].score;
            mcs[nm].item = ls[i].mc[j]; mcs[nm].score = sets[i].score; ++nm;
        }
        if (nm == 0 || (uint64_t)pscore < min_score) continue;
        qsort(mcs, nm, sizeof(scored), cmp_scored);
        uint64_t nd = 0;
        for (uint64_t i = 0; i < nm; ++i) {
            if (i && mcs[i].item == mcs[i - 1].item) { its[nd - 1].score += mcs[i].score; continue; }
            diff_open(&x->md_partial[p], mcs[i].item - md_sets_before(x, p), &its[nd].s);
            its[nd].score = mcs[i].score;
            ++nd;
        }
        tu_add_diff_sets(&x->md_partial[p], its, nd, md_min_color(x, p), scores);
        for (uint64_t i = 0; i < nd; ++i) diff_close(&its[i].s);
    }
    uint64_t cnt = 0;
    for (uint32_t c = 0; c < C; ++c) if ((uint64_t)scores[c] >= min_score) out[cnt++] = c;  /* :315-317 */
    for (uint64_t i = 0; i < n; ++i) { free(ls[i].mc); free(ls[i].part); }
    free(ls); free(scores); free(mcs); free(its);
    return cnt;
}

/* index::pseudoalign_threshold_union (:321-402) */
uint64_t fo_threshold_union(const fo_index* x, const char* seq, uint64_t len, double threshold, uint32_t* out) {
    if (len < x->k) return 0;
    uint64_t nk = len - x->k + 1, nu = 0, npos = 0, prev = FO_INVALID;
    scored* unitigs = (scored*)malloc(sizeof(scored) * nk);
    sq q; sq_init(&q, x);
    for (uint64_t i = 0; i != nk; ++i) {                                  /* :327-347 */
        lookup_result r = sq_lookup(&q, seq + i);
        if (r.kmer_id != FO_INVALID) {
            ++npos;
            if (r.contig_id != prev) { unitigs[nu].item = r.contig_id; unitigs[nu].score = 1; ++nu; prev = r.contig_id; }
            else unitigs[nu - 1].score += 1;
        }
    }
    qsort(unitigs, nu, sizeof(scored), cmp_scored);                        /* :357-372 */
    scored* csets = (scored*)malloc(sizeof(scored) * (nu ? nu : 1));
    uint64_t nc = 0;
    for (uint64_t i = 0; i < nu; ++i) {
        if (i && unitigs[i].item == unitigs[i - 1].item) csets[nc - 1].score += unitigs[i].score;
        else { csets[nc].item = fo_u2c(x, unitigs[i].item); csets[nc].score = unitigs[i].score; ++nc; }
    }
    qsort(csets, nc, sizeof(scored), cmp_scored);                          /* :374-387 */
    uint64_t ns = 0;
    for (uint64_t i = 0; i < nc; ++i) {
        if (i && csets[i].item == csets[ns - 1].item) csets[ns - 1].score += csets[i].score;
        else csets[ns++] = csets[i];
    }
    const uint64_t min_score = (uint64_t)((double)npos * threshold);       /* :389 */
    uint64_t cnt = x->type == 0   ? tu_merge_hybrid(&x->hyb, csets, ns, (int64_t)min_score, out)
                   : x->type == 1 ? tu_merge_meta(x, csets, ns, min_score, out)
                   : x->type == 2 ? tu_merge_diff(x, csets, ns, min_score, out)
                                  : tu_merge_metadiff(x, csets, ns, min_score, out);
    free(unitigs); free(csets);
    return cnt;
}

/* ------------------------------------------------------------------ the per-k-mer tools */
/* index::kmer_conservation (src/kmer_conservation.cpp:7-54): triples {start_pos_in_query, num_kmers, color_set_id} of the maximal
   runs of consecutive positive k-mers with one color-set id. Returns the number of triples (3 uint32 each in out). */
uint64_t fo_kmer_conservation(const fo_index* x, const char* seq, uint64_t len, uint32_t* out) {
    if (len < x->k) return 0;                                            /* :14 */
    const uint64_t nk = len - x->k + 1;
    uint64_t n = 0, prev = FO_INVALID;
    uint32_t start = 0, num = 0;
    sq q; sq_init(&q, x);
    for (uint64_t i = 0; i != nk; ++i) {
        lookup_result r = sq_lookup(&q, seq + i);
        if (r.kmer_id != FO_INVALID) {                                   /* :35-46 */
            uint64_t cid = fo_u2c(x, r.contig_id);
            if (prev != cid) {
                if (prev != FO_INVALID) { out[3 * n] = start; out[3 * n + 1] = num; out[3 * n + 2] = (uint32_t)prev; ++n; }
                num = 0; start = (uint32_t)i;
            }
            num += 1; prev = cid;
        } else {                                                         /* :48-51 */
            if (prev != FO_INVALID) { out[3 * n] = start; out[3 * n + 1] = num; out[3 * n + 2] = (uint32_t)prev; ++n; }
            prev = FO_INVALID;
        }
    }
    if (prev != FO_INVALID) { out[3 * n] = start; out[3 * n + 1] = num; out[3 * n + 2] = (uint32_t)prev; ++n; } /* :54 */
    return n;
}

/* index::kmer_matches (src/kmer_matches.cpp:7-30): positive[i] = k-mer i is in the index (one byte each here), counts[c] +=1 for
   every color of every positive k-mer's color set. Returns the number of k-mers. counts has num_colors entries, zeroed here. */
uint64_t fo_kmer_matches(const fo_index* x, const char* seq, uint64_t len, uint8_t* positive, uint32_t* counts) {
    const uint32_t C = index_num_colors(x);
    memset(counts, 0, sizeof(uint32_t) * C);
    if (len < x->k) return 0;
    const uint64_t nk = len - x->k + 1;
    uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * (C ? C : 1));
    sq q; sq_init(&q, x);
    for (uint64_t i = 0; i != nk; ++i) {
        lookup_result r = sq_lookup(&q, seq + i);
        positive[i] = r.kmer_id != FO_INVALID;
        if (positive[i]) {
            int64_t m = fo_color_set(x, fo_u2c(x, r.contig_id), tmp, C);
            for (int64_t j = 0; j < m; ++j) counts[tmp[j]] += 1;
        }
    }
    free(tmp);
    return nk;
}

int fo_batch_kmer_conservation(const fo_index* x, const char* bases, const uint64_t* read_off, uint32_t n, uint64_t* triple_off,
                               uint32_t* triples, uint64_t cap) {
    uint64_t total = 0;
    uint32_t* tmp = NULL; uint64_t tmp_cap = 0;
    triple_off[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t len = read_off[i + 1] - read_off[i];
        if (3 * (len + 1) > tmp_cap) { tmp_cap = 6 * len + 64; tmp = (uint32_t*)realloc(tmp, sizeof(uint32_t) * tmp_cap); }
        uint64_t c = fo_kmer_conservation(x, bases + read_off[i], len, tmp);
        if (total + c <= cap) memcpy(triples + 3 * total, tmp, sizeof(uint32_t) * 3 * c);
        total += c;
        triple_off[i + 1] = total;
    }
    free(tmp);
    return total > cap ? -7 : 0;
}

/* positive: one byte per k-mer, reads concatenated (kmer_off[i] = first k-mer of read i, n+1 entries); counts: n x num_colors */
int fo_batch_kmer_matches(const fo_index* x, const char* bases, const uint64_t* read_off, uint32_t n, uint64_t* kmer_off,
                          uint8_t* positive, uint64_t cap, uint32_t* counts) {
    const uint32_t C = index_num_colors(x);
    kmer_off[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t len = read_off[i + 1] - read_off[i];
        kmer_off[i + 1] = kmer_off[i] + (len >= x->k ? len - x->k + 1 : 0);
    }
    if (kmer_off[n] > cap) return -7;
    for (uint32_t i = 0; i < n; ++i)
        fo_kmer_matches(x, bases + read_off[i], read_off[i + 1] - read_off[i], positive + kmer_off[i], counts + (uint64_t)i * C);
    return 0;
}

/* ------------------------------------------------------------------ batch drivers */
int fo_batch_fetch_color_set_ids(const fo_index* x, const char* bases, const uint64_t* read_off, uint32_t n,
                                 uint64_t* cid_off, uint32_t* cids, uint64_t cap, uint32_t* num_positive) {
    uint64_t total = 0;
    uint32_t* tmp = NULL; uint64_t tmp_cap = 0;
    cid_off[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t len = read_off[i + 1] - read_off[i];
        if (len + 1 > tmp_cap) { tmp_cap = 2 * len + 64; tmp = (uint32_t*)realloc(tmp, sizeof(uint32_t) * tmp_cap); }
        uint64_t np = 0;
        uint64_t c = fo_fetch_color_set_ids(x, bases + read_off[i], len, tmp, &np);
        if (num_positive) num_positive[i] = (uint32_t)np;
        if (total + c <= cap) memcpy(cids + total, tmp, sizeof(uint32_t) * c);
        total += c;
        cid_off[i + 1] = total;
    }
    free(tmp);
    return total > cap ? -7 : 0;
}

int fo_batch_pseudoalign(const fo_index* x, int algo, double threshold, const char* bases, const uint64_t* read_off,
                         uint32_t n, uint64_t* color_off, uint32_t* colors, uint64_t cap) {
    uint64_t total = 0;
    const uint32_t C = index_num_colors(x);
    uint32_t* res = (uint32_t*)malloc(sizeof(uint32_t) * (C ? C : 1));
    uint32_t* cids = NULL; uint64_t cids_cap = 0;
    color_off[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t len = read_off[i + 1] - read_off[i], c;
        if (algo == 0) {
            if (len + 1 > cids_cap) { cids_cap = 2 * len + 64; cids = (uint32_t*)realloc(cids, sizeof(uint32_t) * cids_cap); }
            uint64_t nc = fo_fetch_color_set_ids(x, bases + read_off[i], len, cids, NULL);
            c = fo_full_intersection(x, cids, nc, res);
        } else {
            c = fo_threshold_union(x, bases + read_off[i], len, threshold, res);
        }
        if (total + c <= cap) memcpy(colors + total, res, sizeof(uint32_t) * c);
        total += c;
        color_off[i + 1] = total;
    }
    free(res); free(cids);
    return total > cap ? -7 : 0;
}
