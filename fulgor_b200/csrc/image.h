/*
 * image.h -- the device index image: ONE contiguous, position-independent blob that the host
 * flattener (fur_reader.cpp) writes once and every GPU holds a replica of (one cudaMemcpy, or one
 * NCCL broadcast of the bytes). All references inside the blob are byte offsets from its start, so
 * the same bytes are valid on every device.
 *
 * What is flattened relative to the reference's on-disk structures, and why:
 *   - PTHash pilots (front/back ranks->dict compact vectors, pthash/include/utils/encoders.hpp:192-195,
 *     391-394) become one u64 per PTHash bucket holding MurmurHash2_64(pilot, seed) already applied
 *     (single_phf.hpp:84-87 hashes the pilot at every query): 1 gather instead of 2 + a hash.
 *   - Elias-Fano sequences that are only ever `access`ed (free slots, bucket sizes, color-set bit
 *     offsets, meta offsets; bits/include/elias_fano.hpp:159-163) become plain arrays.
 *   - each super-k-mer gets an 8-byte record {string offset, window length, color-set id, minimizer
 *     position} which folds buckets::offset_to_id (sshash/include/buckets.hpp:13-40, an Elias-Fano
 *     locate over `pieces`), the window clamp of lookup_canonical_in_super_kmer (buckets.hpp:133-160)
 *     and index::u2c (include/index.hpp:37, rank9) into the one gather that fetches the offset. The
 *     minimizer position (where, relative to the super-k-mer's first base, the canonical minimizer of
 *     ALL its k-mers starts; computed at load with the same arithmetic the kernels use) lets a query
 *     compare at most two window positions instead of all k-m+1 (see scan_super_kmer).
 *   - the 2-bit `strings` and the compressed color-set bit streams stay VERBATIM (bit-identical
 *     words); the kernels decode them in place.
 */
#ifndef FULGOR_B200_IMAGE_H
#define FULGOR_B200_IMAGE_H

#include <stdint.h>

#define FGI_MAGIC 0x3330474D49475546ULL /* "FUGIMG03" */
#define FGI_ALIGN 256

/* one single_phf partition (pthash/include/single_phf.hpp:140-150). The three moduli (table size, dense / sparse bucket
   counts) are < 2^32, so `a mod d` is evaluated as a - mulhi64(a, inv) * d with one conditional correction, where
   inv = floor(2^64 / d): the exact remainder, like the reference's 128-bit fastmod (external/fastmod/fastmod.h:159-162). */
struct fgi_phf_part {
    uint64_t seed;
    uint64_t num_keys;
    uint64_t table_size, inv_table;
    uint64_t num_dense, inv_dense;   /* skew_bucketer (utils/bucketers.hpp:197-206) */
    uint64_t num_sparse, inv_sparse;
    uint64_t offset;      /* partitioned_phf partition offset (partitioned_phf.hpp:23-43) */
    uint64_t pilot_base;  /* first entry of this partition in hashed_pilots[] */
    uint64_t free_base;   /* first entry of this partition in free_slots[] */
    uint64_t pad[5];
};

/* one partitioned_phf (pthash/include/partitioned_phf.hpp:203-210) */
struct fgi_phf {
    uint64_t seed;
    uint64_t num_partitions; /* range_bucketer::num_buckets */
    uint64_t first_part;     /* index into phf_parts[] */
    uint64_t num_keys;
};

/* one color-set container: hybrid (include/color_sets/hybrid.hpp:339-345) or differential
   (include/color_sets/differential.hpp:322-339; sparse_thr / very_dense_thr unused) */
#define FGI_SETS_HYBRID 0
#define FGI_SETS_DIFFERENTIAL 1
struct fgi_hybrid {
    uint32_t num_colors, sparse_thr, very_dense_thr, kind;
    uint64_t num_sets;
    uint64_t set_off_base; /* first entry in set_bit_off[]: num_sets + 1 bit offsets of the sets (hybrid) or of their difference
                              lists (differential); a differential container adds num_sets more entries, the bit offset of
                              the representative each set is coded against */
    uint64_t word_base;    /* first u64 word of this container's bit stream in color_words[] */
};

#define FGI_MAX_SKEW 16

struct fgi_header {
    uint64_t magic;
    uint64_t total_bytes;
    /* sshash::dictionary (sshash/include/dictionary.hpp:141-154) */
    uint32_t k, m;
    uint64_t num_kmers;
    uint64_t hash_magic;      /* mixer_64 (sshash/include/hash_util.hpp:88-111) */
    uint64_t bucketer_T;      /* (uint64_t)(0.6f * (double)UINT64_MAX), utils/bucketers.hpp:163-168 */
    uint64_t num_minimizers;  /* = number of SSHash buckets */
    uint64_t num_super_kmers;
    uint64_t num_unitigs;
    uint64_t num_string_words;
    /* skew index (sshash/include/skew_index.hpp:84-91) */
    uint32_t skew_min_log2, skew_max_log2, skew_log2_max_bucket, num_skew;
    uint32_t skew_phf[FGI_MAX_SKEW];      /* index into phfs[]; phfs[0] is the minimizer MPHF */
    uint64_t skew_pos_base[FGI_MAX_SKEW]; /* first entry in skew_positions[] */
    /* colors */
    uint32_t type;            /* 0 hybrid (.fur), 1 meta (.mfur), 2 differential (.dfur), 3 meta-differential (.mdfur):
                                 bit 0 = a set is a list of partial sets, bit 1 = the containers are differential */
    uint32_t num_colors;
    uint64_t num_color_sets;
    uint32_t num_partitions;  /* meta: number of partial color-set containers; hybrid: 1 */
    uint32_t num_phfs, num_phf_parts, pad0;
    /* section byte offsets from the start of the image */
    uint64_t off_phfs;           /* fgi_phf[num_phfs] */
    uint64_t off_phf_parts;      /* fgi_phf_part[num_phf_parts] */
    uint64_t off_hashed_pilots;  /* u64[] */
    uint64_t off_free_slots;     /* u32[] */
    uint64_t off_bucket_begin;   /* u32[num_minimizers + 1]: first super-k-mer id of bucket b (buckets.hpp:62-67) */
    uint64_t off_sk_records;     /* uint2[num_super_kmers]: {offset, FGI_SK_* fields} */
    uint64_t off_strings;        /* u64[num_string_words + 2 pad] verbatim */
    uint64_t off_skew_positions; /* u32[] */
    uint64_t off_hybrids;        /* fgi_hybrid[num_partitions] */
    uint64_t off_set_bit_off;    /* u64[] */
    uint64_t off_color_words;    /* u64[] verbatim (+2 pad words per container) */
    uint64_t off_meta_off;       /* u64[num_color_sets + 1] element offsets into meta_vals (meta.hpp:275-281) */
    uint64_t off_meta_vals;      /* u32[]: records [n, meta_color_1..n] */
    uint64_t off_part_min_color; /* u32[num_partitions + 1] */
    uint64_t off_part_sets_before; /* u32[num_partitions + 1] */
    uint64_t off_sk_cid;         /* u32[num_super_kmers], only when num_color_sets > 2^21 (else 0): color-set ids that do not fit the record */
    uint64_t num_unpinned;       /* super-k-mers whose k-mers disagree on the minimizer position (full window scan) */
    uint64_t guard_max_hash;     /* 1 iff some homopolymer m-mer hashes to UINT64_MAX (then compute_minimizer's "no minimum found" value matters) */
    uint64_t reserved[5];
};

/* low word of a super-k-mer record: base offset into `strings` (< 2^31) and, for pinned records, whether the stored m-mer at
   the minimizer position reads as the canonical minimizer (1) or as its reverse complement (0) */
#define FGI_SK_OFFSET_MASK 0x7fffffffu
#define FGI_SK_CANON_FWD_SHIFT 31
/* high word of a super-k-mer record */
#define FGI_SK_CID_BITS 21
#define FGI_SK_CID_MASK ((1u << FGI_SK_CID_BITS) - 1u)
#define FGI_SK_WINDOW_SHIFT 21 /* 5 bits: number of k-mers of the super-k-mer */
#define FGI_SK_PM_SHIFT 26     /* 5 bits: base offset of the canonical minimizer from the super-k-mer start */
#define FGI_SK_PINNED_SHIFT 31 /* 1 bit: the minimizer position is the same for all k-mers of the super-k-mer */

#endif
