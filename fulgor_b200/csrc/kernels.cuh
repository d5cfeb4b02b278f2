/*
 * kernels.cuh -- sm_100a device code of the pseudoalignment hot path.
 *
 * Mapping: ONE WARP PER READ, a read handled in segments of 128 k-mers (kmer_tiles below). The reference's sequential
 * seed-and-extend state machine (external/sshash/include/streaming_query.hpp:50-190) is not reproduced step by step; its
 * effect is: the k-mers of a segment are cut into runs that share one minimizer occurrence, the first k-mer of a run is
 * looked up through the minimizer MPHF (one seed per lane), and one comparison of the stored super-k-mer string with the
 * read, aligned on the minimizer, answers every k-mer of the run at once ("extension"). A lookup answer is a pure function of
 * the k-mer (the reference asserts this itself, streaming_query.hpp:107), so the answers are the reference's.
 *
 * Citations: "sshash/" = reference external/sshash/include/, "pthash/" =
 * external/sshash/external/pthash/include/, "bits/" = .../pthash/external/bits/include/.
 */
#ifndef FULGOR_B200_KERNELS_CUH
#define FULGOR_B200_KERNELS_CUH

#ifndef FG_SIMT_EMUL /* tests/simt_emul.h supplies the CUDA vocabulary when the kernels are compiled for the host */
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "image.h"

/* The per-lane functions (hashing, MPHF, string compare, color-set decode) are FG_HD so that the CPU test tier can call them,
   and the kernels themselves compile for the host on the lock-step warp emulator tests/simt_emul.{h,cpp} (FG_SIMT_EMUL).
   That harness is test infrastructure; the library itself has no CPU path. */
#define FG_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define FG_LDG(p) __ldg(p)
#else
#define FG_LDG(p) (*(p))
#endif

namespace fgb {

FG_HD uint64_t fg_mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return uint64_t((unsigned __int128)a * b >> 64);
#endif
}
/* bits 32..63 of a * b (mod 2^64): one high multiply and two multiply-adds */
FG_HD uint32_t fg_mul64_hi32(uint64_t a, uint64_t b) {
    const uint32_t al = uint32_t(a), ah = uint32_t(a >> 32), bl = uint32_t(b), bh = uint32_t(b >> 32);
#ifdef __CUDA_ARCH__
    return __umulhi(al, bl) + al * bh + ah * bl;
#else
    return uint32_t((uint64_t(al) * bl) >> 32) + al * bh + ah * bl;
#endif
}
FG_HD uint64_t fg_brev64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x);
#endif
}
FG_HD uint32_t fg_ffs64(uint64_t x) { /* 1-based index of the lowest set bit, 0 if none */
#ifdef __CUDA_ARCH__
    return uint32_t(__ffsll((long long)x));
#else
    return uint32_t(__builtin_ffsll((long long)x));
#endif
}
FG_HD uint32_t fg_clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return uint32_t(__clz(int(x)));
#else
    return x ? uint32_t(__builtin_clz(x)) : 32u;
#endif
}

#define FG_FULL 0xffffffffu
#define FG_NOT_FOUND 0xffffffffu
#define FG_MAX_ENTRIES 32 /* distinct color sets per read held in registers (one per lane) */
#define FG_DIVERSE_BATCH 10 /* distinct color sets in one batch of 32 items from which a read skips the register table */

/* device view of the image: absolute pointers + the scalars the kernels need */
struct dev_index {
    const fgi_phf* phfs;
    const fgi_phf_part* parts;
    const uint64_t* hashed_pilots;
    const uint32_t* free_slots;
    const uint32_t* bucket_begin;
    const uint2* sk_records;
    const uint32_t* sk_cid; /* nullptr unless the color-set ids do not fit the records */
    const uint64_t* strings;
    const uint32_t* skew_positions;
    const fgi_hybrid* hybrids;
    const uint64_t* set_bit_off;
    const uint64_t* color_words;
    const uint64_t* meta_off;
    const uint32_t* meta_vals;
    const uint32_t* part_min_color;
    const uint32_t* part_sets_before;
    uint64_t hash_magic, bucketer_T;
    uint32_t k, m;
    uint32_t skew_min_log2, skew_max_log2, skew_log2_max_bucket, num_skew;
    uint32_t skew_threshold; /* buckets with more super-k-mers go through the skew index; UINT32_MAX when there is none */
    uint32_t diff;           /* 1: the containers are differential (.dfur / .mdfur), see fgi_hybrid */
    uint32_t skew_phf[FGI_MAX_SKEW];
    uint64_t skew_pos_base[FGI_MAX_SKEW];
    uint32_t type; /* 0: a color set is one set of container 0; 1: a list of partial sets (.mfur / .mdfur) */
    uint32_t num_colors, num_partitions, guard_max_hash;
    uint64_t main_seed, main_nparts; /* the minimizer MPHF (phfs[0]; its partitions are parts[0 .. main_nparts)) */
    fgi_phf_part main_part;          /* parts[0] */
    /* the decoded color-set table: row i = color set i as a bitmap of num_colors bits padded to table_stride 32-bit words.
       Built on each GPU from the compressed sets when it fits the HBM budget (k_expand_color_sets); nullptr otherwise. */
    const uint32_t* set_table;
    uint64_t table_stride;
};

/* ------------------------------------------------------------------ hashing */

/* MurmurHash2_64 of an 8-byte key (pthash/utils/hasher.hpp:53-117) */
FG_HD uint64_t murmur2_64(uint64_t key, uint64_t seed) {
    const uint64_t m = 0xc6a4a7935bd1e995ULL;
    uint64_t h = seed ^ (8 * m);
    uint64_t k = key * m;
    k ^= k >> 47;
    k *= m;
    h ^= k;
    h *= m;
    h ^= h >> 47;
    h *= m;
    h ^= h >> 47;
    return h;
}

/* a mod d for d < 2^32 with inv = floor(2^64 / d) (image.h): q = mulhi(a, inv) is floor(a / d) or one less, so one
   conditional subtraction gives the exact remainder -- the value the reference's fastmod_u64
   (pthash/external/fastmod/fastmod.h:159-162) returns. */
FG_HD uint64_t mod_by_inverse(uint64_t a, uint64_t inv, uint64_t d) {
    if (d == 0) return 0; /* a PTHash partition of <= 15 keys has no dense buckets (0.3 * num_buckets == 0): everything maps to bucket 0 */
    const uint64_t q = fg_mulhi64(a, inv);
    uint64_t r = a - q * d;
    if (r >= d) r -= d;
    return r;
}

/* single_phf::position (pthash/single_phf.hpp:79-101): skew_bucketer (utils/bucketers.hpp:163-168), pre-hashed pilot,
   xor displacement, minimal (free slots) */
FG_HD uint64_t phf_position(const dev_index& I, const fgi_phf_part& P, uint64_t first, uint64_t second) {
    /* one remainder for either side of the skew bucketer: the lanes of a warp fall on both, and two divergent copies of the
       multiply-high sequence cost more than three selects */
    const bool dense = first < I.bucketer_T;
    const uint64_t bucket = (dense ? 0 : P.num_dense) + mod_by_inverse(first, dense ? P.inv_dense : P.inv_sparse, dense ? P.num_dense : P.num_sparse);
    const uint64_t hashed_pilot = FG_LDG(I.hashed_pilots + P.pilot_base + bucket);
    uint64_t pos = mod_by_inverse(second ^ hashed_pilot, P.inv_table, P.table_size);
    if (pos >= P.num_keys) pos = FG_LDG(I.free_slots + P.free_base + (pos - P.num_keys));
    return P.offset + pos;
}

FG_HD fgi_phf_part load_part(const fgi_phf_part* p) {
    fgi_phf_part P;
    P.num_keys = FG_LDG(&p->num_keys);
    P.table_size = FG_LDG(&p->table_size);
    P.inv_table = FG_LDG(&p->inv_table);
    P.num_dense = FG_LDG(&p->num_dense);
    P.inv_dense = FG_LDG(&p->inv_dense);
    P.num_sparse = FG_LDG(&p->num_sparse);
    P.inv_sparse = FG_LDG(&p->inv_sparse);
    P.offset = FG_LDG(&p->offset);
    P.pilot_base = FG_LDG(&p->pilot_base);
    P.free_base = FG_LDG(&p->free_base);
    return P;
}

/* partitioned_phf::operator() (pthash/partitioned_phf.hpp:150-159) with murmurhash2_128 (utils/hasher.hpp:203-207) and
   range_bucketer (utils/bucketers.hpp:216-218) */
FG_HD uint64_t phf_lookup(const dev_index& I, uint32_t phf_id, uint64_t key) {
    const fgi_phf* F = I.phfs + phf_id;
    const uint64_t seed = FG_LDG(&F->seed);
    const uint64_t nparts = FG_LDG(&F->num_partitions);
    const uint64_t first = murmur2_64(key, seed);
    const uint64_t second = murmur2_64(key, ~seed);
    uint64_t p = 0;
    if (nparts > 1) p = (((first ^ second) >> 32) * nparts) >> 32;
    return phf_position(I, load_part(I.parts + FG_LDG(&F->first_part) + p), first, second);
}

/* minimizers::lookup (sshash/minimizers.hpp:36-39): the minimizer MPHF. With a single partition (up to ~3 M minimizers,
   sshash/constants.hpp:13) its descriptor travels in the kernel parameters (constant bank), not through loads. */
FG_HD uint64_t minimizer_bucket(const dev_index& I, uint64_t minimizer) {
    const uint64_t first = murmur2_64(minimizer, I.main_seed);
    const uint64_t second = murmur2_64(minimizer, ~I.main_seed);
    if (I.main_nparts == 1) return phf_position(I, I.main_part, first, second);
    const uint64_t p = (((first ^ second) >> 32) * I.main_nparts) >> 32;
    return phf_position(I, load_part(I.parts + p), first, second);
}

/* ------------------------------------------------------------------ k-mers */

/* reverse complement of a 2-bit packed k-mer (sshash/kmer.hpp:146-170; A=0 C=1 T=2 G=3 so the
   complement is XOR 10b per base) */
FG_HD uint64_t revcomp(uint64_t x, uint32_t k) {
    uint64_t y = fg_brev64(x ^ 0xAAAAAAAAAAAAAAAAULL);
    y = ((y & 0x5555555555555555ULL) << 1) | ((y >> 1) & 0x5555555555555555ULL);
    return y >> (64 - 2 * k);
}

/* canonical minimizer of a k-mer and WHERE it sits. value = min over both strands (compared as integers,
   sshash/streaming_query.hpp:76-79) of util::compute_minimizer (sshash/util.hpp:220-239: the m-mer with the smallest
   mixer_64 hash, sshash/hash_util.hpp:97; first one wins). cpos = base offset of that m-mer inside the k-mer in forward
   coordinates; ambiguous = both strands yield the same value, so the position is not unique. */
struct minimizer_t {
    uint64_t value;
    uint32_t cpos;
    bool ambiguous;
};

FG_HD minimizer_t combine_strands(uint64_t vf, uint32_t jf, uint64_t vr, uint32_t jr, uint32_t k, uint32_t m) {
    minimizer_t r;
    r.ambiguous = vf == vr;
    if (vf <= vr) {
        r.value = vf;
        r.cpos = jf;
    } else {
        r.value = vr;
        r.cpos = (k - m) - jr;
    }
    return r;
}

/* straight evaluation from the k-mer (host emulation, and the specification of the shared-memory pipeline below) */
FG_HD void strand_minimizer(uint64_t x, uint32_t window, uint64_t mmer_mask, uint64_t magic, uint64_t& value, uint32_t& pos) {
    uint64_t best_h = UINT64_MAX;
    value = UINT64_MAX;
    pos = 0;
    for (uint32_t i = 0; i < window; ++i) {
        const uint64_t y = x & mmer_mask;
        const uint64_t h = (y * 0x517cc1b727220a95ULL) ^ magic;
        if (h < best_h) {
            best_h = h;
            value = y;
            pos = i;
        }
        x >>= 2;
    }
}
FG_HD minimizer_t canonical_minimizer(uint64_t fwd, uint64_t rc, uint32_t k, uint32_t m, uint64_t magic) {
    const uint64_t mmer_mask = (1ULL << (2 * m)) - 1;
    uint64_t vf, vr;
    uint32_t jf, jr;
    strand_minimizer(fwd, k - m + 1, mmer_mask, magic, vf, jf);
    strand_minimizer(rc, k - m + 1, mmer_mask, magic, vr, jr);
    return combine_strands(vf, jf, vr, jr, k, m);
}

FG_HD uint32_t ceil_log2_u32(uint32_t v) { /* bits/util.hpp ceil_log2_uint32 */
    return v <= 1 ? 0u : 32u - fg_clz32(v - 1);
}

/* lookup_canonical_in_super_kmer (sshash/buckets.hpp:133-160) on the flattened record: is the k-mer (or its reverse
   complement) one of the `window` consecutive k-mers that start at `offset` in the 2-bit strings? The reference compares
   all of them. Here the record knows where the canonical minimizer of every k-mer of the super-k-mer sits (pm, relative to
   the super-k-mer start; image.h), and the query knows where its own sits (mz.cpos): if the query equals the stored k-mer
   at window index t in forward orientation then t + cpos = pm, in reverse orientation t + (k - m - cpos) = pm -- equal
   k-mers have equal minimizers at equal positions -- so at most two positions can match and only those are compared.
   Records whose k-mers disagree on the position (not pinned) and queries with an ambiguous position fall back to the full
   scan. Returns the color-set id of the enclosing unitig, or FG_NOT_FOUND. */
FG_HD uint32_t scan_super_kmer(const dev_index& I, uint32_t sk, uint64_t fwd, uint64_t rc, uint64_t kmask, const minimizer_t& mz) {
    const uint2 rec = FG_LDG(I.sk_records + sk);
    const uint32_t window = (rec.y >> FGI_SK_WINDOW_SHIFT) & 31u;
    const uint64_t bit = 2 * uint64_t(rec.x & FGI_SK_OFFSET_MASK);
    const uint64_t* w = I.strings + (bit >> 6);
    const uint32_t sh = uint32_t(bit & 63);
    const uint64_t w0 = FG_LDG(w), w1 = FG_LDG(w + 1), w2 = FG_LDG(w + 2);
    bool hit = false;
    if ((rec.y >> FGI_SK_PINNED_SHIFT) && !mz.ambiguous) {
        const uint32_t pm = (rec.y >> FGI_SK_PM_SHIFT) & 31u;
        const uint32_t t_fwd = pm - mz.cpos, t_rc = pm - ((I.k - I.m) - mz.cpos); /* wrap around when negative */
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            const uint32_t t = o ? t_rc : t_fwd;
            if (t < window) {
                const uint32_t s2 = sh + 2 * t;
                const uint64_t a = (s2 & 64) ? w1 : w0, b = (s2 & 64) ? w2 : w1;
                const uint32_t s = s2 & 63;
                const uint64_t cand = (s ? (a >> s) | (b << (64 - s)) : a) & kmask;
                hit |= cand == (o ? rc : fwd);
            }
        }
    } else {
        uint64_t lo = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;
        uint64_t hi = sh ? (w1 >> sh) | (w2 << (64 - sh)) : w1;
        for (uint32_t t = 0; t < window; ++t) {
            const uint64_t cand = lo & kmask;
            hit |= (cand == fwd) | (cand == rc);
            lo = (lo >> 2) | (hi << 62);
            hi >>= 2;
        }
    }
    if (!hit) return FG_NOT_FOUND;
    return I.sk_cid ? FG_LDG(I.sk_cid + sk) : (rec.y & FGI_SK_CID_MASK);
}

/* buckets::lookup_canonical (sshash/buckets.hpp:162-209) + index::u2c on the super-k-mers [begin, begin + n) of one
   bucket, with the skew-index dispatch of dictionary::lookup_uint_canonical (sshash/../src/dictionary.cpp:61-73).
   The "minimizer of the bucket's first k-mer must equal the target" test (buckets.hpp:168-180) is an early-out only:
   equal k-mers have equal minimizers, so a k-mer whose minimizer is absent cannot match any stored k-mer. */
FG_HD uint32_t lookup_in_bucket(const dev_index& I, uint32_t begin, uint32_t n, uint64_t fwd, uint64_t rc, const minimizer_t& mz,
                                uint64_t kmask) {
    if (n > I.skew_threshold) { /* ceil_log2(n) > min_log2 and a skew index exists (sshash/../src/dictionary.cpp:61-63) */
        const uint32_t log2n = ceil_log2_u32(n);
        { /* skew_index::lookup (sshash/skew_index.hpp:40-52) */
            uint32_t pid = log2n - (I.skew_min_log2 + 1);
            if (log2n == I.skew_log2_max_bucket || log2n > I.skew_max_log2) pid = I.num_skew - 1;
            const uint32_t f = I.skew_phf[pid];
            if (FG_LDG(&I.phfs[f].num_partitions) == 0) return FG_NOT_FOUND;
            const uint64_t h = phf_lookup(I, f, fwd < rc ? fwd : rc);
            const uint32_t pos = FG_LDG(I.skew_positions + I.skew_pos_base[pid] + h);
            if (pos < n) return scan_super_kmer(I, begin + pos, fwd, rc, kmask, mz);
            return FG_NOT_FOUND;
        }
    }
    for (uint32_t s = begin; s < begin + n; ++s) {
        const uint32_t cid = scan_super_kmer(I, s, fwd, rc, kmask, mz);
        if (cid != FG_NOT_FOUND) return cid;
    }
    return FG_NOT_FOUND;
}

#ifndef FG_SIMT_EMUL
/* out-of-line copies for the warp pipeline's RARE paths (skew-index buckets, ambiguous or palindromic minimizers, super-k-mers
   without a single minimizer position): the lookup kernels are ~100 KB of code against a 32 KB instruction cache, so what is
   seldom executed is kept out of the instruction stream of what always is */
__device__ __noinline__ uint32_t lookup_in_bucket_rare(const dev_index& I, uint32_t begin, uint32_t n, uint64_t fwd, uint64_t rc, uint32_t cpos, bool ambiguous,
                                                       uint64_t kmask) {
    minimizer_t mz;
    mz.value = 0;
    mz.cpos = cpos;
    mz.ambiguous = ambiguous;
    return lookup_in_bucket(I, begin, n, fwd, rc, mz, kmask);
}
__device__ __noinline__ bool scan_super_kmer_rare(const dev_index& I, uint32_t sk, uint64_t fwd, uint64_t kmask) {
    minimizer_t mz;
    mz.value = 0;
    mz.cpos = 0;
    mz.ambiguous = true;
    return scan_super_kmer(I, sk, fwd, revcomp(fwd, I.k), kmask, mz) != FG_NOT_FOUND;
}
#else
static inline uint32_t lookup_in_bucket_rare(const dev_index& I, uint32_t begin, uint32_t n, uint64_t fwd, uint64_t rc, uint32_t cpos, bool ambiguous, uint64_t kmask) {
    minimizer_t mz;
    mz.value = 0;
    mz.cpos = cpos;
    mz.ambiguous = ambiguous;
    return lookup_in_bucket(I, begin, n, fwd, rc, mz, kmask);
}
static inline bool scan_super_kmer_rare(const dev_index& I, uint32_t sk, uint64_t fwd, uint64_t kmask) {
    minimizer_t mz;
    mz.value = 0;
    mz.cpos = 0;
    mz.ambiguous = true;
    return scan_super_kmer(I, sk, fwd, revcomp(fwd, I.k), kmask, mz) != FG_NOT_FOUND;
}
#endif

/* dictionary::lookup_uint_canonical (sshash/../src/dictionary.cpp:47-77): minimizer -> bucket (minimizers.hpp:36-39,
   buckets::locate_bucket buckets.hpp:62-67) -> lookup inside the bucket */
FG_HD uint32_t lookup_color_set(const dev_index& I, uint64_t fwd, uint64_t rc, const minimizer_t& mz, uint64_t kmask) {
    const uint64_t b = minimizer_bucket(I, mz.value);
    const uint32_t begin = FG_LDG(I.bucket_begin + b), end = FG_LDG(I.bucket_begin + b + 1);
    return lookup_in_bucket(I, begin, end - begin, fwd, rc, mz, kmask);
}

/* ------------------------------------------------------------------ stage 1 for one read, one warp */

/* valid characters are exactly ACGTacgt (sshash/kmer.hpp:214-224,258-260); code = (c >> 1) & 3 (kmer.hpp:199).
   With case folded, A C G T are 0x41 + {0, 2, 6, 19}: one range test and one bit test of 0x80045. */
FG_HD bool base_valid(uint32_t c) {
    const uint32_t d = (c & 0xDFu) - 0x41u;
    return d < 20u && ((0x80045u >> d) & 1u);
}

/* mixer_64 (sshash/hash_util.hpp:88-111) is h(y) = (y * FG_MIX_MUL) ^ magic: a bijection of 64-bit words, so the
   m-mer is recovered from its hash with the inverse multiplier (FG_MIX_MUL * FG_MIX_INV == 1 mod 2^64). */
#define FG_MIX_MUL 0x517cc1b727220a95ULL
FG_HD constexpr uint64_t fg_inverse_odd(uint64_t a) {
    uint64_t x = a; /* Newton: correct to 3 bits, doubles per step */
    for (int i = 0; i < 6; ++i) x *= 2 - a * x;
    return x;
}
#define FG_MIX_INV (fgb::fg_inverse_odd(FG_MIX_MUL))
static_assert(FG_MIX_MUL * fg_inverse_odd(FG_MIX_MUL) == 1ULL, "mixer_64 multiplier inverse");

/* Per-warp shared-memory staging of one read segment of up to FG_SEG_KMERS k-mers.
   The lanes own the k-mers of a segment in BLOCKS: lane l owns k-mers 4l .. 4l+3, so the sliding-window minimum and the
   run detection below are sequential inside a lane and need no exchange.
   Position p of the per-m-mer hash arrays is stored at slot p + p/4: lane l's 16 positions then start at slot 5l, an
   odd stride, which keeps the blocked 64-bit reads free of bank conflicts. */
#define FG_SEG_B 4                        /* k-mers per lane per segment */
#define FG_SEG_KMERS (32 * FG_SEG_B)      /* k-mers per segment */
#define FG_HASH_SLOTS 256                 /* >= slot(FG_SEG_KMERS + 30) + 1; 2 * 8 * FG_HASH_SLOTS bytes also hold seeds and items */
#define FG_SEG_WORDS 6                    /* 64-bit words of packed bases: ceil((FG_SEG_KMERS + 30) / 32) + 1 */
#define FG_WORDS_PAD_BEFORE 1             /* readable words before / after the packed bases, for signed stretch offsets */
#define FG_WORDS_PAD_AFTER 3

/* One run of consecutive k-mers with the same minimizer occurrence, in k-mer order. Written by the run's first k-mer as
   {minimizer, key}; the seed pass replaces the minimizer by its bucket. */
struct seed_slot {
    uint32_t begin; /* in: minimizer low word  | out: first super-k-mer of the bucket */
    uint32_t n;     /* in: minimizer high word | out: number of super-k-mers to compare (0 when the run takes the per-k-mer path) */
    uint32_t key;   /* read position of the minimizer | strand << 8 | index of the run's first k-mer << 16 | per-k-mer path << 31 */
    uint32_t size;  /* out: bucket size */
};
#define FG_SEED_SLOW 0x80000000u
static_assert((sizeof(seed_slot) + sizeof(uint2) + sizeof(uint32_t)) * FG_SEG_KMERS <= 2 * 8 * FG_HASH_SLOTS, "seeds + items + per-k-mer notes must fit in the hash arrays");

struct alignas(16) warp_stage { /* 16-byte stores of seed slots and notes */
    uint64_t h[2 * FG_HASH_SLOTS];    /* [0, FG_HASH_SLOTS): mixer_64 hash of the forward m-mer starting at every position;
                                         [FG_HASH_SLOTS, ..): hash of its reverse complement; afterwards seed slots and items */
    uint64_t words[FG_WORDS_PAD_BEFORE + FG_SEG_WORDS + FG_WORDS_PAD_AFTER];  /* packed bases, base j of the segment at bits 2(j % 32) of word j / 32 */
    uint64_t rcw[FG_WORDS_PAD_BEFORE + FG_SEG_WORDS + FG_WORDS_PAD_AFTER];    /* reverse complement of the 32 * nwords packed bases */
    uint32_t valid[FG_SEG_WORDS + 2]; /* bit j: character j of the segment is one of ACGTacgt (only written when some character is not) */
    uint32_t vk[FG_SEG_B + 2];        /* bit i: k-mer i of the segment is valid */
};

FG_HD constexpr uint32_t fg_hslot(uint32_t p) { return p + (p >> 2); }

/* Minimum of every window of W consecutive m-mer hashes among the W + 3 that start at a lane's base slot, for the lane's four
   k-mers -- on 32-BIT KEYS. The hashing step stores, per position p of the segment and per strand, the pair
       x = top 24 bits of the 64-bit hash | p        y = top 24 bits of the 64-bit hash | (255 - p)
   so min(x) is the smallest hash with the LEFTMOST position among equal top bits and min(y) the same with the RIGHTMOST one:
   one VIMNMX (two positions per VIMNMX3) per comparison instead of the 64-bit compare-and-select chain (2 ISETP + 3 SEL).
   Both minima are taken over every window (one suffix scan over the first window, whose suffix minima are the left parts of
   the other three windows, plus a running prefix over the three extra positions). When they name the same position, exactly
   one position of the window carries the smallest top bits: it IS the minimum under the full 64-bit order, whatever the low
   bits are. When they differ -- the window holds its minimizer twice, or two different m-mers agree on 24 hash bits (4.6e-6
   per window) -- the caller recomputes that window exactly (exact_window_minimizer). Ties of the full hash are ties of the
   m-mer (mixer_64 is a bijection): util::compute_minimizer (sshash/util.hpp:220-239) keeps the FIRST minimum of the strand
   it scans, the leftmost position on the forward strand; the reverse strand's m-mers come in the opposite order, so there
   the rightmost position in forward coordinates wins. */
#ifndef FG_KEY_HASH_MASK /* the CPU test tier also runs the kernels with 3 hash bits, so that most windows take the exact path */
#define FG_KEY_HASH_MASK 0xffffff00u
#endif
template <int W>
__device__ __forceinline__ void window_minima32(const uint2* __restrict__ e, uint32_t (&mx)[FG_SEG_B], uint32_t (&my)[FG_SEG_B]) {
    uint2 c = e[fg_hslot(W - 1)];
#pragma unroll
    for (int j = W - 2; j >= 0; --j) {
        const uint2 v = e[fg_hslot(j)];
        c.x = min(c.x, v.x);
        c.y = min(c.y, v.y);
        if (j < FG_SEG_B) {
            mx[j] = c.x;
            my[j] = c.y;
        }
    }
    uint2 p = e[fg_hslot(W)];
    mx[1] = min(mx[1], p.x);
    my[1] = min(my[1], p.y);
#pragma unroll
    for (int t = 2; t < FG_SEG_B; ++t) {
        const uint2 v = e[fg_hslot(W - 1 + t)];
        p.x = min(p.x, v.x);
        p.y = min(p.y, v.y);
        mx[t] = min(mx[t], p.x);
        my[t] = min(my[t], p.y);
    }
}

/* 64 bits (32 bases) of a packed base array starting at base position p, which may be negative or run past the end by
   as much as the padding words allow; w points at the word of base 0 */
__device__ __forceinline__ uint64_t stretch_at(const uint64_t* w, int p) {
    const int wi = p >> 5; /* floor */
    const uint32_t sh = uint32_t(p & 31) * 2;
    const uint64_t a = w[wi];
    return sh ? (a >> sh) | (w[wi + 1] << (64 - sh)) : a;
}

/* the m-mer (2m <= 64 bits, mask) that starts at base position p >= 0 of a packed base array, read through 32-bit words:
   three loads and two funnel shifts */
__device__ __forceinline__ uint64_t mmer_at(const uint64_t* w, uint32_t p, uint64_t mask) {
    const uint32_t* u = reinterpret_cast<const uint32_t*>(w);
    const uint32_t wi = p >> 4, sh = (p & 15u) * 2;
    const uint32_t a = u[wi], b = u[wi + 1], c = u[wi + 2];
    return (uint64_t(__funnelshift_r(a, b, sh)) | (uint64_t(__funnelshift_r(b, c, sh)) << 32)) & mask;
}

/* util::compute_minimizer (sshash/util.hpp:220-239) on both strands of ONE k-mer, straight from the packed bases, under the
   full 64-bit hash order: the exact answer for the windows where window_minima32's 24-bit keys cannot decide. fw / rc point at
   the packed stream and its reverse complement, fpos = stream position of the k-mer's first base, rc_len = length of the
   reverse-complemented stream in bases. pf / pr = offset of the winning m-mer inside the k-mer, forward coordinates. */
__device__ __noinline__ void exact_window_minimizer(const uint64_t* fw, const uint64_t* rc, uint32_t fpos, uint32_t rc_len, uint32_t window, uint32_t m,
                                                    uint64_t mmer_mask, uint64_t magic, uint32_t& pf, uint32_t& pr) {
    uint64_t bf = 0, br = 0;
    pf = pr = 0;
    for (uint32_t j = 0; j < window; ++j) {
        const uint64_t hf = (mmer_at(fw, fpos + j, mmer_mask) * FG_MIX_MUL) ^ magic;
        const uint64_t hr = (mmer_at(rc, rc_len - (fpos + j) - m, mmer_mask) * FG_MIX_MUL) ^ magic;
        if (j == 0 || hf < bf) bf = hf, pf = j;
        if (j == 0 || hr <= br) br = hr, pr = j;
    }
}

/* Walks one read and reports its positive k-mers as ITEMS {color-set id, number of k-mers}: after next(), lane l holds one
   item (cnt = 0: none). Over the calls until next() returns false every positive k-mer of the read is counted in exactly
   one item; the consumers -- intersection, per-color scores, distinct-set table -- only need that multiset.
   Replaces the per-k-mer loop around streaming_query::lookup_advanced (sshash/streaming_query.hpp:50-109, driven from
   src/ps_full_intersection.cpp:344-353 and src/ps_threshold_union.cpp:327-347): validity test (:53-59), 2-bit packing and
   reverse complement (:62-74), canonical minimizer (:76-83), dictionary lookup by seed (:144-190) and extend (:111-142).

   Per segment of FG_SEG_KMERS k-mers the warp
     0. packs the bases (2 bits each), their reverse complement and their validity bits into shared memory;
     1. hashes every m-mer ONCE, both strands (the k-mers of a read share almost all their m-mers, which the reference
        exploits with its sliding minimizer_enumerator, sshash/minimizer_enumerator.hpp:24-49);
     2. takes the sliding-window minima (window_minima) and from them every k-mer's canonical minimizer;
     3. cuts the k-mers into runs that share one minimizer OCCURRENCE (same value, same read position, same strand) and
        makes the first k-mer of each run a SEED;
     4. seed pass, one seed per lane: minimizer MPHF -> bucket;
     5. pair pass, one (seed, super-k-mer of its bucket) pair per lane: ONE comparison of the stored string with the read,
        aligned on the minimizer (the record knows where the minimizer sits in the super-k-mer, the seed knows where it
        sits in the read). The mismatch-free stretch around the minimizer gives the interval of read k-mers that are in
        the dictionary; clipped to the run and counted over the valid k-mers it becomes one item. This is the device
        analogue of the reference's seed-and-extend: ~1 hash lookup and ~1.2 string comparisons per run of ~5 k-mers.
   A lookup answer is a pure function of the k-mer (streaming_query.hpp:107 asserts it) and every canonical k-mer is
   stored once, so "read bases == stored bases over the k-mer's span, inside one super-k-mer" IS the lookup answer; a
   k-mer can only be stored in the bucket of its own minimizer, so a run needs no other bucket.
   Runs whose bucket is served by the skew index and k-mers whose minimizer value appears on both strands take the
   per-k-mer path (lookup_in_bucket); super-k-mers without a single minimizer position are scanned k-mer by k-mer. */
/* A read in PACKED form (include/fulgor_gpu.h, fulgor_gpu_pack_reads): 2-bit codes, 16 per 32-bit word, the read starting on a
   word; the characters that are not ACGTacgt are listed apart (ascending positions, in bases from the start of the packed
   buffer) and only looked at when the read is flagged. A quarter of the ASCII bytes over PCIe, and no packing in the kernel. */
struct packed_read {
    const uint32_t* words;  /* first word of the read */
    uint32_t len;           /* bases */
    bool flagged;           /* some character is not one of ACGTacgt */
    const uint64_t* invalid; /* ascending positions of such characters (all reads of the launch) */
    uint32_t n_invalid;
    uint64_t pos;           /* position of the read's first base in the coordinates of `invalid` */
};

template <int W, bool PERK = false, bool PACKED = false>
struct kmer_tiles {
    const dev_index& I;
    const uint8_t* __restrict__ seq;
    const uint8_t *buf_begin, *buf_end; /* the bases buffer the read lives in: aligned 4-byte loads stay inside it */
    packed_read pr;                     /* PACKED: the read */
    warp_stage& S;
    uint32_t len, lane, nk, seg_end, nitems, cursor;
    uint32_t shift; /* the packed stream starts at the 4-byte aligned address at or below the segment: base j sits at stream position j + shift */
    uint64_t kmask, mmer_mask;
    uint32_t window, kbits;
    /* PERK: besides the items, the color-set id of EVERY k-mer (FG_NOT_FOUND = negative or invalid) goes to per_kmer[i], i =
       k-mer index in the read -- the per-k-mer view of streaming_query::lookup_advanced that index::kmer_conservation and
       index::kmer_matches consume (src/kmer_conservation.cpp:31-47, src/kmer_matches.cpp:20-28). Set by the caller. */
    uint32_t* per_kmer;
    uint32_t seg_base; /* read index of the current segment's first k-mer */
    bool seg_all_valid; /* every character of the current segment is one of ACGTacgt (warp-uniform) */

    __device__ __forceinline__ kmer_tiles(const dev_index& I_, const uint8_t* seq_, uint32_t len_, const uint8_t* buf_begin_, const uint8_t* buf_end_,
                                          uint32_t lane_, warp_stage& S_)
        : I(I_), seq(seq_), buf_begin(buf_begin_), buf_end(buf_end_), S(S_), len(len_), lane(lane_) {
        init();
    }
    __device__ __forceinline__ kmer_tiles(const dev_index& I_, const packed_read& pr_, uint32_t lane_, warp_stage& S_)
        : I(I_), seq(nullptr), buf_begin(nullptr), buf_end(nullptr), pr(pr_), S(S_), len(pr_.len), lane(lane_) {
        init();
    }
    __device__ __forceinline__ void init() {
        const uint32_t k = I.k;
        nk = len >= k ? len - k + 1 : 0; /* src/ps_full_intersection.cpp:337: shorter reads have no k-mers */
        seg_end = 0;
        nitems = cursor = 0;
        shift = 0;
        kmask = (1ULL << (2 * k)) - 1;
        mmer_mask = (1ULL << (2 * I.m)) - 1;
        window = W ? W : k - I.m + 1;
        kbits = (1u << k) - 1u;
        per_kmer = nullptr;
        seg_base = 0;
    }

    __device__ __forceinline__ const uint64_t* words() const { return S.words + FG_WORDS_PAD_BEFORE; }
    __device__ __forceinline__ const uint64_t* rcwords() const { return S.rcw + FG_WORDS_PAD_BEFORE; }
    /* 2*nbases bits starting at base position p (segment-relative) of the packed words */
    __device__ __forceinline__ uint64_t bases_at(uint32_t p) const { return stretch_at(words(), int(p + shift)); }
    __device__ __forceinline__ uint64_t* hf() const { return S.h; }
    __device__ __forceinline__ uint64_t* hr() const { return S.h + FG_HASH_SLOTS; }
    __device__ __forceinline__ uint2* kf() const { return reinterpret_cast<uint2*>(S.h); } /* the 32-bit key pairs of window_minima32 */
    __device__ __forceinline__ uint2* kr() const { return reinterpret_cast<uint2*>(S.h + FG_HASH_SLOTS); }
    /* position q of the segment: forward m-mer yf, its reverse complement yr */
    __device__ __forceinline__ void store_hashes(uint32_t q, uint64_t yf, uint64_t yr) const {
        if (W) { /* only bits 32..63 of the hashes are needed */
            const uint32_t ka = ((fg_mul64_hi32(yf, FG_MIX_MUL) ^ uint32_t(I.hash_magic >> 32)) & FG_KEY_HASH_MASK) | q;
            const uint32_t kb = ((fg_mul64_hi32(yr, FG_MIX_MUL) ^ uint32_t(I.hash_magic >> 32)) & FG_KEY_HASH_MASK) | q;
            kf()[fg_hslot(q)] = make_uint2(ka, ka ^ 0xffu);
            kr()[fg_hslot(q)] = make_uint2(kb, kb ^ 0xffu);
        } else {
            hf()[fg_hslot(q)] = (yf * FG_MIX_MUL) ^ I.hash_magic;
            hr()[fg_hslot(q)] = (yr * FG_MIX_MUL) ^ I.hash_magic;
        }
    }
    __device__ __forceinline__ seed_slot* seeds() const { return reinterpret_cast<seed_slot*>(S.h); }
    __device__ __forceinline__ uint2* items() const { return reinterpret_cast<uint2*>(seeds() + FG_SEG_KMERS); }
    __device__ __forceinline__ uint32_t* notes() const { return reinterpret_cast<uint32_t*>(items() + FG_SEG_KMERS); } /* per k-mer, for the per-k-mer path */

    /* m-mer behind a window minimum */
    __device__ __forceinline__ uint64_t mmer_of_hash(uint64_t h) const { return (h ^ I.hash_magic) * FG_MIX_INV; }

    /* k-mer i of the segment (i < seg_nk) holds no invalid character; S.vk is only written for segments that have one */
    __device__ __forceinline__ bool kmer_valid(uint32_t i) const { return seg_all_valid || ((S.vk[i >> 5] >> (i & 31)) & 1u); }
    /* number of valid k-mers with index in [lo, hi], hi - lo < 32 */
    __device__ __forceinline__ uint32_t count_valid(uint32_t lo, uint32_t hi) const {
        if (seg_all_valid) return hi - lo + 1; /* [lo, hi] lies inside the segment */
        const uint32_t bits = __funnelshift_r(S.vk[lo >> 5], S.vk[(lo >> 5) + 1], lo & 31);
        const uint32_t len = hi - lo + 1;
        return __popc(len >= 32 ? bits : bits & ((1u << len) - 1u));
    }

    __device__ __forceinline__ void append_items(bool have, uint32_t cid, uint32_t cnt) {
        const uint32_t b = __ballot_sync(FG_FULL, have);
        if (have) items()[nitems + __popc(b & ((1u << lane) - 1u))] = make_uint2(cid, cnt);
        nitems += __popc(b);
    }

    /* step 5 for one (seed, super-k-mer) pair: number of k-mers of the run [i_first, i_last] found in super-k-mer sk */
    __device__ __forceinline__ uint32_t extend_pair(uint32_t sk, int p, bool fw, uint32_t i_first, uint32_t i_last, uint32_t nchars,
                                                    uint32_t nwords, uint32_t& cid, int& found_lo, int& found_hi) const {
        found_lo = 0;
        found_hi = -1; /* PERK: the found k-mers are the valid ones in [found_lo, found_hi] (the unpinned path writes its own) */
        const int k = int(I.k), m = int(I.m);
        const uint2 rec = FG_LDG(I.sk_records + sk);
        cid = I.sk_cid ? FG_LDG(I.sk_cid + sk) : (rec.y & FGI_SK_CID_MASK);
        const int win = int((rec.y >> FGI_SK_WINDOW_SHIFT) & 31u), pm = int((rec.y >> FGI_SK_PM_SHIFT) & 31u);
        if (!(rec.y >> FGI_SK_PINNED_SHIFT)) { /* no single minimizer position: every k-mer of the run against every stored k-mer */
            uint32_t cnt = 0;
            for (uint32_t i = i_first; i <= i_last; ++i) {
                if (!kmer_valid(i)) continue;
                const uint64_t fwd = bases_at(i) & kmask;
                const bool found = scan_super_kmer_rare(I, sk, fwd, kmask);
                cnt += found;
                if (PERK && found) per_kmer[seg_base + i] = cid;
            }
            return cnt;
        }
        const uint64_t bit = 2 * uint64_t(rec.x & FGI_SK_OFFSET_MASK);
        const uint64_t* w = I.strings + (bit >> 6);
        const uint32_t sh = uint32_t(bit & 63);
        const uint64_t w0 = FG_LDG(w), w1 = FG_LDG(w + 1), w2 = FG_LDG(w + 2);
        const uint64_t s_lo = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0; /* the super-k-mer's win + k - 1 <= 61 bases */
        const uint64_t s_hi = sh ? (w1 >> sh) | (w2 << (64 - sh)) : w1;
        /* the stored m-mer at the minimizer position reads as the minimizer itself or as its reverse complement (record bit);
           so does the read's (fw). Same reading on both sides: the read aligns forward, otherwise reverse-complemented. */
        const bool forward = bool(rec.x >> FGI_SK_CANON_FWD_SHIFT) == fw;
        const int len_s = win + k - 1;
        const int d = p - pm;               /* forward: stored base j <-> read base j + d, k-mer index i = t + d */
        const int e = pm + p + m - k;       /* reverse: stored base j <-> complement of read base e + k - 1 - j, i = e - t */
        const uint64_t* src = forward ? words() : rcwords();
        const int off = forward ? d + int(shift) : int(32 * nwords) - int(shift) - e - k;
        const int j_min = forward ? max(0, -d) : max(0, e + k - int(nchars));
        const int j_max = forward ? min(len_s, int(nchars) - d) : min(len_s, e + k);
        const uint64_t x_lo = s_lo ^ stretch_at(src, off), x_hi = s_hi ^ stretch_at(src, off + 32);
        const uint64_t m_lo = (x_lo | (x_lo >> 1)) & 0x5555555555555555ULL, m_hi = (x_hi | (x_hi >> 1)) & 0x5555555555555555ULL;
        /* a = first stored position after the last mismatch LEFT of the minimizer (pm < 32: the low word only); b = first mismatch
           at or right of the minimizer's start. A k-mer at stored position t is present iff a <= t and t + k <= b. (A mismatch
           inside the minimizer makes b < pm + m: what is still "present" then ends before the minimizer does, outside the run,
           and the clip to the run's k-mers below drops it.) */
        const uint64_t left = (1ULL << (2 * pm)) - 1;
        const uint64_t l_lo = m_lo & left, g_lo = m_lo & ~left;
        const int a = l_lo ? ((64 - __clzll((long long)l_lo)) >> 1) + 1 : 0;
        int b = 64;
        if (g_lo) b = (__ffsll((long long)g_lo) - 1) >> 1;
        else if (m_hi) b = 32 + ((__ffsll((long long)m_hi) - 1) >> 1);
        const int t_lo = max(a, j_min), t_hi = min(min(b, j_max) - k, win - 1);
        if (t_lo > t_hi) return 0;
        const int lo = max(forward ? t_lo + d : e - t_hi, int(i_first)), hi = min(forward ? t_hi + d : e - t_lo, int(i_last));
        if (lo > hi) return 0;
        found_lo = lo;
        found_hi = hi;
        return count_valid(uint32_t(lo), uint32_t(hi));
    }

    __device__ __forceinline__ void load_segment() {
        __syncwarp();
        const uint32_t seg0 = seg_end;
        const uint32_t seg_nk = min(uint32_t(FG_SEG_KMERS), nk - seg0);
        seg_end = seg0 + seg_nk;
        nitems = cursor = 0;
        if (PERK) { /* every k-mer starts negative; the warp barriers before the pair pass order these stores before the hits */
            seg_base = seg0;
            for (uint32_t i = lane; i < seg_nk; i += 32) per_kmer[seg0 + i] = FG_NOT_FOUND;
        }
        const uint32_t nchars = seg_nk + I.k - 1, npos = seg_nk + window - 1;
        /* 0. bases. Lane l loads the l-th aligned 4-byte word at or after the segment's aligned floor, packs its four
              characters to one byte (2 bits each) and stores it: the packed stream is the byte array over S.words. */
        const uint8_t* p0 = seq + seg0;
        shift = PACKED ? 0u : uint32_t(reinterpret_cast<uintptr_t>(p0) & 3u);
        const uint32_t nstream = nchars + shift;
        const uint32_t nwords = (nstream + 31) >> 5;
        bool all_valid;
        if (PACKED) {
            /* the segment starts on a word (FG_SEG_KMERS is a multiple of 16): lane l copies 32-bit word l, bits past the
               segment's last character cleared like the 'A' padding of the ASCII path */
            const uint32_t n32 = (nchars + 15) >> 4;
            if (lane < 2 * nwords) {
                uint32_t x = lane < n32 ? __ldg(pr.words + (seg0 >> 4) + lane) : 0u;
                if (lane + 1 == n32 && (nchars & 15u)) x &= (1u << (2 * (nchars & 15u))) - 1u;
                reinterpret_cast<uint32_t*>(S.words + FG_WORDS_PAD_BEFORE)[lane] = x;
            }
            all_valid = !pr.flagged;
            if (!all_valid) { /* rare: validity bits from the list of invalid positions that fall into this segment */
                const uint32_t nvw = (nchars + 32 + 31) >> 5;
                if (lane < nvw) S.valid[lane] = nchars >= 32 * lane + 32 ? ~0u : (nchars > 32 * lane ? (1u << (nchars - 32 * lane)) - 1u : 0u);
                __syncwarp();
                const uint64_t lo_pos = pr.pos + seg0, hi_pos = lo_pos + nchars;
                uint32_t lo = 0, hi = pr.n_invalid; /* first entry >= lo_pos */
                while (lo < hi) {
                    const uint32_t mid = lo + ((hi - lo) >> 1);
                    if (__ldg(pr.invalid + mid) < lo_pos) lo = mid + 1; else hi = mid;
                }
                for (uint32_t i = lo + lane; i < pr.n_invalid; i += 32) {
                    const uint64_t at = __ldg(pr.invalid + i);
                    if (at >= hi_pos) break;
                    const uint32_t j = uint32_t(at - lo_pos);
                    atomicAnd(&S.valid[j >> 5], ~(1u << (j & 31)));
                }
            }
        } else {
        bool bad = false;
        for (uint32_t w0 = 0; w0 < 8 * nwords; w0 += 32) {
            const uint32_t wi = w0 + lane;
            const int bo = int(4 * wi) - int(shift); /* segment index of the word's first character */
            uint32_t x = 0x41414141u;                /* characters outside the segment read as 'A' (code 0) */
            if (bo < int(nchars)) {
                const uint8_t* q = p0 + bo;
                if (q >= buf_begin && q + 4 <= buf_end) {
                    x = __ldg(reinterpret_cast<const uint32_t*>(q));
                } else { /* the word straddles an end of the buffer */
                    x = 0;
                    for (int bb = 0; bb < 4; ++bb) x |= uint32_t(q + bb >= buf_begin && q + bb < buf_end ? q[bb] : uint8_t('A')) << (8 * bb);
                }
                const uint32_t drop_lo = bo < 0 ? uint32_t(-bo) : 0u, drop_hi = bo + 4 > int(nchars) ? uint32_t(bo + 4 - int(nchars)) : 0u;
                if (drop_lo | drop_hi) {
                    const uint32_t keep = (0xffffffffu << (8 * drop_lo)) & (0xffffffffu >> (8 * drop_hi));
                    x = (x & keep) | (0x41414141u & ~keep);
                }
            }
            const uint32_t codes = (x >> 1) & 0x03030303u;                /* (c >> 1) & 3 per character (sshash/kmer.hpp:199) */
            const uint32_t packed = (codes * 0x01041040u) >> 24;         /* c0 | c1 << 2 | c2 << 4 | c3 << 6 */
            /* valid characters are exactly ACGTacgt: rebuild the upper-case letter of every code and compare */
            uint32_t sel = (packed | (packed << 4)) & 0x0f0fu;
            sel = (sel | (sel << 2)) & 0x3333u;
            bad |= __byte_perm(0x47544341u, 0u, sel) != (x & 0xdfdfdfdfu);
            if (wi < 8 * nwords) reinterpret_cast<uint8_t*>(S.words + FG_WORDS_PAD_BEFORE)[wi] = uint8_t(packed);
        }
        all_valid = __ballot_sync(FG_FULL, bad) == 0;
        if (!all_valid) { /* rare: per-character validity bits (segment coordinates) */
            for (uint32_t c = 0; 32 * c < nchars + 32; ++c) {
                const uint32_t j = 32 * c + lane;
                const uint32_t v = __ballot_sync(FG_FULL, j < nchars && base_valid(p0[j < nchars ? j : 0]));
                if (lane == 0) S.valid[c] = v;
            }
        }
        }
        __syncwarp();
        if (lane < nwords) S.rcw[FG_WORDS_PAD_BEFORE + nwords - 1 - lane] = revcomp(S.words[FG_WORDS_PAD_BEFORE + lane], 32);
        if (lane <= FG_WORDS_PAD_AFTER) { /* readable padding: values never reach a valid k-mer */
            if (lane < FG_WORDS_PAD_AFTER) {
                S.words[FG_WORDS_PAD_BEFORE + nwords + lane] = 0;
                S.rcw[FG_WORDS_PAD_BEFORE + nwords + lane] = 0;
            } else {
                S.words[0] = 0;
                S.rcw[0] = 0;
            }
        }
        __syncwarp();
        /* 1. m-mer hashes. Lane l hashes the m-mers at positions 5l .. 5l+4 out of ONE 32-base window (m + 4 <= 32): the
              reverse complement of the window holds the five reverse-complemented m-mers too. The templated windows keep a
              pair of 32-bit keys per position and strand (window_minima32), the generic one the 64-bit hashes. */
        const uint32_t m = I.m;
        /* the templated windows are only dispatched for m >= 16 (engine.cu: dispatch_window): the mask's low word is all ones */
        const uint64_t wmask = W ? (mmer_mask | 0xffffffffULL) : mmer_mask;
        if (W || m <= 28) { /* a templated window means m <= k - 10 */
            const uint32_t q0 = 5 * lane;
            /* The key pairs are written for all 160 positions the lanes' windows can touch: a key carries its own position,
               so whatever bases lie past the segment yield in-range positions for the (invalid) k-mers that see them. */
            if (W || q0 < npos) {
                const uint64_t x = bases_at(q0) & (m + 4 == 32 ? ~0ULL : ((1ULL << (2 * (m + 4))) - 1));
                const uint64_t r = revcomp(x, m + 4);
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    if (W || q0 + j < npos) store_hashes(q0 + j, (x >> (2 * j)) & wmask, (r >> (2 * (4 - j))) & wmask);
                }
            }
        } else {
            for (uint32_t q = lane; q < npos; q += 32) {
                const uint64_t y = bases_at(q) & mmer_mask;
                store_hashes(q, y, revcomp(y, m));
            }
        }
        __syncwarp();
        /* 2. canonical minimizer of the lane's k-mers i0 .. i0+3: value = min over both strands, compared as integers
              (sshash/streaming_query.hpp:76-79); cpos = where it starts inside the k-mer, forward coordinates */
        const uint32_t i0 = FG_SEG_B * lane;
        uint64_t val[FG_SEG_B];
        uint32_t key[FG_SEG_B]; /* read position of the minimizer | strand << 8 | ambiguous << 13 | valid << 14 | cpos << 16 */
        {
            uint32_t pf[FG_SEG_B], pr[FG_SEG_B]; /* where the minimizer of each strand starts inside the k-mer, forward coordinates */
            uint64_t vf[FG_SEG_B], vr[FG_SEG_B]; /* and its value */
            const uint32_t rc_len = 32 * nwords; /* bases of the reverse-complemented stream */
            if (W) {
                uint32_t fx[FG_SEG_B], fy[FG_SEG_B], rx[FG_SEG_B], ry[FG_SEG_B];
                window_minima32<W ? W : 13>(kf() + 5 * lane, fx, fy);
                window_minima32<W ? W : 13>(kr() + 5 * lane, rx, ry);
                uint32_t undecided = 0;
#pragma unroll
                for (int tt = 0; tt < FG_SEG_B; ++tt) {
                    /* (clamped to the segment's m-mers: k-mers past the segment's end see keys of whatever lies there) */
                    pf[tt] = min(fx[tt] & 0xffu, npos - 1) - (i0 + tt);           /* leftmost smallest key */
                    pr[tt] = min((ry[tt] & 0xffu) ^ 0xffu, npos - 1) - (i0 + tt); /* rightmost smallest key */
                    /* x ^ y == 0xff iff both minima sit on the same position */
                    if ((((fx[tt] ^ fy[tt]) ^ 0xffu) | ((rx[tt] ^ ry[tt]) ^ 0xffu)) != 0) undecided |= 1u << tt;
                }
                undecided &= i0 + FG_SEG_B <= seg_nk ? 0xfu : (i0 < seg_nk ? (1u << (seg_nk - i0)) - 1u : 0u); /* the lane's k-mers inside the segment */
                if (__ballot_sync(FG_FULL, undecided != 0)) { /* rare */
#pragma unroll 1
                    for (uint32_t tt = 0; tt < FG_SEG_B; ++tt) {
                        if ((undecided >> tt) & 1u) {
                            uint32_t ef, er;
                            exact_window_minimizer(words(), rcwords(), i0 + tt + shift, rc_len, window, m, mmer_mask, I.hash_magic, ef, er);
#pragma unroll
                            for (int u = 0; u < FG_SEG_B; ++u)
                                if (uint32_t(u) == tt) pf[u] = ef, pr[u] = er;
                        }
                    }
                }
#pragma unroll
                for (int tt = 0; tt < FG_SEG_B; ++tt) {
                    vf[tt] = mmer_at(words(), i0 + tt + pf[tt] + shift, wmask);
                    vr[tt] = mmer_at(rcwords(), rc_len - (i0 + tt + pr[tt] + shift) - m, wmask);
                }
            } else { /* any other (k, m): plain scan of each window */
#pragma unroll
                for (int tt = 0; tt < FG_SEG_B; ++tt) {
                    uint64_t mhf = UINT64_MAX, mhr = UINT64_MAX;
                    pf[tt] = pr[tt] = 0;
                    for (uint32_t j = 0; j < window; ++j) {
                        const uint64_t a = hf()[fg_hslot(i0 + tt + j)], b = hr()[fg_hslot(i0 + tt + j)];
                        if (a < mhf) mhf = a, pf[tt] = j;
                        if (b <= mhr) mhr = b, pr[tt] = j;
                    }
                    vf[tt] = mmer_of_hash(mhf);
                    vr[tt] = mmer_of_hash(mhr);
                    if (I.guard_max_hash) { /* compute_minimizer's "nothing below UINT64_MAX" sentinel when that can occur (image.h);
                                               such indexes are dispatched to this window (engine.cu: dispatch_window) */
                        if (mhf == UINT64_MAX) vf[tt] = UINT64_MAX;
                        if (mhr == UINT64_MAX) vr[tt] = UINT64_MAX;
                    }
                }
            }
            /* i0 is a multiple of 4, so the lane's four k-mers start in the same 32-character word */
            const uint32_t vw0 = all_valid ? ~0u : S.valid[i0 >> 5], vw1 = all_valid ? ~0u : S.valid[(i0 >> 5) + 1];
            uint32_t nibble = 0;
#pragma unroll
            for (int tt = 0; tt < FG_SEG_B; ++tt) {
                const uint32_t i = i0 + tt;
                const uint32_t bits = __funnelshift_r(vw0, vw1, i & 31);
                const bool valid = i < seg_nk && (bits & kbits) == kbits;
                const bool fwd_wins = vf[tt] <= vr[tt];
                val[tt] = fwd_wins ? vf[tt] : vr[tt];
                const uint32_t cpos = fwd_wins ? pf[tt] : pr[tt];
                key[tt] = (i + cpos) | (uint32_t(fwd_wins) << 8) | (uint32_t(vf[tt] == vr[tt]) << 13) | (uint32_t(valid) << 14) | (cpos << 16);
                nibble |= uint32_t(valid) << tt;
            }
            /* valid-k-mer bitmap in k-mer order: 8 lanes per 32-bit word */
            seg_all_valid = all_valid;
            if (!all_valid) {
                const uint32_t part = nibble << (4 * (lane & 7));
#pragma unroll
                for (int w = 0; w < FG_SEG_B; ++w) {
                    const uint32_t word = __reduce_or_sync(FG_FULL, (lane >> 3) == uint32_t(w) ? part : 0u);
                    if (lane == 0) S.vk[w] = word;
                }
                if (lane == 0) S.vk[FG_SEG_B] = S.vk[FG_SEG_B + 1] = 0;
            }
        }
        __syncwarp(); /* every lane is done with the hash arrays: seeds and items may overwrite them */
        /* 3. seeds: the first k-mer of every run of valid k-mers with the same minimizer occurrence -- same value, at the same
              read position, read off the same strand. A k-mer whose minimizer value shows up on both strands (ambiguous)
              is a run of its own. Seeds are numbered in k-mer order. */
        uint32_t nseeds;
        {
            uint32_t prev_key = __shfl_up_sync(FG_FULL, key[FG_SEG_B - 1], 1);
            if (lane == 0) prev_key = 0;
            uint32_t leaders = 0;
#pragma unroll
            for (int tt = 0; tt < FG_SEG_B; ++tt) {
                /* same run: both valid, neither ambiguous, same read position and strand (hence the same m-mer) */
                const bool same = ((key[tt] ^ prev_key) & 0xffffu) == 0 && (key[tt] & 0x6000u) == 0x4000u;
                if (((key[tt] >> 14) & 1u) && !same) leaders |= 1u << tt;
                prev_key = key[tt];
            }
            uint32_t incl = __popc(leaders); /* inclusive prefix sum over the lanes */
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                const uint32_t y = __shfl_up_sync(FG_FULL, incl, dlt);
                if (lane >= uint32_t(dlt)) incl += y;
            }
            nseeds = __shfl_sync(FG_FULL, incl, 31);
            uint32_t s = incl - __popc(leaders); /* seeds before this lane */
            uint32_t note[FG_SEG_B];
#pragma unroll
            for (int tt = 0; tt < FG_SEG_B; ++tt) {
                if ((leaders >> tt) & 1u) { /* one 16-byte store: {minimizer, key, -} */
                    *reinterpret_cast<uint4*>(seeds() + s) =
                        make_uint4(uint32_t(val[tt]), uint32_t(val[tt] >> 32), (key[tt] & 0x1ffu) | ((i0 + tt) << 16) | (((key[tt] >> 13) & 1u) ? FG_SEED_SLOW : 0u), 0u);
                    s += 1;
                }
                /* note for the per-k-mer path: run | ambiguous << 13 | valid << 14 | cpos << 16 (the run wraps to 0xff before the
                   first seed: such k-mers are invalid) */
                note[tt] = ((s - 1) & 0xffu) | (key[tt] & 0x001f6000u);
            }
            reinterpret_cast<uint4*>(notes())[lane] = make_uint4(note[0], note[1], note[2], note[3]); /* k-mer i at notes()[i] */
        }
        __syncwarp();
        /* 4. + 5. in groups of 32 seeds */
        bool any_slow = false;
        for (uint32_t g0 = 0; g0 < nseeds; g0 += 32) {
            /* minimizers::lookup + buckets::locate_bucket, one seed per lane */
            uint32_t n = 0;
            if (g0 + lane < nseeds) {
                seed_slot& slot = seeds()[g0 + lane];
                const uint64_t mini = uint64_t(slot.begin) | (uint64_t(slot.n) << 32);
                const uint64_t b = minimizer_bucket(I, mini);
                const uint32_t begin = FG_LDG(I.bucket_begin + b), size = FG_LDG(I.bucket_begin + b + 1) - begin;
                if (size > I.skew_threshold) slot.key |= FG_SEED_SLOW; /* skew index: per-k-mer path */
                /* A minimizer that is its own reverse complement (possible for even m only) reads the same off both strands, so how
                   it reads says nothing about the orientation of the stored string against the read -- which the pair pass relies
                   on. The per-k-mer path compares both orientations. */
                if (!(I.m & 1u) && revcomp(mini, I.m) == mini) slot.key |= FG_SEED_SLOW;
                n = (slot.key & FG_SEED_SLOW) ? 0u : size;
                any_slow |= (slot.key & FG_SEED_SLOW) != 0;
                slot.begin = begin;
                slot.n = n;
                slot.size = size;
            }
            uint32_t end = n; /* inclusive prefix sum of the pair counts */
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                const uint32_t y = __shfl_up_sync(FG_FULL, end, dlt);
                if (lane >= uint32_t(dlt)) end += y;
            }
            const uint32_t total = __shfl_sync(FG_FULL, end, 31);
            __syncwarp();
            /* one (seed, super-k-mer) pair per lane */
            for (uint32_t q0 = 0; q0 < total; q0 += 32) {
                const uint32_t q = q0 + lane;
                uint32_t owner = 0; /* the lane whose seed owns pair q: the first with end > q */
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const uint32_t probe = __shfl_sync(FG_FULL, end, owner + step - 1);
                    if (probe <= q) owner += step;
                }
                const uint32_t owner_end = __shfl_sync(FG_FULL, end, owner & 31u);
                uint32_t cid = 0, cnt = 0;
                if (q < total) {
                    const seed_slot slot = seeds()[g0 + owner];
                    const uint32_t j = q - (owner_end - slot.n);
                    const uint32_t i_first = (slot.key >> 16) & 0xffu;
                    const uint32_t i_last = (g0 + owner + 1 < nseeds ? ((seeds()[g0 + owner + 1].key >> 16) & 0xffu) : seg_nk) - 1;
                    int found_lo, found_hi;
                    cnt = extend_pair(slot.begin + j, int(slot.key & 0xffu), (slot.key >> 8) & 1u, i_first, i_last, nchars, nwords, cid, found_lo, found_hi);
                    if (PERK && cnt)
                        for (int i = found_lo; i <= found_hi; ++i)
                            if (kmer_valid(uint32_t(i))) per_kmer[seg_base + uint32_t(i)] = cid;
                }
                append_items(cnt != 0, cid, cnt);
            }
        }
        /* per-k-mer path: ambiguous k-mers and runs behind the skew index */
        if (__ballot_sync(FG_FULL, any_slow)) {
#pragma unroll 1
            for (int tt = 0; tt < FG_SEG_B; ++tt) {
                uint32_t cid = FG_NOT_FOUND;
                const uint32_t note = notes()[FG_SEG_B * lane + tt];
                if ((note >> 14) & 1u) {
                    const seed_slot slot = seeds()[note & 0xffu];
                    if (slot.key & FG_SEED_SLOW) {
                        const uint64_t fwd = bases_at(FG_SEG_B * lane + tt) & kmask;
                        cid = lookup_in_bucket_rare(I, slot.begin, slot.size, fwd, revcomp(fwd, I.k), (note >> 16) & 31u, (note >> 13) & 1u, kmask);
                        if (PERK && cid != FG_NOT_FOUND) per_kmer[seg_base + FG_SEG_B * lane + tt] = cid;
                    }
                }
                append_items(cid != FG_NOT_FOUND, cid, 1);
            }
        }
        __syncwarp();
    }

    /* the next batch of items, one per lane (cnt = 0: none); false when the read is exhausted (warp-uniform) */
    __device__ __forceinline__ bool next(uint32_t& cid, uint32_t& cnt) {
        while (cursor >= nitems) {
            if (seg_end >= nk) return false;
            load_segment();
        }
        cid = FG_NOT_FOUND;
        cnt = 0;
        if (cursor + lane < nitems) {
            const uint2 it = items()[cursor + lane];
            cid = it.x;
            cnt = it.y;
        }
        cursor += 32;
        return true;
    }
};

/* a growable list of {color-set id, multiplicity} in global memory, bump-allocated from one pool per launch */
struct entry_pool {
    uint2* base;
    unsigned long long* used; /* entries handed out so far */
    uint64_t cap;             /* entries available */
    uint32_t* exhausted;      /* set when an allocation failed: the host grows the pool and reruns */
};

struct read_hits {
    uint32_t cid;   /* n <= 32: lane j < n holds the j-th distinct color-set id (ascending) ... */
    uint32_t cnt;   /*          ... and the number of positive k-mers that map to it */
    uint32_t n;     /* number of distinct color sets (warp-uniform) */
    uint32_t npos;  /* number of positive k-mers (warp-uniform) */
    uint2* tab;     /* n > 32: the sorted list lives here (shared or pool memory); nullptr otherwise */
    uint32_t cap;   /* capacity of tab (a power of two) */
    bool failed;    /* the pool was exhausted: this read's result is incomplete */
};

__device__ __forceinline__ uint32_t next_pow2(uint32_t v) { return v <= 1 ? 1u : 1u << (32 - __clz(int(v - 1))); }

/* The per-read list beyond 32 distinct color sets (shared memory first, then the pool). Entries are APPENDED without
   looking for their color-set id -- a batch of items costs one ballot and one store per lane -- so the list may hold an id
   several times; table_compact() sorts it and merges equal ids, at the end of the read and whenever the list is full. */

/* room for `extra` more entries; false when the pool is exhausted (the host grows it and reruns) */
__device__ __noinline__ bool table_reserve(read_hits& R, uint32_t extra, uint32_t nk, const entry_pool& pool, uint32_t lane);
__device__ __noinline__ void table_compact(read_hits& R, uint32_t lane);

/* appends {cid, cnt} of every lane with have = true, in lane order */
__device__ __forceinline__ void table_append(read_hits& R, bool have, uint32_t cid, uint32_t cnt, uint32_t nk, const entry_pool& pool, uint32_t lane) {
    const uint32_t b = __ballot_sync(FG_FULL, have);
    if (b == 0) return;
    if (R.n + 32 > R.cap && !table_reserve(R, 32, nk, pool, lane)) return;
    if (have) R.tab[R.n + __popc(b & ((1u << lane) - 1u))] = make_uint2(cid, cnt);
    R.n += __popc(b);
    __syncwarp();
}

/* ascending sort by color-set id of tab[0, n): bitonic network over the next power of two */
__device__ __noinline__ void table_sort(uint2* tab, uint32_t n, uint32_t lane) {
    const uint32_t P = next_pow2(n);
    for (uint32_t i = n + lane; i < P; i += 32) tab[i] = make_uint2(FG_NOT_FOUND, 0);
    __syncwarp();
    for (uint32_t kk = 2; kk <= P; kk <<= 1) {
        for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            for (uint32_t t = lane; t < P / 2; t += 32) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)); /* index with bit j clear */
                const uint32_t l = i | j;
                const uint2 a = tab[i], b = tab[l];
                const bool up = (i & kk) == 0;
                if ((a.x > b.x) == up) {
                    tab[i] = b;
                    tab[l] = a;
                }
            }
            __syncwarp();
        }
    }
}

/* sorts the list by color-set id and merges the entries with equal ids (their counts add up): n becomes the number of
   distinct ids. The head of every run of equal ids sums the run, then the heads are packed to the front in order. */
__device__ __noinline__ void table_compact(read_hits& R, uint32_t lane) {
    uint2* tab = R.tab;
    const uint32_t n = R.n;
    table_sort(tab, n, lane);
    uint32_t out_n = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint2 e = make_uint2(FG_NOT_FOUND, 0);
        bool head = false;
        if (i < n) {
            e = tab[i];
            head = i == 0 || tab[i - 1].x != e.x;
            if (head)
                for (uint32_t j = i + 1; j < n && tab[j].x == e.x; ++j) e.y += tab[j].y;
        }
        const uint32_t b = __ballot_sync(FG_FULL, head);
        __syncwarp(); /* all reads of this block of entries are done before the packed entries overwrite it */
        if (head) tab[out_n + __popc(b & ((1u << lane) - 1u))] = e;
        out_n += __popc(b);
        __syncwarp();
    }
    R.n = out_n;
}

__device__ __noinline__ bool table_reserve(read_hits& R, uint32_t extra, uint32_t nk, const entry_pool& pool, uint32_t lane) {
    if (R.failed) return false;
    table_compact(R, lane);
    if (R.n + extra <= R.cap) return true;
    /* grow into the pool: nk bounds the number of distinct color sets of this read, so a compacted list never needs more */
    const uint32_t want = next_pow2(nk + extra);
    unsigned long long off = 0;
    if (lane == 0) off = atomicAdd(pool.used, (unsigned long long)want);
    off = __shfl_sync(FG_FULL, off, 0);
    if (want <= R.cap || off + want > pool.cap) {
        if (lane == 0) *pool.exhausted = 1;
        R.failed = true;
        return false;
    }
    uint2* nt = pool.base + off;
    for (uint32_t i = lane; i < R.n; i += 32) nt[i] = R.tab[i];
    __syncwarp();
    R.tab = nt;
    R.cap = want;
    return true;
}

/* The per-read table beyond the registers, first stage: an open-addressing HASH TABLE of {color-set id, multiplicity} over the
   warp's scratch entries (shared memory, a power of two). A batch of items costs a few probes per lane (atomicCAS on the key,
   atomicAdd on the count) whatever the number of distinct ids, and nothing is sorted on the way. */
#define FG_HASH_EMPTY FG_NOT_FOUND

__device__ __forceinline__ void hash_clear(uint2* tab, uint32_t cap, uint32_t lane) {
    for (uint32_t i = lane; i < cap; i += 32) tab[i] = make_uint2(FG_HASH_EMPTY, 0);
    __syncwarp();
}

/* inserts {cid, cnt} of every lane with have = true; returns how many NEW ids the batch brought (warp-uniform) */
__device__ __noinline__ uint32_t hash_insert(uint2* tab, uint32_t cap, bool have, uint32_t cid, uint32_t cnt) {
    uint32_t slot = (cid * 0x9E3779B1u) >> 7 & (cap - 1);
    uint32_t fresh = 0;
    bool pending = have;
    while (__ballot_sync(FG_FULL, pending) != 0) {
        if (pending) {
            const uint32_t old = atomicCAS(&tab[slot].x, FG_HASH_EMPTY, cid);
            if (old == FG_HASH_EMPTY || old == cid) {
                atomicAdd(&tab[slot].y, cnt);
                fresh += old == FG_HASH_EMPTY;
                pending = false;
            } else {
                slot = (slot + 1) & (cap - 1);
            }
        }
    }
    __syncwarp();
    return __reduce_add_sync(FG_FULL, fresh);
}

/* packs the occupied slots of the hash table (cap <= 128 entries, at most 4 per lane) into dst[0, n), in slot order; dst may be
   the table itself. Returns n. */
__device__ __noinline__ uint32_t hash_pack(const uint2* tab, uint32_t cap, uint2* dst, uint32_t lane) {
    uint2 e[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) e[g] = uint32_t(32 * g) + lane < cap ? tab[32 * g + lane] : make_uint2(FG_HASH_EMPTY, 0);
    __syncwarp(); /* every slot has been read before dst (possibly the table) is written */
    uint32_t n = 0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const bool have = e[g].x != FG_HASH_EMPTY;
        const uint32_t b = __ballot_sync(FG_FULL, have);
        if (have) dst[n + __popc(b & ((1u << lane) - 1u))] = e[g];
        n += __popc(b);
    }
    __syncwarp();
    return n;
}

/* index::fetch_color_set_ids (src/ps_full_intersection.cpp:335-374) and the counting half of
   index::pseudoalign_threshold_union (src/ps_threshold_union.cpp:327-387) for one read: its DISTINCT color-set ids with the
   number of positive k-mers of each. The reference's two sort+unique passes become three stages, by how many distinct sets
   the read turns out to have:
     1. a 32-entry register table (one entry per lane), fed by a warp match on the ids of each batch of items;
     2. from the first batch with >= FG_DIVERSE_BATCH different ids (or when the registers are full): the hash table above in
        `scratch` (shared memory, scratch_cap entries, a power of two in [64, 128]);
     3. beyond 3/4 of its slots (long reads): an append-only list in the pool, sorted and merged by table_compact() when it
        fills up and at the end.
   SORTED: the list comes out ascending by id, the order index::fetch_color_set_ids returns and the one the deduplication
   compares lists in. The color-set kernels take the ids in any order (an intersection and a sum of multiplicities do not
   care), so the pseudoalignment path skips every sort. */
template <bool SORTED, class TILES>
__device__ __forceinline__ read_hits warp_fetch_color_sets(TILES& tiles, uint32_t lane, uint2* scratch, uint32_t scratch_cap, const entry_pool& pool) {
    read_hits R;
    R.cid = FG_NOT_FOUND;
    R.cnt = 0;
    R.n = 0;
    R.npos = 0;
    R.tab = nullptr;
    R.cap = 0;
    R.failed = false;
    bool hashed = false;   /* stage 2: scratch is the hash table, R.n = ids in it */
    /* stage 1 -> 2: the register entries move into the (cleared) hash table */
    auto to_hash = [&]() {
        hash_clear(scratch, scratch_cap, lane);
        hash_insert(scratch, scratch_cap, lane < R.n, R.cid, R.cnt);
        hashed = true;
    };
    /* stage 2 -> 3: the table's entries become the head of a list in the pool (table_reserve sizes it for the whole read) */
    auto to_list = [&]() {
        hashed = false;
        R.n = hash_pack(scratch, scratch_cap, scratch, lane);
        R.tab = scratch;
        R.cap = scratch_cap;
        table_reserve(R, scratch_cap, tiles.nk, pool, lane); /* sorts the (distinct) entries, then moves them: no room for a batch here */
    };
    uint32_t cid, cnt;
    while (tiles.next(cid, cnt)) {
        const bool found = cnt != 0;
        R.npos += __reduce_add_sync(FG_FULL, cnt);
        if (R.tab != nullptr) { /* stage 3: append, merge later */
            table_append(R, found, cid, cnt, tiles.nk, pool, lane);
            continue;
        }
        if (hashed) {
            R.n += hash_insert(scratch, scratch_cap, found, cid, cnt);
            if (4 * R.n > 3 * scratch_cap) to_list();
            continue;
        }
        const uint32_t grp = __match_any_sync(FG_FULL, cid);
        const bool leader = found && (uint32_t(__ffs(int(grp))) - 1 == lane);
        uint32_t leaders = __ballot_sync(FG_FULL, leader);
        if (__popc(leaders) >= FG_DIVERSE_BATCH) { /* a batch with many different ids: the register table would be walked once per
                                                     id and overflow soon anyway */
            to_hash();
            R.n += hash_insert(scratch, scratch_cap, found, cid, cnt);
            if (4 * R.n > 3 * scratch_cap) to_list();
            continue;
        }
        while (leaders) {
            const int src = __ffs(int(leaders)) - 1;
            leaders &= leaders - 1;
            const uint32_t kk = __shfl_sync(FG_FULL, cid, src);
            const uint32_t cc = __reduce_add_sync(FG_FULL, cid == kk ? cnt : 0u);
            if (!hashed) {
                const uint32_t hit = __ballot_sync(FG_FULL, R.cid == kk);
                if (hit) {
                    if (R.cid == kk) R.cnt += cc;
                    continue;
                }
                if (R.n < FG_MAX_ENTRIES) {
                    if (lane == R.n) {
                        R.cid = kk;
                        R.cnt = cc;
                    }
                    R.n += 1;
                    continue;
                }
                to_hash(); /* registers are full */
            }
            R.n += hash_insert(scratch, scratch_cap, lane == 0, kk, cc); /* the rest of this batch: one merged entry per id */
        }
    }
    if (hashed) { /* out of the table: into the registers when few, else a list in scratch */
        R.n = hash_pack(scratch, scratch_cap, scratch, lane);
        if (R.n <= FG_MAX_ENTRIES) {
            const uint2 e = lane < R.n ? scratch[lane] : make_uint2(FG_NOT_FOUND, 0);
            R.cid = e.x;
            R.cnt = e.y;
        } else {
            R.tab = scratch;
            R.cap = scratch_cap;
            if (SORTED) table_sort(R.tab, R.n, lane);
            return R;
        }
    } else if (R.tab != nullptr) {
        if (!R.failed) {
            table_compact(R, lane);
            if (R.n <= FG_MAX_ENTRIES) { /* few distinct ids after all: back to one entry per lane (already sorted) */
                const uint2 e = lane < R.n ? R.tab[lane] : make_uint2(FG_NOT_FOUND, 0);
                R.cid = e.x;
                R.cnt = e.y;
                R.tab = nullptr;
                R.cap = 0;
            }
        }
        return R;
    }
    if (SORTED && R.n > 1) { /* bitonic sort across lanes; unused lanes hold FG_NOT_FOUND and sink to the end */
#pragma unroll
        for (uint32_t kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
                const uint32_t ok = __shfl_xor_sync(FG_FULL, R.cid, j);
                const uint32_t oc = __shfl_xor_sync(FG_FULL, R.cnt, j);
                const bool up = (lane & kk) == 0;
                const bool lower = (lane & j) == 0;
                const bool take = (up == lower) ? (ok < R.cid) : (ok > R.cid);
                if (take) {
                    R.cid = ok;
                    R.cnt = oc;
                }
            }
        }
    }
    return R;
}

/* ------------------------------------------------------------------ color-set decoding */

/* 64 bits of an LSB-first bit stream starting at bit `pos` (bits/bit_vector.hpp:185-192) */
FG_HD uint64_t bits_at(const uint64_t* __restrict__ words, uint64_t pos) {
    const uint64_t* w = words + (pos >> 6);
    const uint32_t sh = uint32_t(pos & 63);
    const uint64_t a = FG_LDG(w);
    if (sh == 0) return a;
    return (a >> sh) | (FG_LDG(w + 1) << (64 - sh));
}

/* Elias delta (bits/integer_codes.hpp:54-71 over bit_vector::iterator, bit_vector.hpp:234-294):
   gamma(b) then b payload bits; values here are < 2^32 so one 64-bit window always holds a code */
FG_HD uint32_t read_delta(const uint64_t* __restrict__ words, uint64_t& pos) {
    const uint64_t w = bits_at(words, pos);
    const uint32_t u = fg_ffs64(w) - 1u;                /* unary: u zeros, then a one */
    const uint32_t b = uint32_t(((w >> (u + 1)) & ((1ULL << u) - 1)) | (1ULL << u)) - 1u; /* gamma */
    const uint64_t payload = (w >> (2 * u + 1)) & ((1ULL << b) - 1);
    pos += 2 * u + 1 + b;
    return uint32_t((payload | (1ULL << b)) - 1);
}

/* One hybrid color set (include/color_sets/hybrid.hpp:162-188 rewind, :37-95 layout) of a container
   with at most 32 colors, as a bit mask. */
FG_HD uint32_t hybrid_set_mask(const dev_index& I, uint32_t container, uint64_t local_id) {
    const fgi_hybrid* h = I.hybrids + container;
    const uint32_t C = FG_LDG(&h->num_colors);
    const uint32_t cmask = C >= 32 ? ~0u : ((1u << C) - 1u);
    const uint64_t* words = I.color_words + FG_LDG(&h->word_base);
    uint64_t pos = FG_LDG(I.set_bit_off + FG_LDG(&h->set_off_base) + local_id);
    const uint32_t size = read_delta(words, pos);
    if (size >= FG_LDG(&h->sparse_thr) && size < FG_LDG(&h->very_dense_thr)) { /* raw bitmap, unaligned */
        return uint32_t(bits_at(words, pos)) & cmask;
    }
    const bool complement = size >= FG_LDG(&h->very_dense_thr);
    const uint32_t n = complement ? C - size : size;
    uint32_t mask = 0, v = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t d = read_delta(words, pos);
        v = i ? v + d + 1 : d;
        mask |= 1u << (v & 31);
    }
    return complement ? (~mask & cmask) : mask;
}

/* One differential color set (include/color_sets/differential.hpp:256-287) of a container with at most 32 colors: the
   symmetric difference of its difference list and its cluster's representative (both delta-coded gap lists; the
   difference list carries a second header, the size of the decoded set). */
FG_HD uint32_t differential_set_mask(const dev_index& I, uint32_t container, uint64_t local_id) {
    const fgi_hybrid* h = I.hybrids + container;
    const uint64_t* words = I.color_words + FG_LDG(&h->word_base);
    const uint64_t base = FG_LDG(&h->set_off_base);
    uint32_t mask = 0;
    for (uint32_t part = 0; part < 2; ++part) {
        uint64_t pos = FG_LDG(I.set_bit_off + base + (part ? FG_LDG(&h->num_sets) + 1 : 0) + local_id);
        const uint32_t n = read_delta(words, pos);
        if (part == 0) read_delta(words, pos);
        uint32_t v = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t d = read_delta(words, pos);
            v = i ? v + d + 1 : d;
            mask ^= 1u << (v & 31);
        }
    }
    return mask;
}

FG_HD uint32_t partial_set_mask(const dev_index& I, uint32_t container, uint64_t local_id) {
    return I.diff ? differential_set_mask(I, container, local_id) : hybrid_set_mask(I, container, local_id);
}

/* color set `cid` of the index as a mask (num_colors <= 32). Meta (include/color_sets/meta.hpp:93-236):
   the set is the concatenation of its partial sets, each shifted by its partition's min_color. */
FG_HD uint32_t color_set_mask(const dev_index& I, uint32_t cid) {
    if (I.type == 0) return partial_set_mask(I, 0, cid);
    const uint64_t b = FG_LDG(I.meta_off + cid);
    const uint32_t n = FG_LDG(I.meta_vals + b);
    uint32_t mask = 0, p = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t mc = FG_LDG(I.meta_vals + b + 1 + i);
        while (p + 1 < I.num_partitions && mc >= FG_LDG(I.part_sets_before + p + 1)) ++p; /* meta.hpp:227-235 */
        mask |= partial_set_mask(I, p, mc - FG_LDG(I.part_sets_before + p)) << FG_LDG(I.part_min_color + p);
    }
    return mask;
}


/* ------------------------------------------------------------------ color sets of any width */

enum { FG_ENC_DELTA = 0, FG_ENC_BITMAP = 1, FG_ENC_COMPLEMENT = 2, FG_ENC_NONE = 3 };

/* Sequential reader of an LSB-first bit stream (bits/bit_vector.hpp:234-294) that keeps 33..64 bits in a register and
   refills 32 bits at a time, so a run of delta codes costs one 32-bit load per 32 bits instead of two 64-bit loads per code. */
struct bit_cursor {
    const uint32_t* next; /* next 32-bit word to load */
    uint64_t buf;         /* the stream's next `avail` bits, LSB first */
    uint32_t avail;

    __device__ __forceinline__ void open(const uint64_t* words, uint64_t pos) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(words) + (pos >> 5);
        const uint32_t sh = uint32_t(pos & 31);
        buf = uint64_t(FG_LDG(w)) >> sh;
        avail = 32 - sh;
        next = w + 1;
        refill();
    }
    __device__ __forceinline__ void refill() {
        if (avail <= 32) {
            buf |= uint64_t(FG_LDG(next)) << avail;
            next += 1;
            avail += 32;
        }
    }
    __device__ __forceinline__ void skip(uint32_t n) {
        buf >>= n;
        avail -= n;
        refill();
    }
    /* Elias delta (bits/integer_codes.hpp:54-71), values < 2^32: gamma(b) (at most 11 bits) then b <= 32 payload bits */
    __device__ __forceinline__ uint32_t delta() {
        const uint32_t lo = uint32_t(buf);
        const uint32_t u = uint32_t(__ffs(int(lo))) - 1u; /* unary: u zeros, then a one */
        const uint32_t b = (((lo >> (u + 1)) & ((1u << u) - 1u)) | (1u << u)) - 1u;
        skip(2 * u + 1);
        const uint64_t payload = b >= 32 ? (buf & 0xffffffffULL) : (buf & ((1ULL << b) - 1));
        skip(b);
        return uint32_t((payload | (1ULL << b)) - 1);
    }
};

/* one partial color set (a hybrid set of one container) */
struct set_item {
    uint64_t pos;        /* bit position right after the size header */
    const uint64_t* words; /* the container's bit stream */
    uint32_t color_base; /* first global color of the container */
    uint32_t num_colors; /* colors in the container */
    uint32_t nvals;      /* delta-coded values that follow (the set, or its complement) */
    uint32_t enc;
};

/* hybrid::forward_iterator::rewind (include/color_sets/hybrid.hpp:162-188): size header -> encoding */
__device__ __forceinline__ set_item open_set(const dev_index& I, uint32_t container, uint64_t local_id, uint32_t color_base) {
    set_item it;
    const fgi_hybrid* h = I.hybrids + container;
    it.color_base = color_base;
    it.num_colors = FG_LDG(&h->num_colors);
    it.words = I.color_words + FG_LDG(&h->word_base);
    it.pos = FG_LDG(I.set_bit_off + FG_LDG(&h->set_off_base) + local_id);
    const uint32_t size = read_delta(it.words, it.pos);
    if (size < FG_LDG(&h->sparse_thr)) {
        it.enc = FG_ENC_DELTA;
        it.nvals = size;
    } else if (size < FG_LDG(&h->very_dense_thr)) {
        it.enc = FG_ENC_BITMAP;
        it.nvals = 0;
    } else {
        it.enc = FG_ENC_COMPLEMENT;
        it.nvals = it.num_colors - size;
    }
    return it;
}

/* Per-read, per-color counters kept BIT-SLICED in shared memory: plane j holds bit j of every color's counter, one bit per
   color, so a bitmap-coded color set is added to all of its colors with a few word-wide logic operations per 32 colors
   (ripple carry across the planes) instead of one read-modify-write per member, and a whole 4,546-color counter array of
   7-bit counters takes 4 KB instead of 18 KB. Updates are atomic XORs, so lanes may work on different sets at once: a lane
   that flips a bit from 1 to 0 owes the carry to the next plane, whatever other lanes do to the same word in between. */
struct bit_planes {
    uint32_t* base;    /* plane j, word w at base[j * stride + w] */
    uint32_t stride;   /* words per plane */
    uint32_t nplanes;

    /* counter[32 w + b] += weight for every set bit b of `bits` */
    __device__ __forceinline__ void add_word(uint32_t w, uint32_t bits, uint32_t weight) const {
        for (uint32_t j = 0; weight; ++j, weight >>= 1) {
            if (!(weight & 1u)) continue;
            uint32_t carry = bits;
            for (uint32_t k = j; carry && k < nplanes; ++k) carry &= atomicXor(base + k * stride + w, carry);
        }
    }
    __device__ __forceinline__ void add_color(uint32_t c, uint32_t weight) const { add_word(c >> 5, 1u << (c & 31), weight); }
    __device__ __forceinline__ uint32_t plane(uint32_t j, uint32_t w) const { return j < nplanes ? base[j * stride + w] : 0u; }
};

/* 32 bits of a bitmap-coded partial set, aligned to GLOBAL color word w: bit b = color 32 w + b. gmask = the bits of the
   word that belong to the container's color range [color_base, color_base + num_colors). */
__device__ __forceinline__ uint32_t bitmap_word(const set_item& it, uint32_t w, uint32_t& gmask) {
    const int lo = int(32 * w) - int(it.color_base); /* local index of the word's first color */
    const int first = lo < 0 ? -lo : 0;              /* first bit of the word inside the range */
    const int last = min(32, int(it.num_colors) - lo); /* one past the last */
    if (last <= first) {
        gmask = 0;
        return 0;
    }
    gmask = (last >= 32 ? ~0u : ((1u << last) - 1u)) & ~((1u << first) - 1u);
    const uint64_t raw = lo >= 0 ? bits_at(it.words, it.pos + uint64_t(lo)) : bits_at(it.words, it.pos) << first;
    return uint32_t(raw) & gmask;
}

}  // namespace fgb
#endif
