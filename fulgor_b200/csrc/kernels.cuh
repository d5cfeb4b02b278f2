/*
 * kernels.cuh -- sm_100a device code of the pseudoalignment hot path.
 *
 * Mapping: ONE WARP PER READ. Lanes own consecutive k-mer start positions (tiles of 32 k-mers);
 * every lane resolves its own k-mer independently -- canonical minimizer -> minimizer MPHF ->
 * bucket range -> super-k-mer record -> 2-bit string window compare. A lookup answer is a pure
 * function of the k-mer (the reference asserts this itself: external/sshash/include/
 * streaming_query.hpp:107 compares every streamed answer with a from-scratch lookup), so the
 * reference's sequential seed-and-extend state machine is NOT reproduced: neighbouring lanes that
 * share a minimizer issue identical addresses, which the LSU coalesces into one sector request,
 * and that is the device analogue of "extension".
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

/* The per-lane functions (hashing, MPHF, string compare, color-set decode) are FG_HD so that
   tests/host_emul.cu can run the SAME source on the host against the oracle when no GPU is around.
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
    uint32_t skew_threshold, pad1; /* buckets with more super-k-mers go through the skew index; UINT32_MAX when there is none */
    uint32_t skew_phf[FGI_MAX_SKEW];
    uint64_t skew_pos_base[FGI_MAX_SKEW];
    uint32_t type, num_colors, num_partitions, guard_max_hash;
    uint64_t main_seed, main_nparts; /* the minimizer MPHF (phfs[0]; its partitions are parts[0 .. main_nparts)) */
    fgi_phf_part main_part;          /* parts[0] */
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
    const uint64_t q = fg_mulhi64(a, inv);
    uint64_t r = a - q * d;
    if (r >= d) r -= d;
    return r;
}

/* single_phf::position (pthash/single_phf.hpp:79-101): skew_bucketer (utils/bucketers.hpp:163-168), pre-hashed pilot,
   xor displacement, minimal (free slots) */
FG_HD uint64_t phf_position(const dev_index& I, const fgi_phf_part& P, uint64_t first, uint64_t second) {
    uint64_t bucket;
    if (first < I.bucketer_T) {
        bucket = mod_by_inverse(first, P.inv_dense, P.num_dense);
    } else {
        bucket = P.num_dense + mod_by_inverse(first, P.inv_sparse, P.num_sparse);
    }
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
    const uint64_t bit = 2 * uint64_t(rec.x);
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

/* 32 characters (one per lane) -> one 64-bit word of 2-bit codes (char j at bits [2j, 2j+1]) + validity mask */
__device__ __forceinline__ void pack_chars(uint32_t c, bool in_range, uint32_t lane, uint64_t& word, uint32_t& valid) {
    const uint32_t code = (c >> 1) & 3u;
    const uint32_t sh = (lane & 15u) * 2u;
    const uint32_t lo = __reduce_or_sync(FG_FULL, lane < 16 ? code << sh : 0u);
    const uint32_t hi = __reduce_or_sync(FG_FULL, lane >= 16 ? code << sh : 0u);
    word = uint64_t(lo) | (uint64_t(hi) << 32);
    valid = __ballot_sync(FG_FULL, in_range && base_valid(c));
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
   The lanes own the k-mers of a segment in BLOCKS: lane l owns k-mers 4l .. 4l+3 ("tile" t = the t-th k-mer of every
   lane), so the sliding-window minimum and the run detection below are sequential inside a lane and need no exchange.
   Position p of the per-m-mer hash arrays is stored at slot p + p/4: lane l's 16 positions then start at slot 5l, an
   odd stride, which keeps the blocked 64-bit reads free of bank conflicts. */
#define FG_SEG_B 4                        /* k-mers per lane per segment */
#define FG_SEG_KMERS (32 * FG_SEG_B)      /* k-mers per segment */
#define FG_HASH_SLOTS 200                 /* >= slot(FG_SEG_KMERS + 30) + 1 */
#define FG_SEG_WORDS 6                    /* 64-bit words of packed bases: ceil((FG_SEG_KMERS + 30) / 32) + 1 pad */
struct warp_stage {
    uint64_t hf[FG_HASH_SLOTS];       /* mixer_64 hash of the forward m-mer starting at every position; afterwards the
                                         segment's seed list: {minimizer} -> {first super-k-mer, bucket size} */
    uint64_t hr[FG_HASH_SLOTS];       /* hash of its reverse complement */
    uint32_t info[FG_SEG_KMERS];      /* [t * 32 + lane]: seed slot | minimizer position << 8 | ambiguous << 13 | valid << 14 */
    uint64_t words[FG_SEG_WORDS];
    uint32_t valid[FG_SEG_WORDS];
};

FG_HD constexpr uint32_t fg_hslot(uint32_t p) { return p + (p >> 2); }

/* Minimum of every window of W consecutive hashes among the W + 3 that start at a lane's base slot, for the lane's four
   k-mers: h[j] for j in [t, t + W). One suffix scan over the first window (its suffix minima are the left parts of the
   other three windows) plus a running prefix over the three extra positions: W + 4 comparisons for four k-mers.
   Ties: util::compute_minimizer (sshash/util.hpp:220-239) keeps the FIRST minimum of the strand it scans. For the forward
   strand that is the leftmost position; the reverse strand's m-mers come in the opposite order, so there (REV) the
   rightmost position in forward coordinates wins. mh = hash, mp = position relative to the lane's base. */
template <int W, bool REV>
__device__ __forceinline__ void window_minima(const uint64_t* __restrict__ h, uint64_t (&mh)[FG_SEG_B], uint32_t (&mp)[FG_SEG_B]) {
    uint64_t sh[FG_SEG_B];
    uint32_t sp[FG_SEG_B];
    uint64_t ch = h[fg_hslot(W - 1)];
    uint32_t cp = W - 1;
#pragma unroll
    for (int j = W - 2; j >= 0; --j) {
        const uint64_t x = h[fg_hslot(j)];
        const bool take = REV ? (x < ch) : (x <= ch);
        ch = take ? x : ch;
        cp = take ? uint32_t(j) : cp;
        if (j < FG_SEG_B) {
            sh[j] = ch;
            sp[j] = cp;
        }
    }
    mh[0] = sh[0];
    mp[0] = sp[0];
    uint64_t ph = 0;
    uint32_t pp = 0;
#pragma unroll
    for (int t = 1; t < FG_SEG_B; ++t) {
        const uint64_t x = h[fg_hslot(W - 1 + t)];
        const bool take = t == 1 || (REV ? (x <= ph) : (x < ph));
        ph = take ? x : ph;
        pp = take ? uint32_t(W - 1 + t) : pp;
        const bool right = REV ? (ph <= sh[t]) : (ph < sh[t]);
        mh[t] = right ? ph : sh[t];
        mp[t] = right ? pp : sp[t];
    }
}

/* Walks one read: after next(), lane l holds the color-set id of one of its k-mers (FG_NOT_FOUND when negative, invalid or
   past the end); over the calls until done() every k-mer of the read is reported exactly once (in no particular order:
   the consumers -- intersection, per-color scores, distinct-set table -- are order-free).
   Replaces the per-k-mer loop around streaming_query::lookup_advanced (sshash/streaming_query.hpp:50-109, driven from
   src/ps_full_intersection.cpp:344-353): validity test (:53-59), 2-bit packing and reverse complement (:62-74),
   canonical minimizer (:76-83), dictionary lookup (:144-190).

   Per segment of FG_SEG_KMERS k-mers the warp
     0. packs the bases (2 bits each) and their validity bits into shared memory;
     1. hashes every m-mer ONCE, both strands (the k-mers of a read share almost all their m-mers, which the reference
        exploits with its sliding minimizer_enumerator, sshash/minimizer_enumerator.hpp:24-49);
     2. takes the sliding-window minima (window_minima) and from them every k-mer's canonical minimizer;
     3. cuts the k-mers into runs with the same minimizer and makes the first k-mer of each run a SEED: only seeds go
        through the minimizer MPHF and the bucket table, all seeds of the segment at once, one per lane. This is the
        device analogue of the reference's seed-and-extend: ~1 hash lookup per 7 k-mers instead of 1 per k-mer;
     4. (next) every k-mer compares itself with the super-k-mers of its run's bucket.
   A lookup answer is a pure function of the k-mer (streaming_query.hpp:107 asserts it), so sharing the bucket between
   the k-mers of a run changes no result. */
template <int W>
struct kmer_tiles {
    const dev_index& I;
    const uint8_t* __restrict__ seq;
    warp_stage& S;
    uint32_t len, lane, nk, seg_end, t;
    uint64_t kmask, mmer_mask;
    uint32_t window, kbits;

    __device__ __forceinline__ kmer_tiles(const dev_index& I_, const uint8_t* seq_, uint32_t len_, uint32_t lane_, warp_stage& S_)
        : I(I_), seq(seq_), S(S_), len(len_), lane(lane_) {
        const uint32_t k = I.k;
        nk = len >= k ? len - k + 1 : 0; /* src/ps_full_intersection.cpp:337: shorter reads have no k-mers */
        seg_end = 0;
        t = FG_SEG_B;
        kmask = (1ULL << (2 * k)) - 1;
        mmer_mask = (1ULL << (2 * I.m)) - 1;
        window = W ? W : k - I.m + 1;
        kbits = (1u << k) - 1u;
    }
    __device__ __forceinline__ bool done() const { return t == FG_SEG_B && seg_end >= nk; }

    /* 2*nbases bits starting at base position p (segment-relative) of the packed words */
    __device__ __forceinline__ uint64_t bases_at(uint32_t p) const {
        const uint32_t wi = p >> 5, sh = (p & 31) * 2;
        const uint64_t a = S.words[wi];
        return sh ? (a >> sh) | (S.words[wi + 1] << (64 - sh)) : a;
    }
    __device__ __forceinline__ uint2* seeds() const { return reinterpret_cast<uint2*>(S.hf); }

    /* m-mer behind a window minimum; compute_minimizer's "nothing below UINT64_MAX" sentinel when that can occur (image.h) */
    __device__ __forceinline__ uint64_t mmer_of_hash(uint64_t h) const {
        uint64_t y = (h ^ I.hash_magic) * FG_MIX_INV;
        if (I.guard_max_hash && h == UINT64_MAX) y = UINT64_MAX;
        return y;
    }

    __device__ __forceinline__ void load_segment() {
        __syncwarp();
        const uint32_t seg0 = seg_end;
        const uint32_t seg_nk = min(uint32_t(FG_SEG_KMERS), nk - seg0);
        seg_end = seg0 + seg_nk;
        const uint32_t nchars = seg_nk + I.k - 1, npos = seg_nk + window - 1;
        const uint32_t nwords = (nchars + 31) >> 5;
        /* 0. bases */
        for (uint32_t c = 0; c <= nwords; ++c) { /* one extra zero word so that bases_at may read words[wi + 1] */
            const uint32_t p = seg0 + 32 * c + lane;
            const bool in = c < nwords && 32 * c + lane < nchars;
            const uint32_t ch = in ? seq[p] : 0u;
            uint64_t w;
            uint32_t v;
            pack_chars(ch, in, lane, w, v);
            if (lane == 0) {
                S.words[c] = w;
                S.valid[c] = v;
            }
        }
        __syncwarp();
        /* 1. m-mer hashes */
        const uint32_t m = I.m;
        for (uint32_t q = lane; q < npos; q += 32) {
            const uint64_t y = bases_at(q) & mmer_mask;
            S.hf[fg_hslot(q)] = (y * FG_MIX_MUL) ^ I.hash_magic;
            S.hr[fg_hslot(q)] = (revcomp(y, m) * FG_MIX_MUL) ^ I.hash_magic;
        }
        __syncwarp();
        /* 2. canonical minimizer of the lane's k-mers i0 .. i0+3: value = min over both strands, compared as integers
              (sshash/streaming_query.hpp:76-79); cpos = where it starts inside the k-mer, forward coordinates */
        const uint32_t i0 = FG_SEG_B * lane;
        uint64_t val[FG_SEG_B];
        uint32_t inf[FG_SEG_B]; /* cpos << 8 | ambiguous << 13 | valid << 14 */
        {
            uint64_t hf[FG_SEG_B], hr[FG_SEG_B];
            uint32_t pf[FG_SEG_B], pr[FG_SEG_B];
            if (W) {
                window_minima<W ? W : 13, false>(S.hf + 5 * lane, hf, pf);
                window_minima<W ? W : 13, true>(S.hr + 5 * lane, hr, pr);
            } else { /* any other (k, m): plain scan of each window */
#pragma unroll
                for (int tt = 0; tt < FG_SEG_B; ++tt) {
                    hf[tt] = hr[tt] = UINT64_MAX;
                    pf[tt] = pr[tt] = tt;
                    for (uint32_t j = 0; j < window; ++j) {
                        const uint64_t a = S.hf[fg_hslot(i0 + tt + j)], b = S.hr[fg_hslot(i0 + tt + j)];
                        if (a < hf[tt]) hf[tt] = a, pf[tt] = tt + j;
                        if (b <= hr[tt]) hr[tt] = b, pr[tt] = tt + j;
                    }
                }
            }
            /* i0 is a multiple of 4, so the lane's four k-mers start in the same 32-character word */
            const uint32_t vw0 = S.valid[i0 >> 5], vw1 = S.valid[(i0 >> 5) + 1];
#pragma unroll
            for (int tt = 0; tt < FG_SEG_B; ++tt) {
                const uint32_t i = i0 + tt;
                const uint32_t bits = __funnelshift_r(vw0, vw1, i & 31);
                const bool valid = i < seg_nk && (bits & kbits) == kbits;
                const uint64_t vf = mmer_of_hash(hf[tt]), vr = mmer_of_hash(hr[tt]);
                const bool fwd_wins = vf <= vr;
                val[tt] = fwd_wins ? vf : vr;
                const uint32_t cpos = (fwd_wins ? pf[tt] : pr[tt]) - tt;
                inf[tt] = (cpos << 8) | (uint32_t(vf == vr) << 13) | (uint32_t(valid) << 14);
            }
        }
        __syncwarp(); /* every lane is done with the hash arrays: the seed list may overwrite them */
        /* 3. seeds: the first k-mer of every run of valid k-mers with the same minimizer */
        const uint32_t lt_mask = (1u << lane) - 1u;
        uint64_t prev_val = __shfl_up_sync(FG_FULL, val[FG_SEG_B - 1], 1);
        bool prev_valid = __shfl_up_sync(FG_FULL, inf[FG_SEG_B - 1], 1) >> 14;
        if (lane == 0) prev_valid = false;
        uint32_t nseeds = 0, cur_slot = 0xffu;
        bool has_leader = false;
        bool leader[FG_SEG_B];
#pragma unroll
        for (int tt = 0; tt < FG_SEG_B; ++tt) {
            const bool valid = inf[tt] >> 14;
            leader[tt] = valid && !(prev_valid && prev_val == val[tt]);
            prev_valid = valid;
            prev_val = val[tt];
            const uint32_t b = __ballot_sync(FG_FULL, leader[tt]);
            if (leader[tt]) {
                cur_slot = nseeds + __popc(b & lt_mask);
                seeds()[cur_slot] = make_uint2(uint32_t(val[tt]), uint32_t(val[tt] >> 32));
                has_leader = true;
            }
            inf[tt] |= cur_slot; /* 0xff for now when the run started in an earlier lane */
            nseeds += __popc(b);
        }
        { /* runs that continue from an earlier lane: the slot in effect at the end of the nearest lane that has a seed */
            const uint32_t hb = __ballot_sync(FG_FULL, has_leader) & lt_mask;
            const uint32_t inherited = __shfl_sync(FG_FULL, cur_slot, hb ? 31 - __clz(int(hb)) : 0);
#pragma unroll
            for (int tt = 0; tt < FG_SEG_B; ++tt) {
                if ((inf[tt] & 0xffu) == 0xffu) inf[tt] = (inf[tt] & ~0xffu) | (inherited & 0xffu);
                S.info[tt * 32 + lane] = inf[tt];
            }
        }
        __syncwarp();
        /* minimizers::lookup + buckets::locate_bucket for every seed, one per lane */
        for (uint32_t s = lane; s < nseeds; s += 32) {
            const uint2 v = seeds()[s];
            const uint64_t b = minimizer_bucket(I, uint64_t(v.x) | (uint64_t(v.y) << 32));
            const uint32_t begin = FG_LDG(I.bucket_begin + b), end = FG_LDG(I.bucket_begin + b + 1);
            seeds()[s] = make_uint2(begin, end - begin);
        }
        __syncwarp();
    }

    __device__ __forceinline__ uint32_t next() {
        if (t == FG_SEG_B) {
            load_segment();
            t = 0;
        }
        const uint32_t inf = S.info[t * 32 + lane];
        uint32_t cid = FG_NOT_FOUND;
        if (inf >> 14) {
            const uint64_t fwd = bases_at(FG_SEG_B * lane + t) & kmask;
            const uint64_t rc = revcomp(fwd, I.k);
            const uint2 bucket = seeds()[inf & 0xffu];
            minimizer_t mz;
            mz.value = 0;
            mz.cpos = (inf >> 8) & 31u;
            mz.ambiguous = (inf >> 13) & 1u;
            cid = lookup_in_bucket(I, bucket.x, bucket.y, fwd, rc, mz, kmask);
        }
        __syncwarp();
        t += 1;
        return cid;
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

/* slow path of the per-read table: more than 32 distinct color sets. Lanes stride over the list. */
__device__ __noinline__ void table_insert(read_hits& R, uint32_t kk, uint32_t cc, uint32_t nk, const entry_pool& pool, uint32_t lane) {
    uint32_t found = 0;
    for (uint32_t i0 = 0; i0 < R.n && !found; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool eq = i < R.n && R.tab[i].x == kk;
        found = __ballot_sync(FG_FULL, eq);
        if (eq) R.tab[i].y += cc;
    }
    if (!found) {
        if (R.n == R.cap) { /* grow into the pool: nk bounds the number of distinct color sets of this read */
            const uint32_t want = next_pow2(nk);
            unsigned long long off = 0;
            if (lane == 0) off = atomicAdd(pool.used, (unsigned long long)want);
            off = __shfl_sync(FG_FULL, off, 0);
            if (want <= R.cap || off + want > pool.cap) {
                if (lane == 0) *pool.exhausted = 1;
                R.failed = true;
                return;
            }
            uint2* nt = pool.base + off;
            for (uint32_t i = lane; i < R.n; i += 32) nt[i] = R.tab[i];
            R.tab = nt;
            R.cap = want;
        }
        if (lane == 0) R.tab[R.n] = make_uint2(kk, cc);
        R.n += 1;
    }
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

/* index::fetch_color_set_ids (src/ps_full_intersection.cpp:335-374) and the counting half of
   index::pseudoalign_threshold_union (src/ps_threshold_union.cpp:327-387) for one read.
   The reference's two sort+unique passes become: warp match on the color-set id inside a tile, a
   32-entry register table (one entry per lane) across tiles -- spilling to `scratch` (shared memory,
   scratch_cap entries, a power of two) and then to the pool for reads with more distinct color sets --
   and one bitonic sort at the end. */
template <int W>
__device__ __forceinline__ read_hits warp_fetch_color_sets(const dev_index& I, const uint8_t* __restrict__ seq, uint32_t len, uint32_t lane,
                                                           warp_stage& stage, uint2* scratch, uint32_t scratch_cap, const entry_pool& pool) {
    read_hits R;
    R.cid = FG_NOT_FOUND;
    R.cnt = 0;
    R.n = 0;
    R.npos = 0;
    R.tab = nullptr;
    R.cap = 0;
    R.failed = false;
    kmer_tiles<W> tiles(I, seq, len, lane, stage);
    while (!tiles.done()) {
        const uint32_t cid = tiles.next();
        const bool found = cid != FG_NOT_FOUND;
        const uint32_t found_mask = __ballot_sync(FG_FULL, found);
        R.npos += __popc(found_mask);
        if (!found_mask) continue;
        const uint32_t grp = __match_any_sync(FG_FULL, cid);
        const bool leader = found && (uint32_t(__ffs(int(grp))) - 1 == lane);
        const uint32_t gcnt = __popc(grp);
        uint32_t leaders = __ballot_sync(FG_FULL, leader);
        while (leaders) {
            const int src = __ffs(int(leaders)) - 1;
            leaders &= leaders - 1;
            const uint32_t kk = __shfl_sync(FG_FULL, cid, src);
            const uint32_t cc = __shfl_sync(FG_FULL, gcnt, src);
            if (R.tab == nullptr) {
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
                scratch[lane] = make_uint2(R.cid, R.cnt); /* registers are full: move to the list */
                R.tab = scratch;
                R.cap = scratch_cap;
                __syncwarp();
            }
            if (!R.failed) table_insert(R, kk, cc, tiles.nk, pool, lane);
        }
    }
    if (R.tab != nullptr) {
        if (!R.failed) table_sort(R.tab, R.n, lane);
    } else if (R.n > 1) { /* bitonic sort across lanes; unused lanes hold FG_NOT_FOUND and sink to the end */
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

/* color set `cid` of the index as a mask (num_colors <= 32). Meta (include/color_sets/meta.hpp:93-236):
   the set is the concatenation of its partial sets, each shifted by its partition's min_color. */
FG_HD uint32_t color_set_mask(const dev_index& I, uint32_t cid) {
    if (I.type == 0) return hybrid_set_mask(I, 0, cid);
    const uint64_t b = FG_LDG(I.meta_off + cid);
    const uint32_t n = FG_LDG(I.meta_vals + b);
    uint32_t mask = 0, p = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t mc = FG_LDG(I.meta_vals + b + 1 + i);
        while (p + 1 < I.num_partitions && mc >= FG_LDG(I.part_sets_before + p + 1)) ++p; /* meta.hpp:227-235 */
        mask |= hybrid_set_mask(I, p, mc - FG_LDG(I.part_sets_before + p)) << FG_LDG(I.part_min_color + p);
    }
    return mask;
}


/* ------------------------------------------------------------------ color sets of any width */

enum { FG_ENC_DELTA = 0, FG_ENC_BITMAP = 1, FG_ENC_COMPLEMENT = 2, FG_ENC_NONE = 3 };

/* one partial color set to be applied to the per-read score array */
struct set_item {
    uint64_t pos;        /* bit position right after the size header */
    uint32_t container;  /* hybrid container (partition) */
    uint32_t color_base; /* first global color of the container */
    uint32_t num_colors; /* colors in the container */
    uint32_t nvals;      /* delta-coded values that follow (the set, or its complement) */
    uint32_t weight;
    uint32_t enc;
};

/* hybrid::forward_iterator::rewind (include/color_sets/hybrid.hpp:162-188): size header -> encoding */
__device__ __forceinline__ set_item open_set(const dev_index& I, uint32_t container, uint64_t local_id, uint32_t color_base, uint32_t weight) {
    set_item it;
    const fgi_hybrid* h = I.hybrids + container;
    it.container = container;
    it.color_base = color_base;
    it.num_colors = FG_LDG(&h->num_colors);
    it.weight = weight;
    const uint64_t* words = I.color_words + FG_LDG(&h->word_base);
    it.pos = FG_LDG(I.set_bit_off + FG_LDG(&h->set_off_base) + local_id);
    const uint32_t size = read_delta(words, it.pos);
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

/* Adds one round of up to 32 partial sets (one per lane, `valid` lanes only) to the warp's score array:
     scores[c] += weight for every color of a delta-coded or bitmap set,
     base[container] += weight and scores[c] -= weight for every MISSING color of a complement-coded set
   (the same trick as the reference's merge, src/ps_threshold_union.cpp:23-29, so the cost of a set is
   proportional to its encoded length). Bitmap sets are expanded by the whole warp (one 32-bit chunk
   per lane, no conflicts); delta-coded sets are decoded lane-parallel with shared-memory atomics. */
__device__ __forceinline__ void apply_sets(const dev_index& I, bool valid, const set_item& it, int* scores, int* base, uint32_t lane) {
    uint32_t bm = __ballot_sync(FG_FULL, valid && it.enc == FG_ENC_BITMAP);
    while (bm) {
        const int src = __ffs(int(bm)) - 1;
        bm &= bm - 1;
        const uint64_t pos = __shfl_sync(FG_FULL, it.pos, src);
        const uint32_t container = __shfl_sync(FG_FULL, it.container, src);
        const uint32_t cb = __shfl_sync(FG_FULL, it.color_base, src);
        const uint32_t nc = __shfl_sync(FG_FULL, it.num_colors, src);
        const int w = int(__shfl_sync(FG_FULL, it.weight, src));
        const uint64_t* words = I.color_words + FG_LDG(&I.hybrids[container].word_base);
        for (uint32_t c0 = lane * 32; c0 < nc; c0 += 32 * 32) {
            uint32_t bits = uint32_t(bits_at(words, pos + c0));
            if (nc - c0 < 32) bits &= (1u << (nc - c0)) - 1u;
            while (bits) {
                const uint32_t b = uint32_t(__ffs(int(bits))) - 1u;
                bits &= bits - 1;
                scores[cb + c0 + b] += w;
            }
        }
        __syncwarp();
    }
    if (valid && it.enc != FG_ENC_BITMAP) {
        const uint64_t* words = I.color_words + FG_LDG(&I.hybrids[it.container].word_base);
        const int w = it.enc == FG_ENC_COMPLEMENT ? -int(it.weight) : int(it.weight);
        if (it.enc == FG_ENC_COMPLEMENT) atomicAdd(base + it.container, int(it.weight));
        uint64_t pos = it.pos;
        uint32_t v = 0;
        for (uint32_t i = 0; i < it.nvals; ++i) {
            const uint32_t d = read_delta(words, pos);
            v = i ? v + d + 1 : d;
            atomicAdd(scores + it.color_base + v, w);
        }
    }
    __syncwarp();
}

}  // namespace fgb
#endif
