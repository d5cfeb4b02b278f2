/*
 * pipeline_kernels.cuh -- the __global__ entry points of the pseudoalignment pipeline (sm_100a):
 *     k_pseudoalign_small   K1 + fused K2 for indexes with at most 32 colors
 *     k_fetch_color_sets    K1 alone: per read, distinct color-set ids with multiplicities
 *     k_color_sets_general  K2 for any number of colors (hybrid and meta)
 *     k_scan_*, k_emit_*    CSR offsets and ascending color lists
 * engine.cu launches them; tests/simt_emul.cpp compiles the same source for the host (lock-step warp emulation)
 * so that the kernels' logic is checked against the oracle in the CPU-only test tier.
 */
#ifndef FULGOR_B200_PIPELINE_KERNELS_CUH
#define FULGOR_B200_PIPELINE_KERNELS_CUH

#include <type_traits>

#include "../../include/fulgor_gpu.h"
#include "kernels.cuh"

namespace fgb {

#define FG_WARPS_PER_BLOCK 8
#define FG_BLOCK (FG_WARPS_PER_BLOCK * 32)
#ifndef FG_MIN_BLOCKS
#define FG_MIN_BLOCKS 4 /* resident blocks per SM the lookup kernels are compiled for (register budget 65536 / (4 * 256) = 64) */
#endif
#define FG_STAGE_STRIDE FG_MAX_ENTRIES

/* The reads of one launch, in one of two forms. ASCII: the caller's characters as they are (include/fulgor_gpu.h: bases +
   read_off). PACKED: 2-bit codes, every read starting on a 32-bit word, lengths instead of offsets (word_off is their prefix
   sum, computed on the device), invalid characters listed apart. The lookup kernels are templated on the form. */
struct ascii_reads {
    const uint8_t* bases;
    const uint64_t* read_off; /* n + 1 */
    uint64_t read_off_base;   /* bases[0] is character read_off_base of the caller's buffer */
    static constexpr bool packed = false;
    __device__ __forceinline__ uint32_t length(uint32_t r) const { return uint32_t(__ldg(read_off + r + 1) - __ldg(read_off + r)); }
};
struct packed_reads {
    const uint32_t* words;
    const uint64_t* word_off; /* n + 1, first word of every read (chunk-local) */
    const uint32_t* read_len; /* n: bases | FG_READ_FLAGGED */
    const uint64_t* invalid;  /* ascending base positions (16 * word index + base in word, caller's coordinates) */
    uint32_t n_invalid;
    uint64_t pos_base;        /* position of words[0] in those coordinates */
    static constexpr bool packed = true;
    __device__ __forceinline__ uint32_t length(uint32_t r) const { return __ldg(read_len + r) & 0x7fffffffu; }
};
#define FG_READ_FLAGGED 0x80000000u

template <int W, bool PERK>
__device__ __forceinline__ kmer_tiles<W, PERK, false> make_tiles(const dev_index& I, const ascii_reads& in, uint32_t r, uint32_t n_reads, uint32_t lane, warp_stage& S) {
    const uint64_t beg = __ldg(in.read_off + r), end = __ldg(in.read_off + r + 1);
    return kmer_tiles<W, PERK, false>(I, in.bases + (beg - in.read_off_base), uint32_t(end - beg), in.bases,
                                      in.bases + (__ldg(in.read_off + n_reads) - in.read_off_base), lane, S);
}
template <int W, bool PERK>
__device__ __forceinline__ kmer_tiles<W, PERK, true> make_tiles(const dev_index& I, const packed_reads& in, uint32_t r, uint32_t n_reads, uint32_t lane, warp_stage& S) {
    const uint32_t l = __ldg(in.read_len + r);
    const uint64_t w = __ldg(in.word_off + r);
    packed_read pr;
    pr.words = in.words + w;
    pr.len = l & 0x7fffffffu;
    pr.flagged = (l & FG_READ_FLAGGED) != 0;
    pr.invalid = in.invalid;
    pr.n_invalid = in.n_invalid;
    pr.pos = in.pos_base + 16 * w;
    return kmer_tiles<W, PERK, true>(I, pr, lane, S);
}

/* K1 + fused K2 for indexes with at most 32 colors: each read's result is one 32-bit color mask,
   accumulated item by item (no per-read table: AND is idempotent and scores are sums over k-mers).
   Full intersection (src/ps_full_intersection.cpp:377-400 -> intersect :33-127): AND of the hit sets.
   Threshold union (src/ps_threshold_union.cpp:389 + merge :17-40 / merge_meta :43-120): color c is
   reported iff sum over positive k-mers of [c in set(k-mer)] >= uint64(double(npos) * threshold). */
template <int W, class READS>
__global__ void __launch_bounds__(FG_BLOCK, FG_MIN_BLOCKS) k_pseudoalign_small(const __grid_constant__ dev_index I, const READS in,
                                                               uint32_t n_reads, int algo, double threshold,
                                                               uint32_t* __restrict__ masks) {
    __shared__ warp_stage stage[FG_WARPS_PER_BLOCK];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * FG_WARPS_PER_BLOCK;
    for (uint32_t r = blockIdx.x * FG_WARPS_PER_BLOCK + (threadIdx.x >> 5); r < n_reads; r += warps) {
        auto tiles = make_tiles<W, false>(I, in, r, n_reads, lane, stage[threadIdx.x >> 5]);
        uint32_t acc = ~0u, score = 0, npos = 0;
        uint32_t cid, cnt;
        while (tiles.next(cid, cnt)) { /* items {color-set id, number of k-mers}: every lane decodes its own set */
            const bool found = cnt != 0;
            const uint32_t mask = !found ? ~0u : (I.set_table ? __ldg(I.set_table + uint64_t(cid) * I.table_stride) : color_set_mask(I, cid));
            __syncwarp();
            npos += __reduce_add_sync(FG_FULL, cnt);
            if (algo == FULGOR_GPU_FULL_INTERSECTION) {
                acc &= __reduce_and_sync(FG_FULL, mask);
            } else {
                uint32_t todo = __ballot_sync(FG_FULL, found);
                if (2 * __popc(todo) > I.num_colors) { /* more items than colors: one warp reduction per color (lane c keeps color c) */
                    const uint32_t w = found ? cnt : 0u;
                    for (uint32_t c = 0; c < I.num_colors; ++c) {
                        const uint32_t sc = __reduce_add_sync(FG_FULL, ((mask >> c) & 1u) ? w : 0u);
                        if (lane == c) score += sc;
                    }
                } else { /* few items: broadcast each one, lane c tests its color */
                    while (todo) {
                        const int src = __ffs(int(todo)) - 1;
                        todo &= todo - 1;
                        const uint32_t mj = __shfl_sync(FG_FULL, mask, src);
                        const uint32_t wj = __shfl_sync(FG_FULL, cnt, src);
                        score += ((mj >> lane) & 1u) ? wj : 0u;
                    }
                }
            }
        }
        uint32_t res = 0;
        if (npos) {
            if (algo == FULGOR_GPU_FULL_INTERSECTION) {
                res = acc;
            } else {
                const uint64_t min_score = uint64_t(double(npos) * threshold);
                res = __ballot_sync(FG_FULL, lane < I.num_colors && uint64_t(score) >= min_score);
            }
        }
        if (lane == 0) masks[r] = res;
    }
}

#define FG_SCRATCH_ENTRIES 128 /* per-warp shared-memory list for reads with more than 32 distinct color sets */

/* where the sorted {color-set id, multiplicity} list of read r lives: counts[r] <= 32 -> stage[r*32 ..];
   otherwise in the pool at the 64-bit entry offset stored in stage[r*32] */
__device__ __forceinline__ const uint2* entries_of(uint32_t r, uint32_t n, const uint2* __restrict__ stage, const uint2* __restrict__ pool) {
    const uint2* s = stage + uint64_t(r) * FG_STAGE_STRIDE;
    if (n <= FG_STAGE_STRIDE) return s;
    const uint2 o = s[0];
    return pool + (uint64_t(o.x) | (uint64_t(o.y) << 32));
}

/* K1 alone: per read, distinct color-set ids with multiplicities -- ascending when SORTED (index::fetch_color_set_ids' order, which
   the deduplication also relies on), in no particular order otherwise (all the color-set kernels need) */
template <int W, class READS, bool SORTED>
__global__ void __launch_bounds__(FG_BLOCK, FG_MIN_BLOCKS) k_fetch_color_sets(const __grid_constant__ dev_index I, const READS in,
                                                              uint32_t n_reads, uint2* __restrict__ stage, uint32_t* __restrict__ counts,
                                                              uint32_t* __restrict__ num_positive /* nullable */, entry_pool pool,
                                                              uint32_t* __restrict__ max_positive /* nullable: running maximum of num_positive */) {
    __shared__ uint2 scratch[FG_WARPS_PER_BLOCK][FG_SCRATCH_ENTRIES];
    __shared__ warp_stage wstage[FG_WARPS_PER_BLOCK];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warps = gridDim.x * FG_WARPS_PER_BLOCK;
    for (uint32_t r = blockIdx.x * FG_WARPS_PER_BLOCK + wib; r < n_reads; r += warps) {
        auto tiles = make_tiles<W, false>(I, in, r, n_reads, lane, wstage[wib]);
        read_hits R = warp_fetch_color_sets<SORTED>(tiles, lane, scratch[wib], FG_SCRATCH_ENTRIES, pool);
        uint2* s = stage + uint64_t(r) * FG_STAGE_STRIDE;
        if (R.tab == nullptr) {
            if (lane < R.n) s[lane] = make_uint2(R.cid, R.cnt);
        } else if (!R.failed) {
            unsigned long long off = 0;
            if (R.tab == scratch[wib]) { /* the list must outlive this warp's shared memory */
                if (lane == 0) off = atomicAdd(pool.used, (unsigned long long)R.n);
                off = __shfl_sync(FG_FULL, off, 0);
                if (off + R.n > pool.cap) {
                    if (lane == 0) *pool.exhausted = 1;
                    R.failed = true;
                } else {
                    for (uint32_t i = lane; i < R.n; i += 32) pool.base[off + i] = R.tab[i];
                }
            } else {
                off = (unsigned long long)(R.tab - pool.base);
            }
            if (lane == 0) s[0] = make_uint2(uint32_t(off), uint32_t(off >> 32));
        }
        if (lane == 0) {
            counts[r] = R.failed ? 0u : R.n;
            if (num_positive) num_positive[r] = R.npos;
            if (max_positive && R.npos > *max_positive) atomicMax(max_positive, R.npos);
        }
        __syncwarp();
    }
}

/* ---- K2 for indexes with more than 32 colors ---- */

#define FG_K2_SETS_PER_ROUND 256 /* color sets of one read handled per round (unit prefix + deferred list live in shared memory) */
#define FG_K2_WIDE_COLORS 96     /* bitmap-coded partial sets over more colors than this are expanded by the whole warp */

/* launch shape of k_color_sets_general. Shared memory per warp, in 32-bit words: the intersection accumulator (W words,
   W = ceil(num_colors / 32)), the counter planes (planes x W, twice for threshold-union: members and complement-misses),
   two ints per partition, the per-round unit prefix and deferred list, two counters. `planes` = bits of the largest
   possible score = bit_width(max k-mers of a read in the batch). */
struct general_plan {
    uint32_t ints_per_warp, warps_per_block, words_per_read, planes;
    size_t smem_bytes;
    bool ok;
};
static inline general_plan plan_color_sets_general(uint32_t num_colors, uint32_t num_partitions, int algo, uint32_t max_kmers, bool differential = false) {
    general_plan g;
    g.words_per_read = (num_colors + 31) / 32;
    g.planes = 1;
    while (g.planes < 32 && (uint64_t(1) << g.planes) <= max_kmers) g.planes += 1;
    const uint32_t counters = algo == FULGOR_GPU_FULL_INTERSECTION ? 1u : 2u;
    g.ints_per_warp = g.words_per_read * (1 + counters * g.planes) + 2 * num_partitions + (FG_K2_SETS_PER_ROUND + 1) + FG_K2_SETS_PER_ROUND / 2 + 4;
    if (differential) g.ints_per_warp += g.words_per_read; /* one decoded (partial) set at a time */
    g.ints_per_warp = (g.ints_per_warp + 3) & ~3u;
    g.ok = size_t(g.ints_per_warp) * 4 <= 200 * 1024;
    size_t w = (72 * 1024) / (size_t(g.ints_per_warp) * 4); /* ~3 blocks per SM */
    if (w < 1) w = 1;
    if (w > FG_WARPS_PER_BLOCK) w = FG_WARPS_PER_BLOCK;
    g.warps_per_block = uint32_t(w);
    g.smem_bytes = size_t(g.warps_per_block) * g.ints_per_warp * 4;
    return g;
}

/* K2: one warp per read; the read's distinct color sets {id, multiplicity} come from K1.
   Full intersection = colors present in all n hit sets -- the same set as the reference's intersect / meta_intersect
   (src/ps_full_intersection.cpp:33-127, 243-332); threshold union = colors with score >= uint64(double(npos) * threshold)
   (src/ps_threshold_union.cpp:389, merge :17-40, merge_meta :43-120). The result is a bitmap of num_colors bits per read +
   its popcount.

   The UNIT of work is one partial set (hybrid: the set itself; meta: one (partition, partial set) entry of the set's meta
   list, include/color_sets/meta.hpp:93-236). Lanes draw units from a shared ticket, so long delta-coded lists and short
   ones balance out; every lane decodes its own unit with a register-buffered bit cursor:
     full intersection   bitmap: acc &= bitmap (inside the partition's color range)
                         complement-coded: acc &= ~{missing colors}
                         delta-coded (sparse): members counted in the bit-sliced planes, sets counted per partition;
                                               a color survives iff its count equals the partition's number of sparse sets
                         a partition that some set does not list at all is cleared (listed[p] < n)
     threshold union     bitmap / delta-coded: score planes += multiplicity for every member
                         complement-coded: base[p] += multiplicity, miss planes += multiplicity for every MISSING color
                                           (the reference's trick, src/ps_threshold_union.cpp:23-29: cost ~ encoded length)
                         final: score - miss + base[p] >= min_score, evaluated 32 colors at a time with bit-sliced adders
   Bitmaps over more than FG_K2_WIDE_COLORS colors are deferred and expanded by the whole warp, one word per lane. */
__global__ void __launch_bounds__(FG_BLOCK) k_color_sets_general(const __grid_constant__ dev_index I, const uint32_t* __restrict__ counts,
                                                                const uint2* __restrict__ stage, const uint2* __restrict__ pool,
                                                                const uint32_t* __restrict__ num_positive, uint32_t n_reads, int algo,
                                                                double threshold, uint32_t words_per_read, uint32_t planes_cap,
                                                                uint32_t smem_ints_per_warp, uint32_t* __restrict__ res_bits,
                                                                uint32_t* __restrict__ res_counts) {
#ifdef FG_SIMT_EMUL
    uint32_t* smem = static_cast<uint32_t*>(fg_emul_dynamic_smem());
#else
    extern __shared__ uint32_t smem[];
#endif
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const uint32_t C = I.num_colors, P = I.num_partitions, W = words_per_read;
    const bool fi = algo == FULGOR_GPU_FULL_INTERSECTION;
    uint32_t* acc = smem + size_t(wib) * smem_ints_per_warp;
    uint32_t* planes_a = acc + W;
    uint32_t* planes_m = planes_a + planes_cap * W; /* threshold union only */
    uint32_t* pa = planes_m + (fi ? 0 : planes_cap * W); /* per partition: FI sets that list it | TU complement base weight */
    uint32_t* pb = pa + P;                                /* per partition: FI sparse sets */
    uint32_t* pref = pb + P;                              /* units before set j of the round (FG_K2_SETS_PER_ROUND + 1) */
    uint16_t* wide = reinterpret_cast<uint16_t*>(pref + FG_K2_SETS_PER_ROUND + 1);
    uint32_t* ctl = reinterpret_cast<uint32_t*>(wide + FG_K2_SETS_PER_ROUND); /* [0] ticket, [1] deferred units */
    const uint32_t warps = gridDim.x * wpb;
    for (uint32_t r = blockIdx.x * wpb + wib; r < n_reads; r += warps) {
        const uint32_t n = __ldg(counts + r);
        uint32_t* out = res_bits + uint64_t(r) * W;
        if (n == 0) { /* no positive k-mer: empty result (src/ps_full_intersection.cpp:385, ps_threshold_union.cpp:355) */
            for (uint32_t w = lane; w < W; w += 32) out[w] = 0;
            if (lane == 0) res_counts[r] = 0;
            continue;
        }
        const uint2* ents = entries_of(r, n, stage, pool);
        const uint32_t npos = __ldg(num_positive + r);
        const uint64_t min_score = fi ? uint64_t(n) : uint64_t(double(npos) * threshold);
        uint32_t np = 1; /* planes this read needs: scores are at most n (FI) or npos (TU) */
        while (np < planes_cap && (1u << np) <= (fi ? n : npos)) np += 1;
        const bit_planes A{planes_a, W, np}, M{planes_m, W, np};
        for (uint32_t w = lane; w < W; w += 32) acc[w] = w + 1 < W || (C & 31) == 0 ? ~0u : (1u << (C & 31)) - 1u;
        for (uint32_t i = lane; i < np * W; i += 32) {
            planes_a[i] = 0;
            if (!fi) planes_m[i] = 0;
        }
        for (uint32_t p = lane; p < 2 * P; p += 32) pa[p] = 0;
        __syncwarp();

        if (I.diff) {
            /* Differential containers (.dfur / .mdfur, include/color_sets/differential.hpp:256-287): a (partial) set is the symmetric
               difference of its cluster's representative and its own difference list, so it cannot be folded into the result list
               by list; the reference merges the two on the fly (diff_intersect, src/ps_full_intersection.cpp:130-240; merge_diff /
               merge_metadiff, src/ps_threshold_union.cpp:123-318). Here each one is decoded into a scratch bitmap (lane 0 the
               representative, lane 1 the differences, atomic XOR) and then treated as a bitmap. One set at a time: this is the path for
               indexes whose decoded table does not fit the device, not the fast one. */
            uint32_t* tmp = ctl + 4;
            for (uint32_t j = 0; j < n; ++j) {
                const uint2 e = ents[j];
                const uint32_t weight = fi ? 1u : e.y;
                uint64_t list = 0;
                uint32_t units = 1;
                if (I.type != 0) {
                    list = __ldg(I.meta_off + e.x);
                    units = __ldg(I.meta_vals + list);
                }
                for (uint32_t u = 0; u < units; ++u) {
                    uint32_t part = 0, color_base = 0;
                    uint64_t local_id = e.x;
                    if (I.type != 0) { /* meta.hpp:227-235 */
                        const uint32_t mc = __ldg(I.meta_vals + list + 1 + u);
                        uint32_t plo = 0, phi = P;
                        while (phi - plo > 1) {
                            const uint32_t mid = (plo + phi) >> 1;
                            if (__ldg(I.part_sets_before + mid) <= mc) plo = mid; else phi = mid;
                        }
                        part = plo;
                        local_id = mc - __ldg(I.part_sets_before + plo);
                        color_base = __ldg(I.part_min_color + plo);
                    }
                    const fgi_hybrid* h = I.hybrids + part;
                    const uint32_t nc = __ldg(&h->num_colors);
                    const uint32_t w_lo = color_base >> 5, w_hi = (color_base + nc + 31) >> 5; /* words the partition's colors touch */
                    for (uint32_t w = w_lo + lane; w < w_hi; w += 32) tmp[w] = 0;
                    __syncwarp();
                    if (lane < 2) {
                        const uint64_t* words = I.color_words + __ldg(&h->word_base);
                        const uint64_t base = __ldg(&h->set_off_base);
                        bit_cursor cur;
                        cur.open(words, __ldg(I.set_bit_off + base + (lane ? 0 : __ldg(&h->num_sets) + 1) + local_id));
                        const uint32_t cnt = cur.delta();
                        if (lane) cur.delta(); /* the difference list carries the size of the decoded set */
                        uint32_t v = 0;
                        for (uint32_t i = 0; i < cnt; ++i) {
                            const uint32_t d = cur.delta();
                            v = i ? v + d + 1 : d;
                            const uint32_t c = color_base + v;
                            atomicXor(tmp + (c >> 5), 1u << (c & 31));
                        }
                    }
                    __syncwarp();
                    for (uint32_t w = w_lo + lane; w < w_hi; w += 32) {
                        const int lo = int(32 * w) - int(color_base);
                        const int first = lo < 0 ? -lo : 0, last = min(32, int(nc) - lo);
                        const uint32_t gmask = (last >= 32 ? ~0u : ((1u << last) - 1u)) & ~((1u << first) - 1u);
                        const uint32_t bits = tmp[w] & gmask;
                        if (fi) atomicAnd(acc + w, bits | ~gmask); /* edge words are shared with the neighbouring partitions' lanes */
                        else if (bits) A.add_word(w, bits, weight);
                    }
                    if (fi && lane == 0) pa[part] += 1;
                    __syncwarp();
                }
            }
        }
        for (uint32_t j0 = 0; !I.diff && j0 < n; j0 += FG_K2_SETS_PER_ROUND) {
            const uint32_t nb = min(uint32_t(FG_K2_SETS_PER_ROUND), n - j0);
            /* units of this round: prefix of the sets' partial-set counts */
            uint32_t carry = 0;
            for (uint32_t b = 0; b < nb; b += 32) {
                const uint32_t j = b + lane;
                uint32_t units = 0;
                if (j < nb) units = I.type == 0 ? 1u : __ldg(I.meta_vals + __ldg(I.meta_off + ents[j0 + j].x));
                uint32_t incl = units;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(FG_FULL, incl, d);
                    if (lane >= uint32_t(d)) incl += y;
                }
                if (j < nb) pref[j] = carry + incl - units;
                carry += __shfl_sync(FG_FULL, incl, 31);
            }
            const uint32_t total = carry;
            if (lane == 0) {
                pref[nb] = total;
                ctl[0] = 0;
                ctl[1] = 0;
            }
            __syncwarp();
            /* lane-parallel: every lane draws units until none is left */
            for (;;) {
                const uint32_t u = atomicAdd(ctl, 1u);
                if (u >= total) break;
                uint32_t lo = 0, hi = nb; /* the set j with pref[j] <= u < pref[j + 1] */
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (pref[mid] <= u) lo = mid; else hi = mid;
                }
                const uint2 e = ents[j0 + lo];
                const uint32_t weight = fi ? 1u : e.y;
                uint32_t part = 0;
                set_item it;
                if (I.type == 0) {
                    it = open_set(I, 0, e.x, 0);
                } else { /* meta: the (u - pref[lo])-th entry of the set's meta list -> (partition, local set id), meta.hpp:227-235 */
                    const uint32_t mc = __ldg(I.meta_vals + __ldg(I.meta_off + e.x) + 1 + (u - pref[lo]));
                    uint32_t plo = 0, phi = P; /* largest p with sets_before[p] <= mc */
                    while (phi - plo > 1) {
                        const uint32_t mid = (plo + phi) >> 1;
                        if (__ldg(I.part_sets_before + mid) <= mc) plo = mid; else phi = mid;
                    }
                    part = plo;
                    it = open_set(I, part, mc - __ldg(I.part_sets_before + part), __ldg(I.part_min_color + part));
                }
                if (fi) {
                    atomicAdd(pa + part, 1u);
                    if (it.enc == FG_ENC_DELTA) atomicAdd(pb + part, 1u);
                } else if (it.enc == FG_ENC_COMPLEMENT) {
                    atomicAdd(pa + part, weight);
                }
                if (it.enc == FG_ENC_BITMAP) {
                    if (it.num_colors > FG_K2_WIDE_COLORS) {
                        const uint32_t slot = atomicAdd(ctl + 1, 1u);
                        if (slot < FG_K2_SETS_PER_ROUND && u < 65536u) { /* else: expanded right here, by this lane alone */
                            wide[slot] = uint16_t(u);
                            continue;
                        }
                    }
                    for (uint32_t w = it.color_base >> 5; 32 * w < it.color_base + it.num_colors; ++w) {
                        uint32_t gmask;
                        const uint32_t bits = bitmap_word(it, w, gmask);
                        if (fi) atomicAnd(acc + w, bits | ~gmask);
                        else if (bits) A.add_word(w, bits, weight);
                    }
                    continue;
                }
                bit_cursor cur;
                cur.open(it.words, it.pos);
                uint32_t v = 0;
                for (uint32_t i = 0; i < it.nvals; ++i) {
                    const uint32_t d = cur.delta();
                    v = i ? v + d + 1 : d;
                    const uint32_t c = it.color_base + v;
                    if (it.enc == FG_ENC_COMPLEMENT) {
                        if (fi) atomicAnd(acc + (c >> 5), ~(1u << (c & 31)));
                        else M.add_color(c, weight);
                    } else {
                        A.add_color(c, weight);
                    }
                }
            }
            __syncwarp();
            /* the deferred wide bitmaps, one after the other, a word per lane */
            const uint32_t nwide = min(ctl[1], uint32_t(FG_K2_SETS_PER_ROUND));
            for (uint32_t x = 0; x < nwide; ++x) {
                const uint32_t u = wide[x];
                uint32_t lo = 0, hi = nb;
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (pref[mid] <= u) lo = mid; else hi = mid;
                }
                const uint2 e = ents[j0 + lo];
                set_item it;
                if (I.type == 0) {
                    it = open_set(I, 0, e.x, 0);
                } else {
                    const uint32_t mc = __ldg(I.meta_vals + __ldg(I.meta_off + e.x) + 1 + (u - pref[lo]));
                    uint32_t plo = 0, phi = P;
                    while (phi - plo > 1) {
                        const uint32_t mid = (plo + phi) >> 1;
                        if (__ldg(I.part_sets_before + mid) <= mc) plo = mid; else phi = mid;
                    }
                    it = open_set(I, plo, mc - __ldg(I.part_sets_before + plo), __ldg(I.part_min_color + plo));
                }
                for (uint32_t w = (it.color_base >> 5) + lane; 32 * w < it.color_base + it.num_colors; w += 32) {
                    uint32_t gmask;
                    const uint32_t bits = bitmap_word(it, w, gmask);
                    if (fi) atomicAnd(acc + w, bits | ~gmask); /* neighbouring lanes may share the range's edge words with no one: plain AND would do, the atomic keeps it simple */
                    else if (bits) A.add_word(w, bits, e.y);
                }
                __syncwarp();
            }
            __syncwarp();
        }

        /* final: 32 colors per lane and step; a word may straddle partitions */
        uint32_t total = 0;
        for (uint32_t w0 = 0; w0 < W; w0 += 32) {
            const uint32_t w = w0 + lane;
            uint32_t word = 0;
            if (w < W) {
                uint32_t p = 0;
                if (P > 1) { /* the partition of the word's first color: largest p with min_color[p] <= 32 w */
                    uint32_t plo = 0, phi = P;
                    while (phi - plo > 1) {
                        const uint32_t mid = (plo + phi) >> 1;
                        if (__ldg(I.part_min_color + mid) <= 32 * w) plo = mid; else phi = mid;
                    }
                    p = plo;
                }
                for (;; ++p) {
                    const uint32_t cb = P > 1 ? __ldg(I.part_min_color + p) : 0u, ce = P > 1 ? __ldg(I.part_min_color + p + 1) : C;
                    const int first = cb > 32 * w ? int(cb - 32 * w) : 0, last = min(32, int(ce) - int(32 * w));
                    if (last > first) {
                        const uint32_t rm = (last >= 32 ? ~0u : ((1u << last) - 1u)) & ~((1u << first) - 1u);
                        uint32_t ok;
                        if (fi) {
                            ok = pa[p] == n ? acc[w] : 0u;
                            const uint32_t ns = pb[p];
                            for (uint32_t j = 0; j < np; ++j) ok &= ((ns >> j) & 1u) ? A.plane(j, w) : ~A.plane(j, w);
                        } else { /* sign of (A - M + K), K = base[p] - min_score, in np + 3 bits of two's complement */
                            const int64_t K = int64_t(pa[p]) - int64_t(min_score);
                            uint32_t c1 = ~0u, c2 = 0u, sign = 0u; /* A + ~M + 1, then + K */
                            for (uint32_t j = 0; j < np + 3; ++j) {
                                const uint32_t a = A.plane(j, w), m = ~M.plane(j, w);
                                const uint32_t s1 = a ^ m ^ c1;
                                c1 = (a & m) | (a & c1) | (m & c1);
                                const uint32_t kb = ((K >> j) & 1) ? ~0u : 0u;
                                sign = s1 ^ kb ^ c2;
                                c2 = (s1 & kb) | (s1 & c2) | (kb & c2);
                            }
                            ok = ~sign;
                        }
                        word |= ok & rm;
                    }
                    if (P <= 1 || p + 1 >= P || ce >= 32 * w + 32) break;
                }
                out[w] = word;
            }
            total += __reduce_add_sync(FG_FULL, uint32_t(__popc(word)));
        }
        if (lane == 0) res_counts[r] = total;
        __syncwarp();
    }
}

/* ---- the decoded color-set table ---- */

/* words per table row: the color bitmap padded to a whole number of 32-word (128-byte) warp loads */
static inline uint64_t table_stride_words(uint32_t num_colors) { return (uint64_t(num_colors) + 1023) / 1024 * 32; }

/* Decodes color sets [first_set, first_set + n_sets) into bitmap rows, one warp per set: hybrid::forward_iterator
   (include/color_sets/hybrid.hpp:151-305) / meta<hybrid>::forward_iterator (include/color_sets/meta.hpp:93-236) run once per
   set at load time instead of once per read and set at query time. 180 GB of HBM hold the table of any index Fulgor builds
   today (salmonella_4546: 972,178 sets x 640 B = 0.6 GB); the queries then read plain, coalesced bitmaps.
   The row is assembled in shared memory: lanes take the set's partial sets (one for a hybrid index), set member bits with
   atomic ORs -- a complement-coded partial set first fills its partition's color range, then clears the missing colors. */
__global__ void __launch_bounds__(FG_BLOCK) k_expand_color_sets(const __grid_constant__ dev_index I, uint64_t first_set, uint32_t n_sets,
                                                               uint32_t stride, uint32_t* __restrict__ table) {
#ifdef FG_SIMT_EMUL
    uint32_t* smem = static_cast<uint32_t*>(fg_emul_dynamic_smem());
#else
    extern __shared__ uint32_t smem[];
#endif
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    uint32_t* row = smem + size_t(wib) * stride;
    const uint32_t P = I.num_partitions;
    for (uint32_t s = blockIdx.x * wpb + wib; s < n_sets; s += gridDim.x * wpb) {
        const uint64_t cid = first_set + s;
        for (uint32_t w = lane; w < stride; w += 32) row[w] = 0;
        __syncwarp();
        uint64_t list = 0;
        uint32_t units = 1;
        if (I.type != 0) {
            list = __ldg(I.meta_off + cid);
            units = __ldg(I.meta_vals + list);
        }
        for (uint32_t u = lane; u < units; u += 32) {
            uint32_t container = 0, color_base = 0;
            uint64_t local_id = cid;
            if (I.type != 0) { /* meta.hpp:227-235 */
                const uint32_t mc = __ldg(I.meta_vals + list + 1 + u);
                uint32_t plo = 0, phi = P;
                while (phi - plo > 1) {
                    const uint32_t mid = (plo + phi) >> 1;
                    if (__ldg(I.part_sets_before + mid) <= mc) plo = mid; else phi = mid;
                }
                container = plo;
                local_id = mc - __ldg(I.part_sets_before + plo);
                color_base = __ldg(I.part_min_color + plo);
            }
            if (I.diff) { /* differential.hpp:256-287: the set = representative XOR difference list, both gap lists; the row's
                             bits of this partition's color range are zero before, and no other unit touches them */
                const fgi_hybrid* h = I.hybrids + container;
                const uint64_t* words = I.color_words + __ldg(&h->word_base);
                const uint64_t base = __ldg(&h->set_off_base);
                for (uint32_t part = 0; part < 2; ++part) {
                    bit_cursor cur;
                    cur.open(words, __ldg(I.set_bit_off + base + (part ? __ldg(&h->num_sets) + 1 : 0) + local_id));
                    const uint32_t n = cur.delta();
                    if (part == 0) cur.delta(); /* size of the decoded set */
                    uint32_t v = 0;
                    for (uint32_t i = 0; i < n; ++i) {
                        const uint32_t d = cur.delta();
                        v = i ? v + d + 1 : d;
                        const uint32_t c = color_base + v;
                        atomicXor(row + (c >> 5), 1u << (c & 31));
                    }
                }
                continue;
            }
            const set_item it = open_set(I, container, local_id, color_base);
            if (it.enc != FG_ENC_DELTA) { /* bitmap: its bits; complement: the whole range first */
                for (uint32_t w = it.color_base >> 5; 32 * w < it.color_base + it.num_colors; ++w) {
                    uint32_t gmask;
                    const uint32_t bits = bitmap_word(it, w, gmask);
                    atomicOr(row + w, it.enc == FG_ENC_BITMAP ? bits : gmask);
                }
            }
            if (it.enc != FG_ENC_BITMAP) {
                bit_cursor cur;
                cur.open(it.words, it.pos);
                uint32_t v = 0;
                for (uint32_t i = 0; i < it.nvals; ++i) {
                    const uint32_t d = cur.delta();
                    v = i ? v + d + 1 : d;
                    const uint32_t c = it.color_base + v;
                    if (it.enc == FG_ENC_COMPLEMENT) atomicAnd(row + (c >> 5), ~(1u << (c & 31)));
                    else atomicOr(row + (c >> 5), 1u << (c & 31));
                }
            }
        }
        __syncwarp();
        for (uint32_t w = lane; w < stride; w += 32) table[uint64_t(s) * stride + w] = row[w];
        __syncwarp();
    }
}

/* K2 on the decoded table: one warp per read, the read's result accumulated IN REGISTERS, 32 * T colors-words per pass
   (lane l owns words l, l + 32, ...). Every hit set costs T coalesced 128-byte loads and, for full intersection, T ANDs;
   for threshold union the bitmap is added to NP bit-sliced counter planes (kernels.cuh: bit_planes, here in registers,
   no atomics: a lane owns its words) and the threshold test is one bit-sliced subtraction. Same results as
   k_color_sets_general. NP = counter bits (scores < 2^NP), T = words per lane and pass. */
/* Bit-sliced per-color counters in CARRY-SAVE form for the threshold union on the decoded table: plane k holds bit k of 32 T
   colors' scores (acc) plus at most one PENDING vector of the same weight 2^k. A vector arriving at a plane that has no
   pending one is just parked there (no arithmetic); when a second one arrives, the plane's 3:2 compressor folds
   {acc, pending, new} into the new plane bit and one carry for the next plane. Which planes hold a pending vector depends
   only on the multiplicities, which every lane shares, so the control flow is warp-uniform and the planes stay in registers
   (compile-time indices). A hit set costs about one compressor (two LOP3 per word) per set bit of its multiplicity, however
   many planes the counters have -- the ripple-carry adder this replaces paid a full adder per plane and word for every set. */
/* one 3:2 compressor step IN PLACE: a <- a ^ p ^ x (the plane's new bit), x <- majority(a, p, x) (the carry). Two LOP3, the
   second one on the NEW a (0x8e = majority(A ^ B ^ C, B, C) as a function of A, B, C), so no register beyond the three
   operands is live -- written as two plain expressions the compiler kept the old a for the carry and paid a chain of ~20
   register moves where the planes' exits meet (profiles/r02_big_mfur_tu_mixed_k2_*: 21 % IMAD, half of them moves). */
__device__ __forceinline__ void fg_csa(uint32_t& a, uint32_t p, uint32_t& x) {
#if defined(__CUDA_ARCH__)
    asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(p), "r"(x));
    asm("lop3.b32 %0, %1, %2, %0, 0x8e;" : "+r"(x) : "r"(a), "r"(p));
#else
    const uint32_t s = a ^ p ^ x;
    x = (a & p) | (a & x) | (p & x);
    a = s;
#endif
}

/* Planes [0, HI) -- the ones nearly every hit set touches -- keep their accumulated bits in registers and a pending vector in
   shared memory. Planes [HI, NP) live in shared memory altogether, as plain ripple-carry adders without a pending vector: a
   vector only gets there as the carry out of plane HI - 1 (once per 2^HI units of weight) or with a multiplicity >= 2^HI, and
   ripples on while any lane still holds a carry bit. That takes NP x T registers down to HI x T -- the 10-plane kernel ran out
   of registers at 80 (the loads' descriptors were rebuilt from spilled copies around every row, profiles/r02_big_mfur_tu_mixed_k2_*)
   -- for a fourth resident block per SM; the shared memory per warp stays NP x T x 128 bytes. */
#ifndef FG_TU_HI
#define FG_TU_HI 4
#endif
template <int NP, int T>
struct carry_save_counters {
    static constexpr int HI = NP < FG_TU_HI ? NP : FG_TU_HI;
    uint32_t acc[HI][T];
    uint32_t* pend;   /* this lane's column of the warp's shared memory: word (K, t) at pend[(K * T + t) * 32] -- for K < HI the plane's
                         pending vector, for K >= HI the plane itself */
    uint32_t pending; /* bit k: plane k < HI holds a pending vector (warp-uniform) */

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < HI; ++k) {
#pragma unroll
            for (int t = 0; t < T; ++t) acc[k][t] = 0u;
        }
#pragma unroll
        for (int k = HI; k < NP; ++k) {
#pragma unroll
            for (int t = 0; t < T; ++t) pend[(k * T + t) * 32] = 0u;
        }
        pending = 0;
    }
    /* bit k of the scores of this lane's word t (after finish()) */
    __device__ __forceinline__ uint32_t plane(int k, int t) const { return k < HI ? acc[k < HI ? k : 0][t] : pend[(k * T + t) * 32]; }
    template <int K>
    __device__ __forceinline__ void park(const uint32_t (&v)[T]) {
#pragma unroll
        for (int t = 0; t < T; ++t) pend[(K * T + t) * 32] = v[t];
        pending |= 1u << K;
    }
    template <int K>
    __device__ __forceinline__ void compress(uint32_t (&v)[T]) {
#pragma unroll
        for (int t = 0; t < T; ++t) fg_csa(acc[K < HI ? K : 0][t], pend[(K * T + t) * 32], v[t]);
        pending &= ~(1u << K);
    }
    /* plane K >= HI += v, v <- the carry; false when no lane of the warp has a carry left */
    template <int K>
    __device__ __forceinline__ bool ripple(uint32_t (&v)[T]) {
        uint32_t any = 0;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const uint32_t a = pend[(K * T + t) * 32];
            pend[(K * T + t) * 32] = a ^ v[t];
            v[t] &= a;
            any |= v[t];
        }
        return __any_sync(FG_FULL, any != 0);
    }
    /* counters[c] += 2^b for every set bit c of v (v is consumed). The vector enters at plane b: every plane with a pending
       vector compresses and passes the carry on to the next one, the first one without parks it. ONE copy of every plane's
       code: the entry plane is a jump into a chain of cases that fall through (the recursion this replaces instantiated the
       chain once per entry plane, NP^2 / 2 plane bodies). A carry out of the top plane cannot happen (scores < 2^NP). */
    __device__ __forceinline__ void add(uint32_t (&v)[T], uint32_t b) {
#define FG_CS_PLANE(K)                              \
    case K:                                         \
        if constexpr (K < HI) {                     \
            if (!((pending >> K) & 1u)) {           \
                park<K>(v);                         \
                break;                              \
            }                                       \
            compress<K>(v);                         \
        } else if constexpr (K < NP) {              \
            if (!ripple<K>(v)) break;               \
        }                                           \
        [[fallthrough]];
        switch (b) {
            FG_CS_PLANE(0) FG_CS_PLANE(1) FG_CS_PLANE(2) FG_CS_PLANE(3) FG_CS_PLANE(4) FG_CS_PLANE(5) FG_CS_PLANE(6) FG_CS_PLANE(7)
            FG_CS_PLANE(8) FG_CS_PLANE(9) FG_CS_PLANE(10) FG_CS_PLANE(11) FG_CS_PLANE(12) FG_CS_PLANE(13) FG_CS_PLANE(14) FG_CS_PLANE(15)
            FG_CS_PLANE(16) FG_CS_PLANE(17) FG_CS_PLANE(18) FG_CS_PLANE(19) FG_CS_PLANE(20) FG_CS_PLANE(21) FG_CS_PLANE(22) FG_CS_PLANE(23)
            FG_CS_PLANE(24) FG_CS_PLANE(25) FG_CS_PLANE(26) FG_CS_PLANE(27) FG_CS_PLANE(28) FG_CS_PLANE(29) FG_CS_PLANE(30) FG_CS_PLANE(31)
            default: break;
        }
#undef FG_CS_PLANE
    }
    /* folds the pending vectors in, lowest plane first: afterwards plane(k, t) is bit k of the scores */
    template <int K>
    __device__ __forceinline__ void finish_from() {
        if constexpr (K < HI) {
            if ((pending >> K) & 1u) {
                uint32_t c[T];
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const uint32_t p = pend[(K * T + t) * 32];
                    c[t] = acc[K][t] & p;
                    acc[K][t] ^= p;
                }
                pending &= ~(1u << K);
                add(c, K + 1);
            }
            finish_from<K + 1>();
        }
    }
    __device__ __forceinline__ void finish() { finish_from<0>(); }
};

/* dynamic shared memory of k_color_sets_table: the pending vectors of every warp's counters (none for full intersection) */
static inline size_t table_kernel_smem(bool fi, int NP, int T) { return fi ? 0 : size_t(FG_WARPS_PER_BLOCK) * NP * T * 32 * 4; }

/* resident blocks per SM the table kernels are compiled for (register budget 65536 / (blocks * 256)) */
#ifndef FG_TU_BLOCKS
#define FG_TU_BLOCKS 4
#endif
#ifndef FG_FI_DEPTH
#define FG_FI_DEPTH 1 /* rows in flight ahead of the one being ANDed (full intersection) */
#endif
#ifndef FG_FI_BLOCKS
#define FG_FI_BLOCKS 6
#endif
static constexpr int table_kernel_blocks(bool fi, int NP) { return fi ? FG_FI_BLOCKS : (NP <= 10 ? (FG_TU_HI <= 4 ? FG_TU_BLOCKS : 3) : 2); }

template <bool FI, int NP, int T>
__global__ void __launch_bounds__(FG_BLOCK, table_kernel_blocks(FI, NP)) k_color_sets_table(const __grid_constant__ dev_index I, const uint32_t* __restrict__ counts,
                                                              const uint2* __restrict__ stage, const uint2* __restrict__ pool,
                                                              const uint32_t* __restrict__ num_positive, uint32_t n_reads, double threshold,
                                                              uint32_t words_per_read, uint32_t* __restrict__ res_bits,
                                                              uint32_t* __restrict__ res_counts) {
#ifdef FG_SIMT_EMUL
    uint32_t* smem = static_cast<uint32_t*>(fg_emul_dynamic_smem());
#else
    extern __shared__ uint32_t smem[];
#endif
    const uint32_t lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const uint32_t C = I.num_colors, W = words_per_read;
    const uint32_t stride = uint32_t(I.table_stride); /* words per row, a multiple of 32 (table_stride_words) */
    const uint32_t row_bytes = stride * 4; /* < 2^32: a row holds num_colors bits */
    for (uint32_t r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_reads; r += gridDim.x * wpb) {
        const uint32_t n = __ldg(counts + r);
        uint32_t* out = res_bits + uint64_t(r) * W;
        if (n == 0) { /* no positive k-mer: empty result (src/ps_full_intersection.cpp:385, ps_threshold_union.cpp:355) */
            for (uint32_t w = lane; w < W; w += 32) out[w] = 0;
            if (lane == 0) res_counts[r] = 0;
            continue;
        }
        const uint2* ents = entries_of(r, n, stage, pool);
        const uint64_t min_score = FI ? uint64_t(n) : uint64_t(double(__ldg(num_positive + r)) * threshold);
        uint32_t total = 0;
        for (uint32_t w0 = 0; w0 < W; w0 += 32 * T) {
            uint32_t acc[T];                                    /* FI: the running intersection */
            carry_save_counters<FI ? 1 : NP, FI ? 1 : T> cs;   /* TU: the colors' scores, bit-sliced */
#pragma unroll
            for (int t = 0; t < T; ++t) acc[t] = ~0u;
            if (!FI) {
                cs.pend = smem + size_t(threadIdx.x >> 5) * ((FI ? 1 : NP) * (FI ? 1 : T) * 32) + lane;
                cs.clear();
            }
            /* A row load is ONE 64-bit multiply-add (row id x row bytes + this lane's base) and T loads at constant offsets;
               whether the pass needs guards (a last, partial pass over the row) is decided once per pass, not per load -- the
               guarded form cost ~27 instructions per row around its 5 loads. The rows are fetched one entry AHEAD of the
               arithmetic: a warp keeps 2 T row loads in flight instead of T. */
            const char* lane_base = reinterpret_cast<const char*>(I.set_table + w0 + lane);
            const bool whole = w0 + 32 * T <= stride; /* warp-uniform */
            auto load_row = [&](uint32_t id, uint32_t (&x)[T]) {
                const uint32_t* row = reinterpret_cast<const uint32_t*>(lane_base + uint64_t(id) * row_bytes); /* one IMAD.WIDE */
                if (whole) {
#pragma unroll
                    for (int t = 0; t < T; ++t) x[t] = __ldg(row + 32 * t);
                } else {
#pragma unroll
                    for (int t = 0; t < T; ++t) x[t] = w0 + 32 * t < stride ? __ldg(row + 32 * t) : 0u;
                }
            };
            if constexpr (FI) {
                /* full intersection: the rows are fetched FG_FI_DEPTH entries ahead of the ANDs (and the entry list one further):
                   the kernel waits on these loads and on nothing else (long-scoreboard stalls, profiles/r02_big_fi_k2_*) */
                uint2 e[FG_FI_DEPTH + 1];
                uint32_t x[FG_FI_DEPTH + 1][T];
#pragma unroll
                for (int d = 0; d <= FG_FI_DEPTH; ++d) e[d] = uint32_t(d) < n ? ents[d] : make_uint2(0u, 0u);
#pragma unroll
                for (int d = 0; d < FG_FI_DEPTH; ++d)
                    if (uint32_t(d) < n) load_row(e[d].x, x[d]);
                for (uint32_t j = 0; j < n; ++j) {
                    const uint2 e_after = j + FG_FI_DEPTH + 1 < n ? ents[j + FG_FI_DEPTH + 1] : make_uint2(0u, 0u);
                    if (j + FG_FI_DEPTH < n) load_row(e[FG_FI_DEPTH].x, x[FG_FI_DEPTH]);
#pragma unroll
                    for (int t = 0; t < T; ++t) acc[t] &= x[0][t];
#pragma unroll
                    for (int d = 0; d < FG_FI_DEPTH; ++d) {
                        e[d] = e[d + 1];
#pragma unroll
                        for (int t = 0; t < T; ++t) x[d][t] = x[d + 1][t];
                    }
                    e[FG_FI_DEPTH] = e_after;
                }
            } else {
                /* threshold union: the entry list comes 32 entries at a time, one per lane in ONE coalesced load (the batch after
                   it is already on its way), and is handed round with shuffles; the row of entry j + 1 is fetched while the
                   counters take the row of entry j */
                const uint2 none = make_uint2(0u, 0u);
                uint2 cur = lane < n ? ents[lane] : none;
                uint2 nxt = lane + 32 < n ? ents[lane + 32] : none;
                uint32_t x[T], x_next[T];
                load_row(__shfl_sync(FG_FULL, cur.x, 0), x);
                for (uint32_t j = 0; j < n; ++j) {
                    const uint32_t mult = __shfl_sync(FG_FULL, cur.y, j & 31u);
                    if (j + 1 < n) {
                        if (((j + 1) & 31u) == 0) { /* next batch of entries */
                            cur = nxt;
                            nxt = j + 33 + lane < n ? ents[j + 33 + lane] : none;
                        }
                        load_row(__shfl_sync(FG_FULL, cur.x, (j + 1) & 31u), x_next);
                    }
                    /* score += multiplicity for every member: the bitmap enters at the plane of every set bit of the multiplicity */
                    for (uint32_t wt = mult; wt; wt &= wt - 1) {
                        uint32_t v[T];
#pragma unroll
                        for (int t = 0; t < T; ++t) v[t] = x[t];
                        cs.add(v, uint32_t(__ffs(int(wt))) - 1u);
                    }
#pragma unroll
                    for (int t = 0; t < T; ++t) x[t] = x_next[t];
                }
            }
            if (!FI) cs.finish();
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const uint32_t w = w0 + 32 * t + lane;
                uint32_t word = acc[t];
                if (!FI) { /* score >= min_score <=> carry out of score + ~min_score + 1 over NP bits */
                    uint32_t carry = ~0u;
#pragma unroll
                    for (int k = 0; k < (FI ? 1 : NP); ++k) {
                        const uint32_t nb = ((min_score >> k) & 1u) ? 0u : ~0u, pk = cs.plane(k, FI ? 0 : t);
                        carry = (pk & nb) | (pk & carry) | (nb & carry);
                    }
                    word = (min_score >> NP) ? 0u : carry; /* a threshold beyond the counters' range is never met */
                }
                if (w + 1 == W && (C & 31)) word &= (1u << (C & 31)) - 1u;
                if (w < W) out[w] = word; else word = 0;
                total += __popc(word);
            }
        }
        total = __reduce_add_sync(FG_FULL, total);
        if (lane == 0) res_counts[r] = total;
    }
}

/* ---- the per-k-mer tools on the lookup kernel: kmer-conservation and kmer-matches ----
   k_kmer_color_sets runs the same segment pipeline as the pseudoalignment kernels (kmer_tiles) in its per-k-mer mode: the
   color-set id of every k-mer of every read, FG_NOT_FOUND for negative and invalid k-mers -- what the loops around
   streaming_query::lookup_advanced + u2c see (src/kmer_conservation.cpp:31-36, src/kmer_matches.cpp:20-24).
   kmer_off (n_reads + 1, chunk-local) = first k-mer slot of every read. */
template <int W, class READS>
__global__ void __launch_bounds__(FG_BLOCK, FG_MIN_BLOCKS) k_kmer_color_sets(const __grid_constant__ dev_index I, const READS in,
                                                                            uint32_t n_reads, const uint64_t* __restrict__ kmer_off,
                                                                            uint32_t* __restrict__ per_kmer) {
    __shared__ warp_stage wstage[FG_WARPS_PER_BLOCK];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warps = gridDim.x * FG_WARPS_PER_BLOCK;
    for (uint32_t r = blockIdx.x * FG_WARPS_PER_BLOCK + wib; r < n_reads; r += warps) {
        auto tiles = make_tiles<W, true>(I, in, r, n_reads, lane, wstage[wib]);
        tiles.per_kmer = per_kmer + __ldg(kmer_off + r);
        uint32_t cid, cnt;
        while (tiles.next(cid, cnt)) {}
        __syncwarp();
    }
}

/* index::kmer_conservation (src/kmer_conservation.cpp:7-54): maximal runs of consecutive positive k-mers with the same
   color-set id, as triples {start_pos_in_query, num_kmers, color_set_id} (include/util.hpp:74-78). A run starts at a positive
   k-mer whose predecessor is negative or has another color set, and ends likewise; the j-th start of a read pairs with its
   j-th end, so both are placed by a running count. EMIT = false: only the number of runs per read (for the CSR offsets). */
template <bool EMIT>
__global__ void __launch_bounds__(256) k_kmer_runs(const uint32_t* __restrict__ per_kmer, const uint64_t* __restrict__ kmer_off, uint32_t n_reads,
                                                  uint32_t* __restrict__ run_counts, const uint64_t* __restrict__ off,
                                                  const uint64_t* __restrict__ chunk_info, uint32_t* __restrict__ triples, uint64_t cap_triples) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_reads) return;
    const uint64_t k0 = __ldg(kmer_off + r);
    const uint32_t nk = uint32_t(__ldg(kmer_off + r + 1) - k0);
    const uint32_t* c = per_kmer + k0;
    const uint64_t o = EMIT ? __ldg(off + r) - chunk_info[0] : 0;
    uint32_t nstart = 0, nend = 0;
    for (uint32_t i0 = 0; i0 < nk; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint32_t cur = i < nk ? __ldg(c + i) : FG_NOT_FOUND;
        const uint32_t prev = i > 0 && i < nk ? __ldg(c + i - 1) : FG_NOT_FOUND;
        const uint32_t next = i + 1 < nk ? __ldg(c + i + 1) : FG_NOT_FOUND;
        const bool is_start = cur != FG_NOT_FOUND && prev != cur, is_end = cur != FG_NOT_FOUND && next != cur;
        const uint32_t bs = __ballot_sync(FG_FULL, is_start), be = __ballot_sync(FG_FULL, is_end);
        if (EMIT) {
            const uint32_t below = (1u << lane) - 1u;
            if (is_start) {
                const uint64_t t = o + nstart + __popc(bs & below);
                if (t < cap_triples) {
                    triples[3 * t] = i;
                    triples[3 * t + 2] = cur;
                }
            }
            if (is_end) {
                const uint64_t t = o + nend + __popc(be & below);
                if (t < cap_triples) triples[3 * t + 1] = i + 1; /* end (exclusive) for now; the length is fixed up below */
            }
        }
        nstart += __popc(bs);
        nend += __popc(be);
    }
    if (EMIT) {
        __syncwarp();
        for (uint32_t t = lane; t < nstart; t += 32)
            if (o + t < cap_triples) triples[3 * (o + t) + 1] -= triples[3 * (o + t)];
    } else if (lane == 0) {
        run_counts[r] = nstart;
    }
}

/* index::kmer_matches (src/kmer_matches.cpp:7-30), part 1: the positive-k-mer bit vector of every read, bit i = k-mer i is in
   the index; read r owns the 32-bit words [word_off[r], word_off[r+1]) */
__global__ void __launch_bounds__(256) k_kmer_positive_bits(const uint32_t* __restrict__ per_kmer, const uint64_t* __restrict__ kmer_off,
                                                           const uint64_t* __restrict__ word_off, uint32_t n_reads, uint32_t* __restrict__ words) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_reads) return;
    const uint64_t k0 = __ldg(kmer_off + r);
    const uint32_t nk = uint32_t(__ldg(kmer_off + r + 1) - k0);
    uint32_t* out = words + __ldg(word_off + r);
    for (uint32_t i0 = 0; i0 < nk; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint32_t b = __ballot_sync(FG_FULL, i < nk && __ldg(per_kmer + k0 + i) != FG_NOT_FOUND);
        if (lane == 0) out[i0 >> 5] = b;
    }
}

/* index::kmer_matches, part 2: counts[c] = number of positive k-mers whose color set contains c (:25-27), from the read's
   distinct {color-set id, multiplicity} list (K1) and the decoded color-set table: lane l owns color 32 w + l of word w,
   every row word is one broadcast load. counts: n_reads x num_colors. */
__global__ void __launch_bounds__(FG_BLOCK) k_kmer_match_counts(const __grid_constant__ dev_index I, const uint32_t* __restrict__ list_counts,
                                                               const uint2* __restrict__ stage, const uint2* __restrict__ pool, uint32_t n_reads,
                                                               uint32_t* __restrict__ counts) {
    const uint32_t lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const uint32_t C = I.num_colors, W = (C + 31) / 32;
    const uint32_t* __restrict__ table = I.set_table;
    const uint64_t stride = I.table_stride;
    for (uint32_t r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_reads; r += gridDim.x * wpb) {
        const uint32_t n = __ldg(list_counts + r);
        const uint2* ents = n ? entries_of(r, n, stage, pool) : nullptr;
        uint32_t* out = counts + uint64_t(r) * C;
        for (uint32_t w = 0; w < W; ++w) {
            uint32_t acc = 0;
            for (uint32_t j = 0; j < n; ++j) {
                const uint2 e = ents[j];
                acc += ((__ldg(table + uint64_t(e.x) * stride + w) >> lane) & 1u) * e.y;
            }
            if (32 * w + lane < C) out[32 * w + lane] = acc;
        }
    }
}

/* ---- cross-read deduplication of color-set-id lists (the reference's --deduplicate, tools/pseudoalign.cpp:92-226) ----
   The reference sorts the reads' lists lexicographically, intersects each distinct list once and fans the result out to
   every read that has it (preprocessed_query_reader, src/ps_utils.cpp:307-415). Here the distinct lists of a chunk are
   found with an open-addressing table of read indexes: one warp per read hashes its sorted list (order-independent sum of
   mixed ids, so lanes need no prefix), lane 0 claims a slot with a CAS or meets an earlier read there, whose list the
   warp then compares entry by entry (exact: a hash collision only costs one more probe). The first read of a group is its
   representative; every other read gets count 0 for the color-set kernel, so the intersection, the emit and the
   device->host copy happen once per distinct list. rep_of_read[r] = GLOBAL index of r's representative (r itself when it
   is one, and for reads without positive k-mers). slots: 2^log2_slots entries preset to 0xffffffff. */
__global__ void __launch_bounds__(FG_BLOCK) k_group_reads(const uint32_t* __restrict__ counts, const uint2* __restrict__ stage,
                                                         const uint2* __restrict__ pool, uint32_t n_reads, uint32_t read_base,
                                                         uint32_t* __restrict__ slots, uint32_t log2_slots,
                                                         uint32_t* __restrict__ rep_of_read, uint32_t* __restrict__ rep_counts) {
    const uint32_t lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const uint32_t mask = (1u << log2_slots) - 1u;
    for (uint32_t r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_reads; r += gridDim.x * wpb) {
        const uint32_t n = __ldg(counts + r);
        if (n == 0) {
            if (lane == 0) {
                rep_of_read[r] = read_base + r;
                rep_counts[r] = 0;
            }
            continue;
        }
        const uint2* mine = entries_of(r, n, stage, pool);
        uint32_t h = 0;
        for (uint32_t i = lane; i < n; i += 32) {
            uint32_t x = mine[i].x * 0x9E3779B1u;
            x ^= x >> 15;
            x *= 0x85EBCA77u;
            x ^= x >> 13;
            h += x;
        }
        h = __reduce_add_sync(FG_FULL, h) + n * 0xC2B2AE3Du;
        h ^= h >> 16;
        uint32_t slot = h & mask, rep = r;
        for (;;) {
            uint32_t seen = 0;
            if (lane == 0) seen = atomicCAS(slots + slot, 0xffffffffu, r);
            seen = __shfl_sync(FG_FULL, seen, 0);
            if (seen == 0xffffffffu) break; /* claimed: r represents a new list */
            bool same = __ldg(counts + seen) == n;
            if (same) {
                const uint2* theirs = entries_of(seen, n, stage, pool);
                for (uint32_t i = lane; i < n && same; i += 32) same = theirs[i].x == mine[i].x;
            }
            if (__all_sync(FG_FULL, same)) {
                rep = seen;
                break;
            }
            slot = (slot + 1) & mask;
        }
        if (lane == 0) {
            rep_of_read[r] = read_base + rep;
            rep_counts[r] = rep == r ? n : 0;
        }
    }
}

/* counter bits and words per lane for reads of at most max_kmers k-mers */
template <typename F>
static inline void dispatch_table_kernel(int algo, uint32_t max_kmers, F&& f) {
    if (algo == FULGOR_GPU_FULL_INTERSECTION) f(std::true_type(), std::integral_constant<int, 1>(), std::integral_constant<int, 5>());
    else if (max_kmers < (1u << 7)) f(std::false_type(), std::integral_constant<int, 7>(), std::integral_constant<int, 5>());
    else if (max_kmers < (1u << 10)) f(std::false_type(), std::integral_constant<int, 10>(), std::integral_constant<int, 5>());
    else if (max_kmers < (1u << 16)) f(std::false_type(), std::integral_constant<int, 16>(), std::integral_constant<int, 5>());
    else f(std::false_type(), std::integral_constant<int, 32>(), std::integral_constant<int, 1>());
}

/* ---- CSR offsets: exclusive scan of per-read counts (three small kernels) ---- */
#define FG_SCAN_ITEMS 8
#define FG_SCAN_BLOCK 256
#define FG_SCAN_TILE (FG_SCAN_ITEMS * FG_SCAN_BLOCK)

/* MODE 0: the values themselves; 1: their popcounts (color masks -> list lengths); 2: 32-bit words of a packed read of that
   many bases (16 per word, flag bit ignored) */
#define FG_SCAN_PLAIN 0
#define FG_SCAN_POPC 1
#define FG_SCAN_PACKED_WORDS 2
template <int MODE>
__device__ __forceinline__ uint32_t count_of(const uint32_t* __restrict__ in, uint32_t i) {
    const uint32_t v = __ldg(in + i);
    return MODE == FG_SCAN_POPC ? uint32_t(__popc(v)) : (MODE == FG_SCAN_PACKED_WORDS ? ((v & 0x7fffffffu) + 15u) >> 4 : v);
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[FG_SCAN_BLOCK / 32];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(FG_FULL, x, d);
        if (lane >= uint32_t(d)) x += y;
    }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < FG_SCAN_BLOCK / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(FG_FULL, s, d);
            if (lane >= uint32_t(d)) s += y;
        }
        if (lane < FG_SCAN_BLOCK / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    const uint32_t before = w ? warp_sums[w - 1] : 0;
    *total = warp_sums[FG_SCAN_BLOCK / 32 - 1];
    __syncthreads();
    return before + x - v;
}

template <int POPC>
__global__ void __launch_bounds__(FG_SCAN_BLOCK) k_scan_tile_sums(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ tile_sums) {
    const uint32_t base = blockIdx.x * FG_SCAN_TILE + threadIdx.x * FG_SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < FG_SCAN_ITEMS; ++j)
        if (base + j < n) s += count_of<POPC>(in, base + j);
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

/* one block: exclusive scan of the tile sums into 64-bit tile offsets; carries the running CSR
   offset across chunks: carry[0] = running total, chunk_info = {base of this chunk, total of this chunk} */
__global__ void __launch_bounds__(FG_SCAN_BLOCK) k_scan_tile_offsets(const uint32_t* __restrict__ tile_sums, uint32_t n_tiles,
                                                                    uint64_t* __restrict__ tile_off, uint64_t* __restrict__ carry,
                                                                    uint64_t* __restrict__ chunk_info, uint64_t* __restrict__ off_last) {
    __shared__ uint64_t running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < n_tiles; t0 += FG_SCAN_BLOCK) {
        const uint32_t i = t0 + threadIdx.x;
        const uint32_t v = i < n_tiles ? tile_sums[i] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, &total);
        if (i < n_tiles) tile_off[i] = running + ex;
        __syncthreads();
        if (threadIdx.x == 0) running += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint64_t base = carry[0];
        chunk_info[0] = base;
        chunk_info[1] = running;
        carry[0] = base + running;
        *off_last = base + running;
    }
}

template <int POPC>
__global__ void __launch_bounds__(FG_SCAN_BLOCK) k_scan_write(const uint32_t* __restrict__ in, uint32_t n, const uint64_t* __restrict__ tile_off,
                                                             const uint64_t* __restrict__ chunk_info, uint64_t* __restrict__ off) {
    const uint32_t base = blockIdx.x * FG_SCAN_TILE + threadIdx.x * FG_SCAN_ITEMS;
    uint32_t c[FG_SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < FG_SCAN_ITEMS; ++j) {
        c[j] = base + j < n ? count_of<POPC>(in, base + j) : 0;
        s += c[j];
    }
    uint32_t total;
    uint64_t o = chunk_info[0] + tile_off[blockIdx.x] + block_exclusive_scan(s, &total);
#pragma unroll
    for (int j = 0; j < FG_SCAN_ITEMS; ++j) {
        if (base + j < n) off[base + j] = o;
        o += c[j];
    }
}

/* longest read of a packed batch (the host of the device-resident entry point does not know the lengths): sizes the
   threshold-union counters before anything else is enqueued */
__global__ void __launch_bounds__(256) k_max_read_len(const uint32_t* __restrict__ read_len, uint32_t n, uint32_t* __restrict__ out) {
    uint32_t m = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, __ldg(read_len + i) & 0x7fffffffu);
    m = __reduce_max_sync(FG_FULL, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

/* ---- emit ---- */

/* color masks -> ascending color lists at their CSR positions (chunk-local output buffer) */
__global__ void __launch_bounds__(256) k_emit_masks(const uint32_t* __restrict__ masks, const uint64_t* __restrict__ off,
                                                   const uint64_t* __restrict__ chunk_info, uint32_t n, uint32_t* __restrict__ out,
                                                   uint64_t out_cap) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint32_t m = __ldg(masks + r);
    uint64_t o = __ldg(off + r) - chunk_info[0];
    while (m) {
        const uint32_t c = uint32_t(__ffs(int(m))) - 1u;
        m &= m - 1;
        if (o < out_cap) out[o] = c;
        ++o;
    }
}

/* per-read {color-set id, multiplicity} lists (stage or pool) -> CSR of color-set ids */
__global__ void __launch_bounds__(256) k_emit_entries(const uint2* __restrict__ stage, const uint2* __restrict__ pool, const uint32_t* __restrict__ counts,
                                                     const uint64_t* __restrict__ off, const uint64_t* __restrict__ chunk_info, uint32_t n,
                                                     uint32_t* __restrict__ out, uint64_t out_cap) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint32_t c = __ldg(counts + r);
    if (c == 0) return;
    const uint64_t o = __ldg(off + r) - chunk_info[0];
    const uint2* e = entries_of(r, c, stage, pool);
    for (uint32_t i = lane; i < c; i += 32)
        if (o + i < out_cap) out[o + i] = e[i].x;
}

/* per-read color bitmaps -> ascending color lists at their CSR positions. One warp per read; the lanes load 32 words of the row
   at once (one coalesced 128-byte load) and hand them round with register shuffles; word by word, the lane whose bit is set
   writes color 32 w + lane at the running offset + the number of set bits below it: the stores of a word are contiguous.
   Offsets run in 32 bits from the read's base, and the capacity is tested once per read, not per color. */
__global__ void __launch_bounds__(256) k_emit_bits(const uint32_t* __restrict__ res_bits, uint32_t words_per_read, const uint32_t* __restrict__ counts,
                                                  const uint64_t* __restrict__ off, const uint64_t* __restrict__ chunk_info, uint32_t n,
                                                  uint32_t* __restrict__ out, uint64_t out_cap) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint32_t cnt = __ldg(counts + r);
    if (cnt == 0) return;
    const uint64_t o0 = __ldg(off + r) - chunk_info[0];
    if (o0 >= out_cap) return;
    const uint32_t room = uint32_t(min(out_cap - o0, uint64_t(cnt))); /* < cnt only when the caller's buffer is too small (E2BIG) */
    uint32_t* dst = out + o0;
    const uint32_t* bits = res_bits + uint64_t(r) * words_per_read;
    const uint32_t below = (1u << lane) - 1u;
    uint32_t o = 0;
    for (uint32_t w0 = 0; w0 < words_per_read; w0 += 32) {
        const uint32_t mine = w0 + lane < words_per_read ? __ldg(bits + w0 + lane) : 0u;
        if (__ballot_sync(FG_FULL, mine != 0) == 0) continue;
        const uint32_t color0 = 32 * w0 + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const uint32_t word = __shfl_sync(FG_FULL, mine, j);
            const uint32_t at = o + __popc(word & below);
            if (((word >> lane) & 1u) && at < room) dst[at] = color0 + 32 * j;
            o += __popc(word);
        }
    }
}

}  // namespace fgb
#endif
