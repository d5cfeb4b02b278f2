/*
 * simt_emul.cpp -- TEST INFRASTRUCTURE ONLY. Compiles the CUDA kernels of fulgor_b200/csrc for the HOST on top of the
 * lock-step warp emulator in simt_emul.h and drives them in the same order as fulgor_b200/csrc/engine.cu does on the GPU
 * (K1 [+K2] -> scan -> emit), so that the CPU-only test tier checks the kernels' logic -- not just per-lane helpers --
 * against the oracle. Built into build/libfg_simt_emul.so by tests/test_host_logic.py; nothing in the product links it.
 */
#include "simt_emul.h"

#include "../fulgor_b200/csrc/pack_reads.h"
#include "../fulgor_b200/csrc/pipeline_kernels.cuh"

using namespace fgb;

/* emul_set_packed(1): every entry point below first packs its ASCII reads (pack_reads.h, the host half of the packed form) and
   runs the lookup kernels on packed_reads, word offsets from the scan kernels like engine.cu; 0: ascii_reads */
static int g_packed = 0;

template <int MODE>
static void run_scan_into(const uint32_t* counts, uint32_t n, uint64_t* off);

struct emul_reads {
    const uint8_t* bases;
    const uint64_t* read_off;
    uint32_t n;
    std::vector<uint32_t> words, read_len;
    std::vector<uint64_t> invalid, word_off;
    emul_reads(const uint8_t* b, const uint64_t* ro, uint32_t n_) : bases(b), read_off(ro), n(n_) {
        if (!g_packed || n == 0) return;
        words.assign(packed_words_of(ro, n) + 1, 0xdeadbeefu);
        read_len.assign(n, 0);
        pack_reads(reinterpret_cast<const char*>(b), ro, n, words.data(), read_len.data(), invalid, 3);
        word_off.assign(size_t(n) + 1, 0);
        run_scan_into<FG_SCAN_PACKED_WORDS>(read_len.data(), n, word_off.data());
    }
    template <typename F>
    void with(F&& f) const {
        if (g_packed) f(packed_reads{words.data(), word_off.data(), read_len.data(), invalid.data(), uint32_t(invalid.size()), 0});
        else f(ascii_reads{bases, read_off, read_off[0]});
    }
};

static dev_index view_of(const uint8_t* base) {
    fgi_header H;
    std::memcpy(&H, base, sizeof(H));
    dev_index I{};
    I.phfs = reinterpret_cast<const fgi_phf*>(base + H.off_phfs);
    I.parts = reinterpret_cast<const fgi_phf_part*>(base + H.off_phf_parts);
    I.hashed_pilots = reinterpret_cast<const uint64_t*>(base + H.off_hashed_pilots);
    I.free_slots = reinterpret_cast<const uint32_t*>(base + H.off_free_slots);
    I.bucket_begin = reinterpret_cast<const uint32_t*>(base + H.off_bucket_begin);
    I.sk_records = reinterpret_cast<const uint2*>(base + H.off_sk_records);
    I.sk_cid = H.off_sk_cid ? reinterpret_cast<const uint32_t*>(base + H.off_sk_cid) : nullptr;
    I.strings = reinterpret_cast<const uint64_t*>(base + H.off_strings);
    I.skew_positions = reinterpret_cast<const uint32_t*>(base + H.off_skew_positions);
    I.hybrids = reinterpret_cast<const fgi_hybrid*>(base + H.off_hybrids);
    I.set_bit_off = reinterpret_cast<const uint64_t*>(base + H.off_set_bit_off);
    I.color_words = reinterpret_cast<const uint64_t*>(base + H.off_color_words);
    I.meta_off = reinterpret_cast<const uint64_t*>(base + H.off_meta_off);
    I.meta_vals = reinterpret_cast<const uint32_t*>(base + H.off_meta_vals);
    I.part_min_color = reinterpret_cast<const uint32_t*>(base + H.off_part_min_color);
    I.part_sets_before = reinterpret_cast<const uint32_t*>(base + H.off_part_sets_before);
    I.hash_magic = H.hash_magic;
    I.bucketer_T = H.bucketer_T;
    I.k = H.k;
    I.m = H.m;
    I.skew_min_log2 = H.skew_min_log2;
    I.skew_max_log2 = H.skew_max_log2;
    I.skew_log2_max_bucket = H.skew_log2_max_bucket;
    I.num_skew = H.num_skew;
    I.skew_threshold = H.num_skew ? (1u << H.skew_min_log2) : UINT32_MAX;
    I.guard_max_hash = uint32_t(H.guard_max_hash);
    for (int i = 0; i < FGI_MAX_SKEW; ++i) {
        I.skew_phf[i] = H.skew_phf[i];
        I.skew_pos_base[i] = H.skew_pos_base[i];
    }
    I.type = H.type & 1u;
    I.diff = (H.type >> 1) & 1u;
    I.num_colors = H.num_colors;
    I.num_partitions = H.num_partitions;
    I.main_seed = I.phfs[0].seed;
    I.main_nparts = I.phfs[0].num_partitions;
    I.main_part = I.parts[0];
    return I;
}

template <typename F>
static void dispatch_window(const dev_index& I, int force_generic, F&& f) {
    switch (force_generic || I.guard_max_hash || I.m < 16 ? 0 : int(I.k - I.m + 1)) { /* like engine.cu */
        case 13: f(std::integral_constant<int, 13>()); break;
        case 12: f(std::integral_constant<int, 12>()); break;
        case 11: f(std::integral_constant<int, 11>()); break;
        default: f(std::integral_constant<int, 0>()); break;
    }
}

template <int POPC>
static void run_scan_into(const uint32_t* counts, uint32_t n, uint64_t* off) {
    const uint32_t tiles = (n + FG_SCAN_TILE - 1) / FG_SCAN_TILE;
    std::vector<uint32_t> tile_sums(tiles);
    std::vector<uint64_t> tile_off(tiles);
    uint64_t carry = 0, chunk_info[2] = {0, 0};
    simt::launch(tiles, FG_SCAN_BLOCK, 0, [&] { k_scan_tile_sums<POPC>(counts, n, tile_sums.data()); });
    simt::launch(1, FG_SCAN_BLOCK, 0, [&] { k_scan_tile_offsets(tile_sums.data(), tiles, tile_off.data(), &carry, chunk_info, off + n); });
    simt::launch(tiles, FG_SCAN_BLOCK, 0, [&] { k_scan_write<POPC>(counts, n, tile_off.data(), chunk_info, off); });
}

template <bool POPC>
static void run_scan(const uint32_t* counts, uint32_t n, uint64_t* off) {
    run_scan_into<POPC ? FG_SCAN_POPC : FG_SCAN_PLAIN>(counts, n, off);
}

struct k1_out {
    std::vector<uint2> stage, pool;
    std::vector<uint32_t> counts, npos;
};

static int run_k1(const dev_index& I, const emul_reads& rd, uint32_t n, unsigned grid, int force_generic, uint64_t pool_entries, k1_out& o, bool sorted = true) {
    o.stage.assign(size_t(n) * FG_STAGE_STRIDE, uint2{0, 0});
    o.pool.assign(pool_entries, uint2{0, 0});
    o.counts.assign(n, 0);
    o.npos.assign(n, 0);
    unsigned long long used = 0;
    uint32_t exhausted = 0;
    entry_pool pool{o.pool.data(), &used, pool_entries, &exhausted};
    dispatch_window(I, force_generic, [&](auto w) {
        rd.with([&](auto in) {
            simt::launch(grid, FG_BLOCK, 0, [&] {
                if (sorted) k_fetch_color_sets<decltype(w)::value, decltype(in), true>(I, in, n, o.stage.data(), o.counts.data(), o.npos.data(), pool, nullptr);
                else k_fetch_color_sets<decltype(w)::value, decltype(in), false>(I, in, n, o.stage.data(), o.counts.data(), o.npos.data(), pool, nullptr);
            });
        });
    });
    return int(exhausted);
}

extern "C" {

void emul_set_packed(int on) { g_packed = on; }

/* color-set id of every k-mer of one read (0xffffffff = negative / invalid), each k-mer looked up independently with the
   per-lane functions (the specification the warp pipeline must agree with) */
void emul_lookup_read(const uint8_t* image, const char* seq, uint64_t len, uint32_t* cids) {
    const dev_index I = view_of(image);
    const uint32_t k = I.k;
    if (len < k) return;
    const uint64_t kmask = (1ULL << (2 * k)) - 1;
    for (uint64_t i = 0; i + k <= len; ++i) {
        bool valid = true;
        uint64_t fwd = 0;
        for (uint32_t j = 0; j < k; ++j) {
            const uint32_t c = uint8_t(seq[i + j]);
            valid &= base_valid(c);
            fwd |= uint64_t((c >> 1) & 3u) << (2 * j);
        }
        if (!valid) {
            cids[i] = FG_NOT_FOUND;
            continue;
        }
        const uint64_t rc = revcomp(fwd, k);
        const minimizer_t mz = canonical_minimizer(fwd, rc, k, I.m, I.hash_magic);
        cids[i] = lookup_color_set(I, fwd, rc, mz, kmask);
    }
}

/* decode one color set of an index with <= 32 colors as a mask */
uint32_t emul_color_set_mask(const uint8_t* image, uint32_t cid) {
    const dev_index I = view_of(image);
    return color_set_mask(I, cid);
}

int emul_base_valid(uint32_t c) { return base_valid(c) ? 1 : 0; }

/* the whole pseudoalignment pipeline on emulated warps. Returns 0, or -7 when cap is too small (out_off complete). */
int emul_pseudoalign(const uint8_t* image, int algo, double threshold, const uint8_t* bases, const uint64_t* read_off, uint32_t n,
                     uint64_t* out_off, uint32_t* out_vals, uint64_t cap, unsigned grid, int force_generic, int use_table) {
    dev_index I = view_of(image);
    std::vector<uint32_t> table;
    if (use_table) { /* decode every color set once, like fulgor_gpu_index_open does on the device */
        fgi_header H;
        std::memcpy(&H, image, sizeof(H));
        const uint64_t stride = table_stride_words(I.num_colors);
        table.assign(H.num_color_sets * stride, 0xdeadbeefu);
        simt::launch(2, 64, 2 * stride * 4, [&] { k_expand_color_sets(I, 0, uint32_t(H.num_color_sets), uint32_t(stride), table.data()); });
        I.set_table = table.data();
        I.table_stride = stride;
    }
    out_off[0] = 0;
    if (n == 0) return 0;
    const emul_reads rd(bases, read_off, n);
    uint64_t chunk_info[2] = {0, 0};
    if (I.num_colors <= 32) {
        std::vector<uint32_t> masks(n, 0xdeadbeefu);
        dispatch_window(I, force_generic, [&](auto w) {
            rd.with([&](auto in) {
                simt::launch(grid, FG_BLOCK, 0, [&] { k_pseudoalign_small<decltype(w)::value, decltype(in)>(I, in, n, algo, threshold, masks.data()); });
            });
        });
        run_scan<true>(masks.data(), n, out_off);
        if (out_off[n] > cap) return FULGOR_GPU_E2BIG;
        simt::launch((n + 255) / 256, 256, 0, [&] { k_emit_masks(masks.data(), out_off, chunk_info, n, out_vals, cap); });
        return 0;
    }
    k1_out k1;
    uint64_t pool_entries = 1u << 12; /* small on purpose: exercises the grow-and-rerun path */
    while (run_k1(I, rd, n, grid, force_generic, pool_entries, k1, /*sorted=*/false)) pool_entries *= 4; /* like engine.cu: the color-set kernels take any order */
    uint32_t max_kmers = 1;
    for (uint32_t i = 0; i < n; ++i) max_kmers = std::max<uint32_t>(max_kmers, uint32_t(read_off[i + 1] - read_off[i]));
    const general_plan g = plan_color_sets_general(I.num_colors, I.num_partitions, algo, max_kmers, I.diff != 0);
    if (!g.ok) return FULGOR_GPU_EINVAL;
    std::vector<uint32_t> res_bits(size_t(n) * g.words_per_read), res_counts(n);
    if (use_table) {
        dispatch_table_kernel(algo, max_kmers, [&](auto fi, auto np, auto t) {
            simt::launch(grid, FG_BLOCK, table_kernel_smem(decltype(fi)::value, decltype(np)::value, decltype(t)::value), [&] {
                k_color_sets_table<decltype(fi)::value, decltype(np)::value, decltype(t)::value>(
                    I, k1.counts.data(), k1.stage.data(), k1.pool.data(), k1.npos.data(), n, threshold, g.words_per_read, res_bits.data(), res_counts.data());
            });
        });
    } else {
        simt::launch(grid, g.warps_per_block * 32, g.smem_bytes, [&] {
            k_color_sets_general(I, k1.counts.data(), k1.stage.data(), k1.pool.data(), k1.npos.data(), n, algo, threshold, g.words_per_read, g.planes,
                                 g.ints_per_warp, res_bits.data(), res_counts.data());
        });
    }
    run_scan<false>(res_counts.data(), n, out_off);
    if (out_off[n] > cap) return FULGOR_GPU_E2BIG;
    simt::launch(uint32_t((uint64_t(n) * 32 + 255) / 256), 256, 0,
                 [&] { k_emit_bits(res_bits.data(), g.words_per_read, res_counts.data(), out_off, chunk_info, n, out_vals, cap); });
    return 0;
}

/* full intersection once per distinct color-set-id list (engine.cu: enqueue_dedup): K1 -> k_group_reads -> color-set kernel on
   the representatives -> scan -> emit. rep_of_read[i] = representative of read i. */
int emul_pseudoalign_dedup(const uint8_t* image, const uint8_t* bases, const uint64_t* read_off, uint32_t n, uint32_t* rep_of_read,
                           uint64_t* out_off, uint32_t* out_vals, uint64_t cap, unsigned grid, int use_table) {
    dev_index I = view_of(image);
    std::vector<uint32_t> table;
    if (use_table) {
        fgi_header H;
        std::memcpy(&H, image, sizeof(H));
        const uint64_t stride = table_stride_words(I.num_colors);
        table.assign(H.num_color_sets * stride, 0xdeadbeefu);
        simt::launch(2, 64, 2 * stride * 4, [&] { k_expand_color_sets(I, 0, uint32_t(H.num_color_sets), uint32_t(stride), table.data()); });
        I.set_table = table.data();
        I.table_stride = stride;
    }
    out_off[0] = 0;
    if (n == 0) return 0;
    uint64_t chunk_info[2] = {0, 0};
    k1_out k1;
    uint64_t pool_entries = 1u << 12;
    const emul_reads rd(bases, read_off, n);
    while (run_k1(I, rd, n, grid, 0, pool_entries, k1)) pool_entries *= 4;
    uint32_t log2_slots = 4; /* small on purpose: long probe sequences */
    while ((1ull << log2_slots) < 2ull * n) ++log2_slots;
    std::vector<uint32_t> slots(size_t(1) << log2_slots, 0xffffffffu), rep_counts(n, 0xdeadbeefu);
    simt::launch(grid, FG_BLOCK, 0, [&] {
        k_group_reads(k1.counts.data(), k1.stage.data(), k1.pool.data(), n, 0, slots.data(), log2_slots, rep_of_read, rep_counts.data());
    });
    const int algo = FULGOR_GPU_FULL_INTERSECTION;
    uint32_t max_kmers = 1;
    for (uint32_t i = 0; i < n; ++i) max_kmers = std::max<uint32_t>(max_kmers, uint32_t(read_off[i + 1] - read_off[i]));
    const general_plan g = plan_color_sets_general(I.num_colors, I.num_partitions, algo, max_kmers, I.diff != 0);
    if (!g.ok) return FULGOR_GPU_EINVAL;
    std::vector<uint32_t> res_bits(size_t(n) * g.words_per_read), res_counts(n);
    if (use_table) {
        dispatch_table_kernel(algo, max_kmers, [&](auto fi, auto np, auto t) {
            simt::launch(grid, FG_BLOCK, table_kernel_smem(decltype(fi)::value, decltype(np)::value, decltype(t)::value), [&] {
                k_color_sets_table<decltype(fi)::value, decltype(np)::value, decltype(t)::value>(
                    I, rep_counts.data(), k1.stage.data(), k1.pool.data(), k1.npos.data(), n, 1.0, g.words_per_read, res_bits.data(), res_counts.data());
            });
        });
    } else {
        simt::launch(grid, g.warps_per_block * 32, g.smem_bytes, [&] {
            k_color_sets_general(I, rep_counts.data(), k1.stage.data(), k1.pool.data(), k1.npos.data(), n, algo, 1.0, g.words_per_read, g.planes,
                                 g.ints_per_warp, res_bits.data(), res_counts.data());
        });
    }
    run_scan<false>(res_counts.data(), n, out_off);
    if (out_off[n] > cap) return FULGOR_GPU_E2BIG;
    simt::launch(uint32_t((uint64_t(n) * 32 + 255) / 256), 256, 0,
                 [&] { k_emit_bits(res_bits.data(), g.words_per_read, res_counts.data(), out_off, chunk_info, n, out_vals, cap); });
    return 0;
}

/* the per-k-mer tools (engine.cu: run_kmer_tool) on emulated warps. which = 0: kmer-conservation -> out_off = triple offsets,
   out_vals = triples (3 each, cap in triples); which = 1: kmer-matches -> out_off = word offsets, out_vals = positive words,
   counts = n x num_colors (uses the decoded table). */
int emul_kmer_tool(const uint8_t* image, int which, const uint8_t* bases, const uint64_t* read_off, uint32_t n, uint64_t* out_off,
                   uint32_t* out_vals, uint64_t cap, uint32_t* counts, unsigned grid, int force_generic) {
    dev_index I = view_of(image);
    out_off[0] = 0;
    if (n == 0) return 0;
    std::vector<uint64_t> koff(size_t(n) + 1, 0);
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t len = read_off[i + 1] - read_off[i];
        koff[i + 1] = koff[i] + (len >= I.k ? len - I.k + 1 : 0);
    }
    std::vector<uint32_t> per_kmer(koff[n] + 1, 0xdeadbeefu);
    const emul_reads rd(bases, read_off, n);
    dispatch_window(I, force_generic, [&](auto w) {
        rd.with([&](auto in) {
            simt::launch(grid, FG_BLOCK, 0, [&] { k_kmer_color_sets<decltype(w)::value, decltype(in)>(I, in, n, koff.data(), per_kmer.data()); });
        });
    });
    const uint32_t warp_grid = uint32_t((uint64_t(n) * 32 + 255) / 256);
    uint64_t chunk_info[2] = {0, 0};
    if (which == 0) {
        std::vector<uint32_t> run_counts(n, 0xdeadbeefu);
        simt::launch(warp_grid, 256, 0, [&] { k_kmer_runs<false>(per_kmer.data(), koff.data(), n, run_counts.data(), nullptr, nullptr, nullptr, 0); });
        run_scan<false>(run_counts.data(), n, out_off);
        if (out_off[n] > cap) return FULGOR_GPU_E2BIG;
        simt::launch(warp_grid, 256, 0, [&] { k_kmer_runs<true>(per_kmer.data(), koff.data(), n, nullptr, out_off, chunk_info, out_vals, cap); });
        return 0;
    }
    fgi_header H;
    std::memcpy(&H, image, sizeof(H));
    const uint64_t stride = table_stride_words(I.num_colors);
    std::vector<uint32_t> table(H.num_color_sets * stride, 0xdeadbeefu);
    simt::launch(2, 64, 2 * stride * 4, [&] { k_expand_color_sets(I, 0, uint32_t(H.num_color_sets), uint32_t(stride), table.data()); });
    I.set_table = table.data();
    I.table_stride = stride;
    for (uint32_t i = 0; i < n; ++i) out_off[i + 1] = out_off[i] + (koff[i + 1] - koff[i] + 31) / 32;
    if (out_off[n] > cap) return FULGOR_GPU_E2BIG;
    k1_out k1;
    uint64_t pool_entries = 1u << 12;
    while (run_k1(I, rd, n, grid, force_generic, pool_entries, k1)) pool_entries *= 4;
    simt::launch(warp_grid, 256, 0, [&] { k_kmer_positive_bits(per_kmer.data(), koff.data(), out_off, n, out_vals); });
    simt::launch(grid, FG_BLOCK, 0, [&] { k_kmer_match_counts(I, k1.counts.data(), k1.stage.data(), k1.pool.data(), n, counts); });
    return 0;
}

/* stage 1 alone: per read the ascending distinct color-set ids (+ number of positive k-mers) */
int emul_fetch_color_set_ids(const uint8_t* image, const uint8_t* bases, const uint64_t* read_off, uint32_t n, uint64_t* out_off,
                             uint32_t* out_vals, uint64_t cap, uint32_t* num_positive, unsigned grid, int force_generic) {
    const dev_index I = view_of(image);
    out_off[0] = 0;
    if (n == 0) return 0;
    k1_out k1;
    uint64_t pool_entries = 1u << 12;
    const emul_reads rd(bases, read_off, n);
    while (run_k1(I, rd, n, grid, force_generic, pool_entries, k1)) pool_entries *= 4;
    uint64_t chunk_info[2] = {0, 0};
    run_scan<false>(k1.counts.data(), n, out_off);
    if (num_positive) std::memcpy(num_positive, k1.npos.data(), size_t(n) * 4);
    if (out_off[n] > cap) return FULGOR_GPU_E2BIG;
    simt::launch(uint32_t((uint64_t(n) * 32 + 255) / 256), 256, 0,
                 [&] { k_emit_entries(k1.stage.data(), k1.pool.data(), k1.counts.data(), out_off, chunk_info, n, out_vals, cap); });
    return 0;
}
}
