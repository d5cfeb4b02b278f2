/*
 * host_emul.cu -- TEST INFRASTRUCTURE ONLY. Compiles the per-lane device functions of
 * fulgor_b200/csrc/kernels.cuh (they are __host__ __device__) for the HOST, so that the flattened
 * image and the lookup / decode arithmetic can be checked against the oracle in the CPU-only test
 * tier. Nothing in the product links or calls this; it is built into build/libfg_host_emul.so by
 * tests/_checkers.py.
 */
#include <cstring>
#include "../fulgor_b200/csrc/kernels.cuh"

using namespace fgb;

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
    I.type = H.type;
    I.num_colors = H.num_colors;
    I.num_partitions = H.num_partitions;
    I.main_seed = I.phfs[0].seed;
    I.main_nparts = I.phfs[0].num_partitions;
    I.main_part = I.parts[0];
    return I;
}

extern "C" {

/* color-set id of every k-mer of one read (0xffffffff = negative / invalid), each k-mer looked up
   independently exactly like one GPU lane does */
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
}
