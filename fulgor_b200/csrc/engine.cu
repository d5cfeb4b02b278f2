/*
 * engine.cu -- kernels' global entry points, the chunked H2D / compute / D2H pipeline and the C ABI
 * declared in include/fulgor_gpu.h. sm_100a only.
 *
 * Per chunk of reads (host API) the stream order is
 *     H2D(bases, read_off) -> K1(+K2 fused when num_colors <= 32) -> [K2] -> scan -> emit -> D2H
 * with four slots so that the copies of one chunk overlap the kernels of the others. The only
 * cross-chunk dependency is the running CSR offset, carried in device memory.
 */
#include <cuda_runtime.h>
#include <sched.h>

#include <cctype>

#include <algorithm>
#include <cstdio>
#include <deque>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fulgor_gpu.h"
#include "fur_reader.h"
#include "pack_reads.h"
#include "pipeline_kernels.cuh"

namespace fgb {

/* ================================================================== host side */

static thread_local std::string g_error;
static int fail(int code, std::string msg) {
    g_error = std::move(msg);
    return code;
}
#define FG_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_));               \
    } while (0)

struct dev_buffer {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) FG_CUDA(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        FG_CUDA(cudaMalloc(&p, want));
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return static_cast<T*>(p); }
};

#define FG_NUM_SLOTS 4     /* chunks in flight in the host pipeline: copy-in, compute, copy-out + one of slack (see FG_FINALIZE_LAG) */
#define FG_FINALIZE_LAG 2  /* the host reads a chunk's totals (and enqueues its values copy) two chunks behind the one it is
                              enqueueing, so that wait never holds back the next host->device copy: the copy engine, the
                              bottleneck of the host-buffer path, always has the next chunk queued */
struct slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t scanned = nullptr, done = nullptr, info_ready = nullptr;
    dev_buffer bases, read_off, per_read /* masks or counts */, stage, pool, npos, res_bits, res_counts, tile_sums, tile_off, off, out, group_slots, rep, rep_counts, kmer_off, per_kmer, word_off;
    dev_buffer read_len, invalid, pk_word_off; /* packed reads: lengths, invalid positions, first word of every read (device scan) */
    uint64_t* pk_scan = nullptr;           /* device: {carry, chunk base, chunk total} of the word-offset scan */
    uint64_t* chunk_info = nullptr;        /* device: {base, total} */
    uint64_t* h_info = nullptr;            /* pinned host: {base, total, pool exhausted} */
    uint32_t* exhausted = nullptr;         /* device flag: the entry pool was too small */
    uint32_t* max_positive = nullptr;      /* device: largest number of positive k-mers of a read (sizes the color-set kernel's counters) */
    unsigned long long* pool_used = nullptr; /* device: entries handed out */
    bool busy = false;
};

}  // namespace fgb

using namespace fgb;

struct fulgor_gpu_index {
    int device = -1;
    fgi_header H{};
    void* d_image = nullptr;
    bool owns_image = false;
    dev_index I{};
    slot slots[FG_NUM_SLOTS];
    uint64_t* d_carry = nullptr;
    int sm_count = 0;
    uint64_t pool_per_read = 8; /* entry-pool size per read of a chunk; grows (x4) when a launch exhausts it */
    uint32_t* d_table = nullptr; /* decoded color-set table (dev_index::set_table), owned */
    /* deduplication keeps the lists of EVERY read of a call resident until the groups are known (call-wide arrays) */
    fgb::dev_buffer dd_stage, dd_counts, dd_npos, dd_pool, dd_rep, dd_rep_counts, dd_slots;
    /* timing of the last *_device call */
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float last_ms[3] = {0, 0, 0};
    int last_launches = 0;
};

namespace fgb {

static void check_header(const fgi_header& H, uint64_t bytes) {
    if (bytes < sizeof(fgi_header) || H.magic != FGI_MAGIC || H.total_bytes != bytes)
        throw std::runtime_error("not a fulgor-b200 device image (bad magic or size)");
}

static void fill_info(const fgi_header& H, int device, fulgor_gpu_info* out) {
    std::memset(out, 0, sizeof(*out));
    out->k = H.k;
    out->m = H.m;
    out->num_kmers = H.num_kmers;
    out->num_unitigs = H.num_unitigs;
    out->num_color_sets = H.num_color_sets;
    out->num_colors = H.num_colors;
    out->type = H.type;
    out->image_bytes = H.total_bytes;
    out->device = device;
}

static dev_index make_view(const fgi_header& H, const uint8_t* base) {
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
    I.main_seed = 0;
    I.main_nparts = 0;
    return I;
}

static void use_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) throw std::invalid_argument(std::string("no usable CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) throw std::invalid_argument("CUDA device ordinal out of range");
    FG_CUDA(cudaSetDevice(device));
}

/* Decodes every color set into a bitmap row once (k_expand_color_sets) when the table fits the budget: FULGOR_GPU_TABLE_MAX_MB
   (default 65536) and half of the free device memory. Without the table the color-set kernels decode the compressed sets per read. */
static void build_color_set_table(fulgor_gpu_index* x) {
    const char* env = std::getenv("FULGOR_GPU_TABLE_MAX_MB");
    const uint64_t budget = (env ? std::strtoull(env, nullptr, 10) : 65536ull) << 20;
    const uint64_t stride = table_stride_words(x->H.num_colors);
    const uint64_t bytes = x->H.num_color_sets * stride * 4;
    size_t free_b = 0, total_b = 0;
    FG_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (bytes == 0 || bytes > budget || bytes > free_b / 2 || stride * 4 > 48 * 1024) return;
    FG_CUDA(cudaMalloc(&x->d_table, bytes));
    const uint32_t wpb = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(FG_WARPS_PER_BLOCK, (48 * 1024) / (stride * 4))));
    const uint64_t per_launch = 1u << 22;
    for (uint64_t first = 0; first < x->H.num_color_sets; first += per_launch) {
        const uint32_t n = uint32_t(std::min<uint64_t>(per_launch, x->H.num_color_sets - first));
        const uint32_t grid = uint32_t(std::min<uint64_t>((n + wpb - 1) / wpb, uint64_t(x->sm_count) * 16));
        k_expand_color_sets<<<grid, wpb * 32, size_t(wpb) * stride * 4, x->slots[0].stream>>>(x->I, first, n, uint32_t(stride), x->d_table + first * stride);
        FG_CUDA(cudaGetLastError());
    }
    FG_CUDA(cudaStreamSynchronize(x->slots[0].stream));
    x->I.set_table = x->d_table;
    x->I.table_stride = stride;
}

static fulgor_gpu_index* make_handle(const fgi_header& H, void* d_image, bool owns, int device) {
    auto* x = new fulgor_gpu_index();
    x->device = device;
    x->H = H;
    x->d_image = d_image;
    x->owns_image = owns;
    x->I = make_view(H, static_cast<const uint8_t*>(d_image));
    try {
        /* the minimizer MPHF's descriptor goes into the kernel parameters */
        fgi_phf main_phf;
        FG_CUDA(cudaMemcpy(&main_phf, static_cast<const uint8_t*>(d_image) + H.off_phfs, sizeof(main_phf), cudaMemcpyDeviceToHost));
        if (main_phf.first_part != 0 || main_phf.num_partitions == 0) throw std::runtime_error("unexpected minimizer MPHF layout in the image");
        FG_CUDA(cudaMemcpy(&x->I.main_part, static_cast<const uint8_t*>(d_image) + H.off_phf_parts, sizeof(fgi_phf_part), cudaMemcpyDeviceToHost));
        x->I.main_seed = main_phf.seed;
        x->I.main_nparts = main_phf.num_partitions;
        cudaDeviceProp prop;
        FG_CUDA(cudaGetDeviceProperties(&prop, device));
        x->sm_count = prop.multiProcessorCount;
        FG_CUDA(cudaMalloc(&x->d_carry, 8));
        for (auto& s : x->slots) FG_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        build_color_set_table(x); /* on slots[0].stream */
        for (auto& s : x->slots) {
            FG_CUDA(cudaEventCreateWithFlags(&s.scanned, cudaEventDisableTiming));
            FG_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
            FG_CUDA(cudaEventCreateWithFlags(&s.info_ready, cudaEventDisableTiming));
            FG_CUDA(cudaMalloc(&s.chunk_info, 16));
            FG_CUDA(cudaMalloc(&s.exhausted, 4));
            FG_CUDA(cudaMalloc(&s.max_positive, 4));
            FG_CUDA(cudaMalloc(&s.pool_used, 8));
            FG_CUDA(cudaMalloc(&s.pk_scan, 24));
            FG_CUDA(cudaHostAlloc(&s.h_info, 32, cudaHostAllocDefault));
            std::memset(s.h_info, 0, 32);
        }
        for (auto& e : x->ev) FG_CUDA(cudaEventCreate(&e));
    } catch (...) {
        fulgor_gpu_index_close(x);
        throw;
    }
    return x;
}

/* ---- launch helpers ---- */

template <typename F>
static void dispatch_window(const fgi_header& H, F&& f) {
    /* common (k, m) pairs get unrolled minimizer loops on 32-bit keys; an index in which some m-mer hashes to UINT64_MAX
       (compute_minimizer's sentinel, image.h: guard_max_hash) takes the generic window, which knows about it, and so do
       minimizers shorter than 16 bases (the templated code takes the low word of the m-mer mask for all ones) */
    switch (H.guard_max_hash || H.m < 16 ? 0 : H.k - H.m + 1) {
        case 13: f(std::integral_constant<int, 13>()); break; /* k=31, m=19 */
        case 12: f(std::integral_constant<int, 12>()); break; /* k=31, m=20 */
        case 11: f(std::integral_constant<int, 11>()); break; /* k=31, m=21 */
        default: f(std::integral_constant<int, 0>()); break;
    }
}

static uint32_t read_grid(const fulgor_gpu_index* x, uint32_t n_reads) {
    const uint64_t blocks_needed = (uint64_t(n_reads) + FG_WARPS_PER_BLOCK - 1) / FG_WARPS_PER_BLOCK;
    const uint64_t resident = uint64_t(x->sm_count) * FG_MIN_BLOCKS * 2; /* a multiple of the SM count; warps stride over reads */
    return uint32_t(std::max<uint64_t>(1, std::min(blocks_needed, resident)));
}

struct chunk_args {
    const uint8_t* d_bases;
    const uint64_t* d_read_off;
    uint64_t read_off_base;
    uint32_t n;
    uint32_t max_len; /* longest read of the chunk when the host knows it, else 0 */
    /* packed reads (d_words != nullptr) instead of d_bases / d_read_off */
    const uint32_t* d_words = nullptr;
    const uint64_t* d_word_off = nullptr;
    const uint32_t* d_read_len = nullptr;
    const uint64_t* d_invalid = nullptr;
    uint32_t n_invalid = 0;
    uint64_t pos_base = 0;
};

/* calls f with the chunk's reads in the form the lookup kernels are templated on */
template <typename F>
static void with_reads(const chunk_args& a, F&& f) {
    if (a.d_words) f(packed_reads{a.d_words, a.d_word_off, a.d_read_len, a.d_invalid, a.n_invalid, a.pos_base});
    else f(ascii_reads{a.d_bases, a.d_read_off, a.read_off_base});
}

/* enqueue scan (counts -> CSR offsets with the cross-chunk carry) */
template <int POPC>
static int enqueue_scan(fulgor_gpu_index* x, slot& s, const uint32_t* d_counts, uint32_t n, uint64_t* d_off) {
    const uint32_t tiles = (n + FG_SCAN_TILE - 1) / FG_SCAN_TILE;
    s.tile_sums.reserve(size_t(tiles) * 4);
    s.tile_off.reserve(size_t(tiles) * 8);
    k_scan_tile_sums<POPC><<<tiles, FG_SCAN_BLOCK, 0, s.stream>>>(d_counts, n, s.tile_sums.as<uint32_t>());
    k_scan_tile_offsets<<<1, FG_SCAN_BLOCK, 0, s.stream>>>(s.tile_sums.as<uint32_t>(), tiles, s.tile_off.as<uint64_t>(), x->d_carry,
                                                          s.chunk_info, d_off + n);
    k_scan_write<POPC><<<tiles, FG_SCAN_BLOCK, 0, s.stream>>>(d_counts, n, s.tile_off.as<uint64_t>(), s.chunk_info, d_off);
    return 3;
}

/* how a chunk's per-read results become CSR values; kept so that the emit can be repeated into a larger buffer */
struct emit_plan {
    enum kind_t { MASKS, ENTRIES, BITS } kind = MASKS;
    uint32_t n = 0, words_per_read = 0;
};

static void enqueue_emit(slot& s, const emit_plan& e, const uint64_t* d_off, uint32_t* d_out, uint64_t out_cap) {
    if (e.n == 0) return;
    switch (e.kind) {
        case emit_plan::MASKS:
            k_emit_masks<<<(e.n + 255) / 256, 256, 0, s.stream>>>(s.per_read.as<uint32_t>(), d_off, s.chunk_info, e.n, d_out, out_cap);
            break;
        case emit_plan::ENTRIES:
            k_emit_entries<<<uint32_t((uint64_t(e.n) * 32 + 255) / 256), 256, 0, s.stream>>>(s.stage.as<uint2>(), s.pool.as<uint2>(), s.per_read.as<uint32_t>(),
                                                                                          d_off, s.chunk_info, e.n, d_out, out_cap);
            break;
        case emit_plan::BITS:
            k_emit_bits<<<uint32_t((uint64_t(e.n) * 32 + 255) / 256), 256, 0, s.stream>>>(s.res_bits.as<uint32_t>(), e.words_per_read, s.res_counts.as<uint32_t>(),
                                                                                       d_off, s.chunk_info, e.n, d_out, out_cap);
            break;
    }
    FG_CUDA(cudaGetLastError());
}

/* where K1 leaves a chunk's per-read lists: the chunk's own slot buffers, or slices of call-wide arrays (deduplication) */
struct k1_target {
    uint2* stage;
    uint32_t* counts;
    uint32_t* npos; /* nullable */
    entry_pool pool;
};

/* sorted: the lists come out ascending by id (stage 1 results, deduplication); the color-set kernels do not need that */
static int launch_k1(fulgor_gpu_index* x, slot& s, const chunk_args& a, const k1_target& t, bool sorted) {
    const uint32_t grid = read_grid(x, a.n);
    dispatch_window(x->H, [&](auto w) {
        with_reads(a, [&](auto in) {
            if (sorted)
                k_fetch_color_sets<decltype(w)::value, decltype(in), true><<<grid, FG_BLOCK, 0, s.stream>>>(x->I, in, a.n, t.stage, t.counts, t.npos, t.pool,
                                                                                                              a.max_len ? nullptr : s.max_positive);
            else
                k_fetch_color_sets<decltype(w)::value, decltype(in), false><<<grid, FG_BLOCK, 0, s.stream>>>(x->I, in, a.n, t.stage, t.counts, t.npos, t.pool,
                                                                                                               a.max_len ? nullptr : s.max_positive);
        });
    });
    FG_CUDA(cudaGetLastError());
    return 1;
}

/* K1 alone into stage/pool + counts (+ npos) */
static int enqueue_k1(fulgor_gpu_index* x, slot& s, const chunk_args& a, bool want_npos, bool sorted = true) {
    s.per_read.reserve(size_t(a.n) * 4);
    s.stage.reserve(size_t(a.n) * FG_STAGE_STRIDE * sizeof(uint2));
    if (want_npos) s.npos.reserve(size_t(a.n) * 4);
    const uint64_t pool_entries = std::max<uint64_t>(1u << 16, uint64_t(a.n) * x->pool_per_read);
    s.pool.reserve(size_t(pool_entries) * sizeof(uint2));
    FG_CUDA(cudaMemsetAsync(s.exhausted, 0, 4, s.stream));
    FG_CUDA(cudaMemsetAsync(s.pool_used, 0, 8, s.stream));
    FG_CUDA(cudaMemsetAsync(s.max_positive, 0, 4, s.stream));
    return launch_k1(x, s, a, k1_target{s.stage.as<uint2>(), s.per_read.as<uint32_t>(), want_npos ? s.npos.as<uint32_t>() : nullptr,
                                        entry_pool{s.pool.as<uint2>(), s.pool_used, pool_entries, s.exhausted}}, sorted);
}

/* the per-read {color-set id, multiplicity} lists the color-set kernel reads */
struct list_source {
    const uint32_t* counts;
    const uint2* stage;
    const uint2* pool;
    const uint32_t* npos;
};
static emit_plan enqueue_color_sets(fulgor_gpu_index* x, slot& s, const chunk_args& a, const list_source& ls, int algo, double threshold,
                                    uint64_t* d_off, cudaEvent_t after_k2, int* launches);

/* whole path for a device-resident chunk, up to the CSR offsets (d_off, n+1 entries); returns how to emit the values */
static emit_plan enqueue_pseudoalign(fulgor_gpu_index* x, slot& s, const chunk_args& a, int algo, double threshold, uint64_t* d_off,
                                     cudaEvent_t after_k1, cudaEvent_t after_k2, int* launches) {
    emit_plan e;
    e.n = a.n;
    FG_CUDA(cudaMemsetAsync(s.exhausted, 0, 4, s.stream));
    if (x->H.num_colors <= 32) { /* fused: lookup + intersect/union, one 32-bit mask per read */
        s.per_read.reserve(size_t(a.n) * 4);
        const uint32_t grid = read_grid(x, a.n);
        dispatch_window(x->H, [&](auto w) {
            with_reads(a, [&](auto in) {
                k_pseudoalign_small<decltype(w)::value, decltype(in)><<<grid, FG_BLOCK, 0, s.stream>>>(x->I, in, a.n, algo, threshold, s.per_read.as<uint32_t>());
            });
        });
        FG_CUDA(cudaGetLastError());
        *launches += 1;
        if (after_k1) FG_CUDA(cudaEventRecord(after_k1, s.stream));
        if (after_k2) FG_CUDA(cudaEventRecord(after_k2, s.stream));
        e.kind = emit_plan::MASKS;
        if (d_off) *launches += enqueue_scan<FG_SCAN_POPC>(x, s, s.per_read.as<uint32_t>(), a.n, d_off);
        return e;
    }
    *launches += enqueue_k1(x, s, a, true, /*sorted=*/false);
    if (after_k1) FG_CUDA(cudaEventRecord(after_k1, s.stream));
    return enqueue_color_sets(x, s, a, list_source{s.per_read.as<uint32_t>(), s.stage.as<uint2>(), s.pool.as<uint2>(), s.npos.as<uint32_t>()}, algo, threshold,
                              d_off, after_k2, launches);
}

/* the color-set kernel (K2) over the per-read {color-set id, multiplicity} lists K1 left in stage/pool, then the CSR offsets.
   `counts` = entries per read (K1's, or the deduplicated ones where only a group's representative keeps its list). */
static emit_plan enqueue_color_sets(fulgor_gpu_index* x, slot& s, const chunk_args& a, const list_source& ls, int algo, double threshold,
                                    uint64_t* d_off, cudaEvent_t after_k2, int* launches) {
    emit_plan e;
    e.n = a.n;
    /* the counters of the color-set kernels must hold the largest score: at most the k-mers of the longest read. Chunks that
       come from host buffers know that length; for device-resident reads K1 reports the largest number of positive k-mers. */
    uint32_t max_kmers = a.max_len ? (a.max_len >= x->H.k ? a.max_len - x->H.k + 1 : 1u) : 0u;
    if (max_kmers == 0 && algo != FULGOR_GPU_FULL_INTERSECTION) {
        FG_CUDA(cudaMemcpyAsync(s.h_info + 3, s.max_positive, 4, cudaMemcpyDeviceToHost, s.stream));
        FG_CUDA(cudaStreamSynchronize(s.stream));
        max_kmers = std::max<uint32_t>(1, uint32_t(s.h_info[3]));
    }
    e.words_per_read = (x->H.num_colors + 31) / 32;
    s.res_bits.reserve(size_t(a.n) * e.words_per_read * 4);
    s.res_counts.reserve(size_t(a.n) * 4);
    if (x->I.set_table) { /* decoded table: registers only */
        const uint32_t grid = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>((uint64_t(a.n) + FG_WARPS_PER_BLOCK - 1) / FG_WARPS_PER_BLOCK, uint64_t(x->sm_count) * 16)));
        dispatch_table_kernel(algo, max_kmers, [&](auto fi, auto np, auto t) {
            auto kernel = k_color_sets_table<decltype(fi)::value, decltype(np)::value, decltype(t)::value>;
            const size_t smem = table_kernel_smem(decltype(fi)::value, decltype(np)::value, decltype(t)::value);
            if (smem > 48 * 1024) FG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            kernel<<<grid, FG_BLOCK, smem, s.stream>>>(
                x->I, ls.counts, ls.stage, ls.pool, ls.npos, a.n, threshold, e.words_per_read, s.res_bits.as<uint32_t>(), s.res_counts.as<uint32_t>());
        });
    } else { /* compressed sets decoded per read */
        if (max_kmers == 0) { /* full intersection counts sets, at most as many as positive k-mers */
            FG_CUDA(cudaMemcpyAsync(s.h_info + 3, s.max_positive, 4, cudaMemcpyDeviceToHost, s.stream));
            FG_CUDA(cudaStreamSynchronize(s.stream));
            max_kmers = std::max<uint32_t>(1, uint32_t(s.h_info[3]));
        }
        const general_plan g = plan_color_sets_general(x->H.num_colors, x->H.num_partitions, algo, max_kmers, x->I.diff != 0);
        if (!g.ok) throw std::runtime_error("indexes with more than ~50,000 colors are not supported yet");
        const uint32_t wpb = g.warps_per_block, ints = g.ints_per_warp;
        const size_t smem = g.smem_bytes;
        FG_CUDA(cudaFuncSetAttribute(k_color_sets_general, cudaFuncAttributeMaxDynamicSharedMemorySize, int(std::max<size_t>(smem, 48 * 1024))));
        const uint64_t blocks_needed = (uint64_t(a.n) + wpb - 1) / wpb;
        const uint32_t grid = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(blocks_needed, uint64_t(x->sm_count) * 8)));
        k_color_sets_general<<<grid, wpb * 32, smem, s.stream>>>(x->I, ls.counts, ls.stage, ls.pool, ls.npos, a.n, algo, threshold, e.words_per_read, g.planes,
                                                               ints, s.res_bits.as<uint32_t>(), s.res_counts.as<uint32_t>());
    }
    FG_CUDA(cudaGetLastError());
    *launches += 1;
    if (after_k2) FG_CUDA(cudaEventRecord(after_k2, s.stream));
    e.kind = emit_plan::BITS;
    if (d_off) *launches += enqueue_scan<FG_SCAN_PLAIN>(x, s, s.res_counts.as<uint32_t>(), a.n, d_off); /* no offsets: the caller takes the bitmap rows */
    return e;
}

static emit_plan enqueue_fetch(fulgor_gpu_index* x, slot& s, const chunk_args& a, bool want_npos, uint64_t* d_off, int* launches) {
    emit_plan e;
    e.n = a.n;
    e.kind = emit_plan::ENTRIES;
    *launches += enqueue_k1(x, s, a, want_npos);
    *launches += enqueue_scan<FG_SCAN_PLAIN>(x, s, s.per_read.as<uint32_t>(), a.n, d_off);
    return e;
}

/* ---- the chunked host pipeline ---- */

/* Reads per chunk. Small chunks keep the pipeline's fill and drain short (measured on salmonella_10, 10 M reads per call:
   2^20 -> 33.4 ms, 2^18 -> 31.8 ms, the host->device copy of the reads being the bottleneck); deduplication groups reads
   inside a chunk, so it takes the largest one. FULGOR_GPU_CHUNK_READS overrides (tuning, tests of the multi-chunk path). */
static uint64_t chunk_max_reads(bool dedup = false) {
    const char* e = std::getenv("FULGOR_GPU_CHUNK_READS");
    const uint64_t v = e ? std::strtoull(e, nullptr, 10) : 0;
    return v ? std::min<uint64_t>(std::max<uint64_t>(v, 32), 1u << 22) : (dedup ? (1u << 20) : (1u << 18));
}
#define CHUNK_MAX_READS (chunk_max_reads())
static const uint64_t CHUNK_MAX_BASES = 256ull << 20;
static const uint64_t CHUNK_MAX_RESULT_BITS_BYTES = 256ull << 20;
static const int RC_RETRY_LARGER_POOL = 1;

enum class op_kind { FETCH, PSEUDOALIGN, DEDUP };

/* the caller's reads: ASCII (bases + read_off) or packed (words + read_len + invalid positions, include/fulgor_gpu.h) */
struct host_reads {
    const char* bases = nullptr;
    const uint64_t* read_off = nullptr;
    const uint32_t* words = nullptr;
    const uint32_t* read_len = nullptr;
    const uint64_t* invalid = nullptr;
    uint64_t n_invalid = 0;
    bool packed() const { return read_len != nullptr; }
};
/* the caller's result buffers: CSR lists (off + vals, cap entries) or one bitmap row of ceil(num_colors / 32) words per read */
struct host_results {
    uint64_t* off = nullptr;
    uint32_t* vals = nullptr;
    uint64_t cap = 0;
    uint32_t* bitmaps = nullptr;
    uint32_t* per_read = nullptr; /* n_reads entries, nullable: FETCH -> positive k-mer counts; DEDUP -> representatives */
    bool lists() const { return bitmaps == nullptr; }
};

static int run_host_batch_once(fulgor_gpu_index* x, op_kind op, int algo, double threshold, const host_reads& in, uint32_t n_reads,
                               const host_results& out) {
    uint32_t* const num_positive = op == op_kind::FETCH ? out.per_read : nullptr;
    FG_CUDA(cudaMemsetAsync(x->d_carry, 0, 8, x->slots[0].stream));
    FG_CUDA(cudaStreamSynchronize(x->slots[0].stream));

    const bool small = x->H.num_colors <= 32;
    const uint32_t words_per_read = (x->H.num_colors + 31) / 32;
    uint64_t max_reads = chunk_max_reads(false);
    if (op != op_kind::FETCH && !small)
        max_reads = std::max<uint64_t>(1024, std::min<uint64_t>(max_reads, CHUNK_MAX_RESULT_BITS_BYTES / (uint64_t(words_per_read) * 4)));
    struct pending { uint32_t first, n; int slot; emit_plan plan; };
    bool too_big = false, exhausted = false;
    cudaEvent_t prev_scanned = nullptr;

    auto finalize = [&](const pending& c) {
        slot& s = x->slots[c.slot];
        FG_CUDA(cudaEventSynchronize(s.info_ready));
        if (uint32_t(s.h_info[2])) exhausted = true;
        if (out.lists()) {
            const uint64_t base = s.h_info[0], total = s.h_info[1];
            if (base + total > out.cap) too_big = true;
            if (!too_big && !exhausted && total) {
                if (total * 4 > s.out.cap) { /* the first emit ran into the end of the buffer: grow and repeat it */
                    FG_CUDA(cudaStreamSynchronize(s.stream));
                    s.out.reserve(size_t(total) * 4);
                    enqueue_emit(s, c.plan, s.off.as<uint64_t>(), s.out.as<uint32_t>(), s.out.cap / 4);
                }
                FG_CUDA(cudaMemcpyAsync(out.vals + base, s.out.p, total * 4, cudaMemcpyDeviceToHost, s.stream));
            }
        }
        FG_CUDA(cudaEventRecord(s.done, s.stream));
    };

    /* chunks are cut and validated on the fly, so the host-side O(n) work overlaps the GPU work of earlier chunks. Whatever
       goes wrong after the first chunk has been enqueued (a bad offset further down, an allocation, a CUDA error), the copies
       already queued into the caller's buffers must have landed before this call returns: drain every slot, then rethrow. */
    std::deque<pending> inflight;
    uint32_t ci = 0;
    uint64_t word_first = 0, inv_first = 0; /* packed: first word / first invalid-position entry of the next chunk */
    try {
    for (uint32_t first = 0; first < n_reads; ++ci) {
        uint32_t n = uint32_t(std::min<uint64_t>(max_reads, n_reads - first));
        uint32_t chunk_max_len = 1;
        uint64_t chunk_words = 0;
        if (in.packed()) { /* lengths, not offsets: one pass finds the chunk's end, its longest read and its words */
            const uint32_t* rl = in.read_len + first;
            uint32_t longest = 1, i = 0;
            for (; i < n; ++i) {
                const uint32_t len = rl[i] & 0x7fffffffu;
                const uint64_t w = (uint64_t(len) + 15) >> 4;
                if (i && (chunk_words + w) * 4 > CHUNK_MAX_BASES / 4) break; /* the same number of bases per chunk as the ASCII path */
                chunk_words += w;
                longest = std::max(longest, len);
            }
            n = i;
            chunk_max_len = longest;
        } else {
            if (in.read_off[first + n] - in.read_off[first] > CHUNK_MAX_BASES) { /* largest n >= 1 within the byte budget */
                uint32_t lo = 1, hi = n;
                while (lo < hi) {
                    const uint32_t mid = lo + (hi - lo + 1) / 2;
                    if (in.read_off[first + mid] - in.read_off[first] <= CHUNK_MAX_BASES) lo = mid; else hi = mid - 1;
                }
                n = lo;
            }
            const uint64_t* ro = in.read_off + first;
            uint64_t bad = 0, longest = 0;
            for (uint32_t i = 0; i < n; ++i) {
                const uint64_t len = ro[i + 1] - ro[i];
                bad |= len >> 31; /* also catches decreasing offsets (wrap-around) */
                longest = std::max(longest, len);
            }
            if (bad) throw std::invalid_argument("read_off must be non-decreasing and reads shorter than 2^31 characters");
            chunk_max_len = uint32_t(std::max<uint64_t>(1, longest));
        }
        pending c{first, n, int(ci % FG_NUM_SLOTS), emit_plan()};
        slot& s = x->slots[c.slot];
        if (s.busy) FG_CUDA(cudaEventSynchronize(s.done));
        s.busy = true;
        chunk_args a{};
        a.n = c.n;
        a.max_len = chunk_max_len;
        int launches = 0;
        if (in.packed()) {
            uint64_t inv_end = inv_first;
            const uint64_t pos_end = 16 * (word_first + chunk_words);
            while (inv_end < in.n_invalid && in.invalid[inv_end] < pos_end) ++inv_end;
            if (inv_end - inv_first > 0xffffffffull) throw std::invalid_argument("too many invalid characters in one chunk of reads");
            s.bases.reserve(size_t(chunk_words) * 4 + 64);
            s.read_len.reserve(size_t(c.n) * 4);
            s.pk_word_off.reserve(size_t(c.n + 1) * 8);
            s.invalid.reserve(size_t(inv_end - inv_first) * 8 + 8);
            if (chunk_words) FG_CUDA(cudaMemcpyAsync(s.bases.p, in.words + word_first, chunk_words * 4, cudaMemcpyHostToDevice, s.stream));
            FG_CUDA(cudaMemcpyAsync(s.read_len.p, in.read_len + c.first, size_t(c.n) * 4, cudaMemcpyHostToDevice, s.stream));
            if (inv_end > inv_first)
                FG_CUDA(cudaMemcpyAsync(s.invalid.p, in.invalid + inv_first, (inv_end - inv_first) * 8, cudaMemcpyHostToDevice, s.stream));
            /* first word of every read: exclusive scan of the reads' word counts (chunk-local, its own carry) */
            FG_CUDA(cudaMemsetAsync(s.pk_scan, 0, 8, s.stream));
            const uint32_t tiles = (c.n + FG_SCAN_TILE - 1) / FG_SCAN_TILE;
            s.tile_sums.reserve(size_t(tiles) * 4);
            s.tile_off.reserve(size_t(tiles) * 8);
            k_scan_tile_sums<FG_SCAN_PACKED_WORDS><<<tiles, FG_SCAN_BLOCK, 0, s.stream>>>(s.read_len.as<uint32_t>(), c.n, s.tile_sums.as<uint32_t>());
            k_scan_tile_offsets<<<1, FG_SCAN_BLOCK, 0, s.stream>>>(s.tile_sums.as<uint32_t>(), tiles, s.tile_off.as<uint64_t>(), s.pk_scan, s.pk_scan + 1,
                                                                  s.pk_word_off.as<uint64_t>() + c.n);
            k_scan_write<FG_SCAN_PACKED_WORDS><<<tiles, FG_SCAN_BLOCK, 0, s.stream>>>(s.read_len.as<uint32_t>(), c.n, s.tile_off.as<uint64_t>(), s.pk_scan + 1,
                                                                                     s.pk_word_off.as<uint64_t>());
            FG_CUDA(cudaGetLastError());
            launches += 3;
            a.d_words = s.bases.as<uint32_t>();
            a.d_word_off = s.pk_word_off.as<uint64_t>();
            a.d_read_len = s.read_len.as<uint32_t>();
            a.d_invalid = s.invalid.as<uint64_t>();
            a.n_invalid = uint32_t(inv_end - inv_first);
            a.pos_base = 16 * word_first;
            word_first += chunk_words;
            inv_first = inv_end;
        } else {
            const uint64_t b0 = in.read_off[c.first], b1 = in.read_off[c.first + c.n];
            s.bases.reserve(size_t(b1 - b0) + 64);
            s.read_off.reserve(size_t(c.n + 1) * 8);
            if (b1 > b0) FG_CUDA(cudaMemcpyAsync(s.bases.p, in.bases + b0, b1 - b0, cudaMemcpyHostToDevice, s.stream));
            FG_CUDA(cudaMemcpyAsync(s.read_off.p, in.read_off + c.first, size_t(c.n + 1) * 8, cudaMemcpyHostToDevice, s.stream));
            a.d_bases = s.bases.as<uint8_t>();
            a.d_read_off = s.read_off.as<uint64_t>();
            a.read_off_base = b0;
        }
        uint64_t* d_off = nullptr;
        if (out.lists()) {
            s.off.reserve(size_t(c.n + 1) * 8);
            d_off = s.off.as<uint64_t>();
            /* values buffer: exact upper bound when it is small, otherwise a guess that finalize() corrects */
            const uint64_t per_read_guess = (op == op_kind::PSEUDOALIGN && small) ? x->H.num_colors : 64;
            s.out.reserve(size_t(c.n) * per_read_guess * 4);
            if (prev_scanned) FG_CUDA(cudaStreamWaitEvent(s.stream, prev_scanned, 0)); /* the running CSR offset */
        }
        if (op == op_kind::FETCH) {
            c.plan = enqueue_fetch(x, s, a, num_positive != nullptr, d_off, &launches);
        } else {
            c.plan = enqueue_pseudoalign(x, s, a, algo, threshold, d_off, nullptr, nullptr, &launches);
        }
        if (out.lists()) {
            FG_CUDA(cudaEventRecord(s.scanned, s.stream));
            prev_scanned = s.scanned;
            enqueue_emit(s, c.plan, d_off, s.out.as<uint32_t>(), s.out.cap / 4);
            FG_CUDA(cudaMemcpyAsync(s.h_info, s.chunk_info, 16, cudaMemcpyDeviceToHost, s.stream));
        }
        FG_CUDA(cudaMemcpyAsync(s.h_info + 2, s.exhausted, 4, cudaMemcpyDeviceToHost, s.stream));
        FG_CUDA(cudaEventRecord(s.info_ready, s.stream));
        if (out.lists()) {
            /* offsets of reads [first, first+n) and the running total at [first+n] (the next chunk rewrites that entry with the same value) */
            FG_CUDA(cudaMemcpyAsync(out.off + c.first, s.off.p, size_t(c.n + 1) * 8, cudaMemcpyDeviceToHost, s.stream));
        } else { /* bitmap rows: one mask per read (<= 32 colors, the fused kernel's result) or the color-set kernel's rows, as they are */
            const void* rows = c.plan.kind == emit_plan::MASKS ? s.per_read.p : s.res_bits.p;
            FG_CUDA(cudaMemcpyAsync(out.bitmaps + uint64_t(c.first) * words_per_read, rows, size_t(c.n) * words_per_read * 4, cudaMemcpyDeviceToHost, s.stream));
        }
        if (op == op_kind::FETCH && num_positive)
            FG_CUDA(cudaMemcpyAsync(num_positive + c.first, s.npos.p, size_t(c.n) * 4, cudaMemcpyDeviceToHost, s.stream));
        inflight.push_back(c);
        while (inflight.size() > FG_FINALIZE_LAG) {
            finalize(inflight.front());
            inflight.pop_front();
        }
        first += n;
    }
    for (auto const& c : inflight) finalize(c);
    } catch (...) {
        for (auto& s : x->slots) {
            if (s.stream) cudaStreamSynchronize(s.stream); /* best effort: the first error is the one reported */
            s.busy = false;
        }
        throw;
    }
    for (auto& s : x->slots) {
        FG_CUDA(cudaStreamSynchronize(s.stream));
        s.busy = false;
    }
    if (exhausted) return RC_RETRY_LARGER_POOL;
    return too_big ? FULGOR_GPU_E2BIG : 0;
}

/* Full intersection computed once per distinct color-set-id list of the WHOLE call (the reference deduplicates the whole query
   file, tools/pseudoalign.cpp:92-226). Two passes over the chunks on slot 0:
     1. lookup kernel per chunk, the per-read lists kept in call-wide arrays (256 bytes of list head per read + one entry pool);
     2. k_group_reads over all reads of the call: reads with the same list share a representative, the first one found;
     3. color-set kernel, scan, emit and the copy back per chunk, on the representatives only (the others own an empty range).
   The gain of deduplication is the work and the bytes NOT spent on repeated lists, so this path is not software-pipelined. */
static int run_host_dedup_once(fulgor_gpu_index* x, const host_reads& in, uint32_t n_reads, const host_results& out) {
    slot& s = x->slots[0];
    const uint32_t words_per_read = (x->H.num_colors + 31) / 32;
    uint64_t max_reads = chunk_max_reads(true);
    max_reads = std::max<uint64_t>(1024, std::min<uint64_t>(max_reads, CHUNK_MAX_RESULT_BITS_BYTES / (uint64_t(words_per_read) * 4)));
    x->dd_stage.reserve(size_t(n_reads) * FG_STAGE_STRIDE * sizeof(uint2));
    x->dd_counts.reserve(size_t(n_reads) * 4);
    x->dd_npos.reserve(size_t(n_reads) * 4);
    x->dd_rep.reserve(size_t(n_reads) * 4);
    x->dd_rep_counts.reserve(size_t(n_reads) * 4);
    const uint64_t pool_entries = std::max<uint64_t>(1u << 16, uint64_t(n_reads) * x->pool_per_read);
    x->dd_pool.reserve(size_t(pool_entries) * sizeof(uint2));
    uint32_t log2_slots = 10;
    while ((1ull << log2_slots) < 2ull * n_reads) ++log2_slots;
    x->dd_slots.reserve(size_t(4) << log2_slots);
    FG_CUDA(cudaMemsetAsync(x->d_carry, 0, 8, s.stream));
    FG_CUDA(cudaMemsetAsync(s.exhausted, 0, 4, s.stream));
    FG_CUDA(cudaMemsetAsync(s.pool_used, 0, 8, s.stream));
    FG_CUDA(cudaMemsetAsync(x->dd_slots.p, 0xff, size_t(4) << log2_slots, s.stream));
    const entry_pool pool{x->dd_pool.as<uint2>(), s.pool_used, pool_entries, s.exhausted};
    struct piece { uint32_t first, n, max_len; };
    std::vector<piece> pieces;
    try {
        /* 1. the lists of every read */
        for (uint32_t first = 0; first < n_reads;) {
            uint32_t n = uint32_t(std::min<uint64_t>(max_reads, n_reads - first));
            while (n > 1 && in.read_off[first + n] - in.read_off[first] > CHUNK_MAX_BASES) n = (n + 1) / 2;
            uint64_t longest = 1;
            for (uint32_t i = 0; i < n; ++i) {
                const uint64_t len = in.read_off[first + i + 1] - in.read_off[first + i];
                if (len >> 31) throw std::invalid_argument("read_off must be non-decreasing and reads shorter than 2^31 characters");
                longest = std::max(longest, len);
            }
            const uint64_t b0 = in.read_off[first], b1 = in.read_off[first + n];
            s.bases.reserve(size_t(b1 - b0) + 64);
            s.read_off.reserve(size_t(n + 1) * 8);
            if (b1 > b0) FG_CUDA(cudaMemcpyAsync(s.bases.p, in.bases + b0, b1 - b0, cudaMemcpyHostToDevice, s.stream));
            FG_CUDA(cudaMemcpyAsync(s.read_off.p, in.read_off + first, size_t(n + 1) * 8, cudaMemcpyHostToDevice, s.stream));
            chunk_args a{};
            a.d_bases = s.bases.as<uint8_t>();
            a.d_read_off = s.read_off.as<uint64_t>();
            a.read_off_base = b0;
            a.n = n;
            a.max_len = uint32_t(longest);
            launch_k1(x, s, a, k1_target{x->dd_stage.as<uint2>() + uint64_t(first) * FG_STAGE_STRIDE, x->dd_counts.as<uint32_t>() + first,
                                         x->dd_npos.as<uint32_t>() + first, pool}, /*sorted=*/true);
            FG_CUDA(cudaStreamSynchronize(s.stream)); /* the next chunk reuses the input buffers */
            pieces.push_back({first, n, uint32_t(longest)});
            first += n;
        }
        FG_CUDA(cudaMemcpyAsync(s.h_info + 2, s.exhausted, 4, cudaMemcpyDeviceToHost, s.stream));
        FG_CUDA(cudaStreamSynchronize(s.stream));
        if (uint32_t(s.h_info[2])) return RC_RETRY_LARGER_POOL;
        /* 2. groups over the whole call */
        const uint32_t grid = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>((uint64_t(n_reads) + FG_WARPS_PER_BLOCK - 1) / FG_WARPS_PER_BLOCK, uint64_t(x->sm_count) * 16)));
        k_group_reads<<<grid, FG_BLOCK, 0, s.stream>>>(x->dd_counts.as<uint32_t>(), x->dd_stage.as<uint2>(), x->dd_pool.as<uint2>(), n_reads, 0,
                                                       x->dd_slots.as<uint32_t>(), log2_slots, x->dd_rep.as<uint32_t>(), x->dd_rep_counts.as<uint32_t>());
        FG_CUDA(cudaGetLastError());
        FG_CUDA(cudaMemcpyAsync(out.per_read, x->dd_rep.p, size_t(n_reads) * 4, cudaMemcpyDeviceToHost, s.stream));
        /* 3. the representatives' intersections, chunk by chunk */
        bool too_big = false;
        for (const piece& c : pieces) {
            chunk_args a{};
            a.n = c.n;
            a.max_len = c.max_len;
            s.off.reserve(size_t(c.n + 1) * 8);
            s.out.reserve(size_t(c.n) * 64 * 4);
            int launches = 0;
            const list_source ls{x->dd_rep_counts.as<uint32_t>() + c.first, x->dd_stage.as<uint2>() + uint64_t(c.first) * FG_STAGE_STRIDE, x->dd_pool.as<uint2>(),
                                 x->dd_npos.as<uint32_t>() + c.first};
            const emit_plan plan = enqueue_color_sets(x, s, a, ls, FULGOR_GPU_FULL_INTERSECTION, 1.0, s.off.as<uint64_t>(), nullptr, &launches);
            FG_CUDA(cudaMemcpyAsync(s.h_info, s.chunk_info, 16, cudaMemcpyDeviceToHost, s.stream));
            FG_CUDA(cudaMemcpyAsync(out.off + c.first, s.off.p, size_t(c.n + 1) * 8, cudaMemcpyDeviceToHost, s.stream));
            FG_CUDA(cudaStreamSynchronize(s.stream));
            const uint64_t base = s.h_info[0], total = s.h_info[1];
            if (base + total > out.cap) too_big = true;
            if (!too_big && total) {
                s.out.reserve(size_t(total) * 4);
                enqueue_emit(s, plan, s.off.as<uint64_t>(), s.out.as<uint32_t>(), s.out.cap / 4);
                FG_CUDA(cudaMemcpyAsync(out.vals + base, s.out.p, total * 4, cudaMemcpyDeviceToHost, s.stream));
                FG_CUDA(cudaStreamSynchronize(s.stream));
            }
        }
        FG_CUDA(cudaStreamSynchronize(s.stream));
        return too_big ? FULGOR_GPU_E2BIG : 0;
    } catch (...) {
        cudaStreamSynchronize(s.stream);
        throw;
    }
}

static int run_host_batch(fulgor_gpu_index* x, op_kind op, int algo, double threshold, const host_reads& in, uint32_t n_reads, const host_results& out) {
    FG_CUDA(cudaSetDevice(x->device));
    if (out.off) out.off[0] = 0;
    if (n_reads == 0) return 0;
    for (int attempt = 0; attempt < 12; ++attempt) {
        const int rc = op == op_kind::DEDUP ? run_host_dedup_once(x, in, n_reads, out) : run_host_batch_once(x, op, algo, threshold, in, n_reads, out);
        if (rc != RC_RETRY_LARGER_POOL) return rc;
        x->pool_per_read *= 4; /* reads with many distinct color sets: rerun with a larger entry pool (kept for later calls) */
    }
    throw std::runtime_error("entry pool kept overflowing");
}

/* ---- the per-k-mer tools (kmer-conservation, kmer-matches): one chunk at a time on slot 0 ---- */

enum class kmer_tool { CONSERVATION, MATCHES };
static const uint64_t TOOL_MAX_BASES = 64ull << 20;
static const uint64_t TOOL_MAX_COUNT_BYTES = 256ull << 20;

static int run_kmer_tool(fulgor_gpu_index* x, kmer_tool tool, const char* bases, const uint64_t* read_off, uint32_t n_reads, uint64_t* out_off,
                         uint32_t* out_vals, uint64_t cap, uint32_t* counts) {
    FG_CUDA(cudaSetDevice(x->device));
    slot& s = x->slots[0];
    const uint64_t k = x->H.k, C = x->H.num_colors;
    auto kmers_of = [&](uint32_t i) -> uint64_t {
        const uint64_t len = read_off[i + 1] - read_off[i];
        if (len >> 31) throw std::invalid_argument("read_off must be non-decreasing and reads shorter than 2^31 characters");
        return len >= k ? len - k + 1 : 0;
    };
    out_off[0] = 0;
    if (tool == kmer_tool::MATCHES) {
        if (!x->I.set_table)
            throw std::runtime_error("kmer_matches needs the decoded color-set table (it did not fit FULGOR_GPU_TABLE_MAX_MB / the free device memory)");
        for (uint32_t i = 0; i < n_reads; ++i) out_off[i + 1] = out_off[i] + (kmers_of(i) + 31) / 32;
        if (out_off[n_reads] > cap) return FULGOR_GPU_E2BIG;
    }
    if (n_reads == 0) return 0;
    FG_CUDA(cudaMemsetAsync(x->d_carry, 0, 8, s.stream));
    uint64_t max_reads = CHUNK_MAX_READS;
    if (tool == kmer_tool::MATCHES) max_reads = std::max<uint64_t>(256, std::min<uint64_t>(max_reads, TOOL_MAX_COUNT_BYTES / (4 * std::max<uint64_t>(1, C))));
    std::vector<uint64_t> koff, woff;
    bool too_big = false;
    for (uint32_t first = 0; first < n_reads;) {
        uint32_t n = uint32_t(std::min<uint64_t>(max_reads, n_reads - first));
        while (n > 1 && read_off[first + n] - read_off[first] > TOOL_MAX_BASES) n = (n + 1) / 2;
        koff.assign(size_t(n) + 1, 0);
        uint64_t longest = 1;
        for (uint32_t i = 0; i < n; ++i) {
            koff[i + 1] = koff[i] + kmers_of(first + i);
            longest = std::max(longest, read_off[first + i + 1] - read_off[first + i]);
        }
        const uint64_t b0 = read_off[first], b1 = read_off[first + n], total_kmers = koff[n];
        s.bases.reserve(size_t(b1 - b0) + 64);
        s.read_off.reserve(size_t(n + 1) * 8);
        s.kmer_off.reserve(size_t(n + 1) * 8);
        s.per_kmer.reserve(size_t(total_kmers) * 4 + 4);
        if (b1 > b0) FG_CUDA(cudaMemcpyAsync(s.bases.p, bases + b0, b1 - b0, cudaMemcpyHostToDevice, s.stream));
        FG_CUDA(cudaMemcpyAsync(s.read_off.p, read_off + first, size_t(n + 1) * 8, cudaMemcpyHostToDevice, s.stream));
        FG_CUDA(cudaMemcpyAsync(s.kmer_off.p, koff.data(), size_t(n + 1) * 8, cudaMemcpyHostToDevice, s.stream));
        chunk_args a{s.bases.as<uint8_t>(), s.read_off.as<uint64_t>(), b0, n, uint32_t(longest)};
        const uint32_t grid = read_grid(x, n);
        dispatch_window(x->H, [&](auto w) {
            k_kmer_color_sets<decltype(w)::value, ascii_reads><<<grid, FG_BLOCK, 0, s.stream>>>(x->I, ascii_reads{a.d_bases, a.d_read_off, a.read_off_base}, n,
                                                                                                  s.kmer_off.as<uint64_t>(), s.per_kmer.as<uint32_t>());
        });
        FG_CUDA(cudaGetLastError());
        const uint32_t warp_grid = uint32_t((uint64_t(n) * 32 + 255) / 256);
        if (tool == kmer_tool::CONSERVATION) {
            s.per_read.reserve(size_t(n) * 4);
            s.off.reserve(size_t(n + 1) * 8);
            k_kmer_runs<false><<<warp_grid, 256, 0, s.stream>>>(s.per_kmer.as<uint32_t>(), s.kmer_off.as<uint64_t>(), n, s.per_read.as<uint32_t>(), nullptr,
                                                                 nullptr, nullptr, 0);
            FG_CUDA(cudaGetLastError());
            enqueue_scan<FG_SCAN_PLAIN>(x, s, s.per_read.as<uint32_t>(), n, s.off.as<uint64_t>());
            FG_CUDA(cudaMemcpyAsync(s.h_info, s.chunk_info, 16, cudaMemcpyDeviceToHost, s.stream));
            FG_CUDA(cudaMemcpyAsync(out_off + first, s.off.p, size_t(n + 1) * 8, cudaMemcpyDeviceToHost, s.stream));
            FG_CUDA(cudaStreamSynchronize(s.stream));
            const uint64_t base = s.h_info[0], total = s.h_info[1];
            if (base + total > cap) too_big = true;
            if (!too_big && total) {
                s.out.reserve(size_t(total) * 12);
                k_kmer_runs<true><<<warp_grid, 256, 0, s.stream>>>(s.per_kmer.as<uint32_t>(), s.kmer_off.as<uint64_t>(), n, nullptr, s.off.as<uint64_t>(),
                                                                    s.chunk_info, s.out.as<uint32_t>(), total);
                FG_CUDA(cudaGetLastError());
                FG_CUDA(cudaMemcpyAsync(out_vals + 3 * base, s.out.p, size_t(total) * 12, cudaMemcpyDeviceToHost, s.stream));
            }
            FG_CUDA(cudaStreamSynchronize(s.stream));
        } else {
            /* the read's distinct {color-set id, multiplicity} list (K1) feeds the per-color counts */
            for (int attempt = 0;; ++attempt) {
                enqueue_k1(x, s, a, false);
                FG_CUDA(cudaMemcpyAsync(s.h_info + 2, s.exhausted, 4, cudaMemcpyDeviceToHost, s.stream));
                FG_CUDA(cudaStreamSynchronize(s.stream));
                if (!uint32_t(s.h_info[2])) break;
                if (attempt >= 12) throw std::runtime_error("entry pool kept overflowing");
                x->pool_per_read *= 4;
            }
            woff.assign(size_t(n) + 1, 0);
            for (uint32_t i = 0; i <= n; ++i) woff[i] = out_off[first + i] - out_off[first];
            s.word_off.reserve(size_t(n + 1) * 8);
            s.out.reserve(size_t(woff[n]) * 4 + 4);
            s.res_bits.reserve(size_t(n) * C * 4 + 4);
            FG_CUDA(cudaMemcpyAsync(s.word_off.p, woff.data(), size_t(n + 1) * 8, cudaMemcpyHostToDevice, s.stream));
            k_kmer_positive_bits<<<warp_grid, 256, 0, s.stream>>>(s.per_kmer.as<uint32_t>(), s.kmer_off.as<uint64_t>(), s.word_off.as<uint64_t>(), n,
                                                                  s.out.as<uint32_t>());
            FG_CUDA(cudaGetLastError());
            const uint32_t cgrid = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>((uint64_t(n) + FG_WARPS_PER_BLOCK - 1) / FG_WARPS_PER_BLOCK, uint64_t(x->sm_count) * 16)));
            k_kmer_match_counts<<<cgrid, FG_BLOCK, 0, s.stream>>>(x->I, s.per_read.as<uint32_t>(), s.stage.as<uint2>(), s.pool.as<uint2>(), n,
                                                                  s.res_bits.as<uint32_t>());
            FG_CUDA(cudaGetLastError());
            if (woff[n]) FG_CUDA(cudaMemcpyAsync(out_vals + out_off[first], s.out.p, size_t(woff[n]) * 4, cudaMemcpyDeviceToHost, s.stream));
            if (C) FG_CUDA(cudaMemcpyAsync(counts + uint64_t(first) * C, s.res_bits.p, size_t(n) * C * 4, cudaMemcpyDeviceToHost, s.stream));
            FG_CUDA(cudaStreamSynchronize(s.stream));
        }
        first += n;
    }
    return too_big ? FULGOR_GPU_E2BIG : 0;
}

static void check_algo(int algo, double threshold) {
    if (algo != FULGOR_GPU_FULL_INTERSECTION && algo != FULGOR_GPU_THRESHOLD_UNION) throw std::invalid_argument("unknown algorithm");
    /* same domain check as the reference CLI (tools/pseudoalign.cpp:272-281) */
    if (algo == FULGOR_GPU_THRESHOLD_UNION && !(threshold > 0.0 && threshold <= 1.0))
        throw std::invalid_argument("threshold must be a float in (0.0,1.0]");
}

template <typename F>
static int guarded(F&& f) {
    try {
        return f();
    } catch (std::bad_alloc const&) {
        return fail(FULGOR_GPU_ENOMEM, "out of host memory");
    } catch (std::invalid_argument const& e) {
        const bool nodev = std::strstr(e.what(), "no usable CUDA device") != nullptr;
        return fail(nodev ? FULGOR_GPU_ENODEV : FULGOR_GPU_EINVAL, e.what());
    } catch (std::exception const& e) {
        const bool oom = std::strstr(e.what(), "out of memory") != nullptr;
        return fail(oom ? FULGOR_GPU_ENOMEM : FULGOR_GPU_EIO, e.what());
    }
}

}  // namespace fgb

/* ================================================================== C ABI */

extern "C" {

const char* fulgor_gpu_last_error(void) { return g_error.c_str(); }
const char* fulgor_gpu_version(void) { return "fulgor-b200 0.1 (sm_100a)"; }

int fulgor_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

/* CPUs on the same NUMA node / PCIe root as the device, from sysfs (local_cpulist of the device's PCI function) */
static bool cpus_local_to_device(int device, cpu_set_t* set) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, int(sizeof(bus)), device) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    for (char* c = bus; *c; ++c) *c = char(std::tolower(*c));
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return false;
    char line[4096] = {0};
    const bool got = std::fgets(line, sizeof(line), f) != nullptr;
    std::fclose(f);
    if (!got) return false;
    CPU_ZERO(set);
    int n = 0;
    for (char* p = line; *p && *p != '\n';) { /* "0-15,32-47" */
        char* e = nullptr;
        const long a = std::strtol(p, &e, 10);
        if (e == p) break;
        long b = a;
        p = e;
        if (*p == '-') {
            b = std::strtol(p + 1, &e, 10);
            p = e;
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) {
            CPU_SET(int(c), set);
            ++n;
        }
        if (*p == ',') ++p;
    }
    return n > 0;
}

int fulgor_gpu_bind_host_thread(int device) {
    cpu_set_t local, current, both;
    if (!cpus_local_to_device(device, &local)) return 0;
    if (sched_getaffinity(0, sizeof(current), &current) != 0) return 0;
    CPU_AND(&both, &local, &current);
    const int n = CPU_COUNT(&both);
    if (n == 0 || n == CPU_COUNT(&current)) return 0; /* nothing to narrow (one node, or the CPUs are not ours) */
    return sched_setaffinity(0, sizeof(both), &both) == 0 ? n : 0;
}

void* fulgor_gpu_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void fulgor_gpu_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int fulgor_gpu_image_build(const char* index_path, uint8_t** image, uint64_t* image_bytes) {
    return guarded([&]() -> int {
        if (!index_path || !image || !image_bytes) throw std::invalid_argument("null argument");
        if (index_type_from_path(index_path) == -1) throw std::invalid_argument(std::string("Wrong index filename supplied: ") + index_path);
        std::vector<uint8_t> img = build_image_from_file(index_path);
        uint8_t* p = static_cast<uint8_t*>(std::malloc(img.size()));
        if (!p) throw std::bad_alloc();
        std::memcpy(p, img.data(), img.size());
        *image = p;
        *image_bytes = img.size();
        return 0;
    });
}
void fulgor_gpu_image_free(uint8_t* image) { std::free(image); }

int fulgor_gpu_image_info(const uint8_t* image, uint64_t image_bytes, fulgor_gpu_info* out) {
    return guarded([&]() -> int {
        if (!image || !out) throw std::invalid_argument("null argument");
        fgi_header H;
        if (image_bytes < sizeof(H)) throw std::invalid_argument("image too small");
        std::memcpy(&H, image, sizeof(H));
        try {
            check_header(H, image_bytes);
        } catch (std::exception const& e) {
            throw std::invalid_argument(e.what());
        }
        fill_info(H, -1, out);
        return 0;
    });
}

int fulgor_gpu_index_open_image(const uint8_t* image, uint64_t image_bytes, int device, fulgor_gpu_index** out) {
    return guarded([&]() -> int {
        if (!image || !out) throw std::invalid_argument("null argument");
        fgi_header H;
        if (image_bytes < sizeof(H)) throw std::invalid_argument("image too small");
        std::memcpy(&H, image, sizeof(H));
        try {
            check_header(H, image_bytes);
        } catch (std::exception const& e) {
            throw std::invalid_argument(e.what());
        }
        use_device(device);
        void* d = nullptr;
        FG_CUDA(cudaMalloc(&d, image_bytes));
        try {
            FG_CUDA(cudaMemcpy(d, image, image_bytes, cudaMemcpyHostToDevice));
            *out = make_handle(H, d, true, device);
        } catch (...) {
            cudaFree(d);
            throw;
        }
        return 0;
    });
}

int fulgor_gpu_index_open(const char* index_path, int device, fulgor_gpu_index** out) {
    uint8_t* img = nullptr;
    uint64_t bytes = 0;
    int rc = fulgor_gpu_image_build(index_path, &img, &bytes);
    if (rc) return rc;
    rc = fulgor_gpu_index_open_image(img, bytes, device, out);
    fulgor_gpu_image_free(img);
    return rc;
}

int fulgor_gpu_index_adopt_device_image(const void* device_image, uint64_t image_bytes, int device, fulgor_gpu_index** out) {
    return guarded([&]() -> int {
        if (!device_image || !out) throw std::invalid_argument("null argument");
        use_device(device);
        fgi_header H;
        if (image_bytes < sizeof(H)) throw std::invalid_argument("image too small");
        FG_CUDA(cudaMemcpy(&H, device_image, sizeof(H), cudaMemcpyDeviceToHost));
        try {
            check_header(H, image_bytes);
        } catch (std::exception const& e) {
            throw std::invalid_argument(e.what());
        }
        *out = make_handle(H, const_cast<void*>(device_image), false, device);
        return 0;
    });
}

void fulgor_gpu_index_close(fulgor_gpu_index* x) {
    if (!x) return;
    cudaSetDevice(x->device);
    for (auto& s : x->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        for (dev_buffer* b : {&s.bases, &s.read_off, &s.per_read, &s.stage, &s.pool, &s.npos, &s.res_bits, &s.res_counts, &s.tile_sums, &s.tile_off,
                              &s.off, &s.out, &s.group_slots, &s.rep, &s.rep_counts, &s.kmer_off, &s.per_kmer, &s.word_off, &s.read_len, &s.invalid,
                              &s.pk_word_off})
            b->release();
        if (s.chunk_info) cudaFree(s.chunk_info);
        if (s.exhausted) cudaFree(s.exhausted);
        if (s.max_positive) cudaFree(s.max_positive);
        if (s.pool_used) cudaFree(s.pool_used);
        if (s.pk_scan) cudaFree(s.pk_scan);
        if (s.h_info) cudaFreeHost(s.h_info);
        if (s.scanned) cudaEventDestroy(s.scanned);
        if (s.done) cudaEventDestroy(s.done);
        if (s.info_ready) cudaEventDestroy(s.info_ready);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    for (auto& e : x->ev)
        if (e) cudaEventDestroy(e);
    if (x->d_carry) cudaFree(x->d_carry);
    if (x->d_table) cudaFree(x->d_table);
    for (fgb::dev_buffer* b : {&x->dd_stage, &x->dd_counts, &x->dd_npos, &x->dd_pool, &x->dd_rep, &x->dd_rep_counts, &x->dd_slots}) b->release();
    if (x->owns_image && x->d_image) cudaFree(x->d_image);
    delete x;
}

int fulgor_gpu_index_info(const fulgor_gpu_index* x, fulgor_gpu_info* out) {
    if (!x || !out) return fail(FULGOR_GPU_EINVAL, "null argument");
    fill_info(x->H, x->device, out);
    return 0;
}

int fulgor_gpu_fetch_color_set_ids(fulgor_gpu_index* x, const char* bases, const uint64_t* read_off, uint32_t n_reads, uint64_t* cid_off,
                                   uint32_t* cids, uint64_t cids_cap, uint32_t* num_positive) {
    return guarded([&]() -> int {
        if (!x || !read_off || !cid_off || (!bases && n_reads && read_off[n_reads] > read_off[0]) || (!cids && cids_cap))
            throw std::invalid_argument("null argument");
        host_reads in;
        in.bases = bases;
        in.read_off = read_off;
        host_results out;
        out.off = cid_off;
        out.vals = cids;
        out.cap = cids_cap;
        out.per_read = num_positive;
        int rc = run_host_batch(x, op_kind::FETCH, 0, 0.0, in, n_reads, out);
        if (rc == FULGOR_GPU_E2BIG) return fail(rc, "cids_cap too small; cid_off[n_reads] holds the required capacity");
        return rc;
    });
}

int fulgor_gpu_pseudoalign(fulgor_gpu_index* x, int algo, double threshold, const char* bases, const uint64_t* read_off, uint32_t n_reads,
                           uint64_t* color_off, uint32_t* colors, uint64_t colors_cap) {
    return guarded([&]() -> int {
        if (!x || !read_off || !color_off || (!bases && n_reads && read_off[n_reads] > read_off[0]) || (!colors && colors_cap))
            throw std::invalid_argument("null argument");
        check_algo(algo, threshold);
        host_reads in;
        in.bases = bases;
        in.read_off = read_off;
        host_results out;
        out.off = color_off;
        out.vals = colors;
        out.cap = colors_cap;
        int rc = run_host_batch(x, op_kind::PSEUDOALIGN, algo, threshold, in, n_reads, out);
        if (rc == FULGOR_GPU_E2BIG) return fail(rc, "colors_cap too small; color_off[n_reads] holds the required capacity");
        return rc;
    });
}

int fulgor_gpu_pseudoalign_dedup(fulgor_gpu_index* x, const char* bases, const uint64_t* read_off, uint32_t n_reads, uint32_t* rep_of_read,
                                 uint64_t* color_off, uint32_t* colors, uint64_t colors_cap) {
    return guarded([&]() -> int {
        if (!x || !read_off || !color_off || !rep_of_read || (!bases && n_reads && read_off[n_reads] > read_off[0]) || (!colors && colors_cap))
            throw std::invalid_argument("null argument");
        host_reads in;
        in.bases = bases;
        in.read_off = read_off;
        host_results out;
        out.off = color_off;
        out.vals = colors;
        out.cap = colors_cap;
        out.per_read = rep_of_read;
        int rc = run_host_batch(x, op_kind::DEDUP, FULGOR_GPU_FULL_INTERSECTION, 1.0, in, n_reads, out);
        if (rc == FULGOR_GPU_E2BIG) return fail(rc, "colors_cap too small; color_off[n_reads] holds the required capacity");
        return rc;
    });
}

int fulgor_gpu_pack_reads(const char* bases, const uint64_t* read_off, uint32_t n_reads, uint32_t* words, uint64_t words_cap, uint32_t* read_len,
                          uint64_t* invalid_pos, uint64_t invalid_cap, uint64_t* n_words, uint64_t* n_invalid, int threads) {
    return guarded([&]() -> int {
        if (!read_off || !n_words || !n_invalid || (n_reads && (!read_len || (!bases && read_off[n_reads] > read_off[0]))) || (!words && words_cap) ||
            (!invalid_pos && invalid_cap))
            throw std::invalid_argument("null argument");
        for (uint32_t i = 0; i < n_reads; ++i)
            if ((read_off[i + 1] - read_off[i]) >> 31) throw std::invalid_argument("read_off must be non-decreasing and reads shorter than 2^31 characters");
        *n_words = packed_words_of(read_off, n_reads);
        *n_invalid = 0;
        if (*n_words > words_cap) return fail(FULGOR_GPU_E2BIG, "words_cap too small; *n_words holds the required capacity");
        std::vector<uint64_t> inv;
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        pack_reads(bases, read_off, n_reads, words, read_len, inv, threads > 0 ? unsigned(threads) : hw);
        *n_invalid = inv.size();
        if (inv.size() > invalid_cap) return fail(FULGOR_GPU_E2BIG, "invalid_cap too small; *n_invalid holds the required capacity");
        if (!inv.empty()) std::memcpy(invalid_pos, inv.data(), inv.size() * 8);
        return 0;
    });
}

static host_reads packed_input(const uint32_t* words, const uint32_t* read_len, uint32_t n_reads, const uint64_t* invalid_pos, uint64_t n_invalid) {
    if (!read_len || (!invalid_pos && n_invalid)) throw std::invalid_argument("null argument");
    if (!words)
        for (uint32_t i = 0; i < n_reads; ++i)
            if (read_len[i] & 0x7fffffffu) throw std::invalid_argument("null argument");
    for (uint64_t i = 1; i < n_invalid; ++i)
        if (invalid_pos[i] <= invalid_pos[i - 1]) throw std::invalid_argument("invalid_pos must be ascending");
    host_reads in;
    in.words = words;
    in.read_len = read_len;
    in.invalid = invalid_pos;
    in.n_invalid = n_invalid;
    return in;
}

int fulgor_gpu_pseudoalign_packed(fulgor_gpu_index* x, int algo, double threshold, const uint32_t* words, const uint32_t* read_len, uint32_t n_reads,
                                  const uint64_t* invalid_pos, uint64_t n_invalid, uint64_t* color_off, uint32_t* colors, uint64_t colors_cap) {
    return guarded([&]() -> int {
        if (!x || !color_off || (!colors && colors_cap)) throw std::invalid_argument("null argument");
        check_algo(algo, threshold);
        const host_reads in = packed_input(words, read_len, n_reads, invalid_pos, n_invalid);
        host_results out;
        out.off = color_off;
        out.vals = colors;
        out.cap = colors_cap;
        int rc = run_host_batch(x, op_kind::PSEUDOALIGN, algo, threshold, in, n_reads, out);
        if (rc == FULGOR_GPU_E2BIG) return fail(rc, "colors_cap too small; color_off[n_reads] holds the required capacity");
        return rc;
    });
}

int fulgor_gpu_pseudoalign_bitmaps(fulgor_gpu_index* x, int algo, double threshold, const char* bases, const uint64_t* read_off, uint32_t n_reads,
                                   uint32_t* bitmaps) {
    return guarded([&]() -> int {
        if (!x || !read_off || (!bases && n_reads && read_off[n_reads] > read_off[0]) || (!bitmaps && n_reads)) throw std::invalid_argument("null argument");
        check_algo(algo, threshold);
        host_reads in;
        in.bases = bases;
        in.read_off = read_off;
        host_results out;
        out.bitmaps = bitmaps;
        return run_host_batch(x, op_kind::PSEUDOALIGN, algo, threshold, in, n_reads, out);
    });
}

int fulgor_gpu_pseudoalign_packed_bitmaps(fulgor_gpu_index* x, int algo, double threshold, const uint32_t* words, const uint32_t* read_len,
                                          uint32_t n_reads, const uint64_t* invalid_pos, uint64_t n_invalid, uint32_t* bitmaps) {
    return guarded([&]() -> int {
        if (!x || (!bitmaps && n_reads)) throw std::invalid_argument("null argument");
        check_algo(algo, threshold);
        const host_reads in = packed_input(words, read_len, n_reads, invalid_pos, n_invalid);
        host_results out;
        out.bitmaps = bitmaps;
        return run_host_batch(x, op_kind::PSEUDOALIGN, algo, threshold, in, n_reads, out);
    });
}

int fulgor_gpu_kmer_conservation(fulgor_gpu_index* x, const char* bases, const uint64_t* read_off, uint32_t n_reads, uint64_t* triple_off,
                                 uint32_t* triples, uint64_t triples_cap) {
    return guarded([&]() -> int {
        if (!x || !read_off || !triple_off || (!bases && n_reads && read_off[n_reads] > read_off[0]) || (!triples && triples_cap))
            throw std::invalid_argument("null argument");
        int rc = run_kmer_tool(x, kmer_tool::CONSERVATION, bases, read_off, n_reads, triple_off, triples, triples_cap, nullptr);
        if (rc == FULGOR_GPU_E2BIG) return fail(rc, "triples_cap too small; triple_off[n_reads] holds the required capacity");
        return rc;
    });
}

int fulgor_gpu_kmer_matches(fulgor_gpu_index* x, const char* bases, const uint64_t* read_off, uint32_t n_reads, uint64_t* word_off,
                            uint32_t* positive_words, uint64_t words_cap, uint32_t* counts) {
    return guarded([&]() -> int {
        if (!x || !read_off || !word_off || !counts || (!bases && n_reads && read_off[n_reads] > read_off[0]) || (!positive_words && words_cap))
            throw std::invalid_argument("null argument");
        int rc = run_kmer_tool(x, kmer_tool::MATCHES, bases, read_off, n_reads, word_off, positive_words, words_cap, counts);
        if (rc == FULGOR_GPU_E2BIG) return fail(rc, "words_cap too small; word_off[n_reads] holds the required capacity");
        return rc;
    });
}

int fulgor_gpu_pseudoalign_device(fulgor_gpu_index* x, int algo, double threshold, const char* d_bases, const uint64_t* d_read_off,
                                  uint32_t n_reads, uint64_t read_off_base, uint64_t* d_color_off, uint32_t* d_colors, uint64_t colors_cap,
                                  uint64_t* total_out) {
    return guarded([&]() -> int {
        if (!x || !d_read_off || !d_color_off || !total_out) throw std::invalid_argument("null argument");
        check_algo(algo, threshold);
        FG_CUDA(cudaSetDevice(x->device));
        slot& s = x->slots[0];
        FG_CUDA(cudaMemsetAsync(x->d_carry, 0, 8, s.stream));
        if (n_reads == 0) {
            FG_CUDA(cudaMemsetAsync(d_color_off, 0, 8, s.stream));
            FG_CUDA(cudaStreamSynchronize(s.stream));
            *total_out = 0;
            x->last_launches = 0;
            return 0;
        }
        chunk_args a{reinterpret_cast<const uint8_t*>(d_bases), d_read_off, read_off_base, n_reads, 0};
        for (int attempt = 0;; ++attempt) {
            FG_CUDA(cudaMemsetAsync(x->d_carry, 0, 8, s.stream));
            FG_CUDA(cudaEventRecord(x->ev[0], s.stream));
            x->last_launches = 0;
            const emit_plan plan = enqueue_pseudoalign(x, s, a, algo, threshold, d_color_off, x->ev[1], x->ev[2], &x->last_launches);
            enqueue_emit(s, plan, d_color_off, d_colors, colors_cap);
            x->last_launches += 1;
            FG_CUDA(cudaEventRecord(x->ev[3], s.stream));
            FG_CUDA(cudaMemcpyAsync(s.h_info, s.chunk_info, 16, cudaMemcpyDeviceToHost, s.stream));
            FG_CUDA(cudaMemcpyAsync(s.h_info + 2, s.exhausted, 4, cudaMemcpyDeviceToHost, s.stream));
            FG_CUDA(cudaStreamSynchronize(s.stream));
            if (!uint32_t(s.h_info[2])) break;
            if (attempt >= 12) throw std::runtime_error("entry pool kept overflowing");
            x->pool_per_read *= 4;
        }
        for (int i = 0; i < 3; ++i) FG_CUDA(cudaEventElapsedTime(&x->last_ms[i], x->ev[i], x->ev[i + 1]));
        *total_out = s.h_info[1];
        if (s.h_info[1] > colors_cap) return fail(FULGOR_GPU_E2BIG, "colors_cap too small; *total_out holds the required capacity");
        return 0;
    });
}

int fulgor_gpu_pseudoalign_packed_device(fulgor_gpu_index* x, int algo, double threshold, const uint32_t* d_words, const uint32_t* d_read_len,
                                         uint32_t n_reads, const uint64_t* d_invalid_pos, uint32_t n_invalid, int result_bitmaps,
                                         uint64_t* d_color_off, uint32_t* d_colors, uint64_t colors_cap, uint64_t* total_out) {
    return guarded([&]() -> int {
        if (!x || !total_out || (n_reads && !d_read_len) || (!result_bitmaps && !d_color_off) || (n_invalid && !d_invalid_pos))
            throw std::invalid_argument("null argument");
        check_algo(algo, threshold);
        FG_CUDA(cudaSetDevice(x->device));
        slot& s = x->slots[0];
        const uint32_t words_per_read = (x->H.num_colors + 31) / 32;
        *total_out = 0;
        x->last_launches = 0;
        if (n_reads == 0) {
            if (!result_bitmaps) FG_CUDA(cudaMemsetAsync(d_color_off, 0, 8, s.stream));
            FG_CUDA(cudaStreamSynchronize(s.stream));
            return 0;
        }
        if (result_bitmaps && uint64_t(n_reads) * words_per_read > colors_cap) {
            *total_out = uint64_t(n_reads) * words_per_read;
            return fail(FULGOR_GPU_E2BIG, "colors_cap too small; *total_out holds the required capacity");
        }
        s.pk_word_off.reserve(size_t(n_reads + 1) * 8);
        const uint32_t tiles = (n_reads + FG_SCAN_TILE - 1) / FG_SCAN_TILE;
        s.tile_sums.reserve(size_t(tiles) * 4);
        s.tile_off.reserve(size_t(tiles) * 8);
        chunk_args a{};
        a.n = n_reads;
        a.d_words = d_words;
        a.d_word_off = s.pk_word_off.as<uint64_t>();
        a.d_read_len = d_read_len;
        a.d_invalid = d_invalid_pos;
        a.n_invalid = n_invalid;
        uint64_t* d_off = result_bitmaps ? nullptr : d_color_off;
        if (x->H.num_colors > 32) { /* the color-set kernel's counters are sized by the longest read: learn it before the pipeline starts */
            FG_CUDA(cudaMemsetAsync(s.max_positive, 0, 4, s.stream));
            k_max_read_len<<<std::min<uint32_t>((n_reads + 255) / 256, 1024), 256, 0, s.stream>>>(d_read_len, n_reads, s.max_positive);
            FG_CUDA(cudaGetLastError());
            FG_CUDA(cudaMemcpyAsync(s.h_info + 3, s.max_positive, 4, cudaMemcpyDeviceToHost, s.stream));
            FG_CUDA(cudaStreamSynchronize(s.stream));
            a.max_len = std::max<uint32_t>(1, uint32_t(s.h_info[3]));
        }
        for (int attempt = 0;; ++attempt) {
            FG_CUDA(cudaMemsetAsync(x->d_carry, 0, 8, s.stream));
            FG_CUDA(cudaMemsetAsync(s.pk_scan, 0, 8, s.stream));
            FG_CUDA(cudaEventRecord(x->ev[0], s.stream));
            x->last_launches = 3;
            k_scan_tile_sums<FG_SCAN_PACKED_WORDS><<<tiles, FG_SCAN_BLOCK, 0, s.stream>>>(d_read_len, n_reads, s.tile_sums.as<uint32_t>());
            k_scan_tile_offsets<<<1, FG_SCAN_BLOCK, 0, s.stream>>>(s.tile_sums.as<uint32_t>(), tiles, s.tile_off.as<uint64_t>(), s.pk_scan, s.pk_scan + 1,
                                                                  s.pk_word_off.as<uint64_t>() + n_reads);
            k_scan_write<FG_SCAN_PACKED_WORDS><<<tiles, FG_SCAN_BLOCK, 0, s.stream>>>(d_read_len, n_reads, s.tile_off.as<uint64_t>(), s.pk_scan + 1,
                                                                                     s.pk_word_off.as<uint64_t>());
            FG_CUDA(cudaGetLastError());
            const emit_plan plan = enqueue_pseudoalign(x, s, a, algo, threshold, d_off, x->ev[1], x->ev[2], &x->last_launches);
            if (result_bitmaps) { /* the rows as the kernels left them: masks (<= 32 colors) or the color-set kernel's bitmaps */
                const void* rows = plan.kind == emit_plan::MASKS ? s.per_read.p : s.res_bits.p;
                FG_CUDA(cudaMemcpyAsync(d_colors, rows, size_t(n_reads) * words_per_read * 4, cudaMemcpyDeviceToDevice, s.stream));
            } else {
                enqueue_emit(s, plan, d_off, d_colors, colors_cap);
                x->last_launches += 1;
                FG_CUDA(cudaMemcpyAsync(s.h_info, s.chunk_info, 16, cudaMemcpyDeviceToHost, s.stream));
            }
            FG_CUDA(cudaEventRecord(x->ev[3], s.stream));
            FG_CUDA(cudaMemcpyAsync(s.h_info + 2, s.exhausted, 4, cudaMemcpyDeviceToHost, s.stream));
            FG_CUDA(cudaStreamSynchronize(s.stream));
            if (!uint32_t(s.h_info[2])) break;
            if (attempt >= 12) throw std::runtime_error("entry pool kept overflowing");
            x->pool_per_read *= 4;
        }
        for (int i = 0; i < 3; ++i) FG_CUDA(cudaEventElapsedTime(&x->last_ms[i], x->ev[i], x->ev[i + 1]));
        if (result_bitmaps) {
            *total_out = uint64_t(n_reads) * words_per_read;
            return 0;
        }
        *total_out = s.h_info[1];
        if (s.h_info[1] > colors_cap) return fail(FULGOR_GPU_E2BIG, "colors_cap too small; *total_out holds the required capacity");
        return 0;
    });
}

int fulgor_gpu_last_kernel_times(const fulgor_gpu_index* x, float ms[3]) {
    if (!x || !ms) return fail(FULGOR_GPU_EINVAL, "null argument");
    for (int i = 0; i < 3; ++i) ms[i] = x->last_ms[i];
    return x->last_launches;
}

}  // extern "C"
