/*
 * engine.cu -- kernels' global entry points, the chunked H2D / compute / D2H pipeline and the C ABI
 * declared in include/fulgor_gpu.h. sm_100a only.
 *
 * Per chunk of reads (host API) the stream order is
 *     H2D(bases, read_off) -> K1(+K2 fused when num_colors <= 32) -> [K2] -> scan -> emit -> D2H
 * with two slots so that the copies of one chunk overlap the kernels of the other. The only
 * cross-chunk dependency is the running CSR offset, carried in device memory.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fulgor_gpu.h"
#include "fur_reader.h"
#include "kernels.cuh"

namespace fgb {

/* ================================================================== kernels */

#define FG_WARPS_PER_BLOCK 8
#define FG_BLOCK (FG_WARPS_PER_BLOCK * 32)
#define FG_STAGE_STRIDE FG_MAX_ENTRIES

/* K1 + fused K2 for indexes with at most 32 colors: each read's result is one 32-bit color mask.
   Full intersection (src/ps_full_intersection.cpp:377-400 -> intersect :33-127): AND of the hit sets.
   Threshold union (src/ps_threshold_union.cpp:389 + merge :17-40 / merge_meta :43-120): color c is
   reported iff sum of multiplicities of the hit sets containing c >= uint64(double(npos) * threshold). */
template <int W>
__global__ void __launch_bounds__(FG_BLOCK) k_pseudoalign_small(const __grid_constant__ dev_index I, const uint8_t* __restrict__ bases,
                                                               const uint64_t* __restrict__ read_off, uint64_t read_off_base,
                                                               uint32_t n_reads, int algo, double threshold,
                                                               uint32_t* __restrict__ masks, uint32_t* __restrict__ overflow_count) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * FG_WARPS_PER_BLOCK;
    for (uint32_t r = blockIdx.x * FG_WARPS_PER_BLOCK + (threadIdx.x >> 5); r < n_reads; r += warps) {
        const uint64_t beg = __ldg(read_off + r), end = __ldg(read_off + r + 1);
        const read_hits R = warp_fetch_color_sets<W>(I, bases + (beg - read_off_base), uint32_t(end - beg), lane);
        if (R.overflow) {
            if (lane == 0) {
                atomicAdd(overflow_count, 1u);
                masks[r] = 0;
            }
            continue;
        }
        uint32_t res = 0;
        if (R.n) {
            const uint32_t my = lane < R.n ? color_set_mask(I, R.cid) : 0u;
            __syncwarp();
            if (algo == FULGOR_GPU_FULL_INTERSECTION) {
                res = __reduce_and_sync(FG_FULL, lane < R.n ? my : ~0u);
            } else {
                uint32_t score = 0;
                for (uint32_t j = 0; j < R.n; ++j) {
                    const uint32_t mj = __shfl_sync(FG_FULL, my, j);
                    const uint32_t wj = __shfl_sync(FG_FULL, R.cnt, j);
                    score += ((mj >> lane) & 1u) ? wj : 0u;
                }
                const uint64_t min_score = uint64_t(double(R.npos) * threshold);
                res = __ballot_sync(FG_FULL, lane < I.num_colors && uint64_t(score) >= min_score);
            }
        }
        if (lane == 0) masks[r] = res;
    }
}

/* K1 alone: per read, ascending distinct color-set ids (+ multiplicities) into a fixed-stride stage */
template <int W>
__global__ void __launch_bounds__(FG_BLOCK) k_fetch_color_sets(const __grid_constant__ dev_index I, const uint8_t* __restrict__ bases,
                                                              const uint64_t* __restrict__ read_off, uint64_t read_off_base,
                                                              uint32_t n_reads, uint32_t* __restrict__ stage_cid,
                                                              uint32_t* __restrict__ stage_cnt /* nullable */,
                                                              uint32_t* __restrict__ counts, uint32_t* __restrict__ num_positive /* nullable */,
                                                              uint32_t* __restrict__ overflow_count) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * FG_WARPS_PER_BLOCK;
    for (uint32_t r = blockIdx.x * FG_WARPS_PER_BLOCK + (threadIdx.x >> 5); r < n_reads; r += warps) {
        const uint64_t beg = __ldg(read_off + r), end = __ldg(read_off + r + 1);
        const read_hits R = warp_fetch_color_sets<W>(I, bases + (beg - read_off_base), uint32_t(end - beg), lane);
        if (R.overflow && lane == 0) atomicAdd(overflow_count, 1u);
        if (lane < R.n) {
            stage_cid[uint64_t(r) * FG_STAGE_STRIDE + lane] = R.cid;
            if (stage_cnt) stage_cnt[uint64_t(r) * FG_STAGE_STRIDE + lane] = R.cnt;
        }
        if (lane == 0) {
            counts[r] = R.overflow ? 0u : R.n;
            if (num_positive) num_positive[r] = R.npos;
        }
    }
}

/* ---- CSR offsets: exclusive scan of per-read counts (three small kernels) ---- */
#define FG_SCAN_ITEMS 8
#define FG_SCAN_BLOCK 256
#define FG_SCAN_TILE (FG_SCAN_ITEMS * FG_SCAN_BLOCK)

template <bool POPC>
__device__ __forceinline__ uint32_t count_of(const uint32_t* __restrict__ in, uint32_t i) {
    const uint32_t v = __ldg(in + i);
    return POPC ? uint32_t(__popc(v)) : v;
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

template <bool POPC>
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

template <bool POPC>
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

/* staged fixed-stride lists -> CSR */
__global__ void __launch_bounds__(256) k_emit_stage(const uint32_t* __restrict__ stage, const uint32_t* __restrict__ counts,
                                                   const uint64_t* __restrict__ off, const uint64_t* __restrict__ chunk_info, uint32_t n,
                                                   uint32_t* __restrict__ out, uint64_t out_cap) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint32_t c = __ldg(counts + r);
    const uint64_t o = __ldg(off + r) - chunk_info[0];
    if (lane < c && o + lane < out_cap) out[o + lane] = __ldg(stage + uint64_t(r) * FG_STAGE_STRIDE + lane);
}

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

struct slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t scanned = nullptr, done = nullptr, info_ready = nullptr;
    dev_buffer bases, read_off, per_read /* masks or counts */, stage_cid, stage_cnt, npos, tile_sums, tile_off, off, out;
    uint64_t* chunk_info = nullptr;   /* device: {base, total} */
    uint64_t* h_info = nullptr;       /* pinned host: {base, total, overflow_count} */
    uint32_t* overflow = nullptr;     /* device counter */
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
    slot slots[2];
    uint64_t* d_carry = nullptr;
    int sm_count = 0;
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
    for (int i = 0; i < FGI_MAX_SKEW; ++i) {
        I.skew_phf[i] = H.skew_phf[i];
        I.skew_pos_base[i] = H.skew_pos_base[i];
    }
    I.type = H.type;
    I.num_colors = H.num_colors;
    I.num_partitions = H.num_partitions;
    return I;
}

static void use_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) throw std::invalid_argument(std::string("no usable CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) throw std::invalid_argument("CUDA device ordinal out of range");
    FG_CUDA(cudaSetDevice(device));
}

static fulgor_gpu_index* make_handle(const fgi_header& H, void* d_image, bool owns, int device) {
    auto* x = new fulgor_gpu_index();
    x->device = device;
    x->H = H;
    x->d_image = d_image;
    x->owns_image = owns;
    x->I = make_view(H, static_cast<const uint8_t*>(d_image));
    try {
        cudaDeviceProp prop;
        FG_CUDA(cudaGetDeviceProperties(&prop, device));
        x->sm_count = prop.multiProcessorCount;
        FG_CUDA(cudaMalloc(&x->d_carry, 8));
        for (auto& s : x->slots) {
            FG_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            FG_CUDA(cudaEventCreateWithFlags(&s.scanned, cudaEventDisableTiming));
            FG_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
            FG_CUDA(cudaEventCreateWithFlags(&s.info_ready, cudaEventDisableTiming));
            FG_CUDA(cudaMalloc(&s.chunk_info, 16));
            FG_CUDA(cudaMalloc(&s.overflow, 4));
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
    switch (H.k - H.m + 1) { /* common (k, m) pairs get unrolled minimizer loops */
        case 13: f(std::integral_constant<int, 13>()); break; /* k=31, m=19 */
        case 12: f(std::integral_constant<int, 12>()); break; /* k=31, m=20 */
        case 11: f(std::integral_constant<int, 11>()); break; /* k=31, m=21 */
        default: f(std::integral_constant<int, 0>()); break;
    }
}

static uint32_t read_grid(const fulgor_gpu_index* x, uint32_t n_reads) {
    const uint64_t blocks_needed = (uint64_t(n_reads) + FG_WARPS_PER_BLOCK - 1) / FG_WARPS_PER_BLOCK;
    const uint64_t resident = uint64_t(x->sm_count) * 8; /* a multiple of the SM count; warps stride over reads */
    return uint32_t(std::max<uint64_t>(1, std::min(blocks_needed, resident)));
}

struct chunk_args {
    const uint8_t* d_bases;
    const uint64_t* d_read_off;
    uint64_t read_off_base;
    uint32_t n;
};

/* enqueue scan (counts -> CSR offsets with the cross-chunk carry) */
template <bool POPC>
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

/* full path on device-resident chunk; writes CSR offsets to d_off (n+1) and colors to d_out (chunk-local) */
static int enqueue_pseudoalign(fulgor_gpu_index* x, slot& s, const chunk_args& a, int algo, double threshold, uint64_t* d_off,
                               uint32_t* d_out, uint64_t out_cap, cudaEvent_t after_k1, cudaEvent_t after_k2) {
    int launches = 0;
    if (x->H.num_colors > 32) throw std::runtime_error("indexes with more than 32 colors: color-set kernel not built yet");
    s.per_read.reserve(size_t(a.n) * 4);
    FG_CUDA(cudaMemsetAsync(s.overflow, 0, 4, s.stream));
    const uint32_t grid = read_grid(x, a.n);
    dispatch_window(x->H, [&](auto w) {
        k_pseudoalign_small<decltype(w)::value><<<grid, FG_BLOCK, 0, s.stream>>>(x->I, a.d_bases, a.d_read_off, a.read_off_base, a.n, algo,
                                                                                   threshold, s.per_read.as<uint32_t>(), s.overflow);
    });
    ++launches;
    if (after_k1) FG_CUDA(cudaEventRecord(after_k1, s.stream));
    if (after_k2) FG_CUDA(cudaEventRecord(after_k2, s.stream));
    launches += enqueue_scan<true>(x, s, s.per_read.as<uint32_t>(), a.n, d_off);
    k_emit_masks<<<(a.n + 255) / 256, 256, 0, s.stream>>>(s.per_read.as<uint32_t>(), d_off, s.chunk_info, a.n, d_out, out_cap);
    ++launches;
    FG_CUDA(cudaGetLastError());
    return launches;
}

static int enqueue_fetch(fulgor_gpu_index* x, slot& s, const chunk_args& a, bool want_npos, uint64_t* d_off, uint32_t* d_out, uint64_t out_cap) {
    int launches = 0;
    s.per_read.reserve(size_t(a.n) * 4);
    s.stage_cid.reserve(size_t(a.n) * FG_STAGE_STRIDE * 4);
    if (want_npos) s.npos.reserve(size_t(a.n) * 4);
    FG_CUDA(cudaMemsetAsync(s.overflow, 0, 4, s.stream));
    const uint32_t grid = read_grid(x, a.n);
    dispatch_window(x->H, [&](auto w) {
        k_fetch_color_sets<decltype(w)::value><<<grid, FG_BLOCK, 0, s.stream>>>(x->I, a.d_bases, a.d_read_off, a.read_off_base, a.n,
                                                                                  s.stage_cid.as<uint32_t>(), nullptr, s.per_read.as<uint32_t>(),
                                                                                  want_npos ? s.npos.as<uint32_t>() : nullptr, s.overflow);
    });
    ++launches;
    launches += enqueue_scan<false>(x, s, s.per_read.as<uint32_t>(), a.n, d_off);
    k_emit_stage<<<(uint64_t(a.n) * 32 + 255) / 256, 256, 0, s.stream>>>(s.stage_cid.as<uint32_t>(), s.per_read.as<uint32_t>(), d_off, s.chunk_info,
                                                                       a.n, d_out, out_cap);
    ++launches;
    FG_CUDA(cudaGetLastError());
    return launches;
}

/* ---- the chunked host pipeline ---- */

static const uint64_t CHUNK_MAX_READS = 1u << 20;
static const uint64_t CHUNK_MAX_BASES = 256ull << 20;

enum class op_kind { FETCH, PSEUDOALIGN };

static int run_host_batch(fulgor_gpu_index* x, op_kind op, int algo, double threshold, const char* bases, const uint64_t* read_off,
                          uint32_t n_reads, uint64_t* out_off, uint32_t* out_vals, uint64_t cap, uint32_t* num_positive) {
    FG_CUDA(cudaSetDevice(x->device));
    out_off[0] = 0;
    if (n_reads == 0) return 0;
    for (uint32_t i = 0; i < n_reads; ++i) {
        if (read_off[i + 1] < read_off[i]) throw std::invalid_argument("read_off must be non-decreasing");
        if (read_off[i + 1] - read_off[i] >= (1ull << 31)) throw std::invalid_argument("reads of 2^31 characters or more are not supported");
    }
    FG_CUDA(cudaMemsetAsync(x->d_carry, 0, 8, x->slots[0].stream));
    FG_CUDA(cudaStreamSynchronize(x->slots[0].stream));

    const uint32_t max_vals_per_read = op == op_kind::FETCH ? FG_MAX_ENTRIES : x->H.num_colors;
    struct pending { uint32_t first, n; int slot; };
    std::vector<pending> chunks;
    for (uint32_t first = 0; first < n_reads;) {
        uint32_t n = 1;
        while (first + n < n_reads && n < CHUNK_MAX_READS && read_off[first + n + 1] - read_off[first] <= CHUNK_MAX_BASES) ++n;
        chunks.push_back({first, n, int(chunks.size() & 1)});
        first += n;
    }
    bool too_big = false;
    uint64_t overflow_reads = 0;
    cudaEvent_t prev_scanned = nullptr;

    auto finalize = [&](const pending& c) {
        slot& s = x->slots[c.slot];
        FG_CUDA(cudaEventSynchronize(s.info_ready));
        const uint64_t base = s.h_info[0], total = s.h_info[1];
        overflow_reads += s.h_info[2];
        if (base + total > cap) too_big = true;
        if (!too_big && total) FG_CUDA(cudaMemcpyAsync(out_vals + base, s.out.p, total * 4, cudaMemcpyDeviceToHost, s.stream));
        FG_CUDA(cudaEventRecord(s.done, s.stream));
    };

    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        const pending& c = chunks[ci];
        slot& s = x->slots[c.slot];
        if (s.busy) FG_CUDA(cudaEventSynchronize(s.done));
        s.busy = true;
        const uint64_t b0 = read_off[c.first], b1 = read_off[c.first + c.n];
        s.bases.reserve(size_t(b1 - b0) + 64);
        s.read_off.reserve(size_t(c.n + 1) * 8);
        s.off.reserve(size_t(c.n + 1) * 8);
        s.out.reserve(size_t(c.n) * max_vals_per_read * 4);
        if (b1 > b0) FG_CUDA(cudaMemcpyAsync(s.bases.p, bases + b0, b1 - b0, cudaMemcpyHostToDevice, s.stream));
        FG_CUDA(cudaMemcpyAsync(s.read_off.p, read_off + c.first, size_t(c.n + 1) * 8, cudaMemcpyHostToDevice, s.stream));
        if (prev_scanned) FG_CUDA(cudaStreamWaitEvent(s.stream, prev_scanned, 0));
        chunk_args a{s.bases.as<uint8_t>(), s.read_off.as<uint64_t>(), b0, c.n};
        const uint64_t out_cap = uint64_t(c.n) * max_vals_per_read;
        if (op == op_kind::FETCH) {
            enqueue_fetch(x, s, a, num_positive != nullptr, s.off.as<uint64_t>(), s.out.as<uint32_t>(), out_cap);
        } else {
            enqueue_pseudoalign(x, s, a, algo, threshold, s.off.as<uint64_t>(), s.out.as<uint32_t>(), out_cap, nullptr, nullptr);
        }
        FG_CUDA(cudaEventRecord(s.scanned, s.stream));
        prev_scanned = s.scanned;
        FG_CUDA(cudaMemcpyAsync(s.h_info, s.chunk_info, 16, cudaMemcpyDeviceToHost, s.stream));
        FG_CUDA(cudaMemcpyAsync(s.h_info + 2, s.overflow, 4, cudaMemcpyDeviceToHost, s.stream));
        FG_CUDA(cudaEventRecord(s.info_ready, s.stream));
        /* offsets of reads [first, first+n) and the running total at [first+n] (overwritten by the next chunk's first entry with the same value) */
        FG_CUDA(cudaMemcpyAsync(out_off + c.first, s.off.p, size_t(c.n + 1) * 8, cudaMemcpyDeviceToHost, s.stream));
        if (op == op_kind::FETCH && num_positive)
            FG_CUDA(cudaMemcpyAsync(num_positive + c.first, s.npos.p, size_t(c.n) * 4, cudaMemcpyDeviceToHost, s.stream));
        if (ci > 0) finalize(chunks[ci - 1]);
    }
    finalize(chunks.back());
    for (auto& s : x->slots) {
        FG_CUDA(cudaStreamSynchronize(s.stream));
        s.busy = false;
    }
    if (overflow_reads)
        throw std::runtime_error(std::to_string(overflow_reads) + " read(s) hit more than " + std::to_string(FG_MAX_ENTRIES) +
                                 " distinct color sets: not supported yet");
    return too_big ? FULGOR_GPU_E2BIG : 0;
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
        for (dev_buffer* b : {&s.bases, &s.read_off, &s.per_read, &s.stage_cid, &s.stage_cnt, &s.npos, &s.tile_sums, &s.tile_off, &s.off, &s.out})
            b->release();
        if (s.chunk_info) cudaFree(s.chunk_info);
        if (s.overflow) cudaFree(s.overflow);
        if (s.h_info) cudaFreeHost(s.h_info);
        if (s.scanned) cudaEventDestroy(s.scanned);
        if (s.done) cudaEventDestroy(s.done);
        if (s.info_ready) cudaEventDestroy(s.info_ready);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    for (auto& e : x->ev)
        if (e) cudaEventDestroy(e);
    if (x->d_carry) cudaFree(x->d_carry);
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
        int rc = run_host_batch(x, op_kind::FETCH, 0, 0.0, bases, read_off, n_reads, cid_off, cids, cids_cap, num_positive);
        if (rc == FULGOR_GPU_E2BIG) return fail(rc, "cids_cap too small; cid_off[n_reads] holds the required capacity");
        return rc;
    });
}

int fulgor_gpu_pseudoalign(fulgor_gpu_index* x, int algo, double threshold, const char* bases, const uint64_t* read_off, uint32_t n_reads,
                           uint64_t* color_off, uint32_t* colors, uint64_t colors_cap) {
    return guarded([&]() -> int {
        if (!x || !read_off || !color_off || (!bases && n_reads && read_off[n_reads] > read_off[0]) || (!colors && colors_cap))
            throw std::invalid_argument("null argument");
        if (algo != FULGOR_GPU_FULL_INTERSECTION && algo != FULGOR_GPU_THRESHOLD_UNION) throw std::invalid_argument("unknown algorithm");
        /* same domain check as the reference CLI (tools/pseudoalign.cpp:272-281) */
        if (algo == FULGOR_GPU_THRESHOLD_UNION && !(threshold > 0.0 && threshold <= 1.0))
            throw std::invalid_argument("threshold must be a float in (0.0,1.0]");
        int rc = run_host_batch(x, op_kind::PSEUDOALIGN, algo, threshold, bases, read_off, n_reads, color_off, colors, colors_cap, nullptr);
        if (rc == FULGOR_GPU_E2BIG) return fail(rc, "colors_cap too small; color_off[n_reads] holds the required capacity");
        return rc;
    });
}

int fulgor_gpu_pseudoalign_device(fulgor_gpu_index* x, int algo, double threshold, const char* d_bases, const uint64_t* d_read_off,
                                  uint32_t n_reads, uint64_t read_off_base, uint64_t* d_color_off, uint32_t* d_colors, uint64_t colors_cap,
                                  uint64_t* total_out) {
    return guarded([&]() -> int {
        if (!x || !d_read_off || !d_color_off || !total_out) throw std::invalid_argument("null argument");
        if (algo != FULGOR_GPU_FULL_INTERSECTION && algo != FULGOR_GPU_THRESHOLD_UNION) throw std::invalid_argument("unknown algorithm");
        if (algo == FULGOR_GPU_THRESHOLD_UNION && !(threshold > 0.0 && threshold <= 1.0))
            throw std::invalid_argument("threshold must be a float in (0.0,1.0]");
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
        chunk_args a{reinterpret_cast<const uint8_t*>(d_bases), d_read_off, read_off_base, n_reads};
        FG_CUDA(cudaEventRecord(x->ev[0], s.stream));
        x->last_launches = enqueue_pseudoalign(x, s, a, algo, threshold, d_color_off, d_colors, colors_cap, x->ev[1], x->ev[2]);
        FG_CUDA(cudaEventRecord(x->ev[3], s.stream));
        FG_CUDA(cudaMemcpyAsync(s.h_info, s.chunk_info, 16, cudaMemcpyDeviceToHost, s.stream));
        FG_CUDA(cudaMemcpyAsync(s.h_info + 2, s.overflow, 4, cudaMemcpyDeviceToHost, s.stream));
        FG_CUDA(cudaStreamSynchronize(s.stream));
        for (int i = 0; i < 3; ++i) FG_CUDA(cudaEventElapsedTime(&x->last_ms[i], x->ev[i], x->ev[i + 1]));
        *total_out = s.h_info[1];
        if (uint32_t(s.h_info[2]))
            throw std::runtime_error(std::to_string(uint32_t(s.h_info[2])) + " read(s) hit more than " + std::to_string(FG_MAX_ENTRIES) +
                                     " distinct color sets: not supported yet");
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
