/*
 * simt_emul.h -- TEST INFRASTRUCTURE ONLY. A minimal lock-step SIMT emulator that lets g++ compile the CUDA kernels of
 * fulgor_b200/csrc/{kernels,pipeline_kernels}.cuh for the HOST, so that the kernels' warp-level logic (shuffles,
 * ballots, shared-memory staging, divergence) is checked against the oracle in the CPU-only test tier, where no GPU
 * exists. Every CUDA thread of a block is a ucontext fiber on one OS thread; a warp collective parks the calling fiber
 * until all 32 lanes of its warp have arrived (a lane that can never arrive is reported as a deadlock, a lane that
 * arrives at a different collective as divergence). Nothing in the product includes this file.
 */
#ifndef FG_SIMT_EMUL_H
#define FG_SIMT_EMUL_H

#include <stdint.h>
#include <ucontext.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define FG_SIMT_EMUL 1

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static /* one block runs at a time */

struct uint2 {
    uint32_t x, y;
};
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct alignas(16) uint4 {
    uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

namespace simt {

struct dim3e {
    unsigned x = 1, y = 1, z = 1;
};

struct fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
    unsigned tid = 0;
};

struct warp_state {
    uint64_t vals[32];
    uint64_t aux[32];
    uint64_t out[32];
    int tag = 0, arrived = 0;
    uint64_t gen = 0;
    int alive = 32;
};

struct block_state {
    std::vector<fiber> fibers;
    std::vector<warp_state> warps;
    int bar_arrived = 0;
    uint64_t bar_gen = 0;
    int alive = 0;
    uint64_t progress = 0;
    ucontext_t sched;
    fiber* cur = nullptr;
    std::function<void()> body;
    std::vector<char> dyn_smem;
};

inline block_state*& B() {
    static block_state* b = nullptr;
    return b;
}
inline dim3e& tidx() {
    static dim3e d;
    return d;
}
inline dim3e& bidx() {
    static dim3e d;
    return d;
}
inline dim3e& bdim() {
    static dim3e d;
    return d;
}
inline dim3e& gdim() {
    static dim3e d;
    return d;
}

inline void yield() {
    block_state* b = B();
    fiber* f = b->cur;
    swapcontext(&f->ctx, &b->sched);
}

inline void fiber_main() {
    block_state* b = B();
    b->body();
    fiber* f = b->cur;
    f->done = true;
    b->alive -= 1;
    b->warps[f->tid >> 5].alive -= 1;
    b->progress += 1;
    swapcontext(&f->ctx, &b->sched);
}

/* run `body` once per thread of a grid of `grid` blocks of `block` threads (blocks one after another) */
inline void launch(unsigned grid, unsigned block, size_t dyn_smem_bytes, std::function<void()> body) {
    for (unsigned bi = 0; bi < grid; ++bi) {
        block_state bs;
        B() = &bs;
        bs.body = body;
        bs.fibers.resize(block);
        bs.warps.resize((block + 31) / 32);
        bs.dyn_smem.assign(dyn_smem_bytes + 16, 0);
        bs.alive = int(block);
        gdim().x = grid;
        bdim().x = block;
        bidx().x = bi;
        for (unsigned w = 0; w < bs.warps.size(); ++w) bs.warps[w].alive = int(std::min(32u, block - 32 * w));
        for (unsigned t = 0; t < block; ++t) {
            fiber& f = bs.fibers[t];
            f.tid = t;
            f.stack.resize(256 * 1024);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack.data();
            f.ctx.uc_stack.ss_size = f.stack.size();
            f.ctx.uc_link = &bs.sched;
            makecontext(&f.ctx, (void (*)())fiber_main, 0);
        }
        while (bs.alive > 0) {
            const uint64_t before = bs.progress;
            for (unsigned t = 0; t < block; ++t) {
                fiber& f = bs.fibers[t];
                if (f.done) continue;
                bs.cur = &f;
                tidx().x = t;
                swapcontext(&bs.sched, &f.ctx);
            }
            if (bs.progress == before) {
                fprintf(stderr, "simt_emul: deadlock in block %u (a lane waits in a collective that the others never reach)\n", bi);
                abort();
            }
        }
        B() = nullptr;
    }
}

/* warp collective: deposit (v, aux), wait for the 32 lanes, let the last arriver compute every lane's result */
template <typename F>
inline uint64_t collective(int tag, uint64_t v, uint64_t aux, F&& finalize) {
    block_state* b = B();
    const unsigned tid = b->cur->tid, lane = tid & 31;
    warp_state& w = b->warps[tid >> 5];
    if (w.alive != 32 && w.alive != int(std::min<size_t>(32, b->fibers.size() - (tid & ~31u)))) {
        fprintf(stderr, "simt_emul: collective %d after some lanes of the warp exited\n", tag);
        abort();
    }
    if (w.arrived == 0) {
        w.tag = tag;
    } else if (w.tag != tag) {
        fprintf(stderr, "simt_emul: divergent collectives in one warp (%d vs %d)\n", w.tag, tag);
        abort();
    }
    w.vals[lane] = v;
    w.aux[lane] = aux;
    b->progress += 1;
    if (++w.arrived == w.alive) {
        finalize(w);
        w.arrived = 0;
        w.gen += 1;
    } else {
        const uint64_t g = w.gen;
        while (w.gen == g) yield();
    }
    /* re-establish the thread index after being resumed */
    return w.out[lane];
}

}  // namespace simt

#define threadIdx (simt::tidx())
#define blockIdx (simt::bidx())
#define blockDim (simt::bdim())
#define gridDim (simt::gdim())

static inline void* fg_emul_dynamic_smem() {
    return reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(simt::B()->dyn_smem.data()) + 15) & ~uintptr_t(15));
}

/* ---- scalar intrinsics ---- */
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return uint64_t((unsigned __int128)a * b >> 64); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __clz(int x) { return x ? __builtin_clz(unsigned(x)) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const uint64_t v = (uint64_t(y) << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = unsigned(v >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    sh &= 31;
    return unsigned(((uint64_t(hi) << 32) | lo) >> sh);
}
static inline unsigned long long __brevll(unsigned long long x) {
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x);
}
using std::max;
using std::min;

template <typename T>
static inline T atomicAdd(T* p, T v) { /* fibers are cooperative: no preemption inside */
    const T old = *p;
    *p = old + v;
    return old;
}

template <typename T>
static inline T atomicAnd(T* p, T v) {
    const T old = *p;
    *p = old & v;
    return old;
}
template <typename T>
static inline T atomicOr(T* p, T v) {
    const T old = *p;
    *p = old | v;
    return old;
}
template <typename T>
static inline T atomicXor(T* p, T v) {
    const T old = *p;
    *p = old ^ v;
    return old;
}
template <typename T>
static inline T atomicCAS(T* p, T expected, T desired) {
    const T old = *p;
    if (old == expected) *p = desired;
    return old;
}
template <typename T>
static inline T atomicMax(T* p, T v) {
    const T old = *p;
    if (v > old) *p = v;
    return old;
}

/* ---- warp collectives (full mask only, like every call site in the kernels) ---- */
static inline void fg_emul_check_mask(unsigned mask) {
    if (mask != 0xffffffffu) {
        fprintf(stderr, "simt_emul: only full-mask collectives are emulated\n");
        abort();
    }
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    fg_emul_check_mask(mask);
    simt::collective(1, 0, 0, [](simt::warp_state&) {});
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    fg_emul_check_mask(mask);
    return unsigned(simt::collective(2, pred ? 1 : 0, 0, [](simt::warp_state& w) {
        uint64_t b = 0;
        for (int l = 0; l < 32; ++l) b |= (w.vals[l] & 1) << l;
        for (int l = 0; l < 32; ++l) w.out[l] = b;
    }));
}
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src) {
    fg_emul_check_mask(mask);
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    raw = simt::collective(3, raw, uint64_t(src & 31), [](simt::warp_state& w) {
        for (int l = 0; l < 32; ++l) w.out[l] = w.vals[w.aux[l]];
    });
    T r;
    std::memcpy(&r, &raw, sizeof(T));
    return r;
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
    fg_emul_check_mask(mask);
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    raw = simt::collective(4, raw, delta, [](simt::warp_state& w) {
        for (int l = 0; l < 32; ++l) w.out[l] = l >= int(w.aux[l]) ? w.vals[l - int(w.aux[l])] : w.vals[l];
    });
    T r;
    std::memcpy(&r, &raw, sizeof(T));
    return r;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask) {
    fg_emul_check_mask(mask);
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    raw = simt::collective(5, raw, uint64_t(lanemask), [](simt::warp_state& w) {
        for (int l = 0; l < 32; ++l) w.out[l] = w.vals[(l ^ int(w.aux[l])) & 31];
    });
    T r;
    std::memcpy(&r, &raw, sizeof(T));
    return r;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
    fg_emul_check_mask(mask);
    return unsigned(simt::collective(6, v, 0, [](simt::warp_state& w) {
        uint64_t r = 0;
        for (int l = 0; l < 32; ++l) r |= w.vals[l];
        for (int l = 0; l < 32; ++l) w.out[l] = r;
    }));
}
static inline unsigned __reduce_and_sync(unsigned mask, unsigned v) {
    fg_emul_check_mask(mask);
    return unsigned(simt::collective(7, v, 0, [](simt::warp_state& w) {
        uint64_t r = ~0ULL;
        for (int l = 0; l < 32; ++l) r &= w.vals[l];
        for (int l = 0; l < 32; ++l) w.out[l] = r;
    }));
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    fg_emul_check_mask(mask);
    return unsigned(simt::collective(9, v, 0, [](simt::warp_state& w) {
        uint64_t r = 0;
        for (int l = 0; l < 32; ++l) r += w.vals[l];
        for (int l = 0; l < 32; ++l) w.out[l] = r & 0xffffffffu;
    }));
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    fg_emul_check_mask(mask);
    return unsigned(simt::collective(21, v, 0, [](simt::warp_state& w) {
        uint64_t r = 0;
        for (int l = 0; l < 32; ++l) r = w.vals[l] > r ? w.vals[l] : r;
        for (int l = 0; l < 32; ++l) w.out[l] = r;
    }));
}
static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
    fg_emul_check_mask(mask);
    return unsigned(simt::collective(8, v, 0, [](simt::warp_state& w) {
        for (int l = 0; l < 32; ++l) {
            uint64_t m = 0;
            for (int j = 0; j < 32; ++j) m |= uint64_t(w.vals[j] == w.vals[l]) << j;
            w.out[l] = m;
        }
    }));
}
static inline void __syncthreads() {
    simt::block_state* b = simt::B();
    b->progress += 1;
    if (++b->bar_arrived == b->alive) {
        b->bar_arrived = 0;
        b->bar_gen += 1;
    } else {
        const uint64_t g = b->bar_gen;
        while (b->bar_gen == g) simt::yield();
    }
}

#endif
