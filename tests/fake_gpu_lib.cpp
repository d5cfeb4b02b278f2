/* TEST INFRASTRUCTURE ONLY -- never shipped, never loaded by the product.
 *
 * A test double of the C ABI (include/fulgor_gpu.h) whose compute entry points are answered by the oracle
 * (oracle/fulgor_oracle.c). Linked with the tool's host code (fulgor_b200/csrc/pseudoalign_cli.cpp) into
 * build/fulgor_cli_hosttest, it lets the CPU tier run the TOOL's host logic -- argument handling, the feeder, the three-stage
 * batch pipeline, the split of a batch over several devices and the splice of their results, the de-duplication fan-out,
 * the three output formats, the per-k-mer tools' lines -- against the same expectations as the GPU tier (tests/test_cli.py).
 * It says nothing about the kernels: those are checked on the emulator (tests/test_host_logic.py) and on the GPU.
 *
 * "Devices": FAKE_GPU_DEVICES (default 2) handles can be opened, so --gpus 2 exercises the multi-device path.
 * De-duplication: reads with the same sequence share a representative (a coarser grouping than the real one, which is
 * allowed: the ABI only promises that a read's result is found in its representative's range).
 */
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../include/fulgor_gpu.h"
#include "../oracle/fulgor_oracle.h"

struct fulgor_gpu_index {
    fo_index* o;
    int device;
    uint64_t info[8];
};

static thread_local std::string g_err;
static int fail(int code, const char* msg) {
    g_err = msg;
    return code;
}

extern "C" {

const char* fulgor_gpu_last_error(void) { return g_err.c_str(); }
const char* fulgor_gpu_version(void) { return "fake (oracle-backed test double)"; }
int fulgor_gpu_device_count(void) {
    const char* e = std::getenv("FAKE_GPU_DEVICES");
    return e ? std::atoi(e) : 2;
}
void* fulgor_gpu_host_alloc(uint64_t bytes) { return std::malloc(bytes ? bytes : 1); }
void fulgor_gpu_host_free(void* p) { std::free(p); }
int fulgor_gpu_bind_host_thread(int) { return 0; }

/* the "image" of the double is the index path */
int fulgor_gpu_image_build(const char* index_path, uint8_t** image, uint64_t* image_bytes) {
    fo_index* o = fo_open(index_path);
    if (!o) return fail(FULGOR_GPU_EIO, fo_last_error());
    fo_close(o);
    const size_t n = std::strlen(index_path) + 1;
    *image = static_cast<uint8_t*>(std::malloc(n));
    std::memcpy(*image, index_path, n);
    *image_bytes = n;
    return 0;
}
void fulgor_gpu_image_free(uint8_t* image) { std::free(image); }
int fulgor_gpu_image_info(const uint8_t*, uint64_t, fulgor_gpu_info*) { return fail(FULGOR_GPU_EINVAL, "not in the test double"); }

int fulgor_gpu_index_open(const char* index_path, int device, fulgor_gpu_index** out) {
    if (device < 0 || device >= fulgor_gpu_device_count()) return fail(FULGOR_GPU_EINVAL, "CUDA device ordinal out of range");
    fo_index* o = fo_open(index_path);
    if (!o) return fail(FULGOR_GPU_EIO, fo_last_error());
    auto* x = new fulgor_gpu_index{o, device, {0}};
    fo_info(o, x->info);
    *out = x;
    return 0;
}
int fulgor_gpu_index_open_image(const uint8_t* image, uint64_t, int device, fulgor_gpu_index** out) {
    return fulgor_gpu_index_open(reinterpret_cast<const char*>(image), device, out);
}
int fulgor_gpu_index_adopt_device_image(const void*, uint64_t, int, fulgor_gpu_index**) { return fail(FULGOR_GPU_EINVAL, "not in the test double"); }
void fulgor_gpu_index_close(fulgor_gpu_index* x) {
    if (!x) return;
    fo_close(x->o);
    delete x;
}
int fulgor_gpu_index_info(const fulgor_gpu_index* x, fulgor_gpu_info* out) {
    std::memset(out, 0, sizeof(*out));
    out->k = uint32_t(x->info[0]);
    out->m = uint32_t(x->info[1]);
    out->num_kmers = x->info[2];
    out->num_unitigs = x->info[3];
    out->num_colors = uint32_t(x->info[4]);
    out->num_color_sets = x->info[5];
    out->type = uint32_t(x->info[6]);
    out->device = x->device;
    return 0;
}

int fulgor_gpu_fetch_color_set_ids(fulgor_gpu_index* x, const char* bases, const uint64_t* read_off, uint32_t n, uint64_t* cid_off, uint32_t* cids,
                                   uint64_t cap, uint32_t* num_positive) {
    std::vector<uint64_t> rel(read_off, read_off + n + 1);
    return fo_batch_fetch_color_set_ids(x->o, bases, rel.data(), n, cid_off, cids, cap, num_positive) ? fail(FULGOR_GPU_E2BIG, "cap") : 0;
}
int fulgor_gpu_pseudoalign(fulgor_gpu_index* x, int algo, double threshold, const char* bases, const uint64_t* read_off, uint32_t n,
                           uint64_t* color_off, uint32_t* colors, uint64_t cap) {
    return fo_batch_pseudoalign(x->o, algo, threshold, bases, read_off, n, color_off, colors, cap) ? fail(FULGOR_GPU_E2BIG, "cap") : 0;
}
int fulgor_gpu_pseudoalign_dedup(fulgor_gpu_index* x, const char* bases, const uint64_t* read_off, uint32_t n, uint32_t* rep, uint64_t* color_off,
                                 uint32_t* colors, uint64_t cap) {
    /* identical sequences share the first one's result; everything else represents itself */
    std::map<std::string, uint32_t> first;
    std::vector<uint64_t> off(1, 0);
    std::string packed;
    std::vector<uint32_t> owner;
    for (uint32_t i = 0; i < n; ++i) {
        std::string s(bases + read_off[i], read_off[i + 1] - read_off[i]);
        auto it = first.find(s);
        if (it == first.end() || s.size() < x->info[0]) { /* short reads have no k-mers: they represent themselves */
            if (it == first.end()) first.emplace(s, i);
            rep[i] = i;
            owner.push_back(i);
            packed += s;
            off.push_back(packed.size());
        } else {
            rep[i] = it->second;
        }
    }
    std::vector<uint64_t> coff(owner.size() + 1);
    std::vector<uint32_t> tmp(cap ? cap : 1);
    const int rc = fo_batch_pseudoalign(x->o, 0, 1.0, packed.data(), off.data(), uint32_t(owner.size()), coff.data(), tmp.data(), cap);
    uint64_t total = 0;
    size_t j = 0;
    color_off[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (j < owner.size() && owner[j] == i) {
            const uint64_t len = coff[j + 1] - coff[j];
            if (!rc) std::memcpy(colors + total, tmp.data() + coff[j], len * 4);
            total += len;
            ++j;
        }
        color_off[i + 1] = total;
    }
    return rc ? fail(FULGOR_GPU_E2BIG, "cap") : 0;
}
int fulgor_gpu_kmer_conservation(fulgor_gpu_index* x, const char* bases, const uint64_t* read_off, uint32_t n, uint64_t* triple_off, uint32_t* triples,
                                 uint64_t cap) {
    return fo_batch_kmer_conservation(x->o, bases, read_off, n, triple_off, triples, cap) ? fail(FULGOR_GPU_E2BIG, "cap") : 0;
}
int fulgor_gpu_kmer_matches(fulgor_gpu_index* x, const char* bases, const uint64_t* read_off, uint32_t n, uint64_t* word_off, uint32_t* words,
                            uint64_t cap, uint32_t* counts) {
    std::vector<uint64_t> koff(uint64_t(n) + 1);
    std::vector<uint8_t> pos(read_off[n] - read_off[0] + 1);
    if (fo_batch_kmer_matches(x->o, bases, read_off, n, koff.data(), pos.data(), pos.size(), counts)) return fail(FULGOR_GPU_EIO, "oracle");
    word_off[0] = 0;
    for (uint32_t i = 0; i < n; ++i) word_off[i + 1] = word_off[i] + (koff[i + 1] - koff[i] + 31) / 32;
    if (word_off[n] > cap) return fail(FULGOR_GPU_E2BIG, "cap");
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t* w = words + word_off[i];
        for (uint64_t j = 0; j < word_off[i + 1] - word_off[i]; ++j) w[j] = 0;
        for (uint64_t j = 0; j < koff[i + 1] - koff[i]; ++j)
            if (pos[koff[i] + j]) w[j >> 5] |= 1u << (j & 31);
    }
    return 0;
}
int fulgor_gpu_pseudoalign_device(fulgor_gpu_index*, int, double, const char*, const uint64_t*, uint32_t, uint64_t, uint64_t*, uint32_t*, uint64_t, uint64_t*) {
    return fail(FULGOR_GPU_EINVAL, "not in the test double");
}
int fulgor_gpu_last_kernel_times(const fulgor_gpu_index*, float ms[3]) {
    ms[0] = ms[1] = ms[2] = 0;
    return 0;
}
}
