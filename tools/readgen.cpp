/*
 * readgen -- synthetic read generator (test/bench tooling, not on the query path).
 *
 * Implements the generator fixed by SURVEY.md section 8(d): each read is a window of one of the
 * packed genomes (uniform over all valid windows of all contigs), with independent 1 % substitutions
 * by a different base and a 50 % chance of being reverse-complemented.
 *
 * Determinism: read i of a stream with seed S uses its own splitmix64 generator seeded with
 * S ^ (0x9E3779B97F4A7C15 * (i + 1)), so any sub-range [first, first+n) can be produced
 * independently (multi-threaded here, or one range per GPU rank) and always yields the same reads.
 *
 * Input: BASE.gpk written by tools/mkdump (2-bit packed genomes + non-ACGT exception list).
 *
 *   library:  fg_readgen(...)                       (ctypes from bench.py / tests)
 *   CLI:      readgen genomes.gpk N out.fastq [--min-len 150 --max-len 150 --seed 42 --sub 0.01 --first 0]
 */
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct splitmix64 {
    uint64_t s;
    explicit splitmix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    /* unbiased enough for test data: 64-bit multiply-high range reduction */
    uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
    double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

struct genomes {
    uint64_t num_contigs = 0, total = 0;
    std::vector<uint64_t> begin; /* num_contigs + 1 global offsets */
    const uint8_t* packed = nullptr;
    std::vector<uint64_t> exceptions;
    bool parse(const uint8_t* p, uint64_t size) {
        if (size < 24 || memcmp(p, "FGPK1\0\0\0", 8)) return false;
        memcpy(&num_contigs, p + 8, 8);
        memcpy(&total, p + 16, 8);
        uint64_t off = 24;
        if (size < off + num_contigs * 12) return false;
        begin.assign(num_contigs + 1, 0);
        for (uint64_t i = 0; i < num_contigs; ++i) {
            uint64_t len;
            memcpy(&len, p + off + 4, 8);
            begin[i + 1] = begin[i] + len;
            off += 12;
        }
        if (begin[num_contigs] != total) return false;
        packed = p + off;
        off += (total + 3) / 4;
        if (size < off + 8) return false;
        uint64_t ne;
        memcpy(&ne, p + off, 8);
        off += 8;
        if (size < off + ne * 8) return false;
        exceptions.resize(ne);
        if (ne) memcpy(exceptions.data(), p + off, ne * 8);
        return true;
    }
    char base(uint64_t g) const { return "ACGT"[(packed[g >> 2] >> (2 * (g & 3))) & 3]; }
};

inline char complement(char c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; }
}

inline uint64_t read_seed(uint64_t seed, uint64_t i) { return seed ^ (0x9E3779B97F4A7C15ULL * (i + 1)); }

inline uint32_t read_length(uint64_t seed, uint64_t i, uint32_t min_len, uint32_t max_len) {
    if (min_len == max_len) return min_len;
    splitmix64 r(read_seed(seed, i));
    return min_len + (uint32_t)r.below((uint64_t)(max_len - min_len) + 1);
}

void make_read(const genomes& g, uint64_t seed, uint64_t i, uint32_t min_len, uint32_t max_len,
               double sub_rate, char* out) {
    splitmix64 r(read_seed(seed, i));
    uint32_t len = min_len;
    if (min_len != max_len) len = min_len + (uint32_t)r.below((uint64_t)(max_len - min_len) + 1);
    /* uniform over valid windows: draw a global position, reject if the window crosses a contig end */
    uint64_t start = 0;
    for (;;) {
        start = r.below(g.total);
        uint64_t c = (uint64_t)(std::upper_bound(g.begin.begin(), g.begin.end(), start) - g.begin.begin()) - 1;
        if (start + len <= g.begin[c + 1]) break;
    }
    for (uint32_t j = 0; j < len; ++j) out[j] = g.base(start + j);
    if (!g.exceptions.empty()) {
        auto it = std::lower_bound(g.exceptions.begin(), g.exceptions.end(), start);
        for (; it != g.exceptions.end() && *it < start + len; ++it) out[*it - start] = 'N';
    }
    for (uint32_t j = 0; j < len; ++j) {
        if (r.unit() < sub_rate && out[j] != 'N') {
            const char* alphabet = "ACGT";
            int cur = (int)(strchr(alphabet, out[j]) - alphabet);
            out[j] = alphabet[(cur + 1 + (int)r.below(3)) & 3];
        }
    }
    if (r.next() & 1) {
        std::reverse(out, out + len);
        for (uint32_t j = 0; j < len; ++j) out[j] = complement(out[j]);
    }
}

}  // namespace

extern "C" {

/* Fills read_off[0..n] (byte offsets into bases, read_off[0] = 0) and, if bases != NULL, the bases.
   Call once with bases == NULL to learn read_off[n] for mixed lengths. Returns 0, or -1 on a bad .gpk. */
int fg_readgen(const uint8_t* gpk, uint64_t gpk_size, uint64_t first, uint64_t n, uint32_t min_len,
               uint32_t max_len, double sub_rate, uint64_t seed, uint64_t* read_off, char* bases,
               int num_threads) {
    genomes g;
    if (!g.parse(gpk, gpk_size) || min_len == 0 || max_len < min_len) return -1;
    read_off[0] = 0;
    for (uint64_t i = 0; i < n; ++i) read_off[i + 1] = read_off[i] + read_length(seed, first + i, min_len, max_len);
    if (!bases) return 0;
    if (num_threads < 1) num_threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < num_threads; ++t)
        pool.emplace_back([&, t]() {
            uint64_t lo = n * (uint64_t)t / num_threads, hi = n * (uint64_t)(t + 1) / num_threads;
            for (uint64_t i = lo; i < hi; ++i)
                make_read(g, seed, first + i, min_len, max_len, sub_rate, bases + read_off[i]);
        });
    for (auto& th : pool) th.join();
    return 0;
}

}  // extern "C"

#ifdef READGEN_MAIN
int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: readgen genomes.gpk N out.fastq [--min-len L --max-len L --seed S --sub P --first I --fasta]\n");
        return 1;
    }
    uint64_t n = strtoull(argv[2], nullptr, 10), seed = 42, first = 0;
    uint32_t min_len = 150, max_len = 150;
    double sub = 0.01;
    bool fasta = false;
    for (int a = 4; a < argc; ++a) {
        if (!strcmp(argv[a], "--fasta")) { fasta = true; continue; }
        if (a + 1 >= argc) break;
        if (!strcmp(argv[a], "--min-len")) min_len = (uint32_t)atoi(argv[++a]);
        else if (!strcmp(argv[a], "--max-len")) max_len = (uint32_t)atoi(argv[++a]);
        else if (!strcmp(argv[a], "--seed")) seed = strtoull(argv[++a], nullptr, 10);
        else if (!strcmp(argv[a], "--sub")) sub = atof(argv[++a]);
        else if (!strcmp(argv[a], "--first")) first = strtoull(argv[++a], nullptr, 10);
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { fprintf(stderr, "readgen: cannot open %s\n", argv[1]); return 1; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> gpk((size_t)sz);
    if (fread(gpk.data(), 1, (size_t)sz, f) != (size_t)sz) return 1;
    fclose(f);
    std::vector<uint64_t> off(n + 1);
    if (fg_readgen(gpk.data(), gpk.size(), first, n, min_len, max_len, sub, seed, off.data(), nullptr, 1)) return 1;
    std::vector<char> bases(off[n]);
    fg_readgen(gpk.data(), gpk.size(), first, n, min_len, max_len, sub, seed, off.data(), bases.data(),
               (int)std::max(1u, std::thread::hardware_concurrency()));
    FILE* o = fopen(argv[3], "w");
    if (!o) { fprintf(stderr, "readgen: cannot write %s\n", argv[3]); return 1; }
    std::vector<char> qual(max_len, 'I');
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t len = off[i + 1] - off[i];
        if (fasta) {
            fprintf(o, ">r%lu\n%.*s\n", (unsigned long)(first + i), (int)len, bases.data() + off[i]);
        } else {
            fprintf(o, "@r%lu\n%.*s\n+\n%.*s\n", (unsigned long)(first + i), (int)len, bases.data() + off[i], (int)len, qual.data());
        }
    }
    fclose(o);
    return 0;
}
#endif
