/* TEST INFRASTRUCTURE: exposes fulgor_b200/csrc/fastx_io.h (the tool's feeder and formatters, host C++ without CUDA) to
   the CPU-only test tier through a C interface. Built into build/libfg_fastx_io_test.so by tests/test_fastx_io.py. */
#include "../fulgor_b200/csrc/fastx_io.h"

#include <cstdlib>

extern "C" {

/* parses a whole file; *bases / *off are malloc'ed (free with fxio_free). Returns the number of reads, -1 on error. */
long long fxio_parse(const char* path, unsigned threads, unsigned long long span, unsigned long long max_reads_serial, char** bases,
                     unsigned long long** off, int* was_mapped) {
    fgio::fastx_source src;
    if (!src.open(path, threads, span, max_reads_serial)) return -1;
    *was_mapped = src.mapped();
    std::vector<char> all;
    std::vector<unsigned long long> offs(1, 0);
    fgio::read_batch b;
    std::vector<char> bb;
    std::vector<uint64_t> bo;
    b.grow = [&](fgio::read_batch& x, uint64_t nb, uint64_t nr) {
        bb.resize(nb + 16);
        bo.resize(nr + 16);
        x.bases = bb.data();
        x.off = bo.data();
        x.bases_cap = bb.size();
        x.reads_cap = bo.size();
    };
    unsigned batch = 0;
    while (src.next_batch(b)) {
        const unsigned long long base = all.size();
        all.insert(all.end(), b.bases, b.bases + b.off[b.n]);
        for (uint32_t i = 1; i <= b.n; ++i) offs.push_back(base + b.off[i]);
        src.set_threads(1 + (++batch * 3) % (threads + 2)); /* the tool changes the thread count between batches */
    }
    *was_mapped = *was_mapped && src.mapped();
    *bases = static_cast<char*>(std::malloc(all.size() + 1));
    std::memcpy(*bases, all.data(), all.size());
    *off = static_cast<unsigned long long*>(std::malloc(offs.size() * 8));
    std::memcpy(*off, offs.data(), offs.size() * 8);
    return (long long)(offs.size() - 1);
}
void fxio_free(void* p) { std::free(p); }

/* fgio::thread_team: `calls` run() calls of varying width on one team; every f(t) must run exactly once per call and run() must
   not return before all of them have. Returns 0 when every call saw exactly its own width of distinct indices. */
int fxio_team_selftest(unsigned calls, unsigned max_width) {
    fgio::thread_team team;
    std::vector<unsigned> hits(max_width + 1);
    unsigned long long state = 88172645463325252ull;
    for (unsigned c = 0; c < calls; ++c) {
        state ^= state << 13, state ^= state >> 7, state ^= state << 17;
        const unsigned n = unsigned(state % (max_width + 1)); /* 0 .. max_width */
        std::fill(hits.begin(), hits.end(), 0u);
        team.run(n, [&](unsigned t) { hits[t] += 1 + c; });
        for (unsigned t = 0; t <= max_width; ++t)
            if (hits[t] != (t < n ? 1 + c : 0u)) return int(c) + 1;
    }
    return 0;
}

/* like fxio_parse with want_names: *names = the reads' names joined by '\n' (malloc'ed, NUL-terminated); returns the number of reads */
long long fxio_parse_names(const char* path, unsigned threads, unsigned long long span, unsigned long long max_reads_serial, char** names) {
    fgio::fastx_source src;
    if (!src.open(path, threads, span, max_reads_serial, true)) return -1;
    fgio::read_batch b;
    std::vector<char> bb;
    std::vector<uint64_t> bo;
    b.grow = [&](fgio::read_batch& x, uint64_t nb, uint64_t nr) {
        bb.resize(nb + 16);
        bo.resize(nr + 16);
        x.bases = bb.data();
        x.off = bo.data();
        x.bases_cap = bb.size();
        x.reads_cap = bo.size();
    };
    std::string all;
    long long total = 0;
    while (src.next_batch(b)) {
        if (b.name_off.size() != size_t(b.n) + 1 || b.name_off.back() != b.names.size()) return -2;
        for (uint32_t i = 0; i < b.n; ++i) {
            all.append(b.names.data() + b.name_off[i], size_t(b.name_off[i + 1] - b.name_off[i]));
            all += '\n';
        }
        total += b.n;
    }
    *names = static_cast<char*>(std::malloc(all.size() + 1));
    std::memcpy(*names, all.c_str(), all.size() + 1);
    return total;
}

/* formats one CSR batch to a file, in `pieces` calls of write_batch */
int fxio_format(const char* path, int fmt, unsigned num_colors, unsigned threads, unsigned n, const unsigned long long* off, const unsigned* colors,
                unsigned pieces) {
    fgio::result_writer w;
    if (!w.open(path, fmt == 0 ? fgio::out_format::ASCII : fmt == 1 ? fgio::out_format::BINARY : fgio::out_format::COMPRESSED, num_colors, threads))
        return -1;
    if (pieces < 1) pieces = 1;
    for (unsigned p = 0; p < pieces; ++p) {
        const unsigned lo = unsigned((unsigned long long)n * p / pieces), hi = unsigned((unsigned long long)n * (p + 1) / pieces);
        std::vector<uint64_t> o(off + lo, off + hi + 1);
        w.write_batch(lo, hi - lo, o.data(), reinterpret_cast<const uint32_t*>(colors));
        w.set_threads(1 + ((p + 1) * 3) % (threads + 2)); /* the tool changes the thread count between batches */
    }
    w.close();
    return w.ok() ? 0 : -2;
}

/* the per-k-mer tools' lines for n reads named r<first+i>; which = 0: kmer-conservation (off = triple offsets, vals = triples),
   which = 1: kmer-matches (read_off, k, off = word offsets, vals = positive words, counts, num_colors); formatted in `pieces` ranges */
int fxio_format_kmer_tool(const char* path, int which, unsigned n, unsigned first, const unsigned long long* off, const unsigned* vals,
                          const unsigned long long* read_off, unsigned k, const unsigned* counts, unsigned num_colors, unsigned pieces) {
    std::string names;
    std::vector<uint64_t> name_off(1, 0);
    for (unsigned i = 0; i < n; ++i) {
        names += "r" + std::to_string(first + i);
        name_off.push_back(names.size());
    }
    FILE* f = std::fopen(path, "wb");
    if (!f) return -1;
    if (pieces < 1) pieces = 1;
    for (unsigned p = 0; p < pieces; ++p) {
        const unsigned lo = unsigned((unsigned long long)n * p / pieces), hi = unsigned((unsigned long long)n * (p + 1) / pieces);
        std::string line;
        if (which == 0) fgio::format_kmer_conservation(names.data(), name_off.data(), lo, hi, reinterpret_cast<const uint64_t*>(off), vals, line);
        else fgio::format_kmer_matches(names.data(), name_off.data(), lo, hi, reinterpret_cast<const uint64_t*>(read_off), k, reinterpret_cast<const uint64_t*>(off), vals, counts, num_colors, line);
        std::fwrite(line.data(), 1, line.size(), f);
    }
    std::fclose(f);
    return 0;
}

/* one write_batch call with deduplicated results: record i carries the range of read rep[i] */
int fxio_format_dedup(const char* path, int fmt, unsigned num_colors, unsigned threads, unsigned n, const unsigned long long* off,
                      const unsigned* colors, const unsigned* rep) {
    fgio::result_writer w;
    if (!w.open(path, fmt == 0 ? fgio::out_format::ASCII : fmt == 1 ? fgio::out_format::BINARY : fgio::out_format::COMPRESSED, num_colors, threads))
        return -1;
    w.write_batch(7, n, reinterpret_cast<const uint64_t*>(off), reinterpret_cast<const uint32_t*>(colors), reinterpret_cast<const uint32_t*>(rep));
    w.close();
    return 0;
}
}
