/*
 * pseudoalign_cli.cpp -- `fulgor_b200_pseudoalign`: the reference's `fulgor pseudoalign` tool
 * (tools/pseudoalign.cpp:228-369) with the hot path replaced by libfulgor_gpu (C ABI, include/fulgor_gpu.h).
 * Same flags, same output formats (src/ps_utils.cpp:48-243), same status lines on stdout/stderr
 * (tools/pseudoalign.cpp:79-88, 326-334). Host C++; the only CUDA it touches is behind the C ABI.
 *
 *   fulgor_b200_pseudoalign -i index.{fur,mfur} -q reads.{fa,fq}[.gz] -o out [-t T] [-r tau]
 *                           [--format ascii|binary|compressed] [--verbose] [--gpus N] [--batch-reads B]
 *
 * Differences from the reference, all outside the results: records are written in read-id order (the
 * reference's order depends on thread scheduling, README.md:220); -t is accepted and only sizes the
 * host-side formatting; --deduplicate is rejected (it only changes how the reference schedules work).
 * Read ids are 0-based positions in the query file. With --gpus N the index
 * image is uploaded to N devices and every batch is split N ways; no cross-GPU reduction exists.
 */
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fulgor_gpu.h"

namespace {

/* ---------------------------------------------------------------- FASTA/FASTQ(.gz) reader */
struct fastx_reader {
    gzFile f = nullptr;
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;

    bool open(const char* path) {
        f = gzopen(path, "rb");
        if (!f) return false;
        gzbuffer(f, 1 << 20);
        buf.resize(1 << 22);
        return true;
    }
    ~fastx_reader() {
        if (f) gzclose(f);
    }
    bool fill() {
        if (eof) return false;
        const int n = gzread(f, buf.data(), unsigned(buf.size()));
        if (n <= 0) {
            eof = true;
            return false;
        }
        pos = 0;
        end = size_t(n);
        return true;
    }
    /* next line without the terminator; false at end of file */
    bool getline(std::string& line) {
        line.clear();
        bool any = false;
        for (;;) {
            if (pos == end && !fill()) return any;
            any = true;
            const char* p = buf.data() + pos;
            const char* nl = static_cast<const char*>(std::memchr(p, '\n', end - pos));
            if (nl) {
                line.append(p, size_t(nl - p));
                pos += size_t(nl - p) + 1;
                if (!line.empty() && line.back() == '\r') line.pop_back();
                return true;
            }
            line.append(p, end - pos);
            pos = end;
        }
    }
    /* appends the next record's sequence to `bases`; false when the file is exhausted */
    std::string line, pending_header;
    bool next(std::vector<char>& bases) {
        std::string header;
        if (!pending_header.empty()) {
            header.swap(pending_header);
        } else {
            do {
                if (!getline(header)) return false;
            } while (header.empty());
        }
        if (header[0] == '@') { /* FASTQ: sequence, '+', quality (single-line records, like kseq on typical files) */
            if (!getline(line)) return false;
            bases.insert(bases.end(), line.begin(), line.end());
            const size_t seq_len = line.size();
            if (!getline(line)) return true; /* '+' */
            size_t q = 0;
            while (q < seq_len && getline(line)) q += line.size();
            return true;
        }
        if (header[0] == '>') { /* FASTA: possibly multi-line */
            while (getline(line)) {
                if (!line.empty() && (line[0] == '>' || line[0] == '@')) {
                    pending_header = line;
                    break;
                }
                bases.insert(bases.end(), line.begin(), line.end());
            }
            return true;
        }
        return false;
    }
};

/* ---------------------------------------------------------------- output formats (src/ps_utils.cpp:48-243) */
inline char* put_u32(char* p, uint32_t v) {
    char tmp[10];
    int n = 0;
    do {
        tmp[n++] = char('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

struct bit_writer { /* LSB-first, like bits::bit_vector::builder */
    std::vector<uint64_t> words;
    uint64_t num_bits = 0;
    void append(uint64_t v, uint32_t len) {
        if (!len) return;
        if (len < 64) v &= (1ULL << len) - 1;
        const uint32_t sh = uint32_t(num_bits & 63);
        if (sh == 0) words.push_back(v);
        else {
            words.back() |= v << sh;
            if (sh + len > 64) words.push_back(v >> (64 - sh));
        }
        num_bits += len;
    }
    /* bits::util::write_delta (bits/include/integer_codes.hpp:54-71): gamma(len) then the payload */
    void gamma(uint64_t x) {
        const uint64_t xx = x + 1;
        const uint32_t b = 63 - uint32_t(__builtin_clzll(xx));
        append(1ULL << b, b + 1); /* b zeros then a one */
        append(xx ^ (1ULL << b), b);
    }
    void delta(uint64_t x) {
        const uint64_t xx = x + 1;
        const uint32_t b = 63 - uint32_t(__builtin_clzll(xx));
        gamma(b);
        append(xx ^ (1ULL << b), b);
    }
    void clear() {
        words.clear();
        num_bits = 0;
    }
};

enum class out_format { ASCII, BINARY, COMPRESSED };

struct writer {
    FILE* f = nullptr;
    out_format fmt = out_format::ASCII;
    uint32_t num_colors = 0, sparse_thr = 0, dense_thr = 0;
    std::vector<char> text;
    bit_writer bw;

    bool open(const char* path, out_format fm, uint32_t nc) {
        f = std::fopen(path, "wb");
        fmt = fm;
        num_colors = nc;
        if (!f) return false;
        if (fmt == out_format::COMPRESSED) { /* psa_compressed_formatter::set_num_colors, src/ps_utils.cpp:160-166 */
            const uint64_t header = nc;
            std::fwrite(&header, 8, 1, f);
            sparse_thr = uint32_t(0.25 * nc);
            dense_thr = uint32_t(0.75 * nc);
        }
        return true;
    }
    void flush_bits() {
        if (!bw.num_bits) return;
        std::fwrite(&bw.num_bits, 8, 1, f);
        std::fwrite(bw.words.data(), 8, bw.words.size(), f);
        bw.clear();
    }
    void write_batch(uint32_t first_id, uint32_t n, const uint64_t* off, const uint32_t* colors) {
        if (fmt == out_format::ASCII) {
            text.resize(size_t(n) * 24 + size_t(off[n] - off[0]) * 11 + 16);
            char* p = text.data();
            for (uint32_t i = 0; i < n; ++i) {
                const uint64_t b = off[i], e = off[i + 1];
                p = put_u32(p, first_id + i);
                *p++ = '\t';
                p = put_u32(p, uint32_t(e - b));
                for (uint64_t j = b; j < e; ++j) {
                    *p++ = '\t';
                    p = put_u32(p, colors[j]);
                }
                *p++ = '\n';
            }
            std::fwrite(text.data(), 1, size_t(p - text.data()), f);
        } else if (fmt == out_format::BINARY) {
            text.resize((size_t(n) * 2 + size_t(off[n] - off[0])) * 4);
            uint32_t* p = reinterpret_cast<uint32_t*>(text.data());
            for (uint32_t i = 0; i < n; ++i) {
                const uint64_t b = off[i], e = off[i + 1];
                *p++ = first_id + i;
                *p++ = uint32_t(e - b);
                std::memcpy(p, colors + b, size_t(e - b) * 4);
                p += e - b;
            }
            std::fwrite(text.data(), 1, text.size(), f);
        } else {
            for (uint32_t i = 0; i < n; ++i) {
                const uint32_t* c = colors + off[i];
                const uint32_t size = uint32_t(off[i + 1] - off[i]);
                bw.delta(first_id + i);
                bw.delta(size);
                if (size == 0) {
                } else if (size < sparse_thr) {
                    bw.delta(c[0]);
                    for (uint32_t j = 1; j < size; ++j) bw.delta(c[j] - (c[j - 1] + 1));
                } else if (size < dense_thr) {
                    const uint64_t start = bw.num_bits;
                    for (uint32_t w = 0; w < num_colors; w += 64) bw.append(0, std::min<uint32_t>(64, num_colors - w));
                    for (uint32_t j = 0; j < size; ++j) {
                        const uint64_t bit = start + c[j];
                        bw.words[bit >> 6] |= 1ULL << (bit & 63);
                    }
                } else { /* the complement, delta-gap coded */
                    bool first = true;
                    uint32_t prev = 0, j = 0;
                    for (uint32_t v = 0; v < num_colors; ++v) {
                        if (j < size && c[j] == v) {
                            ++j;
                            continue;
                        }
                        bw.delta(first ? v : v - (prev + 1));
                        first = false;
                        prev = v;
                    }
                }
                if (bw.words.size() * 8 > (1u << 14)) flush_bits(); /* formatter_buffer, src/ps_utils.cpp:31-38 */
            }
        }
    }
    void close() {
        if (!f) return;
        if (fmt == out_format::COMPRESSED) flush_bits();
        std::fclose(f);
        f = nullptr;
    }
};

/* ---------------------------------------------------------------- arguments (cmd_line_parser semantics) */
struct args_t {
    std::string index, query, output, format = "ascii";
    uint64_t threads = 1;
    double threshold = -1.0;
    bool has_threshold = false, verbose = false, deduplicate = false;
    int gpus = 1;
    uint64_t batch_reads = 4u << 20;
};

void usage() {
    std::cerr << "Usage: fulgor_b200_pseudoalign [-h,--help] -i index_filename -q query_filename -o output_filename [-t num_threads] "
                 "[--verbose] [-r threshold] [--deduplicate] [--format format] [--gpus n] [--batch-reads n]\n";
}

bool parse(int argc, char** argv, args_t& a) {
    for (int i = 1; i < argc; ++i) {
        std::string k = argv[i];
        auto val = [&](std::string& dst) {
            if (i + 1 >= argc) return false;
            dst = argv[++i];
            return true;
        };
        std::string v;
        if (k == "-i") { if (!val(a.index)) return false; }
        else if (k == "-q") { if (!val(a.query)) return false; }
        else if (k == "-o") { if (!val(a.output)) return false; }
        else if (k == "-t") { if (!val(v)) return false; a.threads = std::strtoull(v.c_str(), nullptr, 10); }
        else if (k == "-r") { if (!val(v)) return false; a.threshold = std::strtod(v.c_str(), nullptr); a.has_threshold = true; }
        else if (k == "--format") { if (!val(a.format)) return false; }
        else if (k == "--gpus") { if (!val(v)) return false; a.gpus = std::atoi(v.c_str()); }
        else if (k == "--batch-reads") { if (!val(v)) return false; a.batch_reads = std::strtoull(v.c_str(), nullptr, 10); }
        else if (k == "--verbose") a.verbose = true;
        else if (k == "--deduplicate") a.deduplicate = true;
        else if (k == "-h" || k == "--help") return false;
        else {
            std::cerr << "== error: unknown argument '" << k << "'\n";
            return false;
        }
    }
    if (a.index.empty() || a.query.empty() || a.output.empty()) {
        std::cerr << "== error: -i, -q and -o are required\n";
        return false;
    }
    return true;
}

bool ends_with(std::string const& s, const char* suf) {
    const size_t n = std::strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

struct pinned {
    void* p = nullptr;
    uint64_t cap = 0;
    template <typename T>
    T* get(uint64_t count) {
        const uint64_t bytes = count * sizeof(T);
        if (bytes > cap) {
            fulgor_gpu_host_free(p);
            cap = bytes + bytes / 4 + 4096;
            p = fulgor_gpu_host_alloc(cap);
            if (!p) {
                std::cerr << "cannot allocate pinned host memory\n";
                std::exit(1);
            }
        }
        return static_cast<T*>(p);
    }
    ~pinned() { fulgor_gpu_host_free(p); }
};

}  // namespace

int main(int argc, char** argv) {
    args_t a;
    if (!parse(argc, argv, a)) {
        usage();
        return 1;
    }
    if (a.threads == 1) { /* tools/pseudoalign.cpp:263-270 */
        a.threads += 1;
        std::cerr << "1 thread was specified, but an additional thread will be allocated for parsing" << std::endl;
    }
    if (a.has_threshold && (a.threshold <= 0.0 || a.threshold > 1.0)) {
        std::cerr << "threshold must be a float in (0.0,1.0]" << std::endl;
        return 1;
    }
    const int algo = a.has_threshold ? FULGOR_GPU_THRESHOLD_UNION : FULGOR_GPU_FULL_INTERSECTION;
    if (a.deduplicate) {
        if (a.has_threshold) std::cerr << "Deduplication not available for threshold < 1.0. Remove --deduplicate flag." << std::endl;
        else std::cerr << "--deduplicate is not needed on the GPU path (results are identical without it). Remove --deduplicate flag." << std::endl;
        return 1;
    }
    if (!(ends_with(a.index, ".fur") || ends_with(a.index, ".mfur") || ends_with(a.index, ".dfur") || ends_with(a.index, ".mdfur"))) { /* tools/util.cpp:5-19 */
        std::cerr << "Wrong index filename supplied." << std::endl;
        return 1;
    }
    out_format fmt;
    if (a.format == "ascii") fmt = out_format::ASCII;
    else if (a.format == "binary") fmt = out_format::BINARY;
    else if (a.format == "compressed") fmt = out_format::COMPRESSED;
    else {
        std::cout << "Unknown output format. Supported formats: ascii, binary, compressed." << std::endl;
        return 1;
    }
    if (a.verbose) {
        for (int i = 0; i < argc; ++i) std::cout << argv[i] << " ";
        std::cout << std::endl;
        std::cout << "\n---------------------------------" << std::endl;
        std::cout << "[Index]     " << a.index << std::endl;
        std::cout << "[Queries]   " << a.query << std::endl;
        std::cout << "[Output]    " << a.output << std::endl;
        std::cout << "[Algorithm] " << (algo == FULGOR_GPU_FULL_INTERSECTION ? std::string("full-intersection")
                                                                             : "threshold-union (threshold = " + std::to_string(a.threshold) + ")")
                  << std::endl;
        std::cout << "---------------------------------\n" << std::endl;
    }

    const int ndev = fulgor_gpu_device_count();
    if (ndev < 1) {
        std::cerr << "no usable CUDA device (this tool has no CPU path)" << std::endl;
        return 1;
    }
    if (a.gpus < 1 || a.gpus > ndev) {
        std::cerr << "--gpus must be in [1," << ndev << "]" << std::endl;
        return 1;
    }
    if (a.verbose) std::cout << "*** START: loading the index" << std::endl;
    uint8_t* image = nullptr;
    uint64_t image_bytes = 0;
    if (fulgor_gpu_image_build(a.index.c_str(), &image, &image_bytes)) {
        std::cerr << fulgor_gpu_last_error() << std::endl;
        return 1;
    }
    std::vector<fulgor_gpu_index*> gpu(a.gpus, nullptr);
    for (int g = 0; g < a.gpus; ++g) {
        if (fulgor_gpu_index_open_image(image, image_bytes, g, &gpu[g])) {
            std::cerr << fulgor_gpu_last_error() << std::endl;
            return 1;
        }
    }
    fulgor_gpu_image_free(image);
    fulgor_gpu_info info;
    fulgor_gpu_index_info(gpu[0], &info);
    if (a.verbose) std::cout << "*** DONE: loading the index" << std::endl;
    if (a.verbose) std::cout << "performing queries from file '" << a.query << "'..." << std::endl;

    fastx_reader reader;
    if (!reader.open(a.query.c_str())) {
        std::cerr << "cannot open query file '" << a.query << "'" << std::endl;
        return 1;
    }
    writer out;
    if (!out.open(a.output.c_str(), fmt, info.num_colors)) {
        std::cerr << "cannot open output file '" << a.output << "'" << std::endl;
        return 1;
    }

    const auto t0 = std::chrono::high_resolution_clock::now();
    if (a.verbose) std::cout << "*** START: pseudoalignment" << std::endl;
    uint64_t num_reads = 0, num_mapped = 0;
    std::vector<char> bases;
    std::vector<uint64_t> read_off;
    pinned p_bases, p_off, p_coff, p_colors;
    uint64_t colors_cap = 0;
    bool more = true;
    int rc_all = 0;
    while (more) {
        bases.clear();
        read_off.assign(1, 0);
        while (read_off.size() - 1 < a.batch_reads && bases.size() < (1ull << 31)) {
            if (!reader.next(bases)) {
                more = false;
                break;
            }
            read_off.push_back(bases.size());
        }
        const uint32_t n = uint32_t(read_off.size() - 1);
        if (n == 0) break;
        if (num_reads + n > (1ull << 32)) {
            std::cerr << "more than 2^32 reads: read ids do not fit the output formats" << std::endl;
            return 1;
        }
        char* hb = p_bases.get<char>(bases.size() + 1);
        std::memcpy(hb, bases.data(), bases.size());
        uint64_t* ho = p_off.get<uint64_t>(n + 1);
        std::memcpy(ho, read_off.data(), (n + 1) * 8);
        uint64_t* hc = p_coff.get<uint64_t>(n + 1);
        if (colors_cap == 0) colors_cap = std::max<uint64_t>(uint64_t(n) * std::min<uint32_t>(info.num_colors, 16), 1024);
        uint32_t* hv = p_colors.get<uint32_t>(colors_cap);
        /* split the batch over the GPUs: contiguous ranges balanced by k-mer count */
        const int G = a.gpus;
        std::vector<uint32_t> cut(G + 1, 0);
        {
            const uint64_t k = info.k;
            std::vector<uint64_t> work(n + 1, 0);
            for (uint32_t i = 0; i < n; ++i) {
                const uint64_t len = read_off[i + 1] - read_off[i];
                work[i + 1] = work[i] + (len >= k ? len - k + 1 : 0) + 8;
            }
            for (int g = 1; g < G; ++g) {
                const uint64_t target = work[n] / G * g;
                cut[g] = uint32_t(std::lower_bound(work.begin(), work.end(), target) - work.begin());
            }
            cut[G] = n;
        }
        for (;;) {
            std::vector<int> rcs(G, 0);
            std::vector<std::string> errs(G);
            std::vector<std::vector<uint64_t>> g_off(G);
            std::vector<std::vector<uint32_t>> g_val(G);
            auto run = [&](int g) {
                const uint32_t lo = cut[g], hi = cut[g + 1];
                if (G == 1) {
                    rcs[g] = fulgor_gpu_pseudoalign(gpu[g], algo, a.threshold, hb, ho, n, hc, hv, colors_cap);
                } else { /* each device fills its own CSR; spliced below */
                    g_off[g].assign(hi - lo + 1, 0);
                    uint64_t cap = std::max<uint64_t>(uint64_t(hi - lo) * std::min<uint32_t>(info.num_colors, 16), 1024);
                    for (;;) {
                        g_val[g].resize(cap);
                        rcs[g] = fulgor_gpu_pseudoalign(gpu[g], algo, a.threshold, hb, ho + lo, hi - lo, g_off[g].data(), g_val[g].data(), cap);
                        if (rcs[g] != FULGOR_GPU_E2BIG) break;
                        cap = g_off[g][hi - lo];
                    }
                }
                if (rcs[g] && rcs[g] != FULGOR_GPU_E2BIG) errs[g] = fulgor_gpu_last_error();
            };
            if (G == 1) run(0);
            else {
                std::vector<std::thread> th;
                for (int g = 0; g < G; ++g) th.emplace_back(run, g);
                for (auto& t : th) t.join();
            }
            bool retry = false;
            for (int g = 0; g < G; ++g) {
                if (rcs[g] == FULGOR_GPU_E2BIG && G == 1) { /* grow the values buffer to the reported size and retry */
                    colors_cap = hc[n] + hc[n] / 8;
                    hv = p_colors.get<uint32_t>(colors_cap);
                    retry = true;
                } else if (rcs[g]) {
                    std::cerr << errs[g] << std::endl;
                    rc_all = 1;
                }
            }
            if (rc_all) return 1;
            if (retry) continue;
            if (G > 1) { /* splice the per-device CSRs in read order */
                uint64_t total = 0;
                for (int g = 0; g < G; ++g) total += g_off[g].back();
                if (total > colors_cap) {
                    colors_cap = total + total / 8;
                    hv = p_colors.get<uint32_t>(colors_cap);
                }
                uint64_t base = 0;
                hc[0] = 0;
                for (int g = 0; g < G; ++g) {
                    const uint32_t lo = cut[g], cnt = cut[g + 1] - cut[g];
                    for (uint32_t i = 1; i <= cnt; ++i) hc[lo + i] = base + g_off[g][i];
                    std::memcpy(hv + base, g_val[g].data(), g_off[g].back() * 4);
                    base += g_off[g].back();
                }
            }
            break;
        }
        for (uint32_t i = 0; i < n; ++i) num_mapped += hc[i + 1] > hc[i];
        out.write_batch(uint32_t(num_reads), n, hc, hv);
        num_reads += n;
        if (a.verbose) std::cout << "processed " << num_reads << " reads" << std::endl;
    }
    out.close();
    const auto t1 = std::chrono::high_resolution_clock::now();
    const double ms = double(std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count());
    if (a.verbose) std::cout << "*** DONE: pseudoalignment" << std::endl;
    if (a.verbose) { /* tools/pseudoalign.cpp:79-88 */
        std::cout << "processed " << num_reads << " reads" << std::endl;
        std::cout << "elapsed = " << ms << " millisec / ";
        std::cout << ms / 1000 << " sec / ";
        std::cout << ms / 1000 / 60 << " min / ";
        std::cout << (ms * 1000) / double(num_reads) << " musec/read" << std::endl;
        std::cout << "num_mapped_reads " << num_mapped << "/" << num_reads << " (" << (double(num_mapped) * 100.0) / double(num_reads) << "%)" << std::endl;
    }
    for (auto* g : gpu) fulgor_gpu_index_close(g);
    return 0;
}
