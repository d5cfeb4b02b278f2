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
 * reference's order depends on thread scheduling, README.md:220); -t sizes the host side (parsing and formatting
 * threads, fastx_io.h); --deduplicate computes one intersection per distinct color-set-id list of a batch on the GPU
 * (fulgor_gpu_pseudoalign_dedup) and fans the result out to every read id while formatting.
 * Three batches are in flight: while the GPU works on batch i, batch i+1 is being parsed and batch i-1 formatted.
 * Read ids are 0-based positions in the query file. With --gpus N the index
 * image is uploaded to N devices and every batch is split N ways; no cross-GPU reduction exists.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fulgor_gpu.h"
#include "fastx_io.h"

namespace {

using fgio::out_format;
using fgio::put_u32;

/* ---------------------------------------------------------------- arguments (cmd_line_parser semantics) */
struct args_t {
    std::string index, query, output, format = "ascii";
    uint64_t threads = 1;
    double threshold = -1.0;
    bool has_threshold = false, verbose = false, deduplicate = false;
    int gpus = 1;
    uint64_t batch_reads = 1u << 17; /* reads per pipeline step: three batches are in flight, so a run costs (batches + 2) steps -- small
                                        batches keep the fill and drain short (4 M reads, 8 host threads, GPU stage stubbed out:
                                        382 ms at 2^20, 250 ms at 2^18, 230 ms at 2^17, 215 ms at 2^16 reads per batch) while one
                                        batch still fills a GPU call (a chunk of the library's own pipeline holds 2^18 reads) */
    bool batch_given = false;
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
        else if (k == "--batch-reads") { if (!val(v)) return false; a.batch_reads = std::strtoull(v.c_str(), nullptr, 10); a.batch_given = true; }
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

/* ---------------------------------------------------------------- `kmer-conservation` / `kmer-matches`
   The reference's two per-k-mer tools (tools/kmer_conservation.cpp, tools/kmer_matches.cpp) with the per-read body
   (index::kmer_conservation / index::kmer_matches) replaced by one C-ABI call per batch. Same flags (-i -q -o -t --verbose),
   same output lines, written in input order:
       kmer-conservation:  name \t n [\t (start_pos_in_query num_kmers color_set_id)]*      (tools/kmer_conservation.cpp:27-37)
       kmer-matches:       "num_colors=C" once, then  name \t num_kmers [\t 0|1]* [\t count]*C  (tools/kmer_matches.cpp:98, 28-34)
   These tools print the read NAMES: the feeder keeps the first word of every header (fastx_io.h, want_names). A read shorter
   than k is a line without k-mers (and, for kmer-matches, zero counts). */
int kmer_tool_main(bool conservation, int argc, char** argv) {
    args_t a;
    if (!parse(argc, argv, a)) {
        std::cerr << "Usage: fulgor_b200_pseudoalign " << (conservation ? "kmer-conservation" : "kmer-matches")
                  << " -i index_filename -q query_filename -o output_filename [-t num_threads] [--verbose] [--batch-reads n]\n";
        return 1;
    }
    if (a.threads == 1) std::cerr << "1 thread was specified, but an additional thread will be allocated for parsing" << std::endl;
    if (!(ends_with(a.index, ".fur") || ends_with(a.index, ".mfur") || ends_with(a.index, ".dfur") || ends_with(a.index, ".mdfur"))) {
        std::cerr << "Wrong index filename supplied." << std::endl;
        return 1;
    }
    if (fulgor_gpu_device_count() < 1) {
        std::cerr << "no usable CUDA device (this tool has no CPU path)" << std::endl;
        return 1;
    }
    fulgor_gpu_index* gpu = nullptr;
    if (fulgor_gpu_index_open(a.index.c_str(), 0, &gpu)) {
        std::cerr << fulgor_gpu_last_error() << std::endl;
        return 1;
    }
    fulgor_gpu_info info;
    fulgor_gpu_index_info(gpu, &info);
    /* kmer-matches returns num_colors counts per read: one C-ABI call takes at most `slice` reads (~64 MB of counts) */
    const uint64_t slice = conservation ? std::min<uint64_t>(a.batch_reads, 1u << 18)
                                        : std::max<uint64_t>(64, std::min<uint64_t>(a.batch_reads, (64ull << 20) / std::max<uint32_t>(1, info.num_colors)));
    const unsigned host_threads = unsigned(std::max<uint64_t>(1, a.threads));
    fgio::fastx_source reader; /* the same feeder as pseudoalign, keeping the record names */
    if (!reader.open(a.query.c_str(), host_threads, std::min<uint64_t>(1ull << 28, slice * 320), slice, /*want_names=*/true)) {
        std::cerr << "error in opening the file '" << a.query << "'" << std::endl;
        return 1;
    }
    FILE* out = std::fopen(a.output.c_str(), "wb");
    if (!out) {
        std::cerr << "could not open output file " << a.output << std::endl;
        return 1;
    }
    if (!conservation) std::fprintf(out, "num_colors=%u\n", info.num_colors); /* tools/kmer_matches.cpp:98 */
    const auto t0 = std::chrono::high_resolution_clock::now();
    std::vector<char> bases_buf;
    std::vector<uint64_t> off_buf, res_off;
    fgio::read_batch rb;
    rb.grow = [&](fgio::read_batch& r, uint64_t nb, uint64_t nr) { /* std::vector::resize keeps what the batch already holds */
        if (nb > r.bases_cap) {
            bases_buf.resize(std::max<uint64_t>(nb + nb / 4 + 64, 2 * r.bases_cap));
            r.bases = bases_buf.data();
            r.bases_cap = bases_buf.size();
        }
        if (nr > r.reads_cap) {
            off_buf.resize(std::max<uint64_t>(nr + nr / 4 + 64, 2 * r.reads_cap));
            r.off = off_buf.data();
            r.reads_cap = off_buf.size();
        }
    };
    std::vector<uint32_t> vals, counts;
    std::vector<std::string> pieces;
    uint64_t num_reads = 0;
    while (reader.next_batch(rb)) {
        for (uint32_t first = 0; first < rb.n; first += uint32_t(slice)) { /* a batch may hold more reads than one call should take */
            const uint32_t n = uint32_t(std::min<uint64_t>(slice, rb.n - first));
            const uint64_t* off = rb.off + first; /* absolute offsets into rb.bases, like the C ABI wants them */
            const uint64_t* name_off = rb.name_off.data() + first;
            res_off.assign(uint64_t(n) + 1, 0);
            int rc;
            if (conservation) {
                if (vals.size() < 3 * (uint64_t(n) * 4 + 64)) vals.resize(3 * (uint64_t(n) * 4 + 64));
                while ((rc = fulgor_gpu_kmer_conservation(gpu, rb.bases, off, n, res_off.data(), vals.data(), vals.size() / 3)) == FULGOR_GPU_E2BIG)
                    vals.resize(3 * res_off[n]);
            } else {
                const uint64_t words = (off[n] - off[0]) / 32 + n + 1;
                if (vals.size() < words) vals.resize(words);
                counts.resize(uint64_t(n) * info.num_colors);
                rc = fulgor_gpu_kmer_matches(gpu, rb.bases, off, n, res_off.data(), vals.data(), vals.size(), counts.data());
            }
            if (rc) {
                std::cerr << fulgor_gpu_last_error() << std::endl;
                return 1;
            }
            /* the lines of the slice, formatted by up to T threads over contiguous read ranges and written in input order */
            const unsigned T = unsigned(std::max<uint64_t>(1, std::min<uint64_t>(host_threads, n / 256)));
            pieces.resize(T);
            fgio::parallel_for(T, [&](unsigned t) {
                const uint32_t lo = uint32_t(uint64_t(n) * t / T), hi = uint32_t(uint64_t(n) * (t + 1) / T);
                pieces[t].clear();
                if (conservation) fgio::format_kmer_conservation(rb.names.data(), name_off, lo, hi, res_off.data(), vals.data(), pieces[t]);
                else fgio::format_kmer_matches(rb.names.data(), name_off, lo, hi, off, info.k, res_off.data(), vals.data(), counts.data(), info.num_colors, pieces[t]);
            });
            for (unsigned t = 0; t < T; ++t) std::fwrite(pieces[t].data(), 1, pieces[t].size(), out);
            num_reads += n;
        }
        if (a.verbose) std::cout << "processed " << num_reads << " reads" << std::endl;
    }
    std::fclose(out);
    const double ms = double(std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - t0).count());
    if (a.verbose) { /* tools/kmer_conservation.cpp:118-124 */
        std::cout << "processed " << num_reads << " reads" << std::endl;
        std::cout << "elapsed = " << ms << " millisec / " << ms / 1000 << " sec / " << ms / 1000 / 60 << " min / "
                  << (ms * 1000) / double(std::max<uint64_t>(1, num_reads)) << " musec/read" << std::endl;
    }
    fulgor_gpu_index_close(gpu);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc > 1 && (std::strcmp(argv[1], "kmer-conservation") == 0 || std::strcmp(argv[1], "kmer-matches") == 0))
        return kmer_tool_main(std::strcmp(argv[1], "kmer-conservation") == 0, argc - 1, argv + 1);
    if (argc > 1 && std::strcmp(argv[1], "pseudoalign") == 0) { /* `fulgor pseudoalign ...` spelling */
        --argc;
        ++argv;
    }
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
    if (a.deduplicate && a.has_threshold) { /* tools/pseudoalign.cpp:283-289 */
        std::cerr << "Deduplication not available for threshold < 1.0. Remove --deduplicate flag." << std::endl;
        return 1;
    }
    /* the library deduplicates over a whole call, the reference over the whole file (tools/pseudoalign.cpp:92-226): larger batches
       bring the two closer (8 M reads = 1.2 GB of 150 bp reads per batch buffer, three batches in flight) */
    if (a.deduplicate && !a.batch_given) a.batch_reads = 1u << 23;
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
                  << (a.deduplicate ? "(dedup.)" : "") << std::endl;
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

    /* host threads: -t is shared by the parsing and the formatting stage (they overlap each other and the GPU call). Tokenising
       150 bp FASTQ records costs more than formatting a handful of colors per read, formatting thousands of colors per read far
       more than tokenising: the split starts at 5 : 3 and follows the measured work of the two stages batch by batch. */
    const unsigned total_threads = unsigned(std::max<uint64_t>(2, a.threads));
    unsigned parse_threads = std::max(1u, std::min(total_threads - 1, (total_threads * 5 + 4) / 8));
    unsigned format_threads = total_threads - parse_threads;
    fgio::fastx_source reader;
    if (!reader.open(a.query.c_str(), parse_threads, std::min<uint64_t>(1ull << 30, a.batch_reads * 320), a.batch_reads)) {
        std::cerr << "cannot open query file '" << a.query << "'" << std::endl;
        return 1;
    }
    fgio::result_writer out;
    if (!out.open(a.output.c_str(), fmt, info.num_colors, format_threads)) {
        std::cerr << "cannot open output file '" << a.output << "'" << std::endl;
        return 1;
    }

    uint64_t num_reads = 0, num_mapped = 0;

    /* one batch in flight per pipeline stage */
    struct batch_ctx {
        pinned p_coff, p_colors, p_rep;
        fgio::read_batch reads; /* pinned, owned here */
        uint64_t* hc = nullptr;
        uint32_t* hv = nullptr;
        uint32_t* hr = nullptr; /* --deduplicate: representative of every read (batch-local index) */
        uint64_t colors_cap = 0, first_id = 0;
        bool full = false, failed = false;
    } ctx[3];
    for (auto& c : ctx) {
        c.reads.grow = [&c](fgio::read_batch& r, uint64_t nb, uint64_t nr) { /* keeps the batch's contents */
            if (nb > r.bases_cap) {
                const uint64_t cap = std::max<uint64_t>(nb + nb / 4 + (1u << 20), 2 * r.bases_cap);
                char* p = static_cast<char*>(fulgor_gpu_host_alloc(cap));
                if (!p) { std::cerr << "cannot allocate pinned host memory\n"; std::exit(1); }
                if (r.bases) std::memcpy(p, r.bases, r.bases_cap);
                fulgor_gpu_host_free(r.bases);
                r.bases = p;
                r.bases_cap = cap;
            }
            if (nr > r.reads_cap) {
                const uint64_t cap = std::max<uint64_t>(nr + nr / 4 + (1u << 16), 2 * r.reads_cap);
                uint64_t* p = static_cast<uint64_t*>(fulgor_gpu_host_alloc(cap * 8));
                if (!p) { std::cerr << "cannot allocate pinned host memory\n"; std::exit(1); }
                if (r.off) std::memcpy(p, r.off, r.reads_cap * 8);
                fulgor_gpu_host_free(r.off);
                r.off = p;
                r.reads_cap = cap;
            }
        };
    }
    uint64_t colors_per_read = std::min<uint32_t>(info.num_colors, 16); /* first guess; E2BIG reports the exact need */
    { /* set-up, like loading the index: the pinned buffers of the three batches in flight are sized for a whole batch before
         the clock starts (a pinned allocation costs about as much as parsing the bytes it will hold) */
        uint64_t eb = 0, er = 0;
        if (reader.estimate_batch(eb, er)) {
            for (auto& c : ctx) {
                c.reads.grow(c.reads, eb, er);
                c.p_coff.get<uint64_t>(er + 1);
                c.colors_cap = er * colors_per_read + 1024;
                c.p_colors.get<uint32_t>(c.colors_cap);
                if (a.deduplicate) c.p_rep.get<uint32_t>(er + 1);
            }
        }
    }

    const auto t0 = std::chrono::high_resolution_clock::now();
    if (a.verbose) std::cout << "*** START: pseudoalignment" << std::endl;

    auto gpu_stage = [&](batch_ctx& c) {
        const uint32_t n = c.reads.n;
        char* hb = c.reads.bases;
        uint64_t* ho = c.reads.off;
        c.hc = c.p_coff.get<uint64_t>(uint64_t(n) + 1);
        if (c.colors_cap < uint64_t(n) * colors_per_read + 1024) c.colors_cap = uint64_t(n) * colors_per_read + 1024;
        c.hv = c.p_colors.get<uint32_t>(c.colors_cap);
        c.hr = a.deduplicate ? c.p_rep.get<uint32_t>(uint64_t(n) + 1) : nullptr;
        uint64_t* hc = c.hc;
        /* one full intersection per distinct color-set-id list (groups never span devices: each device deduplicates its range) */
        auto pseudoalign = [&](int g, uint32_t lo, uint32_t cnt, uint64_t* off, uint32_t* vals, uint64_t cap) {
            if (!a.deduplicate) return fulgor_gpu_pseudoalign(gpu[g], algo, a.threshold, hb, ho + lo, cnt, off, vals, cap);
            const int rc = fulgor_gpu_pseudoalign_dedup(gpu[g], hb, ho + lo, cnt, c.hr + lo, off, vals, cap);
            if (rc == 0 && lo)
                for (uint32_t i = lo; i < lo + cnt; ++i) c.hr[i] += lo;
            return rc;
        };
        /* split the batch over the GPUs: contiguous ranges balanced by k-mer count */
        const int G = a.gpus;
        std::vector<uint32_t> cut(G + 1, 0);
        if (G > 1) {
            const uint64_t k = info.k;
            std::vector<uint64_t> work(uint64_t(n) + 1, 0);
            for (uint32_t i = 0; i < n; ++i) {
                const uint64_t len = ho[i + 1] - ho[i];
                work[i + 1] = work[i] + (len >= k ? len - k + 1 : 0) + 8;
            }
            for (int g = 1; g < G; ++g) {
                const uint64_t target = work[n] / G * g;
                cut[g] = uint32_t(std::lower_bound(work.begin(), work.end(), target) - work.begin());
            }
        }
        cut[G] = n;
        for (;;) {
            std::vector<int> rcs(G, 0);
            std::vector<std::string> errs(G);
            std::vector<std::vector<uint64_t>> g_off(G);
            std::vector<std::vector<uint32_t>> g_val(G);
            auto run = [&](int g) {
                const uint32_t lo = cut[g], hi = cut[g + 1];
                if (G == 1) {
                    rcs[g] = pseudoalign(g, 0, n, hc, c.hv, c.colors_cap);
                } else { /* each device fills its own CSR; spliced below */
                    g_off[g].assign(hi - lo + 1, 0);
                    uint64_t cap = std::max<uint64_t>(uint64_t(hi - lo) * colors_per_read, 1024);
                    for (;;) {
                        g_val[g].resize(cap);
                        rcs[g] = pseudoalign(g, lo, hi - lo, g_off[g].data(), g_val[g].data(), cap);
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
                    c.colors_cap = hc[n] + hc[n] / 8;
                    c.hv = c.p_colors.get<uint32_t>(c.colors_cap);
                    retry = true;
                } else if (rcs[g]) {
                    std::cerr << errs[g] << std::endl;
                    c.failed = true;
                }
            }
            if (c.failed) return;
            if (retry) continue;
            if (G > 1) { /* splice the per-device CSRs in read order */
                uint64_t total = 0;
                for (int g = 0; g < G; ++g) total += g_off[g].back();
                if (total > c.colors_cap) {
                    c.colors_cap = total + total / 8;
                    c.hv = c.p_colors.get<uint32_t>(c.colors_cap);
                }
                uint64_t base = 0;
                hc[0] = 0;
                for (int g = 0; g < G; ++g) {
                    const uint32_t lo = cut[g], cnt = cut[g + 1] - cut[g];
                    for (uint32_t i = 1; i <= cnt; ++i) hc[lo + i] = base + g_off[g][i];
                    std::memcpy(c.hv + base, g_val[g].data(), g_off[g].back() * 4);
                    base += g_off[g].back();
                }
            }
            break;
        }
        if (n) colors_per_read = std::max<uint64_t>(colors_per_read, hc[n] / n + 1);
    };
    double parse_ms = 0, format_ms = 0; /* wall time of the last parsing / formatting stage */
    auto format_stage = [&](batch_ctx& c) {
        const auto f0 = std::chrono::steady_clock::now();
        const uint32_t n = c.reads.n;
        uint64_t mapped = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t r = c.hr ? c.hr[i] : i;
            mapped += c.hc[r + 1] > c.hc[r];
        }
        num_mapped += mapped;
        out.write_batch(uint32_t(c.first_id), n, c.hc, c.hv, c.hr);
        format_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - f0).count();
    };

    /* step i: parse batch i | GPU on batch i-1 | format batch i-2, each on its own thread */
    bool more = true, too_many = false;
    for (uint64_t step = 0;; ++step) {
        batch_ctx& cp = ctx[step % 3];
        batch_ctx& cg = ctx[(step + 2) % 3];
        batch_ctx& cf = ctx[(step + 1) % 3];
        const bool do_gpu = step >= 1 && cg.full, do_fmt = step >= 2 && cf.full;
        if (!more && !do_gpu && !do_fmt) break;
        std::thread tg, tf;
        if (do_gpu) tg = std::thread(gpu_stage, std::ref(cg));
        if (do_fmt) tf = std::thread(format_stage, std::ref(cf));
        cp.full = false;
        if (more) {
            const auto p0 = std::chrono::steady_clock::now();
            more = reader.next_batch(cp.reads);
            parse_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - p0).count();
            if (more) {
                cp.full = true;
                cp.first_id = num_reads;
                num_reads += cp.reads.n;
                if (num_reads > (1ull << 32)) too_many = true;
            }
        }
        if (tg.joinable()) tg.join();
        if (tf.joinable()) tf.join();
        if (do_fmt) {
            cf.full = false;
            if (a.verbose) std::cout << "processed " << cf.first_id + cf.reads.n << " reads" << std::endl;
        }
        if (more && do_fmt && parse_ms > 0 && format_ms > 0) {
            /* both stages ran on full batches: a step lasts as long as its slower stage, so a thread moves to the stage that took
               clearly longer (neither stage scales linearly with its threads: comparing thread-time would starve one of them) */
            unsigned np = parse_threads;
            if (parse_ms > 1.3 * format_ms && format_threads > 1) np += 1;
            else if (format_ms > 1.3 * parse_ms && parse_threads > 1) np -= 1;
            if (np != parse_threads) {
                parse_threads = np;
                format_threads = total_threads - parse_threads;
                reader.set_threads(parse_threads);
                out.set_threads(format_threads);
            }
        }
        if (do_gpu && cg.failed) return 1;
        if (too_many) {
            std::cerr << "more than 2^32 reads: read ids do not fit the output formats" << std::endl;
            return 1;
        }
    }
    out.close();
    if (!out.ok()) {
        std::cerr << "error in writing the output file '" << a.output << "'" << std::endl;
        return 1;
    }
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
