/*
 * TEST INFRASTRUCTURE / INTEGRATION PROOF -- not part of the product.
 *
 * The reference's `fulgor pseudoalign` tool with its worker replaced by the GPU library: what INTEGRATION.md asks a
 * reference maintainer to add, COMPILED against the unmodified reference sources where they lie under $REF and linked with
 * fulgor_b200/libfulgor_gpu.so. Everything around the hot path is the reference's own code, included below like
 * $REF/tools/fulgor.cpp does: its command-line parser, its FASTA/FASTQ feeder (FQFeeder), its three output formatters with
 * their mutex-guarded flush (src/ps_utils.cpp:48-243), its ps_options counters and status lines. Only the body of
 * pseudoalign_worker (tools/pseudoalign.cpp:13-54) -- per read: index.fetch_color_set_ids + index.pseudoalign_* -- is
 * replaced: a worker concatenates the reads of one FQFeeder chunk and makes ONE C-ABI call (include/fulgor_gpu.h).
 *
 * Built by oracle/Makefile into oracle/_ref/fulgor_ref_gpu (git-ignored; needs $REF and the built library);
 * tests/test_cli.py::test_reference_tool_with_the_gpu_worker compares its output with the reference's own binary.
 */
#include <iostream>
#include <filesystem>

#include "external/sshash/external/gz/zip_stream.hpp"
#include "external/sshash/external/gz/zip_stream.cpp"
#include "external/sshash/src/build.cpp"
#include "external/sshash/src/dictionary.cpp"
#include "external/sshash/src/info.cpp"
#include "external/sshash/external/pthash/external/cmd_line_parser/include/parser.hpp"
#include "external/FQFeeder/include/FastxParser.hpp"
#include "external/FQFeeder/src/FastxParser.cpp"

#include "include/index_types.hpp"
#include "src/index.cpp"
#include "src/color_sets.cpp"

#include "tools/util.cpp"
#include "tools/pseudoalign.cpp" /* ps_options, the formatters, and (unused here) the reference's own worker */

#include "fulgor_gpu.h"

/* the replacement of pseudoalign_worker (tools/pseudoalign.cpp:13-54) for fastq_query_reader's role (src/ps_utils.cpp:245-305):
   same read ids (chunk_frag_offset().frag_idx + position in the chunk, ps_utils.cpp:271,286), same counters, same formatter */
template <typename Formatter>
void pseudoalign_worker_gpu(fulgor_gpu_index* gpu, fastx_parser::FastxParser<fastx_parser::ReadSeq>& rparser, Formatter& formatter,
                            const double threshold, ps_options& options) {
    auto rg = rparser.getReadGroup();
    auto output_buffer = formatter.buffer();
    std::string bases; /* concatenated reads of the chunk */
    std::vector<uint64_t> read_off, color_off;
    std::vector<uint32_t> colors, ids, one;
    const int algo = options.algo == pseudoalignment_algorithm::THRESHOLD_UNION ? FULGOR_GPU_THRESHOLD_UNION : FULGOR_GPU_FULL_INTERSECTION;
    while (rparser.refill(rg)) {
        bases.clear();
        read_off.assign(1, 0);
        ids.clear();
        uint32_t i = 0;
        for (auto const& record : rg) {
            bases += record.seq;
            read_off.push_back(bases.size());
            ids.push_back(uint32_t(rg.chunk_frag_offset().frag_idx + i++));
        }
        const uint32_t n = uint32_t(ids.size());
        color_off.resize(n + 1);
        colors.resize(std::max<size_t>(colors.size(), 16 * size_t(n) + 16));
        int rc;
        while ((rc = fulgor_gpu_pseudoalign(gpu, algo, algo == FULGOR_GPU_THRESHOLD_UNION ? threshold : 1.0, bases.data(), read_off.data(), n,
                                            color_off.data(), colors.data(), colors.size())) == FULGOR_GPU_E2BIG)
            colors.resize(color_off[n]); /* the required capacity is reported in color_off[n] */
        if (rc) throw std::runtime_error(fulgor_gpu_last_error());
        for (uint32_t j = 0; j != n; ++j) { /* unchanged from here: tools/pseudoalign.cpp:45-50 */
            one.assign(colors.begin() + color_off[j], colors.begin() + color_off[j + 1]);
            options.increment_processed_reads();
            output_buffer.write(ids[j], one);
            if (!one.empty()) options.increment_mapped_reads();
        }
    }
}

/* tools/pseudoalign.cpp:228-369 with essentials::load replaced by fulgor_gpu_index_open and the orchestrator's workers by the
   GPU worker (one handle per worker: a handle serves one call at a time) */
int pseudoalign_gpu(int argc, char** argv) {
    cmd_line_parser::parser parser(argc, argv);
    parser.add("index_filename", "The Fulgor index filename.", "-i", true);
    parser.add("query_filename", "Query filename in FASTA/FASTQ format (optionally gzipped).", "-q", true);
    parser.add("output_filename", "File where output will be written.", "-o", true);
    parser.add("num_threads", "Number of threads (default is 1).", "-t", false);
    parser.add("verbose", "Verbose output during query (default is false).", "--verbose", false, true);
    parser.add("threshold", "Threshold for threshold_union algorithm. It must be a float in (0.0,1.0].", "-r", false);
    parser.add("format", "Format of the output file. Must either ascii, binary, compressed (default is ascii).", "--format", false);
    if (!parser.parse()) return 1;
    auto index_filename = parser.get<std::string>("index_filename");
    auto query_filename = parser.get<std::string>("query_filename");
    auto output_filename = parser.get<std::string>("output_filename");
    auto output_format = parser.parsed("format") ? parser.get<std::string>("format") : "ascii";
    uint64_t num_threads = parser.parsed("num_threads") ? parser.get<uint64_t>("num_threads") : 1;
    if (num_threads == 1) num_threads += 1;
    double threshold = constants::invalid_threshold;
    if (parser.parsed("threshold")) {
        threshold = parser.get<double>("threshold");
        if (threshold <= 0.0 or threshold > 1.0) {
            std::cerr << "threshold must be a float in (0.0,1.0]" << std::endl;
            return 1;
        }
    }
    const auto ps_alg = threshold != constants::invalid_threshold ? pseudoalignment_algorithm::THRESHOLD_UNION : pseudoalignment_algorithm::FULL_INTERSECTION;
    const bool verbose = parser.get<bool>("verbose");
    if (!is_meta_diff(index_filename) && !is_meta(index_filename) && !is_diff(index_filename) && !is_hybrid(index_filename)) {
        std::cerr << "Wrong index filename supplied." << std::endl;
        return 1;
    }
    std::variant<std::monostate, psa_ascii_formatter, psa_binary_formatter, psa_compressed_formatter> formatter;
    if (output_format == "ascii") formatter.emplace<psa_ascii_formatter>(output_filename);
    else if (output_format == "binary") formatter.emplace<psa_binary_formatter>(output_filename);
    else if (output_format == "compressed") formatter.emplace<psa_compressed_formatter>(output_filename);
    else {
        std::cout << "Unknown output format. Supported formats: ascii, binary, compressed." << std::endl;
        return 1;
    }
    ps_options options(ps_alg, verbose, num_threads);
    const uint64_t num_workers = std::min<uint64_t>(num_threads - 1, 4);
    std::vector<fulgor_gpu_index*> handles(num_workers, nullptr);
    if (verbose) essentials::logger("*** START: loading the index");
    for (auto& h : handles)
        if (fulgor_gpu_index_open(index_filename.c_str(), 0, &h)) { /* parses the same .fur / .mfur / .dfur / .mdfur */
            std::cerr << fulgor_gpu_last_error() << std::endl;
            return 1;
        }
    if (verbose) essentials::logger("*** DONE: loading the index");
    fulgor_gpu_info info;
    fulgor_gpu_index_info(handles[0], &info);
    int rc = 0;
    std::visit(
        [&](auto&& fmt) {
            if constexpr (std::is_same_v<std::decay_t<decltype(fmt)>, psa_compressed_formatter>) fmt.set_num_colors(info.num_colors);
            if constexpr (!std::is_same_v<std::decay_t<decltype(fmt)>, std::monostate>) {
                essentials::timer<std::chrono::high_resolution_clock, std::chrono::milliseconds> t;
                t.start();
                fastx_parser::FastxParser<fastx_parser::ReadSeq> rparser({query_filename}, uint32_t(num_workers), 1);
                rparser.start();
                std::vector<std::thread> workers;
                std::atomic<int> failed{0};
                for (uint64_t i = 0; i != num_workers; ++i)
                    workers.emplace_back([&, i]() {
                        try {
                            pseudoalign_worker_gpu(handles[i], rparser, fmt, threshold, options);
                        } catch (std::exception const& e) {
                            std::cerr << e.what() << std::endl;
                            failed = 1;
                        }
                    });
                for (auto& w : workers) w.join();
                rparser.stop();
                t.stop();
                rc = failed;
                if (verbose) {
                    std::cout << "processed " << options.num_reads << " reads" << std::endl;
                    std::cout << "elapsed = " << t.elapsed() << " millisec / " << (t.elapsed() * 1000) / std::max<uint64_t>(1, options.num_reads)
                              << " musec/read" << std::endl;
                    std::cout << "num_mapped_reads " << options.num_mapped_reads << "/" << options.num_reads << std::endl;
                }
            }
        },
        formatter);
    for (auto h : handles) fulgor_gpu_index_close(h);
    return rc;
}

int main(int argc, char** argv) {
    if (argc < 2 || std::string(argv[1]) != "pseudoalign") {
        std::cerr << "usage: fulgor_ref_gpu pseudoalign -i INDEX -q READS -o OUT [-t T] [-r THRESHOLD] [--format ascii|binary|compressed] [--verbose]\n";
        return 1;
    }
    return pseudoalign_gpu(argc - 1, argv + 1);
}
