/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * Translation unit that compiles the UNMODIFIED reference `fulgor` command line
 * straight from the sources where they lie under $REF (= /root/reference).
 * It mirrors the unity-build include list of $REF/tools/fulgor.cpp:4-22 and its
 * tool dispatch (tools/fulgor.cpp:68-109); nothing is copied into this repo.
 * The three GGCAT (Rust) symbols referenced by tools/build.cpp are left
 * unresolved at link time (-Wl,--unresolved-symbols=ignore-all): `fulgor build`
 * therefore cannot run here, but `load`, `color`, `dump`, `stats`, `verify` and
 * `pseudoalign`, `kmer-conservation` and `kmer-matches` are the reference's own code paths.
 *
 * Built by oracle/Makefile into oracle/_ref/fulgor_ref (git-ignored).
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
#include "tools/build.cpp"
#include "tools/pseudoalign.cpp"
#include "tools/kmer_conservation.cpp"
#include "tools/kmer_matches.cpp"

int main(int argc, char** argv) {
    if (argc < 2) {
        std::cerr << "usage: fulgor_ref <pseudoalign|kmer-conservation|kmer-matches|load|dump|color|stats|verify|print-filenames|check> ...\n";
        return 1;
    }
    auto tool = std::string(argv[1]);
    if (tool == "pseudoalign") return pseudoalign(argc - 1, argv + 1);
    if (tool == "load") return load(argc - 1, argv + 1);
    if (tool == "dump") return dump(argc - 1, argv + 1);
    if (tool == "color") return color(argc - 1, argv + 1);
    if (tool == "stats") return stats(argc - 1, argv + 1);
    if (tool == "verify") return verify(argc - 1, argv + 1);
    if (tool == "check") return check(argc - 1, argv + 1);
    if (tool == "print-filenames") return print_filenames(argc - 1, argv + 1);
    if (tool == "kmer-conservation") return kmer_conservation(argc - 1, argv + 1);
    if (tool == "kmer-matches") return kmer_matches(argc - 1, argv + 1);
    std::cerr << "unsupported tool '" << tool << "'\n";
    return 1;
}
