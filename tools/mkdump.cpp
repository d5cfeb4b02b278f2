/*
 * mkdump -- fixture builder (test/bench tooling, not on the query path).
 *
 * Turns a list of FASTA(.gz) genomes into the four text files that the reference's
 * `fulgor load` consumes (format: reference README.md:295-387, parser: src/index.cpp:123-305):
 *     BASE.metadata.txt  BASE.filenames.txt  BASE.color_sets.txt  BASE.unitigs.fa
 * so that a genuine `.fur` can be built by the reference's own CPU tool without GGCAT/Rust,
 * and writes BASE.gpk, a 2-bit packed copy of the genomes used by tools/readgen to draw
 * synthetic reads on machines where the FASTA files are absent (the GPU box).
 *
 * Method: every canonical k-mer gets the set of genomes ("colors") it occurs in; genomes are then
 * walked left to right and cut into paths such that (i) all k-mers of a path have the same color
 * set and (ii) every canonical k-mer is emitted exactly once overall. These paths are valid
 * (not necessarily maximal) monochromatic unitigs: exactly what `load` requires. Color sets are
 * numbered in order of first appearance after sorting by bitmask; unitigs are sorted by color set.
 *
 *   mkdump [-k 31] BASE genome1.fa[.gz] genome2.fa[.gz] ...          (up to 65535 genomes)
 *   mkdump [-k 31] BASE @list.txt                                    (one path per line)
 */
#include <zlib.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <fstream>
#include <map>
#include <vector>

struct contig_t {
    uint32_t genome;
    std::string name;
    std::string seq;
};

static bool read_fasta(const char* path, uint32_t genome, std::vector<contig_t>& out) {
    gzFile f = gzopen(path, "rb");
    if (!f) return false;
    std::vector<char> buf(1 << 20);
    std::string line;
    auto flush_line = [&](std::string const& l) {
        if (l.empty()) return;
        if (l[0] == '>') {
            out.push_back({genome, l.substr(1), std::string()});
        } else if (!out.empty() && out.back().genome == genome) {
            out.back().seq += l;
        }
    };
    int n;
    while ((n = gzread(f, buf.data(), (unsigned)buf.size())) > 0) {
        for (int i = 0; i < n; ++i) {
            char c = buf[i];
            if (c == '\n' || c == '\r') {
                flush_line(line);
                line.clear();
            } else {
                line.push_back(c);
            }
        }
    }
    flush_line(line);
    gzclose(f);
    return true;
}

/* private 2-bit code for this tool only (A0 C1 G2 T3; complement = 3-x) */
static inline int code(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

struct kmer_walker {
    uint64_t k, mask, fwd = 0, rc = 0, valid = 0;
    explicit kmer_walker(uint64_t k) : k(k), mask(k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1)) {}
    /* push one base; returns true when a full valid k-mer ends here */
    bool push(char ch) {
        int c = code(ch);
        if (c < 0) { valid = 0; fwd = rc = 0; return false; }
        fwd = ((fwd << 2) | uint64_t(c)) & mask;
        rc = (rc >> 2) | (uint64_t(3 - c) << (2 * (k - 1)));
        if (valid < k) ++valid;
        return valid == k;
    }
    uint64_t canon() const { return fwd < rc ? fwd : rc; }
};

int main(int argc, char** argv) {
    uint64_t k = 31;
    int a = 1;
    if (a + 1 < argc && !strcmp(argv[a], "-k")) { k = strtoull(argv[a + 1], nullptr, 10); a += 2; }
    if (argc - a < 2) {
        fprintf(stderr, "usage: mkdump [-k K] BASE genome.fa[.gz] ...\n");
        return 1;
    }
    std::string base = argv[a++];
    std::vector<std::string> filenames;
    for (int i = a; i < argc; ++i) {
        if (argv[i][0] == '@') {
            std::ifstream in(argv[i] + 1);
            std::string line;
            while (std::getline(in, line))
                if (!line.empty()) filenames.push_back(line);
        } else {
            filenames.push_back(argv[i]);
        }
    }
    const int num_genomes = (int)filenames.size();
    if (num_genomes > 65535 || num_genomes < 1 || k > 31 || k < 3) {
        fprintf(stderr, "mkdump: 1..65535 genomes, 3 <= k <= 31\n");
        return 1;
    }
    std::vector<contig_t> contigs;
    for (int g = 0; g < num_genomes; ++g) {
        if (!read_fasta(filenames[g].c_str(), g, contigs)) {
            fprintf(stderr, "mkdump: cannot open %s\n", filenames[g].c_str());
            return 1;
        }
    }
    uint64_t total_bases = 0;
    for (auto const& c : contigs) total_bases += c.seq.size();
    fprintf(stderr, "mkdump: %d genomes, %zu contigs, %lu bases\n", num_genomes, contigs.size(),
            (unsigned long)total_bases);

    /* pass 1: canonical k-mer -> ascending list of genomes, interned as a color-set id */
    std::vector<std::pair<uint64_t, uint16_t>> occ;
    occ.reserve(total_bases);
    for (auto const& c : contigs) {
        kmer_walker w(k);
        for (char ch : c.seq)
            if (w.push(ch)) occ.push_back({w.canon(), (uint16_t)c.genome});
    }
    std::sort(occ.begin(), occ.end());
    std::vector<uint64_t> kmers;
    std::vector<uint32_t> masks; /* provisional color-set id per k-mer */
    std::map<std::vector<uint16_t>, uint32_t> interned;
    for (size_t i = 0; i < occ.size();) {
        size_t j = i;
        std::vector<uint16_t> cs;
        while (j < occ.size() && occ[j].first == occ[i].first) {
            if (cs.empty() || cs.back() != occ[j].second) cs.push_back(occ[j].second);
            ++j;
        }
        kmers.push_back(occ[i].first);
        auto it = interned.emplace(std::move(cs), (uint32_t)interned.size()).first;
        masks.push_back(it->second);
        i = j;
    }
    std::vector<std::pair<uint64_t, uint16_t>>().swap(occ);
    const uint64_t num_kmers = kmers.size();
    fprintf(stderr, "mkdump: %lu distinct canonical %lu-mers\n", (unsigned long)num_kmers,
            (unsigned long)k);

    /* final color-set ids: lexicographic order of the genome lists (deterministic) */
    std::vector<std::vector<uint16_t> const*> distinct;
    std::vector<uint32_t> mask_to_id(interned.size());
    for (auto const& kv : interned) {
        mask_to_id[kv.second] = (uint32_t)distinct.size();
        distinct.push_back(&kv.first);
    }

    /* pass 2: cut genomes into monochromatic paths using every k-mer once */
    std::vector<uint8_t> visited(num_kmers, 0);
    std::vector<std::pair<uint32_t, std::string>> unitigs; /* (color set id, sequence) */
    auto find = [&](uint64_t x) {
        return size_t(std::lower_bound(kmers.begin(), kmers.end(), x) - kmers.begin());
    };
    for (auto const& c : contigs) {
        kmer_walker w(k);
        std::string cur;
        uint32_t cur_mask = 0;
        auto close = [&]() {
            if (!cur.empty()) unitigs.push_back({mask_to_id[cur_mask], cur});
            cur.clear();
        };
        for (size_t i = 0; i < c.seq.size(); ++i) {
            if (!w.push(c.seq[i])) { close(); continue; }
            size_t pos = find(w.canon());
            if (visited[pos]) { close(); continue; }
            visited[pos] = 1;
            if (!cur.empty() && masks[pos] == cur_mask) {
                cur.push_back(c.seq[i]);
            } else {
                close();
                cur = c.seq.substr(i + 1 - k, k);
                cur_mask = masks[pos];
            }
        }
        close();
    }
    for (auto& u : unitigs)
        for (auto& ch : u.second) ch = "ACGT"[code(ch)];
    std::stable_sort(unitigs.begin(), unitigs.end(),
                     [](auto const& x, auto const& y) { return x.first < y.first; });
    uint64_t check = 0;
    for (auto const& u : unitigs) check += u.second.size() - k + 1;
    if (check != num_kmers) {
        fprintf(stderr, "mkdump: internal error, %lu k-mers in unitigs vs %lu\n",
                (unsigned long)check, (unsigned long)num_kmers);
        return 1;
    }
    fprintf(stderr, "mkdump: %zu unitigs, %zu color sets\n", unitigs.size(), distinct.size());

    auto open = [&](std::string const& suffix) {
        FILE* f = fopen((base + suffix).c_str(), "w");
        if (!f) { fprintf(stderr, "mkdump: cannot write %s%s\n", base.c_str(), suffix.c_str()); exit(1); }
        return f;
    };
    FILE* f = open(".metadata.txt");
    fprintf(f, "k=%lu\nnum_kmers=%lu\nnum_colors=%d\nnum_unitigs=%zu\nnum_color_sets=%zu\n",
            (unsigned long)k, (unsigned long)num_kmers, num_genomes, unitigs.size(), distinct.size());
    fclose(f);
    f = open(".filenames.txt");
    for (auto const& fn : filenames) {
        size_t s = fn.find_last_of('/');
        fprintf(f, "%s\n", s == std::string::npos ? fn.c_str() : fn.c_str() + s + 1);
    }
    fclose(f);
    f = open(".color_sets.txt");
    for (auto const* cs : distinct) {
        fprintf(f, "size=%zu", cs->size());
        for (uint16_t g : *cs) fprintf(f, " %u", (unsigned)g);
        fprintf(f, "\n");
    }
    fclose(f);
    f = open(".unitigs.fa");
    for (auto const& u : unitigs) fprintf(f, "> color_set_id=%u\n%s\n", u.first, u.second.c_str());
    fclose(f);

    /* BASE.gpk: "FGPK1\0\0\0" u64 num_contigs, u64 total_bases, then per contig {u32 genome, u64 length},
       then ceil(total_bases/4) bytes of 2-bit codes (A0 C1 G2 T3, base i at bits 2(i%4) of byte i/4),
       then u64 num_exceptions and that many u64 global positions whose base is not ACGT (stored as A). */
    f = fopen((base + ".gpk").c_str(), "wb");
    const char magic[8] = {'F', 'G', 'P', 'K', '1', 0, 0, 0};
    fwrite(magic, 1, 8, f);
    uint64_t nc = contigs.size();
    fwrite(&nc, 8, 1, f);
    fwrite(&total_bases, 8, 1, f);
    for (auto const& c : contigs) {
        uint64_t len = c.seq.size();
        fwrite(&c.genome, 4, 1, f);
        fwrite(&len, 8, 1, f);
    }
    std::vector<uint8_t> packed((total_bases + 3) / 4, 0);
    std::vector<uint64_t> exceptions;
    uint64_t g = 0;
    for (auto const& c : contigs)
        for (char ch : c.seq) {
            int x = code(ch);
            if (x < 0) { exceptions.push_back(g); x = 0; }
            packed[g >> 2] |= uint8_t(x << (2 * (g & 3)));
            ++g;
        }
    fwrite(packed.data(), 1, packed.size(), f);
    uint64_t ne = exceptions.size();
    fwrite(&ne, 8, 1, f);
    if (ne) fwrite(exceptions.data(), 8, ne, f);
    fclose(f);
    fprintf(stderr, "mkdump: wrote %s.{metadata,filenames,color_sets}.txt, .unitigs.fa, .gpk (%lu non-ACGT)\n",
            base.c_str(), (unsigned long)ne);
    return 0;
}
