/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * C-callable shim around the UNMODIFIED reference library code, compiled from the
 * sources where they lie under $REF (= /root/reference); nothing is copied here.
 * It exposes exactly the three `index<ColorSets>` member functions that form the
 * hot path (declared $REF/include/index.hpp:39-46):
 *     fetch_color_set_ids            src/ps_full_intersection.cpp:335-374
 *     pseudoalign_full_intersection  src/ps_full_intersection.cpp:377-400
 *     pseudoalign_threshold_union    src/ps_threshold_union.cpp:321-402
 * the two per-k-mer tools that share the lookup:
 *     kmer_conservation              src/kmer_conservation.cpp:7-54
 *     kmer_matches                   src/kmer_matches.cpp:7-30
 * plus the per-k-mer `sshash::streaming_query::lookup_advanced`
 * (external/sshash/include/streaming_query.hpp:50-109) and color-set decoding, so
 * that tests can pin the C restatement (oracle/fulgor_oracle.c) and the CUDA path
 * against the real thing, and bench.py can time the reference on the host cores.
 *
 * Built by oracle/Makefile into oracle/_ref/libfulgor_ref.so (git-ignored).
 */
#include <iostream>
#include <filesystem>
#include <thread>
#include <variant>
#include <cstring>

#include "external/sshash/external/gz/zip_stream.hpp"
#include "external/sshash/external/gz/zip_stream.cpp"
#include "external/sshash/src/build.cpp"
#include "external/sshash/src/dictionary.cpp"
#include "external/sshash/src/info.cpp"

#include "include/index_types.hpp"
#include "src/index.cpp"
#include "src/color_sets.cpp"
#include "src/ps_full_intersection.cpp"
#include "src/ps_threshold_union.cpp"
#include "src/kmer_conservation.cpp"
#include "src/kmer_matches.cpp"

using namespace fulgor;

namespace {

struct ref_handle {
    int type;  // 0 = hybrid (.fur), 1 = meta (.mfur), 2 = differential (.dfur), 3 = meta-differential (.mdfur)
    std::variant<hfur_index_t, mfur_index_t, dfur_index_t, mdfur_index_t> index;
};

/* run f on the loaded index, whatever its color-set type (like std::visit in tools/pseudoalign.cpp:335) */
template <typename F>
auto on_index(void* hp, F&& f) {
    return std::visit(std::forward<F>(f), static_cast<ref_handle*>(hp)->index);
}

bool ends_with(std::string const& s, std::string const& p) {
    return s.size() >= p.size() && std::equal(p.begin(), p.end(), s.end() - p.size());
}

template <typename Index>
void lookup_read(Index const& index, char const* seq, uint64_t len, uint64_t* contig_ids) {
    auto const& dict = index.get_k2u();
    const uint64_t k = dict.k();
    if (len < k) return;
    sshash::streaming_query<kmer_type, true> query(&dict);
    query.reset();
    for (uint64_t i = 0; i != len - k + 1; ++i) {
        auto answer = query.lookup_advanced(seq + i);
        contig_ids[i] = (answer.kmer_id != sshash::constants::invalid_uint64) ? answer.contig_id
                                                                              : uint64_t(-1);
    }
}

struct range_out {
    std::vector<uint64_t> sizes;
    std::vector<uint32_t> values;
};

/* run fn(read i, out vector) over [0,n) with nthreads contiguous ranges, then splice in order */
template <typename Fn>
int run_batch(uint32_t n, int nthreads, uint64_t* off, uint32_t* vals, uint64_t cap, Fn fn) {
    if (nthreads < 1) nthreads = 1;
    std::vector<range_out> outs(nthreads);
    std::vector<std::thread> threads;
    auto work = [&](int t) {
        uint64_t lo = uint64_t(n) * t / nthreads, hi = uint64_t(n) * (t + 1) / nthreads;
        auto& o = outs[t];
        o.sizes.reserve(hi - lo);
        std::vector<uint32_t> res;
        for (uint64_t i = lo; i != hi; ++i) {
            res.clear();
            fn(i, res);
            o.sizes.push_back(res.size());
            o.values.insert(o.values.end(), res.begin(), res.end());
        }
    };
    if (nthreads == 1) {
        work(0);
    } else {
        for (int t = 0; t != nthreads; ++t) threads.emplace_back(work, t);
        for (auto& t : threads) t.join();
    }
    uint64_t total = 0, r = 0;
    off[0] = 0;
    for (auto const& o : outs) {
        for (auto s : o.sizes) {
            total += s;
            off[++r] = total;
        }
    }
    if (total > cap) return -7; /* E2BIG: off[n] holds the required capacity */
    uint64_t pos = 0;
    for (auto const& o : outs) {
        if (!o.values.empty()) std::memcpy(vals + pos, o.values.data(), o.values.size() * 4);
        pos += o.values.size();
    }
    return 0;
}

}  // namespace

extern "C" {

void* fref_open(const char* path) {
    try {
        auto* h = new ref_handle();
        std::string p(path);
        /* type by suffix, like tools/util.cpp:5-19 */
        if (ends_with(p, ".mdfur")) {
            h->type = 3;
            essentials::load(h->index.emplace<mdfur_index_t>(), path);
        } else if (ends_with(p, ".dfur")) {
            h->type = 2;
            essentials::load(h->index.emplace<dfur_index_t>(), path);
        } else if (ends_with(p, ".mfur")) {
            h->type = 1;
            essentials::load(h->index.emplace<mfur_index_t>(), path);
        } else if (ends_with(p, ".fur")) {
            h->type = 0;
            essentials::load(h->index.emplace<hfur_index_t>(), path);
        } else {
            delete h;
            return nullptr;
        }
        return h;
    } catch (std::exception const& e) {
        std::cerr << "fref_open: " << e.what() << std::endl;
        return nullptr;
    }
}

void fref_close(void* hp) { delete static_cast<ref_handle*>(hp); }

/* out: k, m, num_kmers, num_unitigs, num_colors, num_color_sets, type */
void fref_info(void* hp, uint64_t* out) {
    auto* h = static_cast<ref_handle*>(hp);
    auto fill = [&](auto const& idx) {
        out[0] = idx.k();
        out[1] = idx.get_k2u().m();
        out[2] = idx.num_kmers();
        out[3] = idx.num_unitigs();
        out[4] = idx.num_colors();
        out[5] = idx.num_color_sets();
        out[6] = h->type;
    };
    on_index(hp, fill);
}

/* per-k-mer streaming lookup of one read; contig_ids has len-k+1 slots; -1 = negative */
void fref_lookup_read(void* hp, const char* seq, uint64_t len, uint64_t* contig_ids) {
    on_index(hp, [&](auto const& idx) { lookup_read(idx, seq, len, contig_ids); });
}

/* unitig id -> color set id (include/index.hpp:37) */
uint64_t fref_u2c(void* hp, uint64_t unitig_id) {
    return on_index(hp, [&](auto const& idx) -> uint64_t { return idx.u2c(unitig_id); });
}

/* decode color set `id` with the reference iterator; returns its size (writes min(size,cap)) */
int64_t fref_color_set(void* hp, uint64_t id, uint32_t* out, uint64_t cap) {
    uint64_t n = 0;
    auto dec = [&](auto const& idx) {
        auto it = idx.color_set(id);
        const uint64_t size = it.size();
        for (uint64_t i = 0; i != size; ++i, ++it) {
            if (n < cap) out[n] = *it;
            ++n;
        }
    };
    on_index(hp, dec);
    return int64_t(n);
}

/* stage 1 for a batch: CSR of sorted distinct color-set ids per read */
int fref_fetch_color_set_ids(void* hp, const char* bases, const uint64_t* read_off, uint32_t n,
                             uint64_t* cid_off, uint32_t* cids, uint64_t cap, int nthreads) {
    auto run = [&](auto const& idx) {
        return run_batch(n, nthreads, cid_off, cids, cap, [&](uint64_t i, std::vector<uint32_t>& res) {
            std::string seq(bases + read_off[i], read_off[i + 1] - read_off[i]);
            idx.fetch_color_set_ids(seq, res);
        });
    };
    return on_index(hp, run);
}

/* whole path for a batch; algo 0 = full intersection, 1 = threshold union */
int fref_pseudoalign(void* hp, int algo, double threshold, const char* bases,
                     const uint64_t* read_off, uint32_t n, uint64_t* color_off, uint32_t* colors,
                     uint64_t cap, int nthreads) {
    auto run = [&](auto const& idx) {
        return run_batch(n, nthreads, color_off, colors, cap,
                         [&](uint64_t i, std::vector<uint32_t>& res) {
                             std::string seq(bases + read_off[i], read_off[i + 1] - read_off[i]);
                             if (algo == 0) {
                                 /* same call sequence as src/ps_utils.cpp:275-280 followed by
                                    tools/pseudoalign.cpp:29 */
                                 std::vector<uint32_t> cids, tmp;
                                 idx.fetch_color_set_ids(seq, cids);
                                 idx.pseudoalign_full_intersection(cids, res, tmp);
                             } else {
                                 /* tools/pseudoalign.cpp:32 (the CLI's redundant
                                    fetch_color_set_ids in this mode is NOT charged) */
                                 idx.pseudoalign_threshold_union(seq, res, threshold);
                             }
                         });
    };
    return on_index(hp, run);
}

/* index::kmer_conservation for a batch: CSR of triples {start_pos_in_query, num_kmers, color_set_id} */
int fref_kmer_conservation(void* hp, const char* bases, const uint64_t* read_off, uint32_t n,
                           uint64_t* triple_off, uint32_t* triples, uint64_t cap) {
    auto run = [&](auto const& idx) {
        uint64_t total = 0;
        triple_off[0] = 0;
        std::vector<kmer_conservation_triple> info;
        for (uint32_t i = 0; i != n; ++i) {
            std::string seq(bases + read_off[i], read_off[i + 1] - read_off[i]);
            info.clear(); /* like tools/kmer_conservation.cpp:36 (the early return for short reads leaves it untouched) */
            idx.kmer_conservation(seq, info);
            for (auto const& t : info) {
                if (total < cap) {
                    triples[3 * total] = t.start_pos_in_query;
                    triples[3 * total + 1] = t.num_kmers;
                    triples[3 * total + 2] = t.color_set_id;
                }
                ++total;
            }
            triple_off[i + 1] = total;
        }
        return total > cap ? -7 : 0;
    };
    return on_index(hp, run);
}

/* index::kmer_matches for a batch: positive = one byte per k-mer (reads concatenated, kmer_off = n+1 offsets),
   counts = n x num_colors. Reads shorter than k get zero counts here (the reference leaves its buffers untouched). */
int fref_kmer_matches(void* hp, const char* bases, const uint64_t* read_off, uint32_t n,
                      uint64_t* kmer_off, uint8_t* positive, uint64_t cap, uint32_t* counts) {
    auto run = [&](auto const& idx) {
        const uint64_t k = idx.k(), C = idx.num_colors();
        kmer_off[0] = 0;
        for (uint32_t i = 0; i != n; ++i) {
            const uint64_t len = read_off[i + 1] - read_off[i];
            kmer_off[i + 1] = kmer_off[i] + (len >= k ? len - k + 1 : 0);
        }
        if (kmer_off[n] > cap) return -7;
        std::vector<count_type> c(C);
        for (uint32_t i = 0; i != n; ++i) {
            std::string seq(bases + read_off[i], read_off[i + 1] - read_off[i]);
            bits::bit_vector::builder bvb;
            std::fill(c.begin(), c.end(), 0);
            idx.kmer_matches(seq, bvb, c);
            for (uint64_t j = 0; j != bvb.num_bits(); ++j) positive[kmer_off[i] + j] = bvb.get(j);
            std::memcpy(counts + uint64_t(i) * C, c.data(), C * sizeof(uint32_t));
        }
        return 0;
    };
    return on_index(hp, run);
}

}  // extern "C"
