/*
 * fur_reader.cpp -- loads a reference-built Fulgor index (.fur / .mfur, unchanged on-disk format)
 * and flattens it into the device image described in image.h. Host-only C++; no CUDA here.
 *
 * File format: the `essentials` visitor serialisation (reference
 * external/sshash/external/pthash/external/bits/external/essentials/include/essentials.hpp:287-306):
 * PODs are raw little-endian bytes, std::vector<POD> is u64 n + n elements, everything else is the
 * concatenation of its members in `visit` order. Field order per type: SURVEY.md Appendix A, cited
 * at each reader below ("bits/" = external/sshash/external/pthash/external/bits/include/,
 * "pthash/" = external/sshash/external/pthash/include/, "sshash/" = external/sshash/include/).
 */
#include "fur_reader.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>

namespace fgb {
namespace {

typedef unsigned __int128 u128;

struct byte_reader {
    const uint8_t* p;
    const uint8_t* end;
    const uint8_t* take(uint64_t n) {
        if (uint64_t(end - p) < n) throw std::runtime_error("index file truncated");
        const uint8_t* q = p;
        p += n;
        return q;
    }
    template <typename T>
    T pod() {
        T v;
        std::memcpy(&v, take(sizeof(T)), sizeof(T));
        return v;
    }
    /* std::vector<POD> */
    template <typename T>
    std::vector<T> vec() {
        uint64_t n = pod<uint64_t>();
        if (n > uint64_t(end - p) / sizeof(T)) throw std::runtime_error("index file truncated (vector)");
        std::vector<T> v(n);
        if (n) std::memcpy(v.data(), take(n * sizeof(T)), n * sizeof(T));
        return v;
    }
};

/* bits/bit_vector.hpp:343-351 */
struct bit_vector {
    uint64_t num_bits = 0;
    std::vector<uint64_t> words;
    void read(byte_reader& r) {
        num_bits = r.pod<uint64_t>();
        words = r.vec<uint64_t>();
        if (num_bits > words.size() * 64) throw std::runtime_error("malformed bit vector");
    }
};

/* bits/compact_vector.hpp:290-301; value i sits at bit i*width, LSB first */
struct compact_vector {
    uint64_t size = 0, width = 0, mask = 0;
    std::vector<uint64_t> words;
    void read(byte_reader& r) {
        size = r.pod<uint64_t>();
        width = r.pod<uint64_t>();
        mask = r.pod<uint64_t>();
        words = r.vec<uint64_t>();
        if (width > 64 || (width && size > (words.size() * 64) / width)) throw std::runtime_error("malformed compact vector");
    }
    uint64_t operator[](uint64_t i) const {
        if (i >= size) throw std::runtime_error("compact vector index out of range (damaged index file)");
        if (width == 0) return 0;
        const uint64_t pos = i * width, w = pos >> 6, s = pos & 63;
        uint64_t v = words[w] >> s;
        if (s + width > 64) v |= words[w + 1] << (64 - s);
        return v & mask;
    }
};

/* bits/darray.hpp:146-158 -- a select index; the flattener enumerates sequentially and never needs it */
void skip_darray(byte_reader& r) {
    r.pod<uint64_t>();
    r.vec<int64_t>();
    r.vec<uint16_t>();
    r.vec<uint64_t>();
}

/* bits/elias_fano.hpp:303-317; access(i) = ((select1(high, i) - i) << l) | low[i] (:159-163) */
struct elias_fano {
    uint64_t back = 0;
    bit_vector high;
    compact_vector low;
    void read(byte_reader& r) {
        back = r.pod<uint64_t>();
        high.read(r);
        skip_darray(r);
        skip_darray(r);
        low.read(r);
    }
    uint64_t size() const { return low.size; }
    /* all values, by one pass over the ones of the high-bits vector */
    std::vector<uint64_t> decode() const {
        std::vector<uint64_t> out(size());
        uint64_t i = 0;
        const uint64_t l = low.width;
        if (l >= 64) throw std::runtime_error("malformed Elias-Fano sequence (low-bits width)");
        for (uint64_t w = 0; w < high.words.size() && i < out.size(); ++w) {
            uint64_t x = high.words[w];
            while (x && i < out.size()) {
                const uint64_t pos = (w << 6) + uint64_t(__builtin_ctzll(x));
                out[i] = ((pos - i) << l) | low[i];
                ++i;
                x &= x - 1;
            }
        }
        if (i != out.size()) throw std::runtime_error("malformed Elias-Fano sequence");
        return out;
    }
};

/* pthash/utils/hasher.hpp:53-117 for an 8-byte key */
inline uint64_t murmur2_64(uint64_t key, uint64_t seed) {
    const uint64_t m = 0xc6a4a7935bd1e995ULL;
    uint64_t h = seed ^ (8 * m);
    uint64_t k = key * m;
    k ^= k >> 47;
    k *= m;
    h ^= k;
    h *= m;
    h ^= h >> 47;
    h *= m;
    h ^= h >> 47;
    return h;
}

/* canonical minimizer of a k-mer WITH its position, the same arithmetic as kernels.cuh (util::compute_minimizer,
   sshash/util.hpp:220-239; mixer_64, sshash/hash_util.hpp:97). Returns false when both strands give the same value
   (position ambiguous). pos = base offset of the minimizer inside the k-mer, in forward (string) coordinates. */
inline uint64_t revcomp_bits(uint64_t x, uint32_t n) {
    uint64_t c = x ^ 0xAAAAAAAAAAAAAAAAULL, y = 0;
    for (uint32_t i = 0; i < 32; ++i) {
        y = (y << 2) | (c & 3);
        c >>= 2;
    }
    return y >> (64 - 2 * n);
}
inline void strand_minimizer(uint64_t x, uint32_t window, uint64_t mmer_mask, uint64_t magic, uint64_t& value, uint32_t& pos) {
    uint64_t best_h = UINT64_MAX;
    value = UINT64_MAX;
    pos = 0;
    for (uint32_t i = 0; i < window; ++i) {
        const uint64_t y = x & mmer_mask;
        const uint64_t h = (y * 0x517cc1b727220a95ULL) ^ magic;
        if (h < best_h) {
            best_h = h;
            value = y;
            pos = i;
        }
        x >>= 2;
    }
}
inline bool canonical_minimizer_pos(uint64_t fwd, uint32_t k, uint32_t m, uint64_t magic, uint64_t& value, uint32_t& pos) {
    const uint64_t mmer_mask = (1ULL << (2 * m)) - 1;
    uint64_t vf, vr;
    uint32_t jf, jr;
    strand_minimizer(fwd, k - m + 1, mmer_mask, magic, vf, jf);
    strand_minimizer(revcomp_bits(fwd, k), k - m + 1, mmer_mask, magic, vr, jr);
    value = vf < vr ? vf : vr;
    pos = vf < vr ? jf : (k - m) - jr;
    return vf != vr;
}

/* growing image with aligned sections */
struct image_writer {
    std::vector<uint8_t> bytes;
    uint64_t begin_section() {
        bytes.resize((bytes.size() + FGI_ALIGN - 1) / FGI_ALIGN * FGI_ALIGN, 0);
        return bytes.size();
    }
    template <typename T>
    uint64_t section(std::vector<T> const& v, uint64_t pad_elems = 0) {
        uint64_t off = begin_section();
        bytes.resize(off + (v.size() + pad_elems) * sizeof(T), 0);
        if (!v.empty()) std::memcpy(bytes.data() + off, v.data(), v.size() * sizeof(T));
        return off;
    }
};

struct flattener {
    fgi_header H{};
    std::vector<fgi_phf> phfs;
    std::vector<fgi_phf_part> parts;
    std::vector<uint64_t> hashed_pilots;
    std::vector<uint32_t> free_slots;
    std::vector<uint32_t> skew_positions;
    std::vector<fgi_hybrid> hybrids;
    std::vector<uint64_t> set_bit_off;
    std::vector<uint64_t> color_words;

    static void split(u128 v, uint64_t& lo, uint64_t& hi) {
        lo = uint64_t(v);
        hi = uint64_t(v >> 64);
    }

    /* pthash/single_phf.hpp:140-150; pilots pthash/utils/encoders.hpp:207-215,406-414 */
    void read_single_phf(byte_reader& r, uint64_t partition_offset) {
        fgi_phf_part P{};
        P.seed = r.pod<uint64_t>();
        P.num_keys = r.pod<uint64_t>();
        P.table_size = r.pod<uint64_t>();
        r.pod<uint64_t>(); /* M_128 (fastmod constant of table_size): recomputed below in 64-bit form */
        r.pod<uint64_t>();
        r.pod<uint64_t>(); /* M_64 */
        P.num_dense = r.pod<uint64_t>();
        P.num_sparse = r.pod<uint64_t>();
        r.pod<uint64_t>(); /* M_dense, M_sparse */
        r.pod<uint64_t>();
        r.pod<uint64_t>();
        r.pod<uint64_t>();
        auto inverse = [](uint64_t d) -> uint64_t {
            if (d >= (1ULL << 32)) throw std::runtime_error("MPHF partition with a modulus >= 2^32 is not supported");
            if (d == 0) return 0; /* never used: mod_by_inverse answers 0 for a zero modulus (kernels.cuh) */
            if (d == 1) return UINT64_MAX;
            return uint64_t((u128(1) << 64) / d);
        };
        P.inv_table = inverse(P.table_size);
        P.inv_dense = inverse(P.num_dense);
        P.inv_sparse = inverse(P.num_sparse);
        compact_vector front_ranks, front_dict, back_ranks, back_dict;
        front_ranks.read(r);
        front_dict.read(r);
        back_ranks.read(r);
        back_dict.read(r);
        elias_fano fs;
        fs.read(r);
        P.offset = partition_offset;
        P.pilot_base = hashed_pilots.size();
        P.free_base = free_slots.size();
        /* dual<dictionary,dictionary>::access (encoders.hpp:391-394,192-195), with the pilot hash of
           single_phf::position (single_phf.hpp:84-87) applied once here */
        for (uint64_t i = 0; i < front_ranks.size; ++i)
            hashed_pilots.push_back(murmur2_64(front_dict[front_ranks[i]], P.seed));
        for (uint64_t i = 0; i < back_ranks.size; ++i)
            hashed_pilots.push_back(murmur2_64(back_dict[back_ranks[i]], P.seed));
        if (P.num_keys >= (1ULL << 32)) throw std::runtime_error("MPHF partition with >= 2^32 keys is not supported");
        /* what phf_position dereferences on the GPU: one pilot per PTHash bucket, one free slot per table position beyond the
           keys, every free slot a valid key position */
        if (front_ranks.size + back_ranks.size != P.num_dense + P.num_sparse) throw std::runtime_error("damaged index: MPHF pilot count differs from its bucket count");
        if (P.table_size < P.num_keys || P.table_size == 0) throw std::runtime_error("damaged index: MPHF table smaller than its key set");
        if (fs.size() != P.table_size - P.num_keys) throw std::runtime_error("damaged index: MPHF free-slot count differs from table_size - num_keys");
        for (uint64_t v : fs.decode()) {
            if (v >= P.num_keys) throw std::runtime_error("damaged index: MPHF free slot outside the key range");
            free_slots.push_back(uint32_t(v));
        }
        parts.push_back(P);
    }

    /* pthash/partitioned_phf.hpp:203-210,23-43; range_bucketer utils/bucketers.hpp:244-248 */
    uint32_t read_partitioned_phf(byte_reader& r) {
        fgi_phf F{};
        F.seed = r.pod<uint64_t>();
        F.num_keys = r.pod<uint64_t>();
        r.pod<uint64_t>(); /* table_size */
        F.num_partitions = r.pod<uint64_t>();
        r.pod<uint64_t>();
        r.pod<uint64_t>(); /* range_bucketer M (unused) */
        const uint64_t n = r.pod<uint64_t>();
        if (n > uint64_t(r.end - r.p)) throw std::runtime_error("index file truncated (partitions)");
        F.first_part = parts.size();
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t off = r.pod<uint64_t>();
            read_single_phf(r, off);
        }
        if (n == 0) {
            /* a never-built (empty) skew partition: seed/num_keys/table_size are uninitialised words
               in the file (partitioned_phf.hpp:213-215 has no default initialisers) */
            F.num_partitions = 0;
            F.num_keys = 0;
        } else if (n != F.num_partitions) {
            throw std::runtime_error("partitioned MPHF: bucketer/partition count mismatch");
        }
        phfs.push_back(F);
        return uint32_t(phfs.size() - 1);
    }

    /* include/color_sets/hybrid.hpp:339-345 */
    void read_hybrid(byte_reader& r) {
        fgi_hybrid h{};
        h.num_colors = r.pod<uint32_t>();
        h.sparse_thr = r.pod<uint32_t>();
        h.very_dense_thr = r.pod<uint32_t>();
        elias_fano offs;
        offs.read(r);
        bit_vector bits;
        bits.read(r);
        if (offs.size() == 0) throw std::runtime_error("hybrid color sets: empty offsets");
        h.num_sets = offs.size() - 1;
        h.set_off_base = set_bit_off.size();
        h.word_base = color_words.size();
        for (uint64_t v : offs.decode()) set_bit_off.push_back(v);
        color_words.insert(color_words.end(), bits.words.begin(), bits.words.end());
        color_words.push_back(0); /* the decoders read 128-bit windows */
        color_words.push_back(0);
        hybrids.push_back(h);
    }

    /* include/color_sets/differential.hpp:322-339. A set is stored as its symmetric difference with the representative of
       its cluster (:8-96); color_set(i) pairs m_color_set_offsets[i] with m_representative_offsets[rank1(m_clusters, i)]
       (:289-295), where m_clusters has a one at the last set of every cluster. Both offset sequences become plain arrays
       here: num_sets + 1 bit offsets of the difference lists (the last one = end of the stream), then num_sets bit offsets
       of the representative each set refers to. */
    void read_differential(byte_reader& r) {
        fgi_hybrid h{};
        h.num_colors = r.pod<uint32_t>();
        h.kind = FGI_SETS_DIFFERENTIAL;
        elias_fano rep_offs, set_offs;
        rep_offs.read(r);
        set_offs.read(r);
        bit_vector bits, clusters;
        bits.read(r);
        clusters.read(r);
        r.vec<uint64_t>(); /* rank9 index over m_clusters: replaced by the sequential pass below */
        h.num_sets = set_offs.size();
        if (clusters.num_bits != h.num_sets) throw std::runtime_error("differential color sets: clusters/offsets size mismatch");
        h.set_off_base = set_bit_off.size();
        h.word_base = color_words.size();
        for (uint64_t v : set_offs.decode()) set_bit_off.push_back(v);
        set_bit_off.push_back(bits.num_bits);
        const std::vector<uint64_t> reps = rep_offs.decode();
        uint64_t cluster = 0;
        for (uint64_t i = 0; i < h.num_sets; ++i) {
            if (cluster >= reps.size()) throw std::runtime_error("differential color sets: more clusters than representatives");
            set_bit_off.push_back(reps[cluster]);
            cluster += (clusters.words[i >> 6] >> (i & 63)) & 1;
        }
        color_words.insert(color_words.end(), bits.words.begin(), bits.words.end());
        color_words.push_back(0);
        color_words.push_back(0);
        hybrids.push_back(h);
    }
};

/* LSB-first reader over a bit_vector for the load-time decoding of the meta-differential lists
   (bits/bit_vector.hpp:234-294, Elias delta bits/integer_codes.hpp:54-71) */
struct host_bit_cursor {
    const bit_vector& b;
    uint64_t pos;
    uint64_t take(uint64_t l) {
        if (l == 0) return 0;
        if (pos + l > b.num_bits) throw std::runtime_error("bit stream overrun while decoding color-set lists");
        const uint64_t w = pos >> 6, s = pos & 63;
        uint64_t v = b.words[w] >> s;
        if (s + l > 64) v |= b.words[w + 1] << (64 - s);
        pos += l;
        return l == 64 ? v : (v & ((1ULL << l) - 1));
    }
    uint64_t unary() {
        uint64_t n = 0;
        while (take(1) == 0) ++n;
        return n;
    }
    uint64_t gamma() {
        const uint64_t n = unary();
        return (take(n) | (1ULL << n)) - 1;
    }
    uint64_t delta() {
        const uint64_t n = gamma();
        return (take(n) | (1ULL << n)) - 1;
    }
};

bool ends_with(std::string const& s, const char* suf) {
    const size_t n = std::strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

}  // namespace

std::vector<uint8_t> build_image(const uint8_t* file, uint64_t size, int type) {
    byte_reader r{file, file + size};
    flattener F;
    fgi_header& H = F.H;
    H.magic = FGI_MAGIC;
    H.type = uint32_t(type);

    /* include/index.hpp:94-102 begins with the version (include/util.hpp:31-35,91-95) */
    const uint8_t major = r.pod<uint8_t>();
    r.pod<uint8_t>();
    r.pod<uint8_t>();
    if (major != 4) throw std::runtime_error("MAJOR index version mismatch: Fulgor index must be rebuilt with the current library version");
    /* sshash::dictionary (sshash/dictionary.hpp:141-154); version check sshash/util.hpp:149-153 */
    const uint8_t smajor = r.pod<uint8_t>();
    r.pod<uint8_t>();
    r.pod<uint8_t>();
    if (smajor != 4) throw std::runtime_error("MAJOR index version mismatch: SSHash index must be rebuilt with the current library version");
    H.num_kmers = r.pod<uint64_t>();
    H.k = r.pod<uint16_t>();
    H.m = r.pod<uint16_t>();
    const uint8_t canonical = r.pod<uint8_t>();
    if (!canonical) throw std::runtime_error("only canonical SSHash dictionaries are supported (Fulgor always builds canonical ones)");
    if (H.k == 0 || H.k > 31 || H.m == 0 || H.m > H.k) throw std::runtime_error("unsupported k/m (k must be <= 31)");
    H.hash_magic = r.pod<uint64_t>();
    {
        static const float a = 0.6f; /* pthash/utils/util.hpp:26: a FLOAT constant */
        H.bucketer_T = uint64_t(a * double(UINT64_MAX));
    }
    F.read_partitioned_phf(r); /* phfs[0]: minimizers (sshash/minimizers.hpp) */

    /* buckets (sshash/buckets.hpp:330-336) */
    elias_fano pieces_ef, nskb_ef;
    pieces_ef.read(r);
    nskb_ef.read(r);
    compact_vector offsets;
    offsets.read(r);
    bit_vector strings;
    strings.read(r);

    /* skew index (sshash/skew_index.hpp:84-91) */
    H.skew_min_log2 = r.pod<uint16_t>();
    H.skew_max_log2 = r.pod<uint16_t>();
    H.skew_log2_max_bucket = r.pod<uint32_t>();
    const uint64_t n_skew = r.pod<uint64_t>();
    if (n_skew > FGI_MAX_SKEW) throw std::runtime_error("too many skew-index partitions");
    H.num_skew = uint32_t(n_skew);
    for (uint64_t i = 0; i < n_skew; ++i) H.skew_phf[i] = F.read_partitioned_phf(r);
    if (r.pod<uint64_t>() != n_skew) throw std::runtime_error("skew index: mphfs/positions size mismatch");
    for (uint64_t i = 0; i < n_skew; ++i) {
        compact_vector pos;
        pos.read(r);
        H.skew_pos_base[i] = F.skew_positions.size();
        for (uint64_t j = 0; j < pos.size; ++j) F.skew_positions.push_back(uint32_t(pos[j]));
    }
    { /* weights (sshash/weights.hpp:183-187): empty in Fulgor indexes */
        compact_vector a, c;
        elias_fano b;
        a.read(r);
        b.read(r);
        c.read(r);
    }

    /* u2c + rank9 (include/index.hpp:94-102); the rank index is rebuilt implicitly below */
    bit_vector u2c;
    u2c.read(r);
    r.vec<uint64_t>();

    if (type == 0) {
        F.read_hybrid(r);
        H.num_colors = F.hybrids[0].num_colors;
        H.num_color_sets = F.hybrids[0].num_sets;
        H.num_partitions = 1;
    }
    std::vector<uint64_t> meta_off;
    std::vector<uint32_t> meta_vals, part_min_color, part_sets_before;
    if (type == 1) { /* include/color_sets/meta.hpp:275-281 */
        H.num_colors = r.pod<uint32_t>();
        compact_vector meta_sets;
        meta_sets.read(r);
        elias_fano moff;
        moff.read(r);
        const uint64_t np = r.pod<uint64_t>();
        if (np > uint64_t(r.end - r.p)) throw std::runtime_error("index file truncated (partial color sets)");
        for (uint64_t i = 0; i < np; ++i) F.read_hybrid(r);
        struct endpoint { uint32_t min_color, num_color_sets_before; }; /* meta.hpp:9-17 */
        std::vector<endpoint> eps = r.vec<endpoint>();
        if (eps.size() != np + 1) throw std::runtime_error("meta color sets: endpoints/partitions size mismatch");
        meta_off = moff.decode();
        if (meta_off.empty()) throw std::runtime_error("meta color sets: empty offsets");
        H.num_color_sets = meta_off.size() - 1;
        H.num_partitions = uint32_t(np);
        meta_vals.resize(meta_sets.size);
        for (uint64_t i = 0; i < meta_sets.size; ++i) meta_vals[i] = uint32_t(meta_sets[i]);
        for (auto const& e : eps) {
            part_min_color.push_back(e.min_color);
            part_sets_before.push_back(e.num_color_sets_before);
        }
    }
    if (type == 2) { /* include/color_sets/differential.hpp:322-339 */
        F.read_differential(r);
        H.num_colors = F.hybrids[0].num_colors;
        H.num_color_sets = F.hybrids[0].num_sets;
        H.num_partitions = 1;
    }
    if (type == 3) { /* include/color_sets/meta_differential.hpp:306-331 */
        H.num_colors = r.pod<uint32_t>();
        r.pod<uint32_t>(); /* m_num_partition_sets */
        elias_fano ps_offs, rel_offs;
        ps_offs.read(r);
        rel_offs.read(r);
        struct endpoint { uint64_t min_color, num_color_sets; }; /* meta_differential.hpp:8-16 */
        std::vector<endpoint> eps = r.vec<endpoint>();
        const uint64_t np = r.pod<uint64_t>();
        if (np > uint64_t(r.end - r.p)) throw std::runtime_error("index file truncated (partial color sets)");
        if (eps.size() != np || np == 0) throw std::runtime_error("meta-differential color sets: endpoints/partitions size mismatch");
        for (uint64_t i = 0; i < np; ++i) F.read_differential(r);
        bit_vector relative_colors, partition_sets, ps_partitions;
        relative_colors.read(r);
        partition_sets.read(r);
        ps_partitions.read(r);
        r.vec<uint64_t>(); /* rank9 index over m_partition_sets_partitions */
        const std::vector<uint64_t> pso = ps_offs.decode(), rco = rel_offs.decode();
        if (rco.empty()) throw std::runtime_error("meta-differential color sets: empty offsets");
        H.num_color_sets = rco.size() - 1;
        H.num_partitions = uint32_t(np);
        if (ps_partitions.num_bits != H.num_color_sets) throw std::runtime_error("meta-differential color sets: partition-set marks/offsets size mismatch");
        uint64_t before = 0;
        for (uint64_t p = 0; p < np; ++p) {
            if (eps[p].num_color_sets != F.hybrids[p].num_sets) throw std::runtime_error("meta-differential color sets: endpoint/partition set count mismatch");
            part_min_color.push_back(uint32_t(eps[p].min_color));
            part_sets_before.push_back(uint32_t(before));
            before += eps[p].num_color_sets;
        }
        part_min_color.push_back(H.num_color_sets ? H.num_colors : 0);
        part_sets_before.push_back(uint32_t(before));
        /* every set's list of partial sets, decoded once into the records the meta index uses ([n, meta color 1..n], meta
           color = sets before its partition + relative id): the partition ids come from the delta-coded partition set shared
           by a group of color sets (forward_iterator::init / read_partition_id, meta_differential.hpp:126-193; the group of
           set i is rank1(m_partition_sets_partitions, i), :286-293), the relative ids from fixed-width fields of
           msb(num_color_sets of the partition) + 1 bits (:184-191). */
        uint64_t group = 0;
        std::vector<uint32_t> parts_of_group;
        bool group_loaded = false;
        for (uint64_t i = 0; i < H.num_color_sets; ++i) {
            if (!group_loaded) {
                if (group >= pso.size()) throw std::runtime_error("meta-differential color sets: more groups than partition sets");
                host_bit_cursor c{partition_sets, pso[group]};
                const uint64_t n = c.delta();
                parts_of_group.clear();
                uint64_t pid = 0;
                for (uint64_t j = 0; j < n; ++j) {
                    pid += c.delta();
                    if (pid >= np) throw std::runtime_error("meta-differential color sets: partition id out of range");
                    parts_of_group.push_back(uint32_t(pid));
                }
                group_loaded = true;
            }
            meta_off.push_back(meta_vals.size());
            meta_vals.push_back(uint32_t(parts_of_group.size()));
            host_bit_cursor rc{relative_colors, rco[i]};
            for (uint32_t pid : parts_of_group) {
                if (eps[pid].num_color_sets == 0) throw std::runtime_error("meta-differential color sets: partition without color sets");
                const uint64_t width = 64 - uint64_t(__builtin_clzll(eps[pid].num_color_sets));
                const uint64_t rel = rc.take(width);
                if (rel >= eps[pid].num_color_sets) throw std::runtime_error("meta-differential color sets: relative id out of range");
                meta_vals.push_back(part_sets_before[pid] + uint32_t(rel));
            }
            if ((ps_partitions.words[i >> 6] >> (i & 63)) & 1) {
                ++group;
                group_loaded = false;
            }
        }
        meta_off.push_back(meta_vals.size());
    }
    /* filenames (include/filenames.hpp:37-41) are not needed on the device */
    r.vec<uint32_t>();
    r.vec<char>();
    if (r.p != r.end) throw std::runtime_error("index file has trailing bytes (wrong index type for this suffix?)");

    /* ---- derived arrays ---- */
    const std::vector<uint64_t> pieces = pieces_ef.decode();
    const std::vector<uint64_t> nskb = nskb_ef.decode();
    if (pieces.size() < 2 || nskb.empty()) throw std::runtime_error("empty dictionary");
    H.num_unitigs = pieces.size() - 1;
    H.num_minimizers = nskb.size() - 1;
    H.num_super_kmers = offsets.size;
    /* the differential builders allocate one bit more (include/builders/differential_builder.hpp:335-336) */
    if (u2c.num_bits != H.num_unitigs && !(type >= 2 && u2c.num_bits == H.num_unitigs + 1))
        throw std::runtime_error("u2c size does not match the number of unitigs");
    if (H.num_super_kmers >= (1ULL << 32) || pieces.back() >= (1ULL << 31))
        throw std::runtime_error("dictionary too large for 32-bit super-k-mer ids / 31-bit string offsets");
    if (H.k - H.m + 1 > 31) throw std::runtime_error("k - m + 1 > 31 is not supported");

    /* structural checks the kernels rely on (a reference-built file passes all of them; the reference itself does not look) */
    if (2 * pieces.back() > strings.num_bits || strings.words.size() * 64 < strings.num_bits)
        throw std::runtime_error("unitig end-points run past the strings");
    if (u2c.words.size() * 64 < u2c.num_bits) throw std::runtime_error("u2c bit vector shorter than its size");
    for (size_t i = 0; i < F.hybrids.size(); ++i) { /* color-set bit offsets: ascending, inside the container's stream */
        const fgi_hybrid& h = F.hybrids[i];
        const uint64_t stream_bits = ((i + 1 < F.hybrids.size() ? F.hybrids[i + 1].word_base : F.color_words.size()) - h.word_base - 2) * 64;
        const uint64_t* so = F.set_bit_off.data() + h.set_off_base;
        for (uint64_t j = 0; j <= h.num_sets; ++j)
            if (so[j] > stream_bits || (j && so[j] < so[j - 1])) throw std::runtime_error("color-set offsets are not ascending inside their bit stream");
        if (h.kind == FGI_SETS_DIFFERENTIAL)
            for (uint64_t j = 0; j < h.num_sets; ++j)
                if (so[h.num_sets + 1 + j] >= stream_bits) throw std::runtime_error("representative offset outside the color-set bit stream");
    }
    if (type == 1 || type == 3) { /* meta lists: [n, meta color 1..n] records, meta colors inside the partial sets */
        if (meta_off.empty() || meta_off.back() != meta_vals.size()) throw std::runtime_error("meta color lists do not fill their vector");
        const uint64_t total_partial = part_sets_before.empty() ? 0 : part_sets_before.back();
        for (uint64_t c = 0; c + 1 < meta_off.size(); ++c) {
            const uint64_t b = meta_off[c], e = meta_off[c + 1];
            if (b >= e || e > meta_vals.size() || meta_vals[b] != e - b - 1) throw std::runtime_error("malformed meta color list");
            for (uint64_t j = b + 1; j < e; ++j)
                if (meta_vals[j] >= total_partial) throw std::runtime_error("meta color out of range");
        }
    }

    /* buckets::locate_bucket (sshash/buckets.hpp:62-67): begin(b) = EF[b] + b */
    std::vector<uint32_t> bucket_begin(nskb.size());
    for (uint64_t b = 0; b < nskb.size(); ++b) {
        if (b && nskb[b] < nskb[b - 1]) throw std::runtime_error("damaged index: bucket sizes are not ascending");
        bucket_begin[b] = uint32_t(nskb[b] + b);
    }
    if (nskb.back() + nskb.size() - 1 != H.num_super_kmers) throw std::runtime_error("bucket sizes do not add up to the number of super-k-mers");
    /* what the MPHF lookups index on the GPU: the minimizer MPHF answers in [0, num_minimizers), skew MPHF i in
       [0, its positions table), and every partition's key range lies inside its function's */
    for (size_t f = 0; f < F.phfs.size(); ++f) {
        const fgi_phf& P = F.phfs[f];
        uint64_t limit = H.num_minimizers;
        if (f > 0) {
            limit = 0;
            for (uint32_t i = 0; i < H.num_skew; ++i)
                if (H.skew_phf[i] == f) limit = (i + 1 < H.num_skew ? H.skew_pos_base[i + 1] : F.skew_positions.size()) - H.skew_pos_base[i];
        }
        if (P.num_keys > limit) throw std::runtime_error("damaged index: MPHF with more keys than the table it indexes");
        for (uint64_t j = 0; j < P.num_partitions; ++j) {
            const fgi_phf_part& part = F.parts[P.first_part + j];
            if (part.offset > P.num_keys || part.num_keys > P.num_keys - part.offset) throw std::runtime_error("damaged index: MPHF partition outside its function's key range");
        }
    }

    /* index::u2c (include/index.hpp:37): color-set id of unitig u = number of ones in u2c[0, u) */
    std::vector<uint32_t> unitig_cid(H.num_unitigs);
    {
        uint32_t rank = 0;
        for (uint64_t u = 0; u < H.num_unitigs; ++u) {
            unitig_cid[u] = rank;
            rank += uint32_t((u2c.words[u >> 6] >> (u & 63)) & 1);
        }
    }
    if (H.num_unitigs && unitig_cid.back() >= H.num_color_sets) throw std::runtime_error("u2c marks more color sets than the index holds");
    /* per super-k-mer: offset, window = min(k-m+1, contig_end - offset - k + 1) (buckets.hpp:133-160), the color-set id
       of the unitig that contains it (buckets.hpp:13-40 + u2c), and the position of the canonical minimizer */
    std::vector<uint64_t> sk_records(H.num_super_kmers);
    std::vector<uint32_t> sk_cid;
    /* color-set ids beyond the record's 21 bits go to a side array (sk_cid). FULGOR_GPU_FORCE_WIDE_CIDS=1 takes that layout for any
       index: no index in the test tiers has 2^21 color sets, the tests use the switch to run the kernels' side-array branch */
    const char* force_wide = std::getenv("FULGOR_GPU_FORCE_WIDE_CIDS");
    const bool wide_cids = H.num_color_sets > (uint64_t(FGI_SK_CID_MASK) + 1) || (force_wide && force_wide[0] == '1');
    if (wide_cids) sk_cid.resize(H.num_super_kmers);
    const uint64_t max_window = H.k - H.m + 1;
    const uint64_t kmask = (1ULL << (2 * H.k)) - 1;
    std::vector<uint64_t> strings_padded(strings.words);
    strings_padded.push_back(0);
    strings_padded.push_back(0);
    const unsigned nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<uint64_t> unpinned(nthreads, 0);
    std::vector<std::string> errors(nthreads);
    auto work = [&](unsigned tid) {
        const uint64_t lo = H.num_super_kmers * tid / nthreads, hi = H.num_super_kmers * (tid + 1) / nthreads;
        for (uint64_t s = lo; s < hi; ++s) {
            const uint64_t off = offsets[s];
            const uint64_t u = uint64_t(std::upper_bound(pieces.begin(), pieces.end(), off) - pieces.begin()) - 1;
            if (u >= H.num_unitigs) {
                errors[tid] = "super-k-mer offset outside the strings";
                return;
            }
            const uint64_t contig_end = pieces[u + 1];
            uint64_t window = 0;
            if (contig_end >= off + H.k) window = std::min<uint64_t>(max_window, contig_end - off - H.k + 1);
            /* The reference scans up to k-m+1 k-mers from the offset (buckets.hpp:133-160), which can run past the end of
               the super-k-mer into k-mers of OTHER buckets. Only k-mers whose canonical minimizer equals the one of the
               first k-mer (the bucket's key) can ever match a query routed to this bucket, so the window stops there. */
            bool pinned = window > 0;
            uint32_t pm = 0;
            uint64_t v0 = 0;
            for (uint64_t t = 0; t < window; ++t) {
                const uint64_t bit = 2 * (off + t), w = bit >> 6, sh = bit & 63;
                const uint64_t x = (sh ? (strings_padded[w] >> sh) | (strings_padded[w + 1] << (64 - sh)) : strings_padded[w]) & kmask;
                uint64_t value;
                uint32_t pos;
                const bool unique = canonical_minimizer_pos(x, H.k, H.m, H.hash_magic, value, pos);
                if (t == 0) {
                    v0 = value;
                    pm = pos;
                } else if (value != v0) {
                    window = t;
                    break;
                }
                if (!unique || uint32_t(t) + pos != pm) pinned = false;
            }
            if (!pinned) {
                pm = 0;
                ++unpinned[tid];
            }
            /* does the stored m-mer at the minimizer position read as the canonical minimizer itself (1) or as its reverse complement (0)? */
            uint32_t canon_fwd = 0;
            if (pinned) {
                const uint64_t bit = 2 * (off + pm), w = bit >> 6, sh = bit & 63;
                const uint64_t y = (sh ? (strings_padded[w] >> sh) | (strings_padded[w + 1] << (64 - sh)) : strings_padded[w]) & ((1ULL << (2 * H.m)) - 1);
                canon_fwd = y == v0;
            }
            uint32_t hi32 = (uint32_t(window) << FGI_SK_WINDOW_SHIFT) | (pm << FGI_SK_PM_SHIFT) | (uint32_t(pinned) << FGI_SK_PINNED_SHIFT);
            if (wide_cids) sk_cid[s] = unitig_cid[u];
            else hi32 |= unitig_cid[u];
            sk_records[s] = uint64_t(uint32_t(off)) | (uint64_t(canon_fwd) << FGI_SK_CANON_FWD_SHIFT) | (uint64_t(hi32) << 32);
        }
    };
    {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nthreads; ++t) pool.emplace_back(work, t);
        for (auto& t : pool) t.join();
    }
    for (auto const& e : errors)
        if (!e.empty()) throw std::runtime_error(e);
    for (uint64_t v : unpinned) H.num_unpinned += v;

    /* util::compute_minimizer (sshash/util.hpp:220-239) starts from min_hash = UINT64_MAX with a strict '<', so a k-mer whose
       m-mers ALL hash to UINT64_MAX gets the sentinel UINT64_MAX as its "minimizer". The hash is a bijection, so this needs all
       k-m+1 m-mers of the k-mer to be equal, i.e. one of the four homopolymers. Tell the kernels whether that can happen. */
    for (uint64_t base = 0; base < 4; ++base) {
        uint64_t y = 0;
        for (uint32_t i = 0; i < H.m; ++i) y |= base << (2 * i);
        if (((y * 0x517cc1b727220a95ULL) ^ H.hash_magic) == UINT64_MAX) H.guard_max_hash = 1;
    }

    H.num_string_words = strings.words.size();
    H.num_phfs = uint32_t(F.phfs.size());
    H.num_phf_parts = uint32_t(F.parts.size());

    /* ---- emit ---- */
    image_writer W;
    W.bytes.resize(sizeof(fgi_header), 0);
    H.off_phfs = W.section(F.phfs);
    H.off_phf_parts = W.section(F.parts);
    H.off_hashed_pilots = W.section(F.hashed_pilots);
    H.off_free_slots = W.section(F.free_slots, 1);
    H.off_bucket_begin = W.section(bucket_begin);
    H.off_sk_records = W.section(sk_records);
    H.off_sk_cid = wide_cids ? W.section(sk_cid) : 0;
    H.off_strings = W.section(strings.words, 2);
    H.off_skew_positions = W.section(F.skew_positions, 1);
    H.off_hybrids = W.section(F.hybrids);
    H.off_set_bit_off = W.section(F.set_bit_off);
    H.off_color_words = W.section(F.color_words);
    H.off_meta_off = W.section(meta_off, 1);
    H.off_meta_vals = W.section(meta_vals, 1);
    H.off_part_min_color = W.section(part_min_color, 1);
    H.off_part_sets_before = W.section(part_sets_before, 1);
    W.begin_section();
    H.total_bytes = W.bytes.size();
    std::memcpy(W.bytes.data(), &H, sizeof(H));
    return std::move(W.bytes);
}

int index_type_from_path(const char* path) {
    /* the reference infers the index type from the file suffix only (tools/util.cpp:5-19) */
    std::string p(path);
    if (ends_with(p, ".mdfur")) return 3;
    if (ends_with(p, ".dfur")) return 2;
    if (ends_with(p, ".mfur")) return 1;
    if (ends_with(p, ".fur")) return 0;
    return -1;
}

std::vector<uint8_t> build_image_from_file(const char* path) {
    const int type = index_type_from_path(path);
    if (type < 0) throw std::runtime_error(std::string("Wrong index filename supplied: ") + path);
    FILE* f = std::fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("error in opening binary file: ") + path);
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf(sz > 0 ? size_t(sz) : 0);
    const size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
    std::fclose(f);
    if (got != buf.size()) throw std::runtime_error(std::string("short read on ") + path);
    return build_image(buf.data(), buf.size(), type);
}

}  // namespace fgb
