/*
 * pack_reads.h -- host side of the PACKED read form (include/fulgor_gpu.h): ASCII reads -> 2-bit codes, 16 per 32-bit word,
 * every read starting on a word; code = (c >> 1) & 3, the reference's own encoding (external/sshash/include/kmer.hpp:199:
 * A 0, C 1, T 2, G 3, either case), so the kernels use the words as they are. Characters other than ACGTacgt -- which make
 * every k-mer over them invalid (kmer.hpp:214-224,258-260) -- keep their (meaningless) code and are listed apart, by position.
 * Plain C++ (no CUDA): used by the library's fulgor_gpu_pack_reads and by host code that produces packed batches itself.
 */
#ifndef FULGOR_B200_PACK_READS_H
#define FULGOR_B200_PACK_READS_H

#include <stdint.h>

#include <algorithm>
#include <thread>
#include <vector>

namespace fgb {

#define FG_PACK_FLAGGED 0x80000000u

struct pack_lut {
    uint8_t v[256]; /* bits 0-1: code; bit 7: not one of ACGTacgt */
    pack_lut() {
        for (int c = 0; c < 256; ++c) {
            const bool ok = c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'a' || c == 'c' || c == 'g' || c == 't';
            v[c] = uint8_t(((c >> 1) & 3) | (ok ? 0 : 0x80));
        }
    }
};

/* packs one read into out (ceil(len / 16) words, unused bits zero); invalid character positions (relative to the read) are
   appended to inv shifted by pos0. Returns true when the read has such a character. */
inline bool pack_one_read(const uint8_t* s, uint32_t len, uint32_t* out, uint64_t pos0, std::vector<uint64_t>& inv) {
    static const pack_lut lut;
    uint32_t bad_any = 0;
    const uint32_t full = len >> 4;
    for (uint32_t w = 0; w < full; ++w) {
        const uint8_t* p = s + 16 * w;
        uint32_t x = 0, bad = 0;
        for (int j = 0; j < 16; ++j) {
            const uint32_t t = lut.v[p[j]];
            x |= (t & 3u) << (2 * j);
            bad |= t;
        }
        out[w] = x;
        if (bad & 0x80u) {
            bad_any = 1;
            for (int j = 0; j < 16; ++j)
                if (lut.v[p[j]] & 0x80u) inv.push_back(pos0 + 16 * w + j);
        }
    }
    const uint32_t rest = len & 15u;
    if (rest) {
        const uint8_t* p = s + 16 * full;
        uint32_t x = 0;
        for (uint32_t j = 0; j < rest; ++j) {
            const uint32_t t = lut.v[p[j]];
            x |= (t & 3u) << (2 * j);
            if (t & 0x80u) {
                bad_any = 1;
                inv.push_back(pos0 + 16 * full + j);
            }
        }
        out[full] = x;
    }
    return bad_any != 0;
}

/* words a batch needs: sum over the reads of ceil(len / 16) */
inline uint64_t packed_words_of(const uint64_t* read_off, uint32_t n) {
    uint64_t w = 0;
    for (uint32_t i = 0; i < n; ++i) w += (read_off[i + 1] - read_off[i] + 15) >> 4;
    return w;
}

/* Packs a batch with up to `threads` threads. words must hold packed_words_of() entries, read_len n entries. The invalid
   positions (ascending, in bases from words[0]: 16 * word index + base in word) are returned in `invalid`. */
inline void pack_reads(const char* bases, const uint64_t* read_off, uint32_t n, uint32_t* words, uint32_t* read_len, std::vector<uint64_t>& invalid,
                       unsigned threads) {
    invalid.clear();
    if (n == 0) return;
    const unsigned T = std::max(1u, std::min<unsigned>(threads, (n + 4095) / 4096));
    std::vector<uint32_t> cut(T + 1);
    for (unsigned t = 0; t <= T; ++t) cut[t] = uint32_t(uint64_t(n) * t / T);
    std::vector<uint64_t> first_word(T + 1, 0);
    std::vector<std::vector<uint64_t>> inv(T);
    auto run = [&](auto&& f) {
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < T; ++t) pool.emplace_back(f, t);
        f(0u);
        for (auto& th : pool) th.join();
    };
    run([&](unsigned t) { first_word[t + 1] = packed_words_of(read_off + cut[t], cut[t + 1] - cut[t]); });
    for (unsigned t = 0; t < T; ++t) first_word[t + 1] += first_word[t];
    run([&](unsigned t) {
        uint64_t w = first_word[t];
        for (uint32_t i = cut[t]; i < cut[t + 1]; ++i) {
            const uint64_t len = read_off[i + 1] - read_off[i];
            const bool flagged = pack_one_read(reinterpret_cast<const uint8_t*>(bases) + read_off[i], uint32_t(len), words + w, 16 * w, inv[t]);
            read_len[i] = uint32_t(len) | (flagged ? FG_PACK_FLAGGED : 0u);
            w += (len + 15) >> 4;
        }
    });
    for (unsigned t = 0; t < T; ++t) invalid.insert(invalid.end(), inv[t].begin(), inv[t].end());
}

}  // namespace fgb
#endif
