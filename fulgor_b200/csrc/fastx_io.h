/*
 * fastx_io.h -- host side of the `pseudoalign` tool around the GPU call: the FASTA/FASTQ feeder and the three output
 * formatters, both multi-threaded. Plain C++17, no CUDA: pseudoalign_cli.cpp uses it, and tests/fastx_io_test.cpp
 * exposes it to the CPU-only test tier.
 *
 * Replaces, for this tool only:
 *   - the FQFeeder producer/consumer parser the reference drives from tools/pseudoalign.cpp:54-77
 *     (external/FQFeeder/src/FastxParser.cpp:138-260): here an uncompressed query file is memory-mapped and cut into
 *     slabs at record boundaries that worker threads tokenise concurrently, straight into the (pinned) batch buffers the
 *     GPU call reads; gzip input goes through one inflating reader;
 *   - psa_{ascii,binary,compressed}_formatter (src/ps_utils.cpp:48-243): same bytes per record; a batch is split
 *     between threads by output volume and the pieces are written in read order.
 * Read ids are 0-based record positions in the query file.
 */
#ifndef FULGOR_B200_FASTX_IO_H
#define FULGOR_B200_FASTX_IO_H

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace fgio {

/* runs f(0) .. f(n-1) on n threads (f(0) on the caller's) */
inline void parallel_for(unsigned n, const std::function<void(unsigned)>& f) {
    if (n <= 1) {
        if (n == 1) f(0);
        return;
    }
    std::vector<std::thread> th;
    for (unsigned t = 1; t < n; ++t) th.emplace_back(f, t);
    f(0);
    for (auto& t : th) t.join();
}

/* The same with PERSISTENT workers: the tool calls it a dozen times per batch (tokenising rounds, copies, formatting, writing), and
   creating the threads anew each time cost as much as the work of a round (220-290 us per call of five threads on the build VM).
   One team per stage object; run() is called by one thread at a time. */
class thread_team {
public:
    thread_team() = default;
    thread_team(const thread_team&) = delete;
    thread_team& operator=(const thread_team&) = delete;
    ~thread_team() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        work_.notify_all();
        for (auto& t : th_) t.join();
    }
    /* runs f(0) .. f(n-1), f(0) on the caller's thread, and returns when all are done */
    void run(unsigned n, const std::function<void(unsigned)>& f) {
        if (n <= 1) {
            if (n == 1) f(0);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            while (th_.size() + 1 < n) {
                const unsigned id = unsigned(th_.size()) + 1;
                th_.emplace_back([this, id] { worker(id); });
            }
            job_ = &f;
            width_ = n;
            pending_ = n - 1;
            ++generation_;
        }
        work_.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void worker(unsigned id) {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            work_.wait(lk, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            if (id >= width_) continue; /* this call uses fewer workers */
            const std::function<void(unsigned)>* job = job_;
            lk.unlock();
            (*job)(id);
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::mutex m_;
    std::condition_variable work_, done_;
    std::vector<std::thread> th_;
    const std::function<void(unsigned)>* job_ = nullptr;
    unsigned width_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

/* a batch of reads in the layout the C ABI takes: concatenated bases + CSR offsets. The buffers belong to the caller
   (pinned memory in the tool); grow() is called when a batch needs more room and must keep what the batch already holds. */
struct read_batch {
    char* bases = nullptr;
    uint64_t* off = nullptr;
    uint64_t bases_cap = 0, reads_cap = 0;
    uint32_t n = 0;
    std::function<void(read_batch&, uint64_t /*bases*/, uint64_t /*reads*/)> grow;
    /* record names (first word of the header), only filled when the source was opened with want_names: read i is
       names[name_off[i] .. name_off[i+1]) */
    std::vector<char> names;
    std::vector<uint64_t> name_off;
    void reserve(uint64_t nbases, uint64_t nreads) {
        if (nbases > bases_cap || nreads > reads_cap) grow(*this, nbases, nreads);
    }
};

/* ---------------------------------------------------------------- serial reader (gzip or plain; multi-line records) */
struct serial_fastx_reader {
    gzFile f = nullptr;
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    std::string line, pending_header;

    bool open(const char* path) {
        f = gzopen(path, "rb");
        if (!f) return false;
        gzbuffer(f, 1 << 20);
        buf.resize(1 << 22);
        return true;
    }
    ~serial_fastx_reader() {
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
    bool getline(std::string& out) {
        out.clear();
        bool any = false;
        for (;;) {
            if (pos == end && !fill()) return any;
            any = true;
            const char* p = buf.data() + pos;
            const char* nl = static_cast<const char*>(std::memchr(p, '\n', end - pos));
            if (nl) {
                out.append(p, size_t(nl - p));
                pos += size_t(nl - p) + 1;
                if (!out.empty() && out.back() == '\r') out.pop_back();
                return true;
            }
            out.append(p, end - pos);
            pos = end;
        }
    }
    /* appends the next record's sequence to `bases`; false when the file is exhausted. name (optional) receives the record's
       name: the header without its first character, up to the first white space (klibpp, FQFeeder/include/kseq++.hpp:598) */
    bool next(std::vector<char>& bases, std::string* name = nullptr) {
        std::string header;
        if (!pending_header.empty()) {
            header.swap(pending_header);
        } else {
            do {
                if (!getline(header)) return false;
            } while (header.empty());
        }
        if (name) {
            size_t e = 1;
            while (e < header.size() && !std::isspace(static_cast<unsigned char>(header[e]))) ++e;
            name->assign(header, 1, e - 1);
        }
        if (header[0] == '@') { /* FASTQ: sequence line(s), '+', as many quality characters */
            if (!getline(line)) return false;
            bases.insert(bases.end(), line.begin(), line.end());
            size_t seq_len = line.size();
            for (;;) { /* multi-line sequence until the '+' line */
                if (!getline(line)) return true;
                if (!line.empty() && line[0] == '+') break;
                bases.insert(bases.end(), line.begin(), line.end());
                seq_len += line.size();
            }
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

/* ---------------------------------------------------------------- query source */
class fastx_source {
public:
    ~fastx_source() {
        if (map_ && map_ != MAP_FAILED) munmap(const_cast<char*>(map_), size_);
        if (fd_ >= 0) close(fd_);
    }
    /* threads = tokenising threads for memory-mapped input; span = bytes of file per batch */
    /* tokenising threads of the next batch (the tool rebalances its parsing and formatting threads batch by batch) */
    void set_threads(unsigned threads) { threads_ = std::max(1u, threads); }
    bool open(const char* path, unsigned threads, uint64_t span_bytes, uint64_t max_reads_serial, bool want_names = false) {
        want_names_ = want_names;
        if (const char* e = std::getenv("FULGOR_SLAB_KB")) slab_bytes_ = std::max<size_t>(64, std::strtoull(e, nullptr, 10)) << 10;
        threads_ = std::max(1u, threads);
        span_ = std::max<uint64_t>(span_bytes, 1 << 16);
        max_reads_serial_ = std::max<uint64_t>(1, max_reads_serial);
        path_ = path;
        fd_ = ::open(path, O_RDONLY);
        if (fd_ < 0) return false;
        struct stat st;
        unsigned char magic[2] = {0, 0};
        const bool regular = fstat(fd_, &st) == 0 && S_ISREG(st.st_mode);
        if (regular && st.st_size >= 2 && pread(fd_, magic, 2, 0) == 2 && !(magic[0] == 0x1f && magic[1] == 0x8b)) {
            size_ = size_t(st.st_size);
            map_ = static_cast<const char*>(mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0));
            if (map_ != MAP_FAILED) {
                madvise(const_cast<char*>(map_), size_, MADV_SEQUENTIAL);
                pos_ = skip_blank(0);
                if (pos_ < size_ && (map_[pos_] == '@' || map_[pos_] == '>')) {
                    fastq_ = map_[pos_] == '@';
                    mapped_ = true;
                    return true;
                }
                munmap(const_cast<char*>(map_), size_);
            }
            map_ = nullptr;
        }
        return serial_.open(path); /* gzip, pipes, empty or odd files */
    }
    bool mapped() const { return mapped_; }

    /* upper estimate of what the next batch holds, for sizing its buffers up front: the bases cannot exceed the bytes of the
       span (half of them in a FASTQ file), the reads are estimated from the size of the first record ahead. False (and
       zeros) for inputs that are read serially. */
    bool estimate_batch(uint64_t& nbases, uint64_t& nreads) const {
        nbases = nreads = 0;
        if (!mapped_ || pos_ >= size_) return false;
        const uint64_t bytes = std::min<uint64_t>(size_ - pos_, span_ + uint64_t(threads_) * slab_bytes_);
        const size_t r0 = skip_blank(pos_);
        size_t r1 = r0;
        for (int l = 0; l < (fastq_ ? 4 : 2) && r1 < size_; ++l) r1 = next_line(r1);
        const uint64_t rec = std::max<uint64_t>(fastq_ ? 8 : 4, r1 - r0);
        nbases = bytes / (fastq_ ? 2 : 1) + 64;
        nreads = bytes / rec + bytes / rec / 16 + 1024;
        return true;
    }

    /* fills b with the next reads; returns false when the input is exhausted and b is empty */
    bool next_batch(read_batch& b) {
        b.n = 0;
        b.names.clear();
        b.name_off.assign(1, 0);
        if (mapped_) {
            while (b.n == 0 && pos_ < size_) {
                if (!next_mapped(b)) mapped_ = false; /* a record the slab tokeniser does not handle (multi-line FASTQ): serial from here on */
                if (!mapped_) break;
            }
            if (mapped_) return b.n != 0;
            /* the failed round may follow successful rounds of this batch: the serial reader starts the batch over, so what those
               rounds appended (names in particular) must go */
            b.n = 0;
            b.names.clear();
            b.name_off.assign(1, 0);
        }
        return next_serial(b);
    }

private:
    size_t skip_blank(size_t p) const {
        while (p < size_ && (map_[p] == '\n' || map_[p] == '\r')) ++p;
        return p;
    }
    size_t next_line(size_t p) const { /* start of the line after the one containing p; size_ if none */
        const char* nl = static_cast<const char*>(std::memchr(map_ + p, '\n', size_ - p));
        return nl ? size_t(nl - map_) + 1 : size_;
    }
    /* first record start at or after p (p need not be a line start) */
    size_t find_record(size_t p) const {
        if (p >= size_) return size_;
        if (p > 0 && map_[p - 1] != '\n') p = next_line(p);
        while (p < size_) {
            if (!fastq_) {
                if (map_[p] == '>') return p;
            } else if (map_[p] == '@') { /* a header, unless it is a quality line that happens to start with '@':
                                            two lines below a header there is the '+' line */
                const size_t l1 = next_line(p), l2 = l1 < size_ ? next_line(l1) : size_;
                if (l2 < size_ && map_[l2] == '+') return p;
                if (l2 >= size_) return p; /* truncated tail: let the tokeniser decide */
            }
            p = next_line(p);
        }
        return size_;
    }

    struct slab_out {
        std::vector<char> raw;   /* the slab's bytes (pread scales across threads where first-touch faults of one mapping do not) */
        std::vector<char> bases;
        std::vector<uint32_t> lens;
        std::vector<char> names; /* want_names_: the reads' names, concatenated, and their lengths */
        std::vector<uint32_t> name_lens;
        bool ok = true;
    };

    static size_t line_after(const char* d, size_t n, size_t p) { /* start of the line after the one containing p; n if none */
        const char* nl = static_cast<const char*>(std::memchr(d + p, '\n', n - p));
        return nl ? size_t(nl - d) + 1 : n;
    }

    /* tokenises the whole records in d[0, n): n is a record boundary or the end of the file */
    void tokenise(const char* d, size_t n, slab_out& o) const {
        o.bases.clear(); /* the slabs' buffers persist across batches: no reallocation, no fresh pages after the first batch */
        o.lens.clear();
        o.names.clear();
        o.name_lens.clear();
        o.ok = true;
        o.bases.reserve(n / (fastq_ ? 2 : 1) + 64);
        o.lens.reserve(n / (fastq_ ? 64 : 32) + 64);
        size_t p = 0;
        while (p < n) {
            while (p < n && (d[p] == '\n' || d[p] == '\r')) ++p;
            if (p >= n) break;
            const size_t l1 = line_after(d, n, p); /* header */
            const size_t start = o.bases.size();
            if (want_names_) { /* the first word after '@' / '>' (klibpp: up to the first white space) */
                size_t e = p + 1;
                while (e < l1 && !std::isspace(static_cast<unsigned char>(d[e]))) ++e;
                o.names.insert(o.names.end(), d + p + 1, d + e);
                o.name_lens.push_back(uint32_t(e - (p + 1)));
            }
            if (fastq_) {
                if (d[p] != '@') {
                    o.ok = false;
                    return;
                }
                if (l1 >= n) { /* a header without a sequence line ends the input, like the serial reader */
                    if (want_names_) {
                        o.names.resize(o.names.size() - o.name_lens.back());
                        o.name_lens.pop_back();
                    }
                    return;
                }
                const size_t l2 = line_after(d, n, l1);
                size_t s_end = l2 - (l2 > l1 && d[l2 - 1] == '\n' ? 1 : 0);
                if (s_end > l1 && d[s_end - 1] == '\r') --s_end;
                o.bases.insert(o.bases.end(), d + l1, d + s_end);
                if (l2 >= n) { /* no '+' line: the serial reader accepts this too */
                    o.lens.push_back(uint32_t(o.bases.size() - start));
                    return;
                }
                if (d[l2] != '+') { /* multi-line sequence */
                    o.bases.resize(start);
                    o.ok = false;
                    return;
                }
                const size_t l3 = (l2 + 1 < n && d[l2 + 1] == '\n') ? l2 + 2 : line_after(d, n, l2); /* the separator line is almost always a bare '+' */
                /* the quality line normally is exactly as long as the sequence: look for its newline there before scanning */
                const size_t q_guess = l3 + (s_end - l1);
                const size_t l4 = (q_guess < n && d[q_guess] == '\n') ? q_guess + 1 : (l3 < n ? line_after(d, n, l3) : n);
                size_t q_end = l4 - (l4 > l3 && d[l4 - 1] == '\n' ? 1 : 0);
                if (q_end > l3 && d[q_end - 1] == '\r') --q_end;
                if (q_end - l3 < s_end - l1 && l4 < n) { /* quality shorter than the sequence: multi-line record */
                    o.bases.resize(start);
                    o.ok = false;
                    return;
                }
                p = l4;
            } else {
                if (d[p] != '>') {
                    o.ok = false;
                    return;
                }
                size_t q = l1;
                while (q < n && d[q] != '>' && d[q] != '@') {
                    const size_t nq = line_after(d, n, q);
                    size_t e = nq - (nq > q && d[nq - 1] == '\n' ? 1 : 0);
                    if (e > q && d[e - 1] == '\r') --e;
                    o.bases.insert(o.bases.end(), d + q, d + e);
                    q = nq;
                }
                p = q;
            }
            o.lens.push_back(uint32_t(o.bases.size() - start));
        }
    }

    bool load_slab(size_t begin, size_t end, slab_out& o) const {
        o.raw.resize(end - begin);
        size_t got = 0;
        while (got < o.raw.size()) {
            const ssize_t r = pread(fd_, o.raw.data() + got, o.raw.size() - got, off_t(begin + got));
            if (r <= 0) return false;
            got += size_t(r);
        }
        return true;
    }

    /* One batch = up to span_ bytes of the file, produced in ROUNDS of threads_ cache-sized slabs: a slab is read (pread),
       tokenised while still hot in the cache, and its reads appended to the batch buffers. */
    bool next_mapped(read_batch& b) {
        const unsigned T = threads_;
        const size_t batch_end = std::min<uint64_t>(size_, pos_ + span_);
        std::vector<slab_out>& out = slabs_;
        out.resize(T);
        uint64_t nbases = 0, nreads = 0;
        std::vector<size_t> cut(T + 1);
        std::vector<uint64_t> base_at(T + 1), read_at(T + 1);
        { /* size the batch buffers ONCE (they are pinned memory in the tool: growing them round by round would cost a pinned
             allocation and a copy each time); the rounds below still grow the buffers if the estimate was short */
            uint64_t eb = 0, er = 0;
            estimate_batch(eb, er);
            b.reserve(eb, er);
        }
        while (pos_ < batch_end) {
            const size_t begin = pos_, target = std::min<uint64_t>(size_, begin + uint64_t(T) * slab_bytes_);
            cut[0] = begin;
            for (unsigned t = 1; t <= T; ++t) cut[t] = std::max(cut[t - 1], find_record(begin + size_t((target - begin) * uint64_t(t) / T)));
            if (target == size_) cut[T] = size_;
            team_.run(T, [&](unsigned t) {
                if (load_slab(cut[t], cut[t + 1], out[t])) tokenise(out[t].raw.data(), out[t].raw.size(), out[t]);
                else out[t].ok = false;
            });
            for (auto const& o : out)
                if (!o.ok) { /* hand over to the serial reader from the start of this batch */
                    reads_done_ -= nreads;
                    return false;
                }
            base_at[0] = nbases;
            read_at[0] = nreads;
            for (unsigned t = 0; t < T; ++t) {
                base_at[t + 1] = base_at[t] + out[t].bases.size();
                read_at[t + 1] = read_at[t] + out[t].lens.size();
            }
            if (read_at[T] > 0xffffffffull) return false;
            b.reserve(base_at[T] + 1, read_at[T] + 1);
            team_.run(T, [&](unsigned t) {
                if (!out[t].bases.empty()) std::memcpy(b.bases + base_at[t], out[t].bases.data(), out[t].bases.size());
                uint64_t o = base_at[t];
                uint64_t* dst = b.off + read_at[t];
                for (uint32_t len : out[t].lens) {
                    *dst++ = o;
                    o += len;
                }
            });
            if (want_names_)
                for (unsigned t = 0; t < T; ++t) {
                    b.names.insert(b.names.end(), out[t].names.begin(), out[t].names.end());
                    for (uint32_t len : out[t].name_lens) b.name_off.push_back(b.name_off.back() + len);
                }
            reads_done_ += read_at[T] - nreads;
            nbases = base_at[T];
            nreads = read_at[T];
            pos_ = cut[T];
        }
        b.off[nreads] = nbases;
        b.n = uint32_t(nreads);
        return true;
    }

    bool next_serial(read_batch& b) {
        if (!serial_open_) {
            if (map_) { /* fell back from the mapped path: reopen and skip what was consumed */
                if (!serial_.open(path_.c_str())) return false;
                std::vector<char> sink;
                uint64_t skipped = 0;
                while (skipped < reads_done_ && serial_.next(sink)) {
                    sink.clear();
                    ++skipped;
                }
            }
            serial_open_ = true;
        }
        std::vector<char>& bases = serial_bases_;
        bases.clear();
        serial_off_.assign(1, 0);
        std::string name;
        while (serial_off_.size() - 1 < max_reads_serial_ && bases.size() < (1ull << 31)) {
            if (!serial_.next(bases, want_names_ ? &name : nullptr)) break;
            serial_off_.push_back(bases.size());
            if (want_names_) {
                b.names.insert(b.names.end(), name.begin(), name.end());
                b.name_off.push_back(b.names.size());
            }
        }
        const uint64_t n = serial_off_.size() - 1;
        if (n == 0) return false;
        b.reserve(bases.size() + 1, n + 1);
        std::memcpy(b.bases, bases.data(), bases.size());
        std::memcpy(b.off, serial_off_.data(), (n + 1) * 8);
        b.n = uint32_t(n);
        return true;
    }

    std::string path_;
    int fd_ = -1;
    const char* map_ = nullptr;
    size_t size_ = 0, pos_ = 0;
    bool mapped_ = false, fastq_ = true, serial_open_ = false, want_names_ = false;
    unsigned threads_ = 1;
    thread_team team_;
    uint64_t span_ = 1ull << 28, max_reads_serial_ = 1u << 22, reads_done_ = 0;
    size_t slab_bytes_ = 2u << 20;
    serial_fastx_reader serial_;
    std::vector<slab_out> slabs_;
    std::vector<char> serial_bases_;
    std::vector<uint64_t> serial_off_;
};

/* ---------------------------------------------------------------- output formats (src/ps_utils.cpp:48-243) */
/* decimal digits of v, two at a time from a table; color ids and list sizes are mostly below 10,000 */
static const char fg_digit_pairs[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869"
    "707172737475767778798081828384858687888990919293949596979899";
inline char* put_u32(char* p, uint32_t v) {
    if (v < 10) {
        *p = char('0' + v);
        return p + 1;
    }
    if (v < 100) {
        std::memcpy(p, fg_digit_pairs + 2 * v, 2);
        return p + 2;
    }
    if (v < 10000) {
        const uint32_t q = v / 100, r = v % 100;
        if (q < 10) *p++ = char('0' + q);
        else {
            std::memcpy(p, fg_digit_pairs + 2 * q, 2);
            p += 2;
        }
        std::memcpy(p, fg_digit_pairs + 2 * r, 2);
        return p + 2;
    }
    char tmp[10];
    int n = 0;
    while (v >= 100) {
        n += 2;
        std::memcpy(tmp + 10 - n, fg_digit_pairs + 2 * (v % 100), 2);
        v /= 100;
    }
    if (v >= 10) {
        n += 2;
        std::memcpy(tmp + 10 - n, fg_digit_pairs + 2 * v, 2);
    } else {
        n += 1;
        tmp[10 - n] = char('0' + v);
    }
    std::memcpy(p, tmp + 10 - n, size_t(n));
    return p + n;
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

/* ---------------------------------------------------------------- output lines of the per-k-mer tools
   reads [lo, hi) of a batch, appended to `line`:
     kmer-conservation  name \t n [\t (start_pos_in_query num_kmers color_set_id)]* \n      (tools/kmer_conservation.cpp:27-37)
     kmer-matches       name \t num_kmers [\t 0|1]*num_kmers [\t count]*num_colors \n        (tools/kmer_matches.cpp:28-34) */
inline void append_u32(std::string& line, uint64_t v) {
    char num[16];
    line.append(num, size_t(put_u32(num, uint32_t(v)) - num));
}
inline void format_kmer_conservation(const char* names, const uint64_t* name_off, uint32_t lo, uint32_t hi, const uint64_t* triple_off,
                                     const uint32_t* triples, std::string& line) {
    for (uint32_t i = lo; i < hi; ++i) {
        line.append(names + name_off[i], size_t(name_off[i + 1] - name_off[i]));
        line += '\t';
        append_u32(line, triple_off[i + 1] - triple_off[i]);
        for (uint64_t t = triple_off[i]; t < triple_off[i + 1]; ++t) {
            line += "\t(";
            append_u32(line, triples[3 * t]);
            line += ' ';
            append_u32(line, triples[3 * t + 1]);
            line += ' ';
            append_u32(line, triples[3 * t + 2]);
            line += ')';
        }
        line += '\n';
    }
}
inline void format_kmer_matches(const char* names, const uint64_t* name_off, uint32_t lo, uint32_t hi, const uint64_t* read_off, uint32_t k,
                                const uint64_t* word_off, const uint32_t* words, const uint32_t* counts, uint32_t num_colors, std::string& line) {
    for (uint32_t i = lo; i < hi; ++i) {
        const uint64_t len = read_off[i + 1] - read_off[i], nk = len >= k ? len - k + 1 : 0;
        line.append(names + name_off[i], size_t(name_off[i + 1] - name_off[i]));
        line += '\t';
        append_u32(line, nk);
        const uint32_t* w = words + word_off[i];
        for (uint64_t j = 0; j < nk; ++j) {
            line += '\t';
            line += char('0' + ((w[j >> 5] >> (j & 31)) & 1u));
        }
        const uint32_t* c = counts + uint64_t(i) * num_colors;
        for (uint32_t j = 0; j < num_colors; ++j) {
            line += '\t';
            append_u32(line, c[j]);
        }
        line += '\n';
    }
}

enum class out_format { ASCII, BINARY, COMPRESSED };

class result_writer {
public:
    bool open(const char* path, out_format fm, uint32_t num_colors, unsigned threads) {
        fd_ = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
        pos_ = 0;
        failed_ = false;
        seekable_ = fd_ >= 0 && ::lseek(fd_, 0, SEEK_CUR) != off_t(-1); /* a pipe or a terminal takes the pieces in order instead */
        fmt_ = fm;
        num_colors_ = num_colors;
        set_threads(threads);
        if (fd_ < 0) return false;
        if (fmt_ == out_format::COMPRESSED) { /* psa_compressed_formatter::set_num_colors, src/ps_utils.cpp:160-166 */
            const uint64_t header = num_colors;
            put(&header, 8, 0);
            pos_ = 8;
            sparse_thr_ = uint32_t(0.25 * num_colors);
            dense_thr_ = uint32_t(0.75 * num_colors);
        }
        return true;
    }
    /* formatting threads of the next write_batch (the tool rebalances its parsing and formatting threads batch by batch) */
    void set_threads(unsigned threads) {
        threads_ = std::max(1u, threads);
        if (pieces_.size() < threads_) pieces_.resize(threads_);
    }
    bool ok() const { return !failed_; }
    /* records first_id .. first_id+n-1: colors[off[i] .. off[i+1]) each; with `rep` (deduplicated results: rep[i] = the read
       of this batch whose range holds read i's colors) record i carries colors[off[rep[i]] .. off[rep[i]+1]) -- every read id
       is written with its group's result, like the reference's preprocessed_query_reader path (tools/pseudoalign.cpp:39-44) */
    void write_batch(uint32_t first_id, uint32_t n, const uint64_t* off, const uint32_t* colors, const uint32_t* rep = nullptr) {
        if (n == 0) return;
        std::vector<uint64_t> fanned;
        const uint64_t* vol = off; /* vol[i] - vol[0] = values written by records before i */
        if (rep) {
            fanned.resize(uint64_t(n) + 1);
            fanned[0] = 0;
            for (uint32_t i = 0; i < n; ++i) fanned[i + 1] = fanned[i] + (off[rep[i] + 1] - off[rep[i]]);
            vol = fanned.data();
        }
        /* cut the batch where the output volume (records + values) splits evenly */
        const unsigned T = unsigned(std::min<uint64_t>(threads_, std::max<uint64_t>(1, n / 1024)));
        std::vector<uint32_t> cut(T + 1, 0);
        const uint64_t total = (vol[n] - vol[0]) + 2 * uint64_t(n);
        for (unsigned t = 1; t < T; ++t) {
            const uint64_t want = total * t / T;
            uint32_t lo = cut[t - 1], hi = n;
            while (lo < hi) { /* first i with volume(i) >= want */
                const uint32_t mid = lo + (hi - lo) / 2;
                if ((vol[mid] - vol[0]) + 2 * uint64_t(mid) < want) lo = mid + 1; else hi = mid;
            }
            cut[t] = lo;
        }
        cut[T] = n;
        /* every thread formats its records, then writes its piece at its own file position (the sizes of the pieces before it
           are known after the formatting): no serial pass over the formatted bytes */
        std::vector<uint64_t> at(T + 1, pos_);
        if (T == 1) {
            format(first_id, 0, n, off, colors, rep, vol[n] - vol[0], pieces_[0]);
            put(pieces_[0].data(), pieces_[0].size(), pos_);
            at[1] = pos_ + pieces_[0].size();
        } else {
            team_.run(T, [&](unsigned t) { format(first_id, cut[t], cut[t + 1], off, colors, rep, vol[cut[t + 1]] - vol[cut[t]], pieces_[t]); });
            for (unsigned t = 0; t < T; ++t) at[t + 1] = at[t] + pieces_[t].size();
            if (seekable_) team_.run(T, [&](unsigned t) { put(pieces_[t].data(), pieces_[t].size(), at[t]); });
            else
                for (unsigned t = 0; t < T; ++t) put(pieces_[t].data(), pieces_[t].size(), at[t]);
        }
        pos_ = at[T];
    }
    void close() {
        if (fd_ >= 0) ::close(fd_);
        fd_ = -1;
    }

private:
    void format(uint32_t first_id, uint32_t lo, uint32_t hi, const uint64_t* off, const uint32_t* colors, const uint32_t* rep, uint64_t num_values,
                std::vector<char>& out) const {
        out.clear();
        if (lo >= hi) return;
        auto src = [rep](uint32_t i) { return rep ? rep[i] : i; };
        if (fmt_ == out_format::ASCII) { /* "id \t n [\t color]* \n", src/ps_utils.cpp:48-100 */
            out.resize(size_t(hi - lo) * 24 + size_t(num_values) * 11 + 16);
            char* p = out.data();
            for (uint32_t i = lo; i < hi; ++i) {
                const uint64_t b = off[src(i)], e = off[src(i) + 1];
                p = put_u32(p, first_id + i);
                *p++ = '\t';
                p = put_u32(p, uint32_t(e - b));
                for (uint64_t j = b; j < e; ++j) {
                    *p++ = '\t';
                    p = put_u32(p, colors[j]);
                }
                *p++ = '\n';
            }
            out.resize(size_t(p - out.data()));
        } else if (fmt_ == out_format::BINARY) { /* u32 id, u32 n, n x u32, src/ps_utils.cpp:102-147 */
            out.resize((size_t(hi - lo) * 2 + size_t(num_values)) * 4);
            uint32_t* p = reinterpret_cast<uint32_t*>(out.data());
            for (uint32_t i = lo; i < hi; ++i) {
                const uint64_t b = off[src(i)], e = off[src(i) + 1];
                *p++ = first_id + i;
                *p++ = uint32_t(e - b);
                std::memcpy(p, colors + b, size_t(e - b) * 4);
                p += e - b;
            }
        } else { /* blocks of {u64 num_bits, words}: delta(id) delta(n) then the list like a hybrid color set, src/ps_utils.cpp:149-243 */
            bit_writer bw;
            auto flush = [&]() {
                if (!bw.num_bits) return;
                const size_t at = out.size();
                out.resize(at + 8 + bw.words.size() * 8);
                std::memcpy(out.data() + at, &bw.num_bits, 8);
                std::memcpy(out.data() + at + 8, bw.words.data(), bw.words.size() * 8);
                bw.clear();
            };
            for (uint32_t i = lo; i < hi; ++i) {
                const uint32_t* c = colors + off[src(i)];
                const uint32_t size = uint32_t(off[src(i) + 1] - off[src(i)]);
                bw.delta(first_id + i);
                bw.delta(size);
                if (size == 0) {
                } else if (size < sparse_thr_) {
                    bw.delta(c[0]);
                    for (uint32_t j = 1; j < size; ++j) bw.delta(c[j] - (c[j - 1] + 1));
                } else if (size < dense_thr_) {
                    const uint64_t start = bw.num_bits;
                    for (uint32_t w = 0; w < num_colors_; w += 64) bw.append(0, std::min<uint32_t>(64, num_colors_ - w));
                    for (uint32_t j = 0; j < size; ++j) {
                        const uint64_t bit = start + c[j];
                        bw.words[bit >> 6] |= 1ULL << (bit & 63);
                    }
                } else { /* the complement, delta-gap coded */
                    bool first = true;
                    uint32_t prev = 0, j = 0;
                    for (uint32_t v = 0; v < num_colors_; ++v) {
                        if (j < size && c[j] == v) {
                            ++j;
                            continue;
                        }
                        bw.delta(first ? v : v - (prev + 1));
                        first = false;
                        prev = v;
                    }
                }
                if (bw.words.size() * 8 > (1u << 14)) flush(); /* formatter_buffer, src/ps_utils.cpp:31-38 */
            }
            flush();
        }
    }

    void put(const void* data, size_t bytes, uint64_t at) {
        const char* p = static_cast<const char*>(data);
        while (bytes) {
            const ssize_t w = seekable_ ? ::pwrite(fd_, p, bytes, off_t(at)) : ::write(fd_, p, bytes);
            if (w <= 0) {
                failed_ = true;
                return;
            }
            p += w;
            at += uint64_t(w);
            bytes -= size_t(w);
        }
    }

    int fd_ = -1;
    uint64_t pos_ = 0; /* file position of the next batch */
    std::atomic<bool> failed_{false};
    bool seekable_ = true;
    out_format fmt_ = out_format::ASCII;
    uint32_t num_colors_ = 0, sparse_thr_ = 0, dense_thr_ = 0;
    unsigned threads_ = 1;
    thread_team team_;
    std::vector<std::vector<char>> pieces_;
};

}  // namespace fgio
#endif
