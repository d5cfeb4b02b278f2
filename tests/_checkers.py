"""ctypes bindings for the TEST-ONLY checkers: the plain-C oracle (oracle/libfulgor_oracle.so), the
unmodified reference compiled from /root/reference (oracle/_ref/libfulgor_ref.so, present only where
it was built) and the synthetic read generator (build/libfg_tools.so).

Nothing in the product package imports this module."""
import ctypes as C
import lzma
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data")
# large generated fixtures (tools/make_standin_4546.sh), git-ignored: in the tree when they should travel to the GPU box, else parked outside
BIG_DIRS = [os.path.join(ROOT, "fixtures_big"), os.environ.get("FG_FIXTURES_BIG", "/tmp/fg_fixtures/big")]
ORACLE_SO = os.path.join(ROOT, "oracle", "libfulgor_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libfulgor_ref.so")
REF_DBG_SO = os.path.join(ROOT, "oracle", "_ref", "libfulgor_ref_dbg.so")  # the same shim with the reference's asserts compiled in
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "fulgor_ref")
TOOLS_SO = os.path.join(ROOT, "build", "libfg_tools.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


def _p(a, t):
    return a.ctypes.data_as(t)


def build_checkers():
    """(re)build the oracle .so and the tools .so if missing or stale."""
    src = os.path.join(ROOT, "oracle", "fulgor_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
    tsrc = os.path.join(ROOT, "tools", "readgen.cpp")
    if not os.path.exists(TOOLS_SO) or os.path.getmtime(TOOLS_SO) < os.path.getmtime(tsrc):
        os.makedirs(os.path.dirname(TOOLS_SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", tsrc, "-o", TOOLS_SO])


class _Batch:
    """Shared CSR batch plumbing for both checkers."""

    def _csr(self, fn, reads, *args):
        bases, off = reads
        n = len(off) - 1
        out_off = np.zeros(n + 1, dtype=np.uint64)
        cap = max(1024, 16 * n)
        while True:
            vals = np.zeros(cap, dtype=np.uint32)
            rc = fn(out_off, vals, cap)
            if rc == 0:
                return out_off, vals[: int(out_off[n])].copy()
            if rc != -7:
                raise RuntimeError(f"checker batch call failed rc={rc}")
            cap = int(out_off[n])


    def _kmer_tools(self, conservation_fn, matches_fn, reads, which):
        """(triple_off, triples[n,3]) for which == 'conservation'; (kmer_off, positive uint8 per k-mer, counts[n, C]) for 'matches'"""
        bases, off = reads
        n = len(off) - 1
        bp = bases.ctypes.data_as(C.c_char_p)
        if which == "conservation":
            toff = np.zeros(n + 1, dtype=np.uint64)
            cap = max(1, int(off[n]))
            tr = np.zeros(3 * cap, dtype=np.uint32)
            assert conservation_fn(self.h, bp, _p(off, u64p), n, _p(toff, u64p), _p(tr, u32p), cap) == 0
            return toff, tr[: 3 * int(toff[n])].reshape(-1, 3).copy()
        koff = np.zeros(n + 1, dtype=np.uint64)
        cap = max(1, int(off[n]))
        pos = np.zeros(cap, dtype=np.uint8)
        counts = np.zeros((n, self.num_colors), dtype=np.uint32)
        assert matches_fn(self.h, bp, _p(off, u64p), n, _p(koff, u64p), _p(pos, u8p), cap, _p(counts, u32p)) == 0
        return koff, pos[: int(koff[n])].copy(), counts


class Oracle(_Batch):
    """The plain-C restatement (oracle/fulgor_oracle.c)."""

    def __init__(self, path):
        build_checkers()
        L = C.CDLL(ORACLE_SO)
        L.fo_open.restype = C.c_void_p
        L.fo_open.argtypes = [C.c_char_p]
        L.fo_last_error.restype = C.c_char_p
        L.fo_close.argtypes = [C.c_void_p]
        L.fo_info.argtypes = [C.c_void_p, u64p]
        L.fo_lookup_read.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, u64p]
        L.fo_lookup_kmer.restype = C.c_uint64
        L.fo_lookup_kmer.argtypes = [C.c_void_p, C.c_char_p]
        L.fo_u2c.restype = C.c_uint64
        L.fo_u2c.argtypes = [C.c_void_p, C.c_uint64]
        L.fo_color_set.restype = C.c_int64
        L.fo_color_set.argtypes = [C.c_void_p, C.c_uint64, u32p, C.c_uint64]
        L.fo_batch_fetch_color_set_ids.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_uint32, u64p, u32p, C.c_uint64, u32p]
        L.fo_batch_pseudoalign.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_char_p, u64p, C.c_uint32, u64p, u32p, C.c_uint64]
        L.fo_batch_kmer_conservation.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_uint32, u64p, u32p, C.c_uint64]
        L.fo_batch_kmer_matches.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_uint32, u64p, u8p, C.c_uint64, u32p]
        self.L = L
        self.h = L.fo_open(path.encode())
        if not self.h:
            raise RuntimeError("fo_open: " + L.fo_last_error().decode())
        info = np.zeros(8, dtype=np.uint64)
        L.fo_info(self.h, _p(info, u64p))
        self.k, self.m, self.num_kmers, self.num_unitigs, self.num_colors, self.num_color_sets, self.type = (int(v) for v in info[:7])

    def close(self):
        if self.h:
            self.L.fo_close(self.h)
            self.h = None

    def lookup_read(self, seq: bytes):
        nk = max(0, len(seq) - self.k + 1)
        out = np.full(nk, np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
        if nk:
            self.L.fo_lookup_read(self.h, seq, len(seq), _p(out, u64p))
        return out

    def lookup_kmer(self, kmer: bytes):
        return int(self.L.fo_lookup_kmer(self.h, kmer))

    def u2c(self, u):
        return int(self.L.fo_u2c(self.h, u))

    def color_set(self, i):
        out = np.zeros(self.num_colors, dtype=np.uint32)
        n = self.L.fo_color_set(self.h, i, _p(out, u32p), self.num_colors)
        return out[:n].copy()

    def fetch_color_set_ids(self, reads, want_positive=False):
        bases, off = reads
        n = len(off) - 1
        npos = np.zeros(n, dtype=np.uint32)
        res = self._csr(lambda o, v, cap: self.L.fo_batch_fetch_color_set_ids(self.h, bases.ctypes.data_as(C.c_char_p), _p(off, u64p), n, _p(o, u64p), _p(v, u32p), cap, _p(npos, u32p)), reads)
        return res + (npos,) if want_positive else res

    def pseudoalign(self, reads, algo=0, threshold=1.0):
        bases, off = reads
        n = len(off) - 1
        return self._csr(lambda o, v, cap: self.L.fo_batch_pseudoalign(self.h, algo, threshold, bases.ctypes.data_as(C.c_char_p), _p(off, u64p), n, _p(o, u64p), _p(v, u32p), cap), reads)

    def kmer_conservation(self, reads):
        return self._kmer_tools(self.L.fo_batch_kmer_conservation, self.L.fo_batch_kmer_matches, reads, "conservation")

    def kmer_matches(self, reads):
        return self._kmer_tools(self.L.fo_batch_kmer_conservation, self.L.fo_batch_kmer_matches, reads, "matches")


def reference_available():
    return os.path.exists(REF_SO)


class Reference(_Batch):
    """The unmodified reference (oracle/_ref/libfulgor_ref.so, built by `make -C oracle ref`)."""

    def __init__(self, path, self_checking=False):
        L = C.CDLL(REF_DBG_SO if self_checking else REF_SO)
        L.fref_open.restype = C.c_void_p
        L.fref_open.argtypes = [C.c_char_p]
        L.fref_close.argtypes = [C.c_void_p]
        L.fref_info.argtypes = [C.c_void_p, u64p]
        L.fref_lookup_read.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, u64p]
        L.fref_u2c.restype = C.c_uint64
        L.fref_u2c.argtypes = [C.c_void_p, C.c_uint64]
        L.fref_color_set.restype = C.c_int64
        L.fref_color_set.argtypes = [C.c_void_p, C.c_uint64, u32p, C.c_uint64]
        L.fref_fetch_color_set_ids.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_uint32, u64p, u32p, C.c_uint64, C.c_int]
        L.fref_pseudoalign.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_char_p, u64p, C.c_uint32, u64p, u32p, C.c_uint64, C.c_int]
        L.fref_kmer_conservation.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_uint32, u64p, u32p, C.c_uint64]
        L.fref_kmer_matches.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_uint32, u64p, u8p, C.c_uint64, u32p]
        self.L = L
        self.h = L.fref_open(path.encode())
        if not self.h:
            raise RuntimeError("fref_open failed for " + path)
        info = np.zeros(8, dtype=np.uint64)
        L.fref_info(self.h, _p(info, u64p))
        self.k, self.m, self.num_kmers, self.num_unitigs, self.num_colors, self.num_color_sets, self.type = (int(v) for v in info[:7])

    def close(self):
        if self.h:
            self.L.fref_close(self.h)
            self.h = None

    def lookup_read(self, seq: bytes):
        nk = max(0, len(seq) - self.k + 1)
        out = np.full(nk, np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
        if nk:
            self.L.fref_lookup_read(self.h, seq, len(seq), _p(out, u64p))
        return out

    def u2c(self, u):
        return int(self.L.fref_u2c(self.h, u))

    def color_set(self, i):
        out = np.zeros(self.num_colors, dtype=np.uint32)
        n = self.L.fref_color_set(self.h, i, _p(out, u32p), self.num_colors)
        return out[:n].copy()

    def fetch_color_set_ids(self, reads, threads=1):
        bases, off = reads
        n = len(off) - 1
        return self._csr(lambda o, v, cap: self.L.fref_fetch_color_set_ids(self.h, bases.ctypes.data_as(C.c_char_p), _p(off, u64p), n, _p(o, u64p), _p(v, u32p), cap, threads), reads)

    def pseudoalign(self, reads, algo=0, threshold=1.0, threads=1):
        bases, off = reads
        n = len(off) - 1
        return self._csr(lambda o, v, cap: self.L.fref_pseudoalign(self.h, algo, threshold, bases.ctypes.data_as(C.c_char_p), _p(off, u64p), n, _p(o, u64p), _p(v, u32p), cap, threads), reads)

    def kmer_conservation(self, reads):
        return self._kmer_tools(self.L.fref_kmer_conservation, self.L.fref_kmer_matches, reads, "conservation")

    def kmer_matches(self, reads):
        return self._kmer_tools(self.L.fref_kmer_conservation, self.L.fref_kmer_matches, reads, "matches")


_GPK_CACHE = {}


def load_gpk(name="salmonella_10"):
    """Packed genomes (written by tools/mkdump) as a uint8 array; stored xz-compressed in data/."""
    if name not in _GPK_CACHE:
        path = os.path.join(DATA, name + ".gpk")
        for d in BIG_DIRS:
            if not os.path.exists(path) and os.path.exists(os.path.join(d, name + ".gpk")):
                path = os.path.join(d, name + ".gpk")
        if os.path.exists(path):
            raw = open(path, "rb").read()
        else:
            raw = lzma.open(path + ".xz", "rb").read()
        _GPK_CACHE[name] = np.frombuffer(raw, dtype=np.uint8)
    return _GPK_CACHE[name]


def gen_reads(n, min_len=150, max_len=150, seed=42, first=0, sub_rate=0.01, genomes="salmonella_10", threads=8):
    """Synthetic reads per SURVEY 8(d) via tools/readgen.cpp. Returns (bases uint8[...], read_off uint64[n+1])."""
    build_checkers()
    L = C.CDLL(TOOLS_SO)
    L.fg_readgen.argtypes = [u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, C.c_uint64, u64p, C.c_char_p, C.c_int]
    gpk = load_gpk(genomes)
    off = np.zeros(n + 1, dtype=np.uint64)
    rc = L.fg_readgen(_p(gpk, u8p), gpk.size, first, n, min_len, max_len, sub_rate, seed, _p(off, u64p), None, 1)
    assert rc == 0, "bad .gpk"
    bases = np.zeros(int(off[n]) + 1, dtype=np.uint8)
    L.fg_readgen(_p(gpk, u8p), gpk.size, first, n, min_len, max_len, sub_rate, seed, _p(off, u64p), bases.ctypes.data_as(C.c_char_p), threads)
    return bases[:-1], off


def tile_genomes(name, read_len=400, k=31, every=1):
    """Every contig of the packed genomes `name` (tools/mkdump .gpk) cut into reads of read_len bases that overlap by k-1, so
    that EVERY k-mer of every genome is a k-mer of exactly one read: a walk over the whole dictionary the index was built
    from (each of these k-mers is in the index). `every` keeps one genome in `every`. Returns (bases, read_off)."""
    gpk = load_gpk(name)
    assert gpk[:5].tobytes() == b"FGPK1"
    nc, total = (int(v) for v in gpk[8:24].view(np.uint64))
    rec = np.frombuffer(gpk[24:24 + 12 * nc].tobytes(), dtype=np.dtype([("genome", "<u4"), ("len", "<u8")]))
    packed = gpk[24 + 12 * nc: 24 + 12 * nc + (total + 3) // 4]
    codes = np.empty(packed.size * 4, dtype=np.uint8)
    for j in range(4):
        codes[j::4] = (packed >> (2 * j)) & 3
    ascii_ = np.frombuffer(b"ACGT", dtype=np.uint8)[codes[:total]]
    pieces, lens = [], []
    start = 0
    step = read_len - k + 1
    for c in range(nc):
        n = int(rec["len"][c])
        if int(rec["genome"][c]) % every == 0:
            for a in range(0, max(1, n - k + 1), step):
                b = min(n, a + read_len)
                pieces.append(ascii_[start + a:start + b])
                lens.append(b - a)
        start += n
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(np.array(lens, dtype=np.uint64))
    return np.concatenate(pieces), off


def reads_from_list(seqs):
    """Pack a list of bytes objects into (bases, read_off)."""
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    for i, s in enumerate(seqs):
        off[i + 1] = off[i] + len(s)
    bases = np.frombuffer(b"".join(seqs) + b"\0", dtype=np.uint8)[:-1].copy() if seqs else np.zeros(0, dtype=np.uint8)
    if bases.size == 0:
        bases = np.zeros(1, dtype=np.uint8)[:0]
    return bases, off


def index_path(name):
    """data/<name>; `.mfur` files are stored as a tail relative to the sibling `.fur` (the k-mer
    dictionary and u2c parts are byte-identical, reference include/builders/meta_builder.hpp:356-365)
    and materialised on first use."""
    path = os.path.join(DATA, name)
    if os.path.exists(path):
        return path
    for d in BIG_DIRS:
        if os.path.exists(os.path.join(d, name)):
            return os.path.join(d, name)
    tail = path + ".tail"
    if name.endswith(".mfur") and os.path.exists(tail):
        base = path[: -len(".mfur")] + ".fur"
        raw = open(tail, "rb").read()
        prefix_len = int.from_bytes(raw[:8], "little")
        with open(base, "rb") as f:
            head = f.read(prefix_len)
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(head)
            f.write(raw[8:])
        os.replace(tmp, path)
        return path
    raise FileNotFoundError(path)


def check_dedup(rep, off, vals, expected, cids):
    """a deduplicated full-intersection result (include/fulgor_gpu.h: fulgor_gpu_pseudoalign_dedup) against the per-read
    results `expected` (CSR) and the per-read color-set-id lists `cids` (CSR): every read's result is its representative's
    range, a representative represents itself, reads share a representative iff their (non-empty) lists are equal, and
    only representatives own values."""
    eoff, evals = expected
    coff, cvals = cids
    n = len(eoff) - 1
    assert len(rep) == n and len(off) == n + 1
    groups = {}
    for i in range(n):
        r = int(rep[i])
        assert 0 <= r < n and int(rep[r]) == r, (i, r)
        got = vals[int(off[r]):int(off[r + 1])]
        assert np.array_equal(got, evals[int(eoff[i]):int(eoff[i + 1])]), i
        if r != i:
            assert off[i] == off[i + 1], i
        key = cvals[int(coff[i]):int(coff[i + 1])].tobytes()
        if len(key) == 0:
            assert r == i
        else:
            assert groups.setdefault(key, r) == r, i
    assert len(set(groups.values())) == len(groups)
    return len(groups)


def unpack_positive_words(word_off, words, kmer_off):
    """fulgor_gpu_kmer_matches' per-read 32-bit words -> one byte per k-mer, reads concatenated (the checkers' layout)"""
    n = len(kmer_off) - 1
    out = np.zeros(int(kmer_off[n]), dtype=np.uint8)
    for i in range(n):
        nk = int(kmer_off[i + 1] - kmer_off[i])
        w = words[int(word_off[i]):int(word_off[i + 1])]
        assert w.size == (nk + 31) // 32
        bits = np.unpackbits(w.view(np.uint8), bitorder="little")[:nk]
        assert not np.unpackbits(w.view(np.uint8), bitorder="little")[nk:].any()
        out[int(kmer_off[i]):int(kmer_off[i + 1])] = bits
    return out
