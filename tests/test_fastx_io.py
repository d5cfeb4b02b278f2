"""CPU tier: the tool's multi-threaded feeder (memory-mapped slab tokeniser + serial/gzip reader) and its three output
formatters (fulgor_b200/csrc/fastx_io.h) -- the data formats either side of the GPU path. The slab tokeniser must give
exactly what a straightforward parser gives, for any thread count and batch span; the formatters must produce the
reference's record layouts (src/ps_utils.cpp:48-243) regardless of how a batch is split between threads."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

import _checkers as ck
from test_cli import ascii_records, binary_records, compressed_records

SO = os.path.join(ck.ROOT, "build", "libfg_fastx_io_test.so")


@pytest.fixture(scope="module")
def fx():
    src = os.path.join(ck.ROOT, "tests", "fastx_io_test.cpp")
    hdr = os.path.join(ck.ROOT, "fulgor_b200", "csrc", "fastx_io.h")
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in (src, hdr)):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-o", SO, src, "-lz"])
    L = C.CDLL(SO)
    L.fxio_parse.restype = C.c_longlong
    L.fxio_parse.argtypes = [C.c_char_p, C.c_uint, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    L.fxio_free.argtypes = [C.c_void_p]
    L.fxio_format.argtypes = [C.c_char_p, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint]
    L.fxio_parse_names.argtypes = [C.c_char_p, C.c_uint, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_void_p)]
    L.fxio_parse_names.restype = C.c_longlong
    L.fxio_format_kmer_tool.argtypes = [C.c_char_p, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_uint]
    L.fxio_format_dedup.argtypes = [C.c_char_p, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def parse(fx, path, threads=4, span=1 << 16, max_serial=1000):
    b, o, m = C.c_void_p(), C.c_void_p(), C.c_int(0)
    n = fx.fxio_parse(str(path).encode(), threads, span, max_serial, C.byref(b), C.byref(o), C.byref(m))
    assert n >= 0
    off = np.ctypeslib.as_array(C.cast(o, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
    bases = C.string_at(b, int(off[n]))
    fx.fxio_free(b)
    fx.fxio_free(o)
    return [bases[int(off[i]):int(off[i + 1])] for i in range(n)], bool(m.value)


def random_seqs(n, rng, lo=0, hi=400):
    return [bytes(rng.choice(list(b"ACGTNacgt"), int(rng.integers(lo, hi))).astype(np.uint8)) for _ in range(n)]


@pytest.mark.parametrize("threads,span", [(1, 1 << 16), (3, 1 << 16), (8, 1 << 17), (5, 1 << 30)])
def test_fastq_slabs_equal_a_plain_parse(fx, tmp_path, threads, span):
    rng = np.random.default_rng(threads)
    seqs = random_seqs(5000, rng, 1, 400)
    p = tmp_path / "r.fq"
    with open(p, "wb") as f:
        for i, s in enumerate(seqs):
            qual = bytes(rng.choice(list(b"@+I#>"), len(s)).astype(np.uint8))  # quality lines that start with '@', '+' or '>'
            f.write(b"@r%d extra\n%s\n+\n%s\n" % (i, s, qual))
    got, mapped = parse(fx, p, threads, span)
    assert mapped and got == seqs


def test_fastq_variants(fx, tmp_path):
    rng = np.random.default_rng(1)
    seqs = random_seqs(300, rng, 1, 200)
    # CRLF, repeated header after '+', blank lines between records, no trailing newline
    p = tmp_path / "crlf.fq"
    body = b"".join(b"@r%d\r\n%s\r\n+r%d\r\n%s\r\n\r\n" % (i, s, i, b"I" * len(s)) for i, s in enumerate(seqs))
    p.write_bytes(body[:-4])
    got, mapped = parse(fx, p, 4, 1 << 16)
    assert mapped and got == seqs
    # empty reads are records too
    p = tmp_path / "empty.fq"
    seqs2 = [b"ACGT", b"", b"GGA", b""]
    p.write_bytes(b"".join(b"@x\n%s\n+\n%s\n" % (s, b"I" * len(s)) for s in seqs2))
    got, _ = parse(fx, p, 2, 1 << 16)
    assert got == seqs2
    # multi-line FASTQ: the slab tokeniser hands over to the serial reader, ids stay consistent
    p = tmp_path / "multi.fq"
    with open(p, "wb") as f:
        for i, s in enumerate(seqs):
            h = len(s) // 2
            if i >= 200:
                f.write(b"@r%d\n%s\n%s\n+\n%s\n%s\n" % (i, s[:h], s[h:], b"I" * h, b"I" * (len(s) - h)))
            else:
                f.write(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)))
    got, mapped = parse(fx, p, 4, 1 << 14)
    assert not mapped and got == seqs
    # gzip goes through the inflating reader
    p = tmp_path / "r.fq.gz"
    with gzip.open(p, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)))
    got, mapped = parse(fx, p, 4, 1 << 16, max_serial=77)
    assert not mapped and got == seqs


@pytest.mark.parametrize("threads,span", [(1, 1 << 16), (4, 1 << 16), (7, 1 << 30)])
def test_fasta_multiline(fx, tmp_path, threads, span):
    rng = np.random.default_rng(7)
    seqs = random_seqs(2000, rng, 0, 700)
    p = tmp_path / "r.fa"
    with open(p, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">r%d\n" % i)
            for j in range(0, len(s), 60):
                f.write(s[j:j + 60] + b"\n")
    got, mapped = parse(fx, p, threads, span)
    assert mapped and got == seqs


@pytest.mark.parametrize("threads,span", [(1, 1 << 16), (4, 1 << 16), (7, 1 << 30)])
def test_record_names(fx, tmp_path, threads, span):
    """want_names: the first word of every header (klibpp's rule, FQFeeder/include/kseq++.hpp:598), in record order, from the slab
    tokeniser (plain FASTQ / FASTA) and from the serial reader (gzip, multi-line FASTQ)"""
    import gzip

    rng = np.random.default_rng(threads)
    seqs = random_seqs(3000, rng, 1, 300)
    names = [b"read_%d/%d" % (i, i % 3) for i in range(len(seqs))]

    def check(path, mapped_expected=None):
        out = C.c_void_p()
        n = fx.fxio_parse_names(str(path).encode(), threads, span, 700, C.byref(out))
        got = C.string_at(out).split(b"\n")[:-1]
        fx.fxio_free(out)
        assert n == len(names) and got == names

    p = tmp_path / "n.fq"
    with open(p, "wb") as f:
        for nm, s in zip(names, seqs):
            f.write(b"@%s some comment\tx\n%s\n+\n%s\n" % (nm, s, b"I" * len(s)))
    check(p)
    p = tmp_path / "n.fa"
    with open(p, "wb") as f:
        for nm, s in zip(names, seqs):
            f.write(b">%s\n" % nm + b"".join(s[j:j + 60] + b"\n" for j in range(0, len(s), 60)))
    check(p)
    p = tmp_path / "n.fq.gz"
    with gzip.open(p, "wb") as f:
        for nm, s in zip(names, seqs):
            f.write(b"@%s\tc\r\n%s\r\n+\r\n%s\r\n" % (nm, s, b"I" * len(s)))
    check(p)
    p = tmp_path / "multi.fq"  # multi-line FASTQ: the slab tokeniser hands over to the serial reader
    with open(p, "wb") as f:
        for nm, s in zip(names, seqs):
            h = len(s) // 2
            f.write(b"@%s\n%s\n%s\n+\n%s\n%s\n" % (nm, s[:h], s[h:], b"I" * h, b"I" * (len(s) - h)))
    check(p)


def test_empty_and_missing_input(fx, tmp_path):
    p = tmp_path / "empty.fq"
    p.write_bytes(b"")
    got, _ = parse(fx, p)
    assert got == []
    b, o, m = C.c_void_p(), C.c_void_p(), C.c_int(0)
    assert fx.fxio_parse(str(tmp_path / "nope.fq").encode(), 1, 1 << 16, 10, C.byref(b), C.byref(o), C.byref(m)) == -1


@pytest.mark.parametrize("num_colors", [10, 200])
@pytest.mark.parametrize("threads,pieces", [(1, 1), (4, 1), (6, 3)])
def test_formatters(fx, tmp_path, num_colors, threads, pieces):
    """ascii / binary / compressed records (src/ps_utils.cpp:48-243) for lists of every hybrid density class"""
    rng = np.random.default_rng(num_colors + threads)
    n = 20000
    sizes = rng.choice([0, 1, 2, num_colors // 5, num_colors // 2, num_colors - 1, num_colors], n)
    lists = [np.sort(rng.choice(num_colors, int(s), replace=False)).astype(np.uint32) for s in sizes]
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(x) for x in lists])
    colors = np.concatenate(lists + [np.zeros(1, dtype=np.uint32)])
    exp = {i: lists[i].tolist() for i in range(n)}
    outs = {}
    for fmt, name in enumerate(("ascii", "binary", "compressed")):
        path = str(tmp_path / name)
        assert fx.fxio_format(path.encode(), fmt, num_colors, threads, n, off.ctypes.data, colors.ctypes.data, pieces) == 0
        outs[name] = path
    assert ascii_records(outs["ascii"]) == exp
    want = b"".join(b"\t".join([b"%d" % i, b"%d" % len(exp[i])] + [b"%d" % c for c in exp[i]]) + b"\n" for i in range(n))
    assert open(outs["ascii"], "rb").read() == want
    assert binary_records(outs["binary"]) == exp
    if num_colors == 10:  # the pure-Python bit reader is slow on wide bitmaps
        assert compressed_records(outs["compressed"]) == (num_colors, exp)
    else:
        C_, recs = compressed_records_prefix(outs["compressed"], 1500)
        assert C_ == num_colors and all(recs[i] == exp[i] for i in recs)


@pytest.mark.parametrize("threads", [1, 5])
def test_formatters_fan_out_deduplicated_results(fx, tmp_path, threads):
    """write_batch with a representative map (--deduplicate): every read id is written with its group's list, like the
    reference's preprocessed_query_reader path (tools/pseudoalign.cpp:39-44)"""
    rng = np.random.default_rng(threads)
    n, num_colors = 30000, 10
    rep = np.arange(n, dtype=np.uint32)
    members = rng.random(n) < 0.7
    members[:10] = False
    rep[members] = rng.choice(np.flatnonzero(~members), int(members.sum())).astype(np.uint32)
    lists = [np.sort(rng.choice(num_colors, int(rng.integers(0, num_colors + 1)), replace=False)).astype(np.uint32) if not members[i]
             else np.zeros(0, dtype=np.uint32) for i in range(n)]
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(x) for x in lists])
    colors = np.concatenate(lists + [np.zeros(1, dtype=np.uint32)])
    exp = {7 + i: lists[int(rep[i])].tolist() for i in range(n)}
    for fmt, (name, parse_fn) in enumerate((("ascii", ascii_records), ("binary", binary_records), ("compressed", compressed_records))):
        path = str(tmp_path / name)
        assert fx.fxio_format_dedup(path.encode(), fmt, num_colors, threads, n, off.ctypes.data, colors.ctypes.data, rep.ctypes.data) == 0
        got = parse_fn(path)
        assert (got[1] if fmt == 2 else got) == exp


@pytest.mark.parametrize("pieces", [1, 4])
def test_kmer_tool_lines(fx, tmp_path, pieces):
    """the output lines of kmer-conservation / kmer-matches (tools/kmer_conservation.cpp:27-37, tools/kmer_matches.cpp:28-34) from
    the C ABI's batch results, whatever the split into formatting ranges"""
    rng = np.random.default_rng(pieces)
    n, k, C_ = 700, 31, 7
    lens = rng.integers(0, 200, n)
    read_off = np.zeros(n + 1, dtype=np.uint64)
    read_off[1:] = np.cumsum(lens)
    nk = np.maximum(0, lens - k + 1)
    # kmer-conservation
    ntr = rng.integers(0, 5, n)
    toff = np.zeros(n + 1, dtype=np.uint64)
    toff[1:] = np.cumsum(ntr)
    tr = rng.integers(0, 100000, (int(toff[n]), 3)).astype(np.uint32)
    path = str(tmp_path / "cons.txt")
    assert fx.fxio_format_kmer_tool(path.encode(), 0, n, 5, toff.ctypes.data, tr.ctypes.data, None, k, None, 0, pieces) == 0
    want = ["\t".join(["r%d" % (5 + i), str(int(ntr[i]))] + ["(%d %d %d)" % tuple(t) for t in tr[int(toff[i]):int(toff[i + 1])]]) for i in range(n)]
    assert open(path).read().split("\n")[:-1] == want
    # kmer-matches
    woff = np.zeros(n + 1, dtype=np.uint64)
    woff[1:] = np.cumsum((nk + 31) // 32)
    bits = [rng.integers(0, 2, int(x)).astype(np.uint8) for x in nk]
    words = np.zeros(int(woff[n]) + 1, dtype=np.uint32)
    for i in range(n):
        for j, b in enumerate(bits[i]):
            if b:
                words[int(woff[i]) + j // 32] |= np.uint32(1 << (j % 32))
    counts = rng.integers(0, 300, (n, C_)).astype(np.uint32)
    path = str(tmp_path / "match.txt")
    assert fx.fxio_format_kmer_tool(path.encode(), 1, n, 0, woff.ctypes.data, words.ctypes.data, read_off.ctypes.data, k, counts.ctypes.data, C_, pieces) == 0
    want = ["\t".join(["r%d" % i, str(int(nk[i]))] + [str(int(b)) for b in bits[i]] + [str(int(c)) for c in counts[i]]) for i in range(n)]
    assert open(path).read().split("\n")[:-1] == want


def compressed_records_prefix(path, limit):
    """first `limit` records of every block"""
    from test_cli import BitReader

    raw = open(path, "rb").read()
    Cn = int.from_bytes(raw[:8], "little")
    sparse, dense = int(0.25 * Cn), int(0.75 * Cn)
    recs, p = {}, 8
    while p < len(raw) and len(recs) < limit:
        nbits = int.from_bytes(raw[p:p + 8], "little")
        nwords = (nbits + 63) // 64
        words = np.frombuffer(raw, dtype="<u8", count=nwords, offset=p + 8)
        p += 8 + 8 * nwords
        r = BitReader(words, nbits)
        while r.p < nbits and len(recs) < limit:
            rid, size = r.delta(), r.delta()
            if size == 0:
                vals = []
            elif size < sparse:
                vals = [r.delta()]
                for _ in range(size - 1):
                    vals.append(vals[-1] + r.delta() + 1)
            elif size < dense:
                bits = r.take(Cn)
                vals = [c for c in range(Cn) if (bits >> c) & 1]
            else:
                missing = []
                for i in range(Cn - size):
                    missing.append(r.delta() if i == 0 else missing[-1] + r.delta() + 1)
                ms = set(missing)
                vals = [c for c in range(Cn) if c not in ms]
            recs[rid] = vals
    return Cn, recs


def test_thread_team_runs_every_index_once(fx):
    """fgio::thread_team (the persistent workers of the parsing and formatting stages): 20,000 calls of random width 0..9 on one
    team -- each index of a call runs exactly once, and run() returns only after all of them did"""
    fx.fxio_team_selftest.argtypes = [C.c_uint, C.c_uint]
    assert fx.fxio_team_selftest(20000, 9) == 0

