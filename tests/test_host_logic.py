"""CPU tier: the index loader / flattener (host C++ inside libfulgor_gpu.so); the CUDA kernels themselves compiled for the
host on a lock-step warp emulator (tests/simt_emul.{h,cpp}: per-lane arithmetic AND the whole K1 -> K2 -> scan -> emit
pipeline with its shuffles, ballots and shared-memory staging), all against the oracle; and the C ABI surface."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import _checkers as ck

ROOT = ck.ROOT
EMUL_SO = os.path.join(ROOT, "build", "libfg_simt_emul.so")
INDEXES = ["salmonella_10.fur", "salmonella_10.mfur", "salmonella_10.dfur", "salmonella_10.mdfur", "synth_200.fur", "synth_200.mfur", "synth_200.dfur", "synth_200.mdfur", "synth_skew.fur"]


def _load_emul(so, flags=()):
    src = os.path.join(ROOT, "tests", "simt_emul.cpp")
    csrc = os.path.join(ROOT, "fulgor_b200", "csrc")
    deps = [src, os.path.join(ROOT, "tests", "simt_emul.h")] + [os.path.join(csrc, f) for f in ("kernels.cuh", "pipeline_kernels.cuh", "image.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", *flags, "-o", so, src])
    E = C.CDLL(so)
    E.emul_lookup_read.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p]
    E.emul_color_set_mask.restype = C.c_uint32
    E.emul_color_set_mask.argtypes = [C.c_void_p, C.c_uint32]
    E.emul_base_valid.argtypes = [C.c_uint32]
    E.emul_pseudoalign.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64,
                                   C.c_uint, C.c_int, C.c_int]
    E.emul_fetch_color_set_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                           C.c_uint, C.c_int]
    E.emul_kmer_tool.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint, C.c_int]
    E.emul_pseudoalign_dedup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_int]
    E.emul_set_packed.argtypes = [C.c_int]
    return E


@pytest.fixture(scope="module")
def emul():
    return _load_emul(EMUL_SO)


@pytest.fixture(scope="module")
def weak_key_emul():
    """the same kernels with THREE hash bits in the 32-bit minimizer keys (kernels.cuh: window_minima32) instead of 24: different
    m-mers then share their key bits in most windows, so the `undecided` detection and exact_window_minimizer carry the result"""
    return _load_emul(os.path.join(ROOT, "build", "libfg_simt_emul_weak.so"), ["-DFG_KEY_HASH_MASK=0xe0000000u"])


def emul_pseudoalign(E, img, reads, algo, thr, num_colors, grid=1, generic=0, table=0):
    bases, off = reads
    n = len(off) - 1
    out_off = np.zeros(n + 1, dtype=np.uint64)
    cap = max(1, n * num_colors)
    vals = np.zeros(cap, dtype=np.uint32)
    rc = E.emul_pseudoalign(img.ctypes.data, algo, thr, bases.ctypes.data, off.ctypes.data, n, out_off.ctypes.data, vals.ctypes.data, cap, grid, generic, table)
    assert rc == 0, rc
    return out_off, vals[: int(out_off[n])]


def emul_fetch(E, img, reads, grid=1, generic=0):
    bases, off = reads
    n = len(off) - 1
    out_off = np.zeros(n + 1, dtype=np.uint64)
    cap = max(1, int(off[n]))
    vals = np.zeros(cap, dtype=np.uint32)
    npos = np.zeros(n, dtype=np.uint32)
    rc = E.emul_fetch_color_set_ids(img.ctypes.data, bases.ctypes.data, off.ctypes.data, n, out_off.ctypes.data, vals.ctypes.data, cap,
                                    npos.ctypes.data, grid, generic)
    assert rc == 0, rc
    return out_off, vals[: int(out_off[n])], npos


@pytest.fixture(scope="module", params=INDEXES)
def loaded(request, built_lib):
    import fulgor_b200 as fg

    path = ck.index_path(request.param)
    img = fg.build_image(path)
    o = ck.Oracle(path)
    o.name = request.param
    yield fg, img, o
    o.close()


def test_image_header_matches_index(loaded):
    fg, img, o = loaded
    i = fg.image_info(img)
    assert (i.k, i.m, i.num_kmers, i.num_unitigs, i.num_colors, i.num_color_sets, i.type) == (
        o.k, o.m, o.num_kmers, o.num_unitigs, o.num_colors, o.num_color_sets, o.type)
    assert i.image_bytes == img.size and i.device == -1
    assert img.size % 256 == 0


def test_image_build_is_deterministic(loaded):
    fg, img, o = loaded
    assert np.array_equal(fg.build_image(ck.index_path(o.name)), img)


def test_every_color_set_decodes_like_the_oracle(loaded, emul):
    fg, img, o = loaded
    if o.num_colors > 32:
        pytest.skip("mask decode is the <= 32 colors path")
    for cid in range(o.num_color_sets):
        mask = emul.emul_color_set_mask(img.ctypes.data, cid)
        exp = 0
        for c in o.color_set(cid):
            exp |= 1 << int(c)
        assert mask == exp, cid


def test_per_kmer_lookup_like_the_oracle(loaded, emul):
    """every k-mer looked up independently (what one GPU lane does) == the reference's streaming answer"""
    fg, img, o = loaded
    bases, off = ck.gen_reads(1500, 75, 300, seed=7, genomes=o.name.split(".")[0])
    for i in range(1500):
        seq = bases[int(off[i]):int(off[i + 1])].tobytes()
        if i % 50 == 0:
            seq = seq[:40] + b"N" + seq[41:]
        if i % 50 == 1:
            seq = seq.lower()
        contigs = o.lookup_read(seq)
        exp = np.array([0xFFFFFFFF if c == 0xFFFFFFFFFFFFFFFF else o.u2c(int(c)) for c in contigs], dtype=np.uint32)
        got = np.full(len(exp), 0xFFFFFFFF, dtype=np.uint32)
        emul.emul_lookup_read(img.ctypes.data, seq, len(seq), got.ctypes.data)
        assert np.array_equal(got, exp), i


def test_base_validity_is_exactly_ACGTacgt(emul):
    """sshash/kmer.hpp:214-224,258-260"""
    for c in range(256):
        assert bool(emul.emul_base_valid(c)) == (chr(c) in "ACGTacgt"), c


def _edge_reads(genomes):
    rng = np.random.default_rng(5)
    g = ck.gen_reads(48, 150, 150, seed=99, genomes=genomes)
    seqs = [g[0][int(g[1][i]):int(g[1][i + 1])].tobytes() for i in range(48)]
    out = [b"", b"A", b"ACGT" * 7 + b"AC", seqs[0][:31], seqs[1][:32], seqs[2].lower(), seqs[3][:75] + b"N" + seqs[3][76:],
           b"N" * 150, seqs[4][:30] + b"n" + seqs[4][31:], b"A" * 200, b"ACGT" * 50, seqs[5] + seqs[6] + seqs[7],
           bytes(rng.choice(list(b"ACGT"), 500).astype(np.uint8)), seqs[8][:149] + b"X", b"-" + seqs[9][1:],
           seqs[10][:158], seqs[11] + seqs[12][:9], seqs[13] + seqs[14][:8], seqs[15][:127 + 30], seqs[16] + seqs[17][:128 + 30 - 150 + 1]]
    return ck.reads_from_list(out + seqs[18:30])


@pytest.mark.parametrize("generic", [0, 1])
def test_emulated_kernels_stage1_like_the_oracle(loaded, emul, generic):
    """k_fetch_color_sets on emulated warps (segment pipeline: sliding-window minimizers, seeds, bucket compare, the
    per-read table with its shared-memory and pool spill) == index::fetch_color_set_ids"""
    fg, img, o = loaded
    n = 400 if o.num_colors <= 32 else 150
    for reads in (ck.gen_reads(n, 75, 300, seed=11, genomes=o.name.split(".")[0]), _edge_reads(o.name.split(".")[0])):
        got = emul_fetch(emul, img, reads, grid=2, generic=generic)
        exp = o.fetch_color_set_ids(reads, want_positive=True)
        for a, b in zip(got, exp):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("table", [0, 1])
@pytest.mark.parametrize("algo,thr", [(0, 1.0), (1, 0.8), (1, 0.25)])
def test_emulated_kernels_pseudoalign_like_the_oracle(loaded, emul, algo, thr, table):
    """the whole kernel pipeline (fused small-color kernel, or K1 + color-set kernel, then scan + emit) on emulated warps ==
    pseudoalign_full_intersection / pseudoalign_threshold_union; table = with the decoded color-set table
    (k_expand_color_sets + k_color_sets_table) or decoding the compressed sets per read (k_color_sets_general)"""
    fg, img, o = loaded
    n = 400 if o.num_colors <= 32 else 120
    for reads in (ck.gen_reads(n, 150, 150, seed=12, genomes=o.name.split(".")[0]), _edge_reads(o.name.split(".")[0])):
        got = emul_pseudoalign(emul, img, reads, algo, thr, o.num_colors, table=table)
        exp = o.pseudoalign(reads, algo, thr)
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])


def test_emulated_threshold_union_wide_counters(emul, built_lib):
    """long reads with REPEATED content: multiplicities of dozens to thousands drive the carry-save counters of the table kernel
    through their 10-, 16- and 32-plane variants (reads of up to 2^10, 2^16 and more k-mers)"""
    import fulgor_b200 as fg

    path = ck.index_path("synth_200.mfur")
    img, o = fg.build_image(path), ck.Oracle(path)
    base = ck.gen_reads(12, 120, 200, seed=33, genomes="synth_200")
    seqs = [base[0][int(base[1][i]):int(base[1][i + 1])].tobytes() for i in range(12)]
    for batch in ([seqs[0] * 3 + seqs[1], seqs[2] * 5],                       # < 2^10 k-mers
                  [seqs[3] * 9 + seqs[4] * 2, (seqs[5] + seqs[6]) * 6, seqs[7]],   # < 2^16
                  [(seqs[8] + seqs[9][:77]) * 300, seqs[10]]):                  # >= 2^16 k-mers: 32 planes
        reads = ck.reads_from_list(batch)
        for thr in (0.9, 0.3):
            got = emul_pseudoalign(emul, img, reads, 1, thr, o.num_colors, table=1)
            exp = o.pseudoalign(reads, 1, thr)
            assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
    o.close()


@pytest.mark.parametrize("table", [0, 1])
def test_emulated_kernels_deduplicate_like_the_reference(loaded, emul, table):
    """K1 -> k_group_reads -> color-set kernel on the representatives only -> scan -> emit on emulated warps: every read's
    result (through its representative) == pseudoalign_full_intersection, and the groups are exactly the distinct
    color-set-id lists (tools/pseudoalign.cpp:92-226). The reads are drawn with repeats so that groups have several members."""
    fg, img, o = loaded
    base = ck.gen_reads(60, 100, 200, seed=21, genomes=o.name.split(".")[0])
    seqs = [base[0][int(base[1][i]):int(base[1][i + 1])].tobytes() for i in range(60)]
    rng = np.random.default_rng(5)
    picks = [seqs[j] for j in rng.integers(0, 60, 150)] + [b"", b"ACGT" * 10, b"N" * 80]
    reads = ck.reads_from_list(picks)
    bases, off = reads
    n = len(off) - 1
    rep = np.zeros(n, dtype=np.uint32)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    cap = max(1, n * o.num_colors)
    vals = np.zeros(cap, dtype=np.uint32)
    rc = emul.emul_pseudoalign_dedup(img.ctypes.data, bases.ctypes.data, off.ctypes.data, n, rep.ctypes.data, out_off.ctypes.data,
                                     vals.ctypes.data, cap, 2, table)
    assert rc == 0
    groups = ck.check_dedup(rep, out_off, vals, o.pseudoalign(reads, 0), o.fetch_color_set_ids(reads))
    assert groups <= 60


@pytest.mark.parametrize("generic", [0, 1])
def test_emulated_kernels_kmer_tools_like_the_oracle(loaded, emul, generic):
    """k_kmer_color_sets (the segment pipeline in per-k-mer mode) + k_kmer_runs / k_kmer_positive_bits / k_kmer_match_counts on
    emulated warps == index::kmer_conservation / index::kmer_matches (src/kmer_conservation.cpp:7-54, src/kmer_matches.cpp:7-30)"""
    fg, img, o = loaded
    genomes = o.name.split(".")[0]
    n = 120 if o.num_colors <= 32 else 50
    for reads in (ck.gen_reads(n, 75, 300, seed=13, genomes=genomes), _edge_reads(genomes)):
        bases, off = reads
        nr = len(off) - 1
        toff = np.zeros(nr + 1, dtype=np.uint64)
        cap = max(1, int(off[nr]))
        tr = np.zeros(3 * cap, dtype=np.uint32)
        assert emul.emul_kmer_tool(img.ctypes.data, 0, bases.ctypes.data, off.ctypes.data, nr, toff.ctypes.data, tr.ctypes.data, cap, None, 2, generic) == 0
        exp_off, exp_tr = o.kmer_conservation(reads)
        assert np.array_equal(toff, exp_off) and np.array_equal(tr[: 3 * int(toff[nr])].reshape(-1, 3), exp_tr)
        if generic:
            continue
        woff = np.zeros(nr + 1, dtype=np.uint64)
        words = np.zeros(cap, dtype=np.uint32)
        counts = np.full((nr, o.num_colors), 0xdeadbeef, dtype=np.uint32)
        assert emul.emul_kmer_tool(img.ctypes.data, 1, bases.ctypes.data, off.ctypes.data, nr, woff.ctypes.data, words.ctypes.data, cap,
                                   counts.ctypes.data, 2, 0) == 0
        koff, pos, exp_counts = o.kmer_matches(reads)
        assert np.array_equal(ck.unpack_positive_words(woff, words, koff), pos)
        assert np.array_equal(counts, exp_counts)


def test_loader_rejects_bad_input(built_lib, tmp_path):
    import fulgor_b200 as fg

    with pytest.raises(fg.FulgorGpuError) as e:
        fg.build_image(str(tmp_path / "missing.fur"))
    assert "opening" in str(e.value)
    with pytest.raises(fg.FulgorGpuError):
        fg.build_image(str(tmp_path / "index.txt"))  # wrong suffix (reference: "Wrong index filename supplied.")
    raw = open(ck.index_path("salmonella_10.fur"), "rb").read()
    trunc = tmp_path / "trunc.fur"
    trunc.write_bytes(raw[: len(raw) // 2])
    with pytest.raises(fg.FulgorGpuError):
        fg.build_image(str(trunc))
    badver = tmp_path / "badver.fur"
    badver.write_bytes(b"\x03" + raw[1:])
    with pytest.raises(fg.FulgorGpuError) as e:
        fg.build_image(str(badver))
    assert "version mismatch" in str(e.value)
    wrong = tmp_path / "wrong.mfur"  # a hybrid file under the meta suffix must not load silently
    wrong.write_bytes(raw)
    with pytest.raises(fg.FulgorGpuError):
        fg.build_image(str(wrong))


def test_c_abi_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "fulgor_gpu.h")).read()
    names = sorted(set(re.findall(r"\b(fulgor_gpu_[a-z_]+)\s*\(", header)))
    assert len(names) >= 15
    L = C.CDLL(built_lib)
    for n in names:
        assert hasattr(L, n), n
    out = subprocess.check_output(["nm", "-D", "--defined-only", built_lib], text=True)
    exported = set(re.findall(r" T (fulgor_gpu_[a-z_]+)", out))
    assert set(names) <= exported


@pytest.mark.parametrize("index", ["synth_200.fur", "synth_200.mfur", "synth_200.dfur", "synth_200.mdfur"])
def test_loader_survives_damaged_files(index, built_lib, tmp_path):
    """truncated and bit-flipped index files either load or are rejected with a message -- never a crash (the loader runs in
    the caller's process); every truncation must be rejected"""
    import fulgor_b200 as fg

    raw = open(ck.index_path(index), "rb").read()
    ext = index.split(".")[1]
    rng = np.random.default_rng(len(raw))
    for i, cut in enumerate((3, 10, 1000, len(raw) // 3, len(raw) // 2, len(raw) - 100, len(raw) - 1)):
        p = tmp_path / f"t{i}.{ext}"
        p.write_bytes(raw[:cut])
        with pytest.raises(fg.FulgorGpuError):
            fg.build_image(str(p))
    loaded = 0
    for i in range(25):
        b = bytearray(raw)
        for _ in range(3):
            b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        p = tmp_path / f"f{i}.{ext}"
        p.write_bytes(bytes(b))
        try:
            fg.build_image(str(p))
            loaded += 1
        except fg.FulgorGpuError as e:
            assert str(e)
    assert loaded <= 25


def test_no_cpu_fallback_without_a_gpu(built_lib):
    """on a box without a GPU the compute entry points must fail loudly (ENODEV), never compute on the CPU"""
    import fulgor_b200 as fg

    if fg.lib().fulgor_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fg.FulgorGpuError) as e:
        fg.Index.open(ck.index_path("salmonella_10.fur"), 0)
    assert e.value.code == fg.index.ENODEV


def test_entry_points_reject_a_null_handle(built_lib):
    """every compute entry point of include/fulgor_gpu.h returns EINVAL on a null handle instead of touching it (no GPU needed)"""
    import fulgor_b200 as fg

    L = fg.lib()
    off = np.zeros(2, dtype=np.uint64)
    out = np.zeros(2, dtype=np.uint64)
    vals = np.zeros(8, dtype=np.uint32)
    EINVAL = -22
    assert L.fulgor_gpu_fetch_color_set_ids(None, b"A", off.ctypes.data, 1, out.ctypes.data, vals.ctypes.data, 8, None) == EINVAL
    assert L.fulgor_gpu_pseudoalign(None, 0, 1.0, b"A", off.ctypes.data, 1, out.ctypes.data, vals.ctypes.data, 8) == EINVAL
    assert L.fulgor_gpu_pseudoalign_dedup(None, b"A", off.ctypes.data, 1, vals.ctypes.data, out.ctypes.data, vals.ctypes.data, 8) == EINVAL
    assert L.fulgor_gpu_kmer_conservation(None, b"A", off.ctypes.data, 1, out.ctypes.data, vals.ctypes.data, 2) == EINVAL
    assert L.fulgor_gpu_kmer_matches(None, b"A", off.ctypes.data, 1, out.ctypes.data, vals.ctypes.data, 8, vals.ctypes.data) == EINVAL
    assert b"null" in L.fulgor_gpu_last_error()
    if L.fulgor_gpu_device_count() == 0:
        assert fg.bind_host_thread(0) == 0  # unknown topology: nothing is narrowed, never an error


def test_product_does_not_reference_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may touch oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fulgor_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "fulgor_oracle" not in text and "oracle/" not in text and "_checkers" not in text, os.path.join(dirpath, f)


def test_emulated_kernels_walk_the_multi_partition_skew_dictionary(emul, built_lib):
    """synth_skew.fur: 8 minimizer-MPHF partitions (minimizer_bucket's load_part branch), 7 skew size classes, one of them
    empty, the last one absorbing buckets beyond 2^max_l and itself a 3-partition MPHF (lookup_in_bucket / phf_lookup). Reads
    tile the genomes the index was built from, so every k-mer is positive; the image header must show what the test is for."""
    import fulgor_b200 as fg
    from fulgor_b200 import imageview as iv

    path = ck.index_path("synth_skew.fur")
    img, o = fg.build_image(path), ck.Oracle(path)
    h = iv.header(img)
    phfs = iv.section(img, h.off_phfs, "<u8", 4 * h.num_phfs).reshape(-1, 4)  # seed, num_partitions, first_part, num_keys
    assert phfs[0][1] >= 2, "the minimizer MPHF must have several partitions"
    assert h.num_skew >= 3 and h.skew_log2_max_bucket > h.skew_max_log2, "skew index with the absorbing last partition"
    assert any(phfs[h.skew_phf[i]][1] == 0 for i in range(h.num_skew)), "an empty skew partition"
    assert phfs[h.skew_phf[h.num_skew - 1]][1] >= 2, "a partitioned skew MPHF"
    sizes = iv.bucket_sizes(img)
    assert sizes.max() > 4096 and len({int(np.ceil(np.log2(s))) for s in sizes[sizes > 64]}) >= 5
    reads = ck.tile_genomes("synth_skew", every=4)
    got = emul_fetch(emul, img, reads, grid=4)
    exp = o.fetch_color_set_ids(reads, want_positive=True)
    for a, b in zip(got, exp):
        assert np.array_equal(a, b)
    assert np.array_equal(got[2], np.maximum(0, np.diff(reads[1].astype(np.int64)) - (o.k - 1)))
    o.close()


@pytest.fixture
def packed_emul(emul):
    """the emulator with its lookup kernels fed PACKED reads (pack_reads.h + packed_reads, the scan kernels for the word offsets)"""
    emul.emul_set_packed.argtypes = [C.c_int]
    emul.emul_set_packed(1)
    yield emul
    emul.emul_set_packed(0)


def _nasty_reads(genomes):
    """edge reads plus reads with invalid characters at word boundaries, first / last positions and in runs"""
    g = ck.gen_reads(24, 150, 150, seed=31, genomes=genomes)
    s = [g[0][int(g[1][i]):int(g[1][i + 1])].tobytes() for i in range(24)]
    out = [s[0][:15] + b"N" + s[0][16:], s[1][:16] + b"n" + s[1][17:], b"N" + s[2][1:], s[3][:149] + b"N", s[4][:31] + b"NNNNNNNNNNNNNNNNNN" + s[4][49:],
           s[5][:47] + b"." + s[5][48:100] + b"R" + s[5][101:], s[6] + s[7][:10] + b"N" + s[7][11:], (s[8] + s[9])[:159] + b"N" + s[10],
           s[11][:16], s[12][:32], s[13][:33], s[14].lower(), b"N" * 33, b"", b"ACGTN", s[15] + s[16] + s[17][:17]]
    e = _edge_reads(genomes)
    seqs = [e[0][int(e[1][i]):int(e[1][i + 1])].tobytes() for i in range(len(e[1]) - 1)]
    return ck.reads_from_list(out + seqs)


@pytest.mark.parametrize("index", ["salmonella_10.fur", "salmonella_10.mdfur", "synth_200.mfur", "synth_skew.fur"])
def test_emulated_kernels_on_packed_reads(index, packed_emul, built_lib):
    """the lookup kernels on PACKED reads (2-bit codes + list of invalid positions) == the oracle on the ASCII reads they were
    packed from: stage 1, full intersection / threshold union, the per-k-mer view; both minimizer-window code paths"""
    import fulgor_b200 as fg

    path = ck.index_path(index)
    img, o = fg.build_image(path), ck.Oracle(path)
    genomes = index.split(".")[0]
    for reads in (ck.gen_reads(200 if o.num_colors <= 32 else 80, 75, 300, seed=17, genomes=genomes), _nasty_reads(genomes)):
        for generic in (0, 1):
            got = emul_fetch(packed_emul, img, reads, grid=2, generic=generic)
            exp = o.fetch_color_set_ids(reads, want_positive=True)
            for a, b in zip(got, exp):
                assert np.array_equal(a, b)
        for algo, thr in ((0, 1.0), (1, 0.6)):
            got = emul_pseudoalign(packed_emul, img, reads, algo, thr, o.num_colors, table=1)
            exp = o.pseudoalign(reads, algo, thr)
            assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
        bases, off = reads
        nr = len(off) - 1
        cap = int(off[nr]) + 1
        toff, tr = np.zeros(nr + 1, dtype=np.uint64), np.zeros(3 * cap, dtype=np.uint32)
        assert packed_emul.emul_kmer_tool(img.ctypes.data, 0, bases.ctypes.data, off.ctypes.data, nr, toff.ctypes.data, tr.ctypes.data, cap, None, 2, 0) == 0
        exp_off, exp_tr = o.kmer_conservation(reads)
        assert np.array_equal(toff, exp_off) and np.array_equal(tr[: 3 * int(toff[nr])].reshape(-1, 3), exp_tr)
    o.close()


@pytest.mark.parametrize("index", ["salmonella_10.fur", "synth_200.mfur", "synth_skew.fur"])
def test_emulated_kernels_with_color_set_ids_in_the_side_array(index, emul, built_lib, monkeypatch):
    """Indexes with more than 2^21 color sets keep the super-k-mer -> color-set id map in a side array (sk_cid) instead of the
    record's 21 bits (image.h). No index of the test tiers is that large: FULGOR_GPU_FORCE_WIDE_CIDS=1 makes the loader take
    that layout (records carry no id at all then), and the kernels must give the oracle's answers through it: stage 1, full
    intersection / threshold union, the per-k-mer view (seed-and-extend items AND the per-k-mer paths of synth_skew.fur)."""
    import fulgor_b200 as fg
    from fulgor_b200 import imageview as iv

    path = ck.index_path(index)
    plain = fg.build_image(path)
    monkeypatch.setenv("FULGOR_GPU_FORCE_WIDE_CIDS", "1")
    img = fg.build_image(path)
    monkeypatch.delenv("FULGOR_GPU_FORCE_WIDE_CIDS")
    h = iv.header(img)
    assert iv.header(plain).off_sk_cid == 0 and h.off_sk_cid != 0
    rec_hi = iv.section(img, h.off_sk_records, "<u4", 2 * h.num_super_kmers)[1::2]
    assert not np.any(rec_hi & ((1 << 21) - 1)), "the records of a wide image carry no color-set id"
    o = ck.Oracle(path)
    genomes = index.split(".")[0]
    for reads in (ck.gen_reads(150 if o.num_colors <= 32 else 60, 75, 300, seed=23, genomes=genomes), _nasty_reads(genomes)):
        got = emul_fetch(emul, img, reads, grid=2)
        exp = o.fetch_color_set_ids(reads, want_positive=True)
        for a, b in zip(got, exp):
            assert np.array_equal(a, b)
        for algo, thr in ((0, 1.0), (1, 0.6)):
            got = emul_pseudoalign(emul, img, reads, algo, thr, o.num_colors, table=1)
            exp = o.pseudoalign(reads, algo, thr)
            assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
        bases, off = reads
        nr = len(off) - 1
        cap = int(off[nr]) + 1
        toff, tr = np.zeros(nr + 1, dtype=np.uint64), np.zeros(3 * cap, dtype=np.uint32)
        assert emul.emul_kmer_tool(img.ctypes.data, 0, bases.ctypes.data, off.ctypes.data, nr, toff.ctypes.data, tr.ctypes.data, cap, None, 2, 0) == 0
        exp_off, exp_tr = o.kmer_conservation(reads)
        assert np.array_equal(toff, exp_off) and np.array_equal(tr[: 3 * int(toff[nr])].reshape(-1, 3), exp_tr)
    o.close()


@pytest.mark.parametrize("index", ["salmonella_10.fur", "synth_skew.fur"])
def test_emulated_kernels_with_weak_minimizer_keys(index, weak_key_emul, built_lib):
    """32-bit window minima: when the key bits cannot tell two m-mers of a window apart the kernel must notice and recompute the
    window under the full 64-bit hash order -- forced here by keys of 3 hash bits; ASCII and packed reads, lists and per-k-mer view"""
    import fulgor_b200 as fg

    path = ck.index_path(index)
    img, o = fg.build_image(path), ck.Oracle(path)
    genomes = index.split(".")[0]
    E = weak_key_emul
    try:
        for packed in (0, 1):
            E.emul_set_packed(packed)
            for reads in (ck.gen_reads(250, 75, 300, seed=23, genomes=genomes), _nasty_reads(genomes)):
                got = emul_fetch(E, img, reads, grid=2, generic=0)
                exp = o.fetch_color_set_ids(reads, want_positive=True)
                for a, b in zip(got, exp):
                    assert np.array_equal(a, b)
                bases, off = reads
                nr = len(off) - 1
                cap = int(off[nr]) + 1
                toff, tr = np.zeros(nr + 1, dtype=np.uint64), np.zeros(3 * cap, dtype=np.uint32)
                assert E.emul_kmer_tool(img.ctypes.data, 0, bases.ctypes.data, off.ctypes.data, nr, toff.ctypes.data, tr.ctypes.data, cap, None, 2, 0) == 0
                exp_off, exp_tr = o.kmer_conservation(reads)
                assert np.array_equal(toff, exp_off) and np.array_equal(tr[: 3 * int(toff[nr])].reshape(-1, 3), exp_tr)
    finally:
        E.emul_set_packed(0)
        o.close()


def test_pack_reads_host_function(built_lib):
    """fulgor_gpu_pack_reads: 2-bit codes (c >> 1) & 3 at 16 bases per word, every read on a word boundary, invalid characters
    listed by position and flagged in read_len, E2BIG with the required capacities"""
    import fulgor_b200 as fg

    seqs = [b"ACGTacgt", b"", b"N", b"ACGTACGTACGTACGTA", b"TTTTTTTTTTTTTTTTGGGGGGGGGGGGGGGG", b"ACNNGT" * 7, b"A" * 16 + b"x"]
    reads = ck.reads_from_list(seqs)
    words, lens, inv = fg.pack_reads(reads, threads=3)
    exp_words, exp_inv, w = [], [], 0
    for s in seqs:
        nw = (len(s) + 15) // 16
        for j in range(nw):
            x = 0
            for t, c in enumerate(s[16 * j:16 * j + 16]):
                x |= ((c >> 1) & 3) << (2 * t)
            exp_words.append(x)
        exp_inv += [16 * w + p for p, c in enumerate(s) if chr(c) not in "ACGTacgt"]
        w += nw
    assert words.tolist() == exp_words and inv.tolist() == exp_inv
    assert [int(x) & 0x7fffffff for x in lens] == [len(s) for s in seqs]
    assert [bool(int(x) >> 31) for x in lens] == [any(chr(c) not in "ACGTacgt" for c in s) for s in seqs]
    big = ck.gen_reads(20000, 75, 300, seed=5)
    w1, l1, i1 = fg.pack_reads(big, threads=1)
    w8, l8, i8 = fg.pack_reads(big, threads=8)
    assert np.array_equal(w1, w8) and np.array_equal(l1, l8) and np.array_equal(i1, i8)
    L = fg.lib()
    bases, off = reads
    nw, ni = C.c_uint64(0), C.c_uint64(0)
    lens2 = np.zeros(len(seqs), dtype=np.uint32)
    rc = L.fulgor_gpu_pack_reads(bases.ctypes.data, off.ctypes.data, len(seqs), None, 0, lens2.ctypes.data, None, 0, C.byref(nw), C.byref(ni), 1)
    assert rc == fg.index.E2BIG and nw.value == len(exp_words)
