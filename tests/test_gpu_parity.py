"""GPU parity tests proper: the CUDA path, called through the C ABI (fulgor_b200.Index -> libfulgor_gpu.so),
against the oracle on the same seeded inputs. Bit-exact: same per-read lists, same order."""
import numpy as np
import pytest

import _checkers as ck

pytestmark = pytest.mark.gpu

INDEXES = ["salmonella_10.fur", "salmonella_10.mfur", "salmonella_10.dfur", "salmonella_10.mdfur", "synth_200.fur", "synth_200.mfur", "synth_200.dfur", "synth_200.mdfur", "synth_skew.fur"]  # 10 colors (fused kernel) and 200 colors (general kernels)


@pytest.fixture(scope="module", params=INDEXES)
def pair(request, built_lib):
    import fulgor_b200 as fg

    path = ck.index_path(request.param)
    gpu = fg.Index.open(path, 0)
    oracle = ck.Oracle(path)
    gpu.genomes = oracle.genomes = request.param.split(".")[0]
    yield gpu, oracle
    gpu.close()
    oracle.close()


def _same(a, b):
    return a[0].shape == b[0].shape and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def _sized(gpu, n):
    """reads for one oracle comparison: the single-threaded oracle is ~20x slower on the 200-color indexes (long color-set
    lists per read), so those get 40 % of the reads -- the whole GPU tier stays within a few minutes"""
    return n if gpu.num_colors <= 32 else (2 * n) // 5


def edge_reads(genomes, k=31):
    rng = np.random.default_rng(5)
    g = ck.gen_reads(64, 150, 150, seed=99, genomes=genomes)
    seqs = [g[0][int(g[1][i]):int(g[1][i + 1])].tobytes() for i in range(64)]
    out = [b"", b"A", b"ACGT" * 7 + b"AC", seqs[0][:31], seqs[1][:32], seqs[2].lower(), seqs[3][:75] + b"N" + seqs[3][76:],
           b"N" * 150, seqs[4][:30] + b"n" + seqs[4][31:], b"A" * 200, b"ACGT" * 50, seqs[5] + seqs[6] + seqs[7],
           bytes(rng.choice(list(b"ACGT"), 500).astype(np.uint8)), seqs[8][:149] + b"X", b"-" + seqs[9][1:]]
    return ck.reads_from_list(out + seqs[10:40])


def test_info(pair):
    gpu, o = pair
    assert (gpu.k, gpu.m, gpu.num_kmers, gpu.num_unitigs, gpu.num_colors, gpu.num_color_sets, gpu.type) == (
        o.k, o.m, o.num_kmers, o.num_unitigs, o.num_colors, o.num_color_sets, o.type)


@pytest.mark.parametrize("lens", [(150, 150), (75, 300)])
def test_fetch_color_set_ids(pair, lens):
    gpu, o = pair
    reads = ck.gen_reads(_sized(gpu, 20000), lens[0], lens[1], seed=42, genomes=gpu.genomes)
    got = gpu.fetch_color_set_ids(reads, want_positive=True)
    exp = o.fetch_color_set_ids(reads, want_positive=True)
    assert _same(got, exp)
    assert np.array_equal(got[2], exp[2])


@pytest.mark.parametrize("algo,thr", [(0, 1.0), (1, 0.8), (1, 1.0), (1, 0.05), (1, 0.5)])
@pytest.mark.parametrize("lens", [(150, 150), (75, 300)])
def test_pseudoalign(pair, algo, thr, lens):
    gpu, o = pair
    reads = ck.gen_reads(_sized(gpu, 20000), lens[0], lens[1], seed=1234, genomes=gpu.genomes)
    assert _same(gpu.pseudoalign(reads, algo, thr), o.pseudoalign(reads, algo, thr))


@pytest.mark.parametrize("algo,thr", [(0, 1.0), (1, 0.8), (1, 0.001)])
def test_edge_cases(pair, algo, thr):
    gpu, o = pair
    reads = edge_reads(gpu.genomes)
    assert _same(gpu.pseudoalign(reads, algo, thr), o.pseudoalign(reads, algo, thr))
    assert _same(gpu.fetch_color_set_ids(reads), o.fetch_color_set_ids(reads))


@pytest.mark.parametrize("index", ["salmonella_10.fur", "synth_200.mfur", "synth_skew.fur"])
def test_color_set_ids_in_the_side_array(index, built_lib, monkeypatch):
    """the layout of indexes with more than 2^21 color sets (ids in the sk_cid side array, none in the records), forced on small
    indexes by FULGOR_GPU_FORCE_WIDE_CIDS=1: every kernel that maps a super-k-mer to its color set must read the side array"""
    import fulgor_b200 as fg

    path = ck.index_path(index)
    monkeypatch.setenv("FULGOR_GPU_FORCE_WIDE_CIDS", "1")
    gpu = fg.Index.open(path, 0)
    monkeypatch.delenv("FULGOR_GPU_FORCE_WIDE_CIDS")
    o = ck.Oracle(path)
    genomes = index.split(".")[0]
    try:
        for reads in (ck.gen_reads(_sized(gpu, 6000), 75, 300, seed=77, genomes=genomes), edge_reads(genomes)):
            got = gpu.fetch_color_set_ids(reads, want_positive=True)
            exp = o.fetch_color_set_ids(reads, want_positive=True)
            assert _same(got, exp) and np.array_equal(got[2], exp[2])
            for algo, thr in ((0, 1.0), (1, 0.8)):
                assert _same(gpu.pseudoalign(reads, algo, thr), o.pseudoalign(reads, algo, thr))
            got_off, got_tr = gpu.kmer_conservation(reads)
            exp_off, exp_tr = o.kmer_conservation(reads)
            assert np.array_equal(got_off, exp_off) and np.array_equal(got_tr, exp_tr)
    finally:
        gpu.close()
        o.close()


def test_empty_batch(pair):
    gpu, o = pair
    reads = ck.reads_from_list([])
    off, vals = gpu.pseudoalign(reads, 0)
    assert off.tolist() == [0] and vals.size == 0


def test_e2big_reports_required_capacity(pair):
    import fulgor_b200 as fg

    gpu, o = pair
    reads = ck.gen_reads(5000, seed=3, genomes=gpu.genomes)
    exp = o.pseudoalign(reads, 0)
    bases, off = reads
    color_off = np.zeros(len(off), dtype=np.uint64)
    colors = np.zeros(8, dtype=np.uint32)
    rc = gpu.pseudoalign_raw(0, 1.0, bases.ctypes.data, off.ctypes.data, len(off) - 1, color_off.ctypes.data, colors.ctypes.data, 8)
    assert rc == fg.index.E2BIG
    assert int(color_off[-1]) == exp[1].size
    assert np.array_equal(color_off, exp[0])


def test_multi_chunk_large_batch(pair):
    """many pipeline chunks (2^18 reads each, more chunks than slots): exercises the cross-chunk CSR carry and the slot reuse; checked against the oracle on
    a sample and through size-independent properties on the whole batch"""
    gpu, o = pair
    n = (1 << 20) + 70000 if gpu.num_colors <= 32 else 300000
    reads = ck.gen_reads(n, seed=77, genomes=gpu.genomes)
    off, vals = gpu.pseudoalign(reads, 0)
    assert off.size == n + 1 and off[0] == 0 and int(off[-1]) == vals.size
    assert np.all(np.diff(off.astype(np.int64)) >= 0)
    assert vals.max() < gpu.num_colors
    # ascending inside each read: a drop may only happen at a read boundary
    drops = np.nonzero(np.diff(vals.astype(np.int64)) <= 0)[0] + 1
    assert np.all(np.isin(drops, off))
    # idempotence + the batch split does not matter: the tail, recomputed alone, equals the slice of the whole
    lo = n - 20000
    bases, roff = reads
    sub = (bases[int(roff[lo]):], roff[lo:] - roff[lo])
    soff, svals = gpu.pseudoalign(sub, 0)
    assert np.array_equal(soff, off[lo:] - off[lo]) and np.array_equal(svals, vals[int(off[lo]):])
    eoff, evals = o.pseudoalign(sub, 0)
    assert np.array_equal(soff, eoff) and np.array_equal(svals, evals)


def test_long_reads_many_color_sets(pair):
    """reads of 2-5 kbp hit far more than 32 (and more than 256) distinct color sets on the fragmented synthetic index:
    exercises the register -> shared-memory -> pool spill of the per-read table and its retry-with-a-larger-pool path"""
    gpu, o = pair
    reads = ck.gen_reads(600, 2000, 5000, seed=8, genomes=gpu.genomes)
    got = gpu.fetch_color_set_ids(reads, want_positive=True)
    exp = o.fetch_color_set_ids(reads, want_positive=True)
    assert _same(got, exp) and np.array_equal(got[2], exp[2])
    for algo, thr in ((0, 1.0), (1, 0.7)):
        assert _same(gpu.pseudoalign(reads, algo, thr), o.pseudoalign(reads, algo, thr))


BIG = ["synth_4546_big.fur", "synth_4546_big.mfur", "synth_4546_big.dfur", "synth_4546_big.mdfur", "synth_4546.fur", "synth_4546.mfur",
       "synth_4546.dfur", "synth_4546.mdfur", "synth_4546_dense.fur", "synth_4546_dense.mfur"]


@pytest.mark.parametrize("index", BIG)
def test_4546_color_standin(index, built_lib):
    """the 4,546-color stand-ins for salmonella_4546 (tools/make_standin_4546.sh; generated, git-ignored, skipped when absent):
    wide bitmaps, complemented sets and 145+ meta partitions through the general color-set kernel"""
    import fulgor_b200 as fg

    try:
        path = ck.index_path(index)
    except FileNotFoundError:
        pytest.skip("fixtures_big fixture not generated on this machine")
    genomes = index.split(".")[0]
    reads = ck.gen_reads(3000 if "big" not in index else 1500, 75, 300, seed=21, genomes=genomes)
    o = ck.Oracle(path)
    with fg.Index.open(path, 0) as gpu:
        got = gpu.fetch_color_set_ids(reads, want_positive=True)
        exp = o.fetch_color_set_ids(reads, want_positive=True)
        assert _same(got, exp) and np.array_equal(got[2], exp[2])
        for algo, thr in ((0, 1.0), (1, 0.8), (1, 0.3)):
            assert _same(gpu.pseudoalign(reads, algo, thr), o.pseudoalign(reads, algo, thr))
    o.close()


@pytest.mark.parametrize("index", ["salmonella_10.mfur", "salmonella_10.dfur", "salmonella_10.mdfur", "synth_200.fur", "synth_200.mfur", "synth_4546.mfur"])
def test_without_the_decoded_table(index, built_lib, monkeypatch):
    """FULGOR_GPU_TABLE_MAX_MB=0: no decoded color-set table, the kernels decode the compressed sets per read
    (color_set_mask in the fused kernel, k_color_sets_general otherwise)"""
    import fulgor_b200 as fg

    try:
        path = ck.index_path(index)
    except FileNotFoundError:
        pytest.skip("fixtures_big not generated on this machine")
    monkeypatch.setenv("FULGOR_GPU_TABLE_MAX_MB", "0")
    reads = ck.gen_reads(4000, 75, 300, seed=23, genomes=index.split(".")[0])
    o = ck.Oracle(path)
    with fg.Index.open(path, 0) as gpu:
        for algo, thr in ((0, 1.0), (1, 0.8), (1, 0.2)):
            assert _same(gpu.pseudoalign(reads, algo, thr), o.pseudoalign(reads, algo, thr))
    o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", [0, 1500])
def test_deduplicated_full_intersection(pair, chunk, monkeypatch):
    """fulgor_gpu_pseudoalign_dedup (the reference's --deduplicate, tools/pseudoalign.cpp:92-226): reads drawn WITH repeats;
    every read's colors through its representative == pseudoalign_full_intersection, the groups are exactly the distinct
    color-set-id lists, only representatives own values (so the intersection ran once per group). chunk = 1500: the call is
    processed in several chunks and the groups must still span the WHOLE call, like the reference's whole-file deduplication"""
    gpu, o = pair
    if chunk:
        monkeypatch.setenv("FULGOR_GPU_CHUNK_READS", str(chunk))
    base = ck.gen_reads(800, 100, 250, seed=77, genomes=gpu.genomes)
    seqs = [base[0][int(base[1][i]):int(base[1][i + 1])].tobytes() for i in range(800)]
    rng = np.random.default_rng(3)
    picks = [seqs[j] for j in rng.integers(0, 800, 5000)] + [b"", b"ACGT" * 10, b"N" * 80, seqs[0].lower()]
    reads = ck.reads_from_list(picks)
    rep, off, vals = gpu.pseudoalign_dedup(reads)
    groups = ck.check_dedup(rep, off, vals, o.pseudoalign(reads, 0), o.fetch_color_set_ids(reads))
    assert groups <= 800
    # E2BIG protocol: exact capacity reported, second call succeeds
    rep2, off2, vals2 = gpu.pseudoalign_dedup(reads, cap=1)
    ck.check_dedup(rep2, off2, vals2, o.pseudoalign(reads, 0), o.fetch_color_set_ids(reads))


def test_differential_sets_without_the_table(built_lib, monkeypatch):
    """differential containers of more than 32 colors queried WITHOUT the decoded table (k_color_sets_general decodes representative
    XOR differences per read): diff_intersect / merge_diff / merge_metadiff of the reference (src/ps_full_intersection.cpp:130-240,
    src/ps_threshold_union.cpp:123-318)"""
    import fulgor_b200 as fg

    monkeypatch.setenv("FULGOR_GPU_TABLE_MAX_MB", "0")
    for index in ("synth_200.dfur", "synth_200.mdfur"):
        reads = ck.gen_reads(3000, 75, 300, seed=29, genomes="synth_200")
        path = ck.index_path(index)
        o = ck.Oracle(path)
        with fg.Index.open(path, 0) as gpu:
            for algo, thr in ((0, 1.0), (1, 0.8), (1, 0.2)):
                assert _same(gpu.pseudoalign(reads, algo, thr), o.pseudoalign(reads, algo, thr))
            rep, off, vals = gpu.pseudoalign_dedup(reads)
            ck.check_dedup(rep, off, vals, o.pseudoalign(reads, 0), o.fetch_color_set_ids(reads))
        o.close()


def test_kmer_conservation_and_matches(pair):
    """fulgor_gpu_kmer_conservation / fulgor_gpu_kmer_matches == index::kmer_conservation / index::kmer_matches
    (src/kmer_conservation.cpp:7-54, src/kmer_matches.cpp:7-30) on mixed-length reads and the edge reads"""
    gpu, o = pair
    seqs = [b"", b"A", b"ACGT" * 8, b"N" * 100]
    r0 = ck.gen_reads(50, 150, 150, seed=3, genomes=gpu.genomes)
    one = r0[0][: int(r0[1][1])].tobytes()
    seqs += [one[:31], one[:75] + b"N" + one[76:], one.lower(), one + one[:40]]
    for reads in (ck.gen_reads(3000, 75, 300, seed=91, genomes=gpu.genomes), ck.reads_from_list(seqs)):
        toff, tr = gpu.kmer_conservation(reads)
        eoff, etr = o.kmer_conservation(reads)
        assert np.array_equal(toff, eoff) and np.array_equal(tr, etr)
        toff2, tr2 = gpu.kmer_conservation(reads, cap=1)  # E2BIG protocol
        assert np.array_equal(toff2, eoff) and np.array_equal(tr2, etr)
        woff, words, counts = gpu.kmer_matches(reads)
        koff, pos, ecounts = o.kmer_matches(reads)
        assert np.array_equal(ck.unpack_positive_words(woff, words, koff), pos)
        assert np.array_equal(counts, ecounts)


def test_whole_dictionary_walk_multi_partition_and_skew(built_lib):
    """synth_skew.fur on the GPU: reads tile the genomes the index was built from, so EVERY k-mer of the dictionary is looked up --
    through each of the 8 minimizer-MPHF partitions (partitioned_phf.hpp:155-159) and, for the ~10 % of the k-mers that live in
    buckets of more than 64 super-k-mers, through every non-empty skew partition including the absorbing, itself partitioned,
    last one (skew_index.hpp:40-52; a k-mer of such a bucket can only be found that way). All positive, ids == the oracle's."""
    import fulgor_b200 as fg

    path = ck.index_path("synth_skew.fur")
    o = ck.Oracle(path)
    reads = ck.tile_genomes("synth_skew")
    with fg.Index.open(path, 0) as gpu:
        got = gpu.fetch_color_set_ids(reads, want_positive=True)
        exp = o.fetch_color_set_ids(reads, want_positive=True)
        assert _same(got, exp) and np.array_equal(got[2], exp[2])
        assert np.array_equal(got[2], np.maximum(0, np.diff(reads[1].astype(np.int64)) - (o.k - 1)))
        toff, tr = gpu.kmer_conservation(reads)  # the per-k-mer view: runs cover every k-mer of every read
        eoff, etr = o.kmer_conservation(reads)
        assert np.array_equal(toff, eoff) and np.array_equal(tr, etr)
        for algo, thr in ((0, 1.0), (1, 0.8)):
            assert _same(gpu.pseudoalign(reads, algo, thr), o.pseudoalign(reads, algo, thr))
    o.close()


def test_salmonella_4546_scale_standin_walk(built_lib):
    """synth_4546_big.fur, the salmonella_4546-SCALE stand-in (tools/make_standin_4546.sh with SYNTH_EXTRA, built by the
    reference's own `load -m 20`): ~47 M k-mers, 3 minimizer-MPHF partitions, 7 skew size classes up to buckets of > 2^13
    super-k-mers, dictionary image and decoded color-set table larger than the L2. Reads tile every 20th genome (the part of
    the collection that travels with the repo): every one of their ~45 M k-mers must be positive (a size-independent
    property), and on a slice of the walk the per-read ids equal the oracle's."""
    import fulgor_b200 as fg
    from fulgor_b200 import imageview as iv

    try:
        path = ck.index_path("synth_4546_big.fur")
    except FileNotFoundError:
        pytest.skip("fixtures_big/synth_4546_big.fur not generated on this machine")
    img = fg.build_image(path)
    h = iv.header(img)
    phfs = iv.section(img, h.off_phfs, "<u8", 4 * h.num_phfs).reshape(-1, 4)
    assert phfs[0][1] >= 2 and h.num_skew >= 3 and h.num_kmers >= 40_000_000
    assert h.off_hybrids > 126 << 20, "the SSHash part of the image must not fit the L2"
    reads = ck.tile_genomes("synth_4546_big", read_len=1000)
    n = len(reads[1]) - 1
    o = ck.Oracle(path)
    with fg.Index.from_image(img, 0) as gpu:
        off, vals, npos = gpu.fetch_color_set_ids(reads, want_positive=True, cap=64 * n)
        assert np.array_equal(npos, np.maximum(0, np.diff(reads[1].astype(np.int64)) - (o.k - 1))), "a dictionary k-mer was not found"
        lo, hi = n // 3, n // 3 + 4000
        sub = (reads[0][int(reads[1][lo]):int(reads[1][hi])], reads[1][lo:hi + 1] - reads[1][lo])
        eoff, evals, enpos = o.fetch_color_set_ids(sub, want_positive=True)
        assert np.array_equal(off[lo:hi + 1] - off[lo], eoff) and np.array_equal(vals[int(off[lo]):int(off[hi])], evals)
    o.close()


def _nasty_reads(genomes):
    """invalid characters at word boundaries (the packed form has 16 bases per word), at the ends, in runs; lengths around 16 / 32"""
    g = ck.gen_reads(24, 150, 150, seed=31, genomes=genomes)
    s = [g[0][int(g[1][i]):int(g[1][i + 1])].tobytes() for i in range(24)]
    out = [s[0][:15] + b"N" + s[0][16:], s[1][:16] + b"n" + s[1][17:], b"N" + s[2][1:], s[3][:149] + b"N", s[4][:31] + b"N" * 18 + s[4][49:],
           s[5][:47] + b"." + s[5][48:100] + b"R" + s[5][101:], s[6] + s[7][:10] + b"N" + s[7][11:], (s[8] + s[9])[:159] + b"N" + s[10],
           s[11][:16], s[12][:32], s[13][:33], s[14].lower(), b"N" * 33, b"", b"ACGTN", s[15] + s[16] + s[17][:17]]
    return out


@pytest.mark.parametrize("chunk", [0, 700])
def test_packed_reads_and_bitmap_results(pair, chunk, monkeypatch):
    """the compact forms of include/fulgor_gpu.h: packed reads in (fulgor_gpu_pack_reads -> fulgor_gpu_pseudoalign_packed) and
    bitmap rows out (fulgor_gpu_pseudoalign_bitmaps / _packed_bitmaps) give exactly the lists of fulgor_gpu_pseudoalign == the
    oracle's; chunk = 700 cuts the batch into many pipeline chunks (the invalid-position list is sliced per chunk)"""
    import fulgor_b200 as fg

    gpu, o = pair
    if chunk:
        monkeypatch.setenv("FULGOR_GPU_CHUNK_READS", str(chunk))
    g = ck.gen_reads(_sized(gpu, 6000), 75, 300, seed=77, genomes=gpu.genomes)
    seqs = [g[0][int(g[1][i]):int(g[1][i + 1])].tobytes() for i in range(len(g[1]) - 1)]
    nasty = _nasty_reads(gpu.genomes)
    for i, s in enumerate(nasty):  # spread the reads with invalid characters over the batch (and over the chunks)
        seqs.insert((i * 397) % len(seqs), s)
    e = edge_reads(gpu.genomes)
    seqs += [e[0][int(e[1][i]):int(e[1][i + 1])].tobytes() for i in range(len(e[1]) - 1)]
    reads = ck.reads_from_list(seqs)
    packed = fg.pack_reads(reads)
    assert packed[2].size > 30
    for algo, thr in ((0, 1.0), (1, 0.8), (1, 0.3)):
        exp = o.pseudoalign(reads, algo, thr)
        assert _same(gpu.pseudoalign_packed(packed, algo, thr), exp)
        assert _same(gpu.pseudoalign_packed(packed, algo, thr, cap=1), exp)  # E2BIG protocol
        assert _same(fg.unpack_bitmaps(gpu.pseudoalign_bitmaps(reads, algo, thr), gpu.num_colors), exp)
        assert _same(fg.unpack_bitmaps(gpu.pseudoalign_bitmaps(packed, algo, thr, packed=True), gpu.num_colors), exp)
    empty = ck.reads_from_list([])
    assert gpu.pseudoalign_bitmaps(empty).shape[0] == 0
    off, vals = gpu.pseudoalign_packed(fg.pack_reads(empty))
    assert off.tolist() == [0] and vals.size == 0
