"""Pins the plain-C oracle (oracle/fulgor_oracle.c) against the UNMODIFIED reference compiled from /root/reference
(oracle/_ref/libfulgor_ref.so, built by `make -C oracle ref`). Skipped where that library was not built."""
import os

import numpy as np
import pytest

import _checkers as ck

pytestmark = pytest.mark.skipif(not ck.reference_available(), reason="oracle/_ref/libfulgor_ref.so not built (needs /root/reference)")

INDEXES = ["salmonella_10.fur", "salmonella_10.mfur", "salmonella_10.dfur", "salmonella_10.mdfur", "synth_200.fur", "synth_200.mfur", "synth_200.dfur", "synth_200.mdfur", "synth_skew.fur"]


@pytest.fixture(scope="module", params=INDEXES)
def pair(request):
    path = ck.index_path(request.param)
    o, r = ck.Oracle(path), ck.Reference(path)
    o.genomes = request.param.split(".")[0]
    yield o, r
    o.close()
    r.close()


def _same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_info(pair):
    o, r = pair
    assert (o.k, o.m, o.num_kmers, o.num_unitigs, o.num_colors, o.num_color_sets, o.type) == (
        r.k, r.m, r.num_kmers, r.num_unitigs, r.num_colors, r.num_color_sets, r.type)


def test_every_color_set(pair):
    o, r = pair
    for i in range(0, o.num_color_sets, 1 if o.num_color_sets < 1000 else 5):
        assert np.array_equal(o.color_set(i), r.color_set(i)), i


def test_u2c_sample(pair):
    o, r = pair
    rng = np.random.default_rng(1)
    for u in list(rng.integers(0, o.num_unitigs, 2000)) + [0, o.num_unitigs - 1]:
        assert o.u2c(int(u)) == r.u2c(int(u))


def test_streaming_lookup_per_kmer(pair):
    o, r = pair
    bases, off = ck.gen_reads(400, 75, 300, seed=11, genomes=o.genomes)
    for i in range(400):
        seq = bases[int(off[i]):int(off[i + 1])].tobytes()
        if i % 7 == 0:
            seq = seq[:50] + b"N" + seq[51:]
        assert np.array_equal(o.lookup_read(seq), r.lookup_read(seq)), i


@pytest.mark.parametrize("lens", [(150, 150), (75, 300)])
def test_fetch_color_set_ids(pair, lens):
    o, r = pair
    reads = ck.gen_reads(6000, lens[0], lens[1], seed=42, genomes=o.genomes)
    assert _same(o.fetch_color_set_ids(reads), r.fetch_color_set_ids(reads, threads=4))


@pytest.mark.parametrize("algo,thr", [(0, 1.0), (1, 0.8), (1, 1.0), (1, 0.05)])
def test_pseudoalign(pair, algo, thr):
    o, r = pair
    reads = ck.gen_reads(6000, 75, 300, seed=1234, genomes=o.genomes)
    assert _same(o.pseudoalign(reads, algo, thr), r.pseudoalign(reads, algo, thr, threads=4))


def test_edge_reads(pair):
    o, r = pair
    g = ck.gen_reads(8, seed=99, genomes=o.genomes)
    s = [g[0][int(g[1][i]):int(g[1][i + 1])].tobytes() for i in range(8)]
    reads = ck.reads_from_list([b"", b"A", s[0][:30], s[0][:31], s[1].lower(), b"N" * 150, s[2][:75] + b"N" + s[2][76:], b"A" * 200,
                                s[3] + s[4], s[5][:149] + b"X"])
    for algo, thr in ((0, 1.0), (1, 0.8), (1, 0.001)):
        assert _same(o.pseudoalign(reads, algo, thr), r.pseudoalign(reads, algo, thr))
    assert _same(o.fetch_color_set_ids(reads), r.fetch_color_set_ids(reads))


def test_kmer_tools(pair):
    """index::kmer_conservation / index::kmer_matches (src/kmer_conservation.cpp:7-54, src/kmer_matches.cpp:7-30)"""
    o, r = pair
    reads = ck.gen_reads(300, 75, 300, seed=61, genomes=o.genomes)
    seqs = [b"", b"ACGT" * 5, reads[0][: int(reads[1][1])].tobytes()[:40] + b"N" + reads[0][: int(reads[1][1])].tobytes()[41:]]
    extra = ck.reads_from_list(seqs)
    for batch in (reads, extra):
        a, b = o.kmer_conservation(batch), r.kmer_conservation(batch)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        a, b = o.kmer_matches(batch), r.kmer_matches(batch)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.skipif(not os.path.exists(ck.REF_DBG_SO), reason="oracle/_ref/libfulgor_ref_dbg.so not built")
@pytest.mark.parametrize("index", INDEXES)
def test_reference_self_checks_run_clean(index):
    """SURVEY.md section 4, mechanism 1: the reference compiled WITHOUT -DNDEBUG checks itself on the hot path (every streamed
    k-mer answer against an independent lookup, streaming_query.hpp:107; every full intersection against decode-all +
    std::set_intersection, src/ps_full_intersection.cpp:399; every threshold union against a naive score array,
    src/ps_threshold_union.cpp:401) and aborts on any disagreement. It must run clean on our fixtures and reads, and give
    the oracle's answers."""
    path = ck.index_path(index)
    r, o = ck.Reference(path, self_checking=True), ck.Oracle(path)
    reads = ck.gen_reads(1500 if index.startswith("salmonella") else 400, 75, 300, seed=123, genomes=index.split(".")[0])
    assert _same(r.fetch_color_set_ids(reads), o.fetch_color_set_ids(reads))
    for algo, thr in ((0, 1.0), (1, 0.8), (1, 0.3)):
        assert _same(r.pseudoalign(reads, algo, thr), o.pseudoalign(reads, algo, thr))
    r.close()
    o.close()


def test_whole_dictionary_walk_multi_partition_and_skew():
    """synth_skew.fur (tools/make_skew_fixture.sh): 8 minimizer-MPHF partitions (partitioned_phf.hpp:155-159), a skew index
    with 7 size classes of which one is EMPTY and the last one absorbs buckets beyond 2^max_l and is itself a 3-partition
    MPHF (skew_index.hpp:40-52, dictionary.cpp:61-73). Every k-mer of every genome the index was built from is looked up
    (reads tile the genomes): all positive, and the oracle == the reference on every read."""
    path = ck.index_path("synth_skew.fur")
    o, r = ck.Oracle(path), ck.Reference(path)
    reads = ck.tile_genomes("synth_skew")
    a = o.fetch_color_set_ids(reads, want_positive=True)
    assert _same(a, r.fetch_color_set_ids(reads, threads=4))
    assert np.array_equal(a[2], np.maximum(0, np.diff(reads[1].astype(np.int64)) - (o.k - 1)))
    assert _same(o.pseudoalign(reads, 0), r.pseudoalign(reads, 0, threads=4))
    bases, off = reads
    for i in range(0, len(off) - 1, 97):  # per-k-mer streaming answers on a sample of the tiles
        seq = bases[int(off[i]):int(off[i + 1])].tobytes()
        assert np.array_equal(o.lookup_read(seq), r.lookup_read(seq)), i
    o.close()
    r.close()
