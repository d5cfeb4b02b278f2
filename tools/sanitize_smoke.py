"""dev helper (GPU box, under compute-sanitizer): a small pass over every kernel of the library -- fused and unfused
pseudoalignment (table and per-read decoding), stage 1, deduplication, the per-k-mer tools -- on the four index types."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _checkers as ck
import fulgor_b200 as fg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
for name in ("salmonella_10.fur", "salmonella_10.mdfur", "synth_200.fur", "synth_200.mfur", "synth_200.dfur", "synth_skew.fur"):
    genomes = name.split(".")[0]
    reads = ck.gen_reads(n, 75, 300, seed=5, genomes=genomes)
    long_reads = ck.gen_reads(20, 2000, 4000, seed=6, genomes=genomes)
    for table in ("", "0"):
        if table:
            os.environ["FULGOR_GPU_TABLE_MAX_MB"] = table
        else:
            os.environ.pop("FULGOR_GPU_TABLE_MAX_MB", None)
        with fg.Index.open(ck.index_path(name), 0) as idx:
            o = ck.Oracle(ck.index_path(name))
            for r in (reads, long_reads):
                assert all(np.array_equal(a, b) for a, b in zip(idx.fetch_color_set_ids(r), o.fetch_color_set_ids(r)))
                for algo, thr in ((0, 1.0), (1, 0.7)):
                    assert all(np.array_equal(a, b) for a, b in zip(idx.pseudoalign(r, algo, thr), o.pseudoalign(r, algo, thr)))
            packed = fg.pack_reads(reads)  # the compact forms: packed reads in, bitmap rows out
            for algo, thr in ((0, 1.0), (1, 0.7)):
                exp = o.pseudoalign(reads, algo, thr)
                assert all(np.array_equal(a, b) for a, b in zip(idx.pseudoalign_packed(packed, algo, thr), exp))
                assert all(np.array_equal(a, b) for a, b in zip(fg.unpack_bitmaps(idx.pseudoalign_bitmaps(packed, algo, thr, packed=True), idx.num_colors), exp))
            rep, off, vals = idx.pseudoalign_dedup(reads)
            ck.check_dedup(rep, off, vals, o.pseudoalign(reads, 0), o.fetch_color_set_ids(reads))
            toff, tr = idx.kmer_conservation(reads)
            eoff, etr = o.kmer_conservation(reads)
            assert np.array_equal(toff, eoff) and np.array_equal(tr, etr)
            if not table:
                woff, words, counts = idx.kmer_matches(reads)
                koff, pos, ecounts = o.kmer_matches(reads)
                assert np.array_equal(ck.unpack_positive_words(woff, words, koff), pos) and np.array_equal(counts, ecounts)
            o.close()
    print("ok", name, flush=True)
