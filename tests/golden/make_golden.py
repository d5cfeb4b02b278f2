"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference (oracle/_ref/libfulgor_ref.so,
compiled from /root/reference by `make -C oracle ref`) on seeded synthetic reads. Run in the build container only:

    python tests/golden/make_golden.py

Each .npz holds the read generator parameters (the reads are regenerated from data/salmonella_10.gpk.xz by
tools/readgen.cpp, not stored) and the reference's outputs: per-read sorted distinct color-set ids, and the
full-intersection / threshold-union color lists, in CSR form."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _checkers as ck  # noqa: E402

CASES = [  # name, index, n, min_len, max_len, seed
    ("s10_fur_150", "salmonella_10.fur", 4000, 150, 150, 42),
    ("s10_fur_mixed", "salmonella_10.fur", 4000, 75, 300, 43),
    ("s10_mfur_150", "salmonella_10.mfur", 4000, 150, 150, 42),
    ("s10_mfur_mixed", "salmonella_10.mfur", 4000, 75, 300, 43),
    ("synth200_fur_mixed", "synth_200.fur", 3000, 75, 300, 44),
    ("synth200_mfur_mixed", "synth_200.mfur", 3000, 75, 300, 44),
    ("s10_dfur_mixed", "salmonella_10.dfur", 4000, 75, 300, 45),
    ("s10_mdfur_mixed", "salmonella_10.mdfur", 4000, 75, 300, 45),
    ("synth200_dfur_mixed", "synth_200.dfur", 3000, 75, 300, 46),
    ("synth200_mdfur_mixed", "synth_200.mdfur", 3000, 75, 300, 46),
    ("synthskew_fur_mixed", "synth_skew.fur", 4000, 75, 300, 47),  # multi-partition MPHF, every skew class, palindromic minimizers
]
THRESHOLDS = [0.8, 1.0, 0.3]


def main():
    assert ck.reference_available(), "build the reference first: make -C oracle ref"
    only = set(sys.argv[1:])
    for name, index, n, lo, hi, seed in CASES:
        if only and name not in only:
            continue
        ref = ck.Reference(ck.index_path(index))
        reads = ck.gen_reads(n, lo, hi, seed=seed, genomes=index.split(".")[0])
        out = {"index": index, "n": n, "min_len": lo, "max_len": hi, "seed": seed, "thresholds": np.array(THRESHOLDS),
               "reads_checksum": np.uint64(int(reads[0].astype(np.uint64).sum()))}
        out["cid_off"], out["cids"] = ref.fetch_color_set_ids(reads)
        out["fi_off"], out["fi"] = ref.pseudoalign(reads, 0)
        for j, t in enumerate(THRESHOLDS):
            out[f"tu{j}_off"], out[f"tu{j}"] = ref.pseudoalign(reads, 1, t)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "reads", n, "cids", out["cids"].size, "fi", out["fi"].size)
        ref.close()


if __name__ == "__main__":
    main()
