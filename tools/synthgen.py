"""synthgen -- synthetic many-color genome collection (test/bench fixture tooling, not on the query path).

Stand-in for collections like salmonella_4546 that cannot be downloaded here: N genomes evolved from one random ancestor
along a random binary tree (per-edge substitutions, clade-specific insertions and deletions), so that k-mers are shared by
clades of every size and the color sets of the resulting index span all three hybrid encodings (sparse delta-gaps, bitmap,
complemented delta-gaps). Deterministic for a given seed.

    python tools/synthgen.py OUTDIR N GENOME_LEN [--seed 1] [--sub 0.004] [--indel 0.15] [--hgt 0]
writes OUTDIR/g00000.fa ... and OUTDIR/list.txt (for `mkdump BASE @OUTDIR/list.txt`)."""
import argparse
import os

import numpy as np


def evolve(seq, rng, sub, indel, pool=None, hgt=0.0):
    seq = seq.copy()
    if pool and rng.random() < hgt:  # horizontal transfer: a stretch of another lineage replaces the same coordinates
        donor = pool[int(rng.integers(0, len(pool)))]
        n = int(rng.integers(500, 5000))
        lim = min(seq.size, donor.size) - n
        if lim > 0:
            at = int(rng.integers(0, lim))
            seq[at:at + n] = donor[at:at + n]
    nsub = rng.binomial(seq.size, sub)
    pos = rng.integers(0, seq.size, nsub)
    seq[pos] = (seq[pos] + rng.integers(1, 4, nsub)) % 4
    if rng.random() < indel:  # clade-specific insertion of novel sequence
        at = int(rng.integers(0, seq.size))
        seq = np.concatenate([seq[:at], rng.integers(0, 4, int(rng.integers(200, 2000)), dtype=np.uint8), seq[at:]])
    if rng.random() < indel and seq.size > 8000:  # deletion
        at = int(rng.integers(0, seq.size - 3000))
        seq = np.concatenate([seq[:at], seq[at + int(rng.integers(200, 2000)):]])
    return seq


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("outdir")
    ap.add_argument("n", type=int)
    ap.add_argument("length", type=int)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--sub", type=float, default=0.004)
    ap.add_argument("--indel", type=float, default=0.15)
    ap.add_argument("--hgt", type=float, default=0.0, help="probability per new lineage of a horizontal transfer from a random lineage")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    os.makedirs(a.outdir, exist_ok=True)
    pool = [rng.integers(0, 4, a.length, dtype=np.uint8)]
    while len(pool) < a.n:  # split a random lineage into two children (Yule tree)
        i = int(rng.integers(0, len(pool)))
        parent = pool.pop(i)
        if a.hgt > 0:
            pool.append(evolve(parent, rng, a.sub, a.indel, pool, a.hgt))
            pool.append(evolve(parent, rng, a.sub, a.indel, pool, a.hgt))
        else:  # (kept separate so that fixtures generated before --hgt existed stay byte-identical)
            pool.append(evolve(parent, rng, a.sub, a.indel))
            pool.append(evolve(parent, rng, a.sub, a.indel))
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    names = []
    for g, seq in enumerate(pool):
        name = os.path.join(a.outdir, f"g{g:05d}.fa")
        with open(name, "wb") as f:
            f.write(f">g{g:05d}\n".encode())
            f.write(lut[seq].tobytes())
            f.write(b"\n")
        names.append(name)
    with open(os.path.join(a.outdir, "list.txt"), "w") as f:
        f.write("\n".join(names) + "\n")


if __name__ == "__main__":
    main()
