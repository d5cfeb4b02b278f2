"""synthgen -- synthetic many-color genome collection (test/bench fixture tooling, not on the query path).

Stand-in for collections like salmonella_4546 that cannot be downloaded here: N genomes evolved from one random ancestor
along a random binary tree (per-edge substitutions, clade-specific insertions and deletions), so that k-mers are shared by
clades of every size and the color sets of the resulting index span all three hybrid encodings (sparse delta-gaps, bitmap,
complemented delta-gaps). Deterministic for a given seed.

    python tools/synthgen.py OUTDIR N GENOME_LEN [--seed 1] [--sub 0.004] [--indel 0.15] [--hgt 0] [--novel LO HI] [--plant M]
writes OUTDIR/g00000.fa ... and OUTDIR/list.txt (for `mkdump BASE @OUTDIR/list.txt`).

--novel LO HI: every new lineage gains one stretch of LO..HI novel bases and loses as many (an open pangenome: the number of
distinct k-mers grows with the number of lineages, which is what makes a dictionary of salmonella_4546's size).
--plant M: plants "skew" motifs for minimizer length M into the novel stretches: a handful of M-mers R that win the
minimizer of (almost) every k-mer that contains them -- tiny mixer_64 hash (sshash/include/hash_util.hpp:97, default seed)
and tiny integer value -- each embedded in many different random contexts, so that the SSHash buckets of these minimizers
hold 65 .. 4096+ super-k-mers and the index gets a skew index with every size class (sshash/include/skew_index.hpp:40-52),
including the last partition that absorbs buckets beyond 2^max_l."""
import argparse
import os

import numpy as np


MIX_MUL = 0x517cc1b727220a95
MIX_MAGIC_SEED1 = 0x8fbb8d815c9e092e  # mixer_64's magic for the default build seed 1 (stored in every index file)
SKEW_TARGETS = (90, 180, 360, 720, 1500, 3000, 6000)  # contexts per motif: one bucket per skew size class 2^7 .. > 2^12


def skew_motifs(m, rng, targets=SKEW_TARGETS):
    """one m-mer per target (as base codes in this tool's A0 C1 G2 T3 alphabet) whose hash is below 2^46 and whose six
    most significant bases are A: it is the canonical minimizer of every k-mer around it but for a ~2^-12 chance"""
    enc = np.array([0, 1, 3, 2], dtype=np.uint64)  # A C G T -> sshash codes A0 C1 T2 G3 (kmer.hpp:199)
    out = []
    while len(out) < len(targets):
        cand = rng.integers(0, 4, (1 << 20, m - 6), dtype=np.uint8)
        val = np.zeros(cand.shape[0], dtype=np.uint64)
        for i in range(m - 6):
            val |= enc[cand[:, i]] << np.uint64(2 * i)
        h = (val * np.uint64(MIX_MUL)) ^ np.uint64(MIX_MAGIC_SEED1)
        for j in np.nonzero(h < np.uint64(1 << 46))[0]:
            if len(out) < len(targets):
                out.append(np.concatenate([cand[j], np.zeros(6, dtype=np.uint8)]))
    return out


def palindrome_traps(m, k, rng, count=48):
    """(k + m - 1)-base stretches around an m-mer P that is its OWN reverse complement (even m) such that some k-mer of the
    stretch has P as canonical minimizer read off ONE strand only: the other strand holds P too (same hash) but has an m-mer
    with a smaller hash whose value is larger than P's. How the minimizer reads then says nothing about the orientation of the
    indexed string against a query -- a lookup that infers the orientation from the minimizer must try both."""
    if m % 2:
        return []
    sscode = [0, 1, 3, 2]  # this tool's A0 C1 G2 T3 -> sshash A0 C1 T2 G3
    mask = (1 << 64) - 1

    def value(b):
        v = 0
        for i, c in enumerate(b):
            v |= sscode[c] << (2 * i)
        return v

    def strand_min(b):  # (value, hash) of the first m-mer with the smallest hash (util::compute_minimizer)
        best = None
        for j in range(len(b) - m + 1):
            v = value(b[j:j + m])
            h = ((v * MIX_MUL) & mask) ^ MIX_MAGIC_SEED1
            if best is None or h < best[1]:
                best = (v, h)
        return best[0]

    out = []
    while len(out) < count:
        x = [3, 3, 3] + [int(c) for c in rng.integers(0, 4, m // 2 - 3)]  # TTT...: P ends in ...AAA, a small value
        p = x + [3 - c for c in reversed(x)]
        s = [int(c) for c in rng.integers(0, 4, k - m)] + p + [int(c) for c in rng.integers(0, 4, k - m)]
        pv = value(p)
        for i in range(len(s) - k + 1):
            kmer = s[i:i + k]
            vf, vr = strand_min(kmer), strand_min([3 - c for c in reversed(kmer)])
            if vf != vr and min(vf, vr) == pv:
                out.append(np.array(s, dtype=np.uint8))
                break
    return out


def novel_stretch(rng, n, plant):
    """n random bases; with --plant, motifs in fresh random contexts at a rate that reaches the targets over all lineages"""
    s = rng.integers(0, 4, n, dtype=np.uint8)
    if plant is not None:
        motifs, weights, rate = plant[:3]
        m = motifs[0].size
        for _ in range(int(rng.poisson(n * rate))):
            r = motifs[int(rng.choice(len(motifs), p=weights))]
            at = int(rng.integers(16, max(17, n - m - 16)))
            if at + m + 16 <= n:
                s[at:at + m] = r
        traps = plant[3] if len(plant) > 3 else []
        if traps and n > 200:
            for _ in range(2):
                t = traps[int(rng.integers(0, len(traps)))]
                at = int(rng.integers(0, n - t.size))
                s[at:at + t.size] = t
    return s


def evolve_open(seq, rng, sub, lo, hi, pool, hgt, plant):
    """like evolve(), for --novel: one novel stretch in, one stretch of the same length out"""
    seq = seq.copy()
    if pool and rng.random() < hgt:
        donor = pool[int(rng.integers(0, len(pool)))]
        n = int(rng.integers(500, 5000))
        lim = min(seq.size, donor.size) - n
        if lim > 0:
            at = int(rng.integers(0, lim))
            seq[at:at + n] = donor[at:at + n]
    nsub = rng.binomial(seq.size, sub)
    pos = rng.integers(0, seq.size, nsub)
    seq[pos] = (seq[pos] + rng.integers(1, 4, nsub)) % 4
    n = int(rng.integers(lo, hi + 1))
    at = int(rng.integers(0, seq.size))
    seq = np.concatenate([seq[:at], novel_stretch(rng, n, plant), seq[at:]])
    at = int(rng.integers(0, seq.size - n))
    return np.concatenate([seq[:at], seq[at + n:]])


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
    ap.add_argument("--novel", type=int, nargs=2, metavar=("LO", "HI"), help="open pangenome: novel bases gained (and lost) per new lineage")
    ap.add_argument("--plant", type=int, metavar="M", help="plant skew-bucket motifs for minimizer length M (needs --novel)")
    ap.add_argument("--plant-scale", type=float, default=1.0, help="scale of the planted bucket sizes")
    ap.add_argument("--plant-palindromes", action="store_true",
                    help="also plant stretches around reverse-complement-palindromic minimizers (see palindrome_traps)")
    ap.add_argument("--plant-targets", type=str, default=None, help="comma-separated contexts per motif (default: one per skew size class)")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    os.makedirs(a.outdir, exist_ok=True)
    pool = [rng.integers(0, 4, a.length, dtype=np.uint8)]
    plant = None
    if a.plant:
        base_targets = [int(t) for t in a.plant_targets.split(",")] if a.plant_targets else SKEW_TARGETS
        targets = [max(70, int(t * a.plant_scale)) for t in base_targets]
        motifs = skew_motifs(a.plant, rng, targets)
        w = np.array(targets, dtype=np.float64)
        novel_total = 2 * (a.n - 1) * (a.novel[0] + a.novel[1]) / 2
        plant = (motifs, w / w.sum(), 1.15 * w.sum() / novel_total, palindrome_traps(a.plant, 31, rng) if a.plant_palindromes else [])
    while len(pool) < a.n:  # split a random lineage into two children (Yule tree)
        i = int(rng.integers(0, len(pool)))
        parent = pool.pop(i)
        if a.novel:
            pool.append(evolve_open(parent, rng, a.sub, a.novel[0], a.novel[1], pool, a.hgt, plant))
            pool.append(evolve_open(parent, rng, a.sub, a.novel[0], a.novel[1], pool, a.hgt, plant))
        elif a.hgt > 0:
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
