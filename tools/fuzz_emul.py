"""dev helper (CPU): differential fuzzing of the CUDA kernels on the lock-step SIMT emulator (tests/simt_emul.cpp) against the
oracle: reads of random lengths with substitutions, invalid characters, lower case, repeats and chimeras.
    python tools/fuzz_emul.py [index] [rounds] [seed]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _checkers as ck  # noqa: E402
import fulgor_b200 as fg  # noqa: E402

index = sys.argv[1] if len(sys.argv) > 1 else "salmonella_10.fur"
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 5
seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 1
E = C.CDLL(os.path.join(ROOT, "build", "libfg_simt_emul.so"))
E.emul_pseudoalign.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_int, C.c_int]
E.emul_fetch_color_set_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint, C.c_int]
E.emul_pseudoalign_dedup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_int]
E.emul_kmer_tool.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint, C.c_int]
path = ck.index_path(index)
img, o = fg.build_image(path), ck.Oracle(path)
genomes = index.split(".")[0]
small = o.num_colors <= 32


def mutate(rng, s):
    s = bytearray(s)
    kind = rng.integers(0, 8)
    if kind == 0 and len(s) > 40:  # invalid characters
        for p in rng.integers(0, len(s), rng.integers(1, 4)):
            s[p] = rng.choice(list(b"NnXRY-*.@"))
    elif kind == 1:  # lower / mixed case
        for p in range(len(s)):
            if rng.random() < 0.5:
                s[p] = s[p] | 0x20
    elif kind == 2 and len(s) > 60:  # truncate around k
        s = s[: rng.integers(25, 40)]
    elif kind == 3:  # low complexity
        s = bytearray(bytes(rng.choice(list(b"ACGT"), 2)) * rng.integers(20, 120))
    elif kind == 4 and len(s) > 80:  # internal repeat
        a = rng.integers(0, len(s) - 40)
        s = s + s[a:a + 40] * rng.integers(1, 4)
    return bytes(s)


bad = 0
for rd in range(rounds):
    rng = np.random.default_rng(seed0 * 1000 + rd)
    n = 160 if small else 60
    base = ck.gen_reads(n, 31, 420, seed=int(rng.integers(1, 1 << 30)), sub_rate=float(rng.choice([0.0, 0.01, 0.05])), genomes=genomes)
    seqs = [base[0][int(base[1][i]):int(base[1][i + 1])].tobytes() for i in range(n)]
    seqs = [mutate(rng, s) for s in seqs]
    for _ in range(6):  # chimeras
        a, b = rng.integers(0, n, 2)
        seqs.append(seqs[a][: rng.integers(10, 100)] + seqs[b][rng.integers(0, 50):])
    reads = ck.reads_from_list(seqs)
    bases, off = reads
    nr = len(off) - 1
    cap = max(1, nr * max(o.num_colors, 64), int(off[nr]))
    grid = int(rng.integers(1, 4))
    generic = int(rng.integers(0, 2))
    # stage 1
    oo, vv, npos = np.zeros(nr + 1, dtype=np.uint64), np.zeros(cap, dtype=np.uint32), np.zeros(nr, dtype=np.uint32)
    assert E.emul_fetch_color_set_ids(img.ctypes.data, bases.ctypes.data, off.ctypes.data, nr, oo.ctypes.data, vv.ctypes.data, cap, npos.ctypes.data, grid, generic) == 0
    e = o.fetch_color_set_ids(reads, want_positive=True)
    ok = np.array_equal(oo, e[0]) and np.array_equal(vv[: int(oo[nr])], e[1]) and np.array_equal(npos, e[2])
    # pseudoalign
    for algo, thr in ((0, 1.0), (1, float(rng.choice([0.05, 0.3, 0.8, 1.0])))):
        for table in ((1,) if o.type >= 2 and not small else (0, 1)):
            oo2, vv2 = np.zeros(nr + 1, dtype=np.uint64), np.zeros(cap, dtype=np.uint32)
            assert E.emul_pseudoalign(img.ctypes.data, algo, thr, bases.ctypes.data, off.ctypes.data, nr, oo2.ctypes.data, vv2.ctypes.data, cap, grid, generic, table) == 0
            e2 = o.pseudoalign(reads, algo, thr)
            ok = ok and np.array_equal(oo2, e2[0]) and np.array_equal(vv2[: int(oo2[nr])], e2[1])
    # kmer-conservation
    to, tr = np.zeros(nr + 1, dtype=np.uint64), np.zeros(3 * cap, dtype=np.uint32)
    assert E.emul_kmer_tool(img.ctypes.data, 0, bases.ctypes.data, off.ctypes.data, nr, to.ctypes.data, tr.ctypes.data, cap, None, grid, generic) == 0
    e3 = o.kmer_conservation(reads)
    ok = ok and np.array_equal(to, e3[0]) and np.array_equal(tr[: 3 * int(to[nr])].reshape(-1, 3), e3[1])
    # kmer-matches (decoded table)
    wo, ww, cc = np.zeros(nr + 1, dtype=np.uint64), np.zeros(cap, dtype=np.uint32), np.zeros((nr, o.num_colors), dtype=np.uint32)
    assert E.emul_kmer_tool(img.ctypes.data, 1, bases.ctypes.data, off.ctypes.data, nr, wo.ctypes.data, ww.ctypes.data, cap, cc.ctypes.data, grid, 0) == 0
    koff, pos, ecounts = o.kmer_matches(reads)
    ok = ok and np.array_equal(ck.unpack_positive_words(wo, ww, koff), pos) and np.array_equal(cc, ecounts)
    # deduplicated full intersection: the batch twice over, so every list has at least two reads
    twice = ck.reads_from_list(seqs + seqs)
    b2, f2 = twice
    n2 = len(f2) - 1
    rep, do, dv = np.zeros(n2, dtype=np.uint32), np.zeros(n2 + 1, dtype=np.uint64), np.zeros(2 * cap, dtype=np.uint32)
    assert E.emul_pseudoalign_dedup(img.ctypes.data, b2.ctypes.data, f2.ctypes.data, n2, rep.ctypes.data, do.ctypes.data, dv.ctypes.data, 2 * cap, grid, 1) == 0
    try:
        ck.check_dedup(rep, do, dv, o.pseudoalign(twice, 0), o.fetch_color_set_ids(twice))
    except AssertionError:
        ok = False
    print(f"round {rd}: {nr} reads, grid {grid}, generic {generic}: {'ok' if ok else 'MISMATCH'}", flush=True)
    bad += not ok
print("mismatching rounds:", bad)
sys.exit(1 if bad else 0)
