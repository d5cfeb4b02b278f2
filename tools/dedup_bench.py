"""dev helper (GPU box): full intersection with and without cross-read deduplication (fulgor_gpu_pseudoalign_dedup) through the
host-buffer C ABI with pinned buffers, on reads drawn WITH repeats (real read sets repeat the same color-set-id list massively,
SURVEY.md 8(f) rank 1). Prints one JSON line.
    python tools/dedup_bench.py [index] [n_reads] [n_distinct] [steps]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _checkers as ck  # noqa: E402  (read generator only)
import fulgor_b200 as fg  # noqa: E402

index = sys.argv[1] if len(sys.argv) > 1 else "salmonella_10.fur"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
distinct = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
genomes = index.split(".")[0]
base, boff = ck.gen_reads(distinct, 150, 150, seed=7, genomes=genomes)
pick = np.random.default_rng(1).integers(0, distinct, n)
bases = base.reshape(distinct, 150)[pick].reshape(-1)
off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(150))
L = fg.lib()
with fg.Index.open(ck.index_path(index), 0) as idx:
    so, sv = idx.pseudoalign((bases[: 150 * 20000], off[:20001]), 0)
    cap = int(int(so[20000]) / 20000 * n * 1.3) + (1 << 20)
    pb, po, pc, pv, pr = (fg.PinnedBuffer(x) for x in (bases.size + 64, 8 * (n + 1), 8 * (n + 1), 4 * cap, 4 * n))
    pb.view(np.uint8, bases.size)[:] = bases
    po.view(np.uint64, n + 1)[:] = off

    def plain():
        rc = L.fulgor_gpu_pseudoalign(idx._h, 0, 1.0, pb.ptr, po.ptr, n, pc.ptr, pv.ptr, cap)
        assert rc == 0, rc

    def dedup():
        rc = L.fulgor_gpu_pseudoalign_dedup(idx._h, pb.ptr, po.ptr, n, pr.ptr, pc.ptr, pv.ptr, cap)
        assert rc == 0, rc

    out = {"index": index, "reads": n, "distinct_reads": distinct, "steps": steps}
    for name, fn in (("plain", plain), ("dedup", dedup)):
        fn()
        fn()
        t = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = (time.perf_counter() - t) / steps
        total = int(pc.view(np.uint64, n + 1)[n])
        out[name] = {"ms_per_step": dt * 1e3, "reads_per_s": n / dt, "colors_copied_back": total}
        if name == "plain":
            ref_off = pc.view(np.uint64, n + 1).copy()
            ref_val = pv.view(np.uint32, total).copy()
    rep = pr.view(np.uint32, n)
    coff = pc.view(np.uint64, n + 1)
    vals = pv.view(np.uint32, int(coff[n]))
    sample = np.random.default_rng(2).integers(0, n, 20000)
    ok = all(np.array_equal(vals[int(coff[rep[i]]):int(coff[rep[i] + 1])], ref_val[int(ref_off[i]):int(ref_off[i + 1])]) for i in sample)
    out["groups"] = int((rep == np.arange(n, dtype=np.uint32)).sum())
    out["dedup_matches_plain_on_sample"] = bool(ok)
    print(json.dumps(out))
