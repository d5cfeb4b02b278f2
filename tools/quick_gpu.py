"""dev helper: quick kernel timing on device-resident reads (not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import fulgor_b200 as fg
import _checkers as ck

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
name = sys.argv[2] if len(sys.argv) > 2 else "salmonella_10.fur"
idx = fg.Index.open(ck.index_path(name), 0)
t = time.time(); bases, off = ck.gen_reads(n, seed=42, threads=32, genomes=name.split(".")[0]); print("gen", time.time() - t)
db = torch.from_numpy(bases).cuda(); do = torch.from_numpy(off.view(np.int64)).cuda()
dco = torch.zeros(n + 1, dtype=torch.int64, device="cuda"); dc = torch.zeros(n * idx.num_colors, dtype=torch.int32, device="cuda")
for algo, thr in ((0, 1.0), (1, 0.8)):
    for it in range(4):
        tot = idx.pseudoalign_device(algo, thr, db.data_ptr(), do.data_ptr(), n, 0, dco.data_ptr(), dc.data_ptr(), dc.numel())
        l, ms = idx.last_kernel_times()
        print(f"algo {algo} n={n} total={tot} launches={l} ms={ms} reads/s={n / (sum(ms) / 1e3):.3e}")
# host API
for it in range(3):
    t = time.time(); o, v = idx.pseudoalign((bases, off), 0); dt = time.time() - t
    print(f"host API (pageable) {dt*1e3:.1f} ms  {n/dt:.3e} reads/s")
