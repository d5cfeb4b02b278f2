"""dev helper (GPU box): host-buffer e2e (pinned buffers) of salmonella_10 full intersection for several FULGOR_GPU_CHUNK_READS."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fulgor_b200 as fg
import _checkers as ck

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
bases, off = ck.gen_reads(n, seed=42, threads=16)
idx = fg.Index.open(ck.index_path("salmonella_10.fur"), 0)
cap = n * idx.num_colors
pb, po, pc, pv = (fg.PinnedBuffer(x) for x in (bases.size + 64, 8 * (n + 1), 8 * (n + 1), 4 * cap))
pb.view(np.uint8, bases.size)[:] = bases
po.view(np.uint64, n + 1)[:] = off
L = fg.lib()
ref = None
for chunk in sys.argv[2:] or ["1048576", "524288", "262144"]:
    os.environ["FULGOR_GPU_CHUNK_READS"] = chunk
    for it in range(3):
        assert L.fulgor_gpu_pseudoalign(idx._h, 0, 1.0, pb.ptr, po.ptr, n, pc.ptr, pv.ptr, cap) == 0
    t = time.perf_counter()
    for it in range(10):
        assert L.fulgor_gpu_pseudoalign(idx._h, 0, 1.0, pb.ptr, po.ptr, n, pc.ptr, pv.ptr, cap) == 0
    dt = (time.perf_counter() - t) / 10
    tot = int(pc.view(np.uint64, n + 1)[n])
    chk = int(pv.view(np.uint32, tot).astype(np.uint64).sum()) ^ int(pc.view(np.uint64, n + 1).sum())
    ref = chk if ref is None else ref
    print(f"chunk {chunk}: {dt*1e3:.2f} ms/step  {n/dt/1e6:.1f} M reads/s  H2D-equivalent {(bases.size + 8*n)/dt/1e9:.1f} GB/s  checksum_ok={chk == ref}")
