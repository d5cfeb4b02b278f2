"""N > 1 host logic on CPU: world_size-2 gloo run of the replication path (rank 0 builds the image, one broadcast, every rank
gets identical bytes) and of the read sharding (per-rank generated shards == slices of the global batch)."""
import os
import subprocess
import sys

import numpy as np

import _checkers as ck

WORKER = r'''
import os, sys, hashlib
sys.path.insert(0, os.environ["FG_ROOT"]); sys.path.insert(0, os.path.join(os.environ["FG_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
import _checkers as ck
import fulgor_b200 as fg
from fulgor_b200 import replicate
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
image = replicate.broadcast_image(ck.index_path("synth_200.mfur"), torch.device("cpu"))
arr = image.numpy()
info = fg.image_info(arr)
digest = hashlib.sha256(arr.tobytes()).hexdigest()
gathered = [None] * world
dist.all_gather_object(gathered, (digest, int(info.num_colors), int(info.num_color_sets), int(arr.size)))
assert len(set(gathered)) == 1, gathered
assert info.num_colors == 200 and info.type == 1
# weak-scaling shards: rank r generates reads [r*n, (r+1)*n) of the global stream
n = 500
mine = ck.gen_reads(n, 75, 300, seed=42, first=rank * n, genomes="synth_200")
whole = ck.gen_reads(n * world, 75, 300, seed=42, genomes="synth_200")
lo, hi = int(whole[1][rank * n]), int(whole[1][(rank + 1) * n])
assert np.array_equal(mine[0], whole[0][lo:hi])
# strong-scaling split of one batch, balanced by k-mer count
cuts = replicate.shard_by_kmers(whole[1], 31, world)
assert cuts[0] == 0 and cuts[-1] == n * world and all(a <= b for a, b in zip(cuts, cuts[1:]))
L = np.diff(whole[1].astype(np.int64)); work = np.maximum(L - 30, 0)
parts = [work[cuts[g]:cuts[g + 1]].sum() for g in range(world)]
assert max(parts) - min(parts) <= 0.02 * sum(parts), parts
assert replicate.shard_range(10, rank, world) == ((0, 5) if rank == 0 else (5, 10))
dist.barrier()
if rank == 0:
    print("MULTI_OK", digest[:12])
'''


def test_world_size_2_gloo(tmp_path, built_lib):
    ck.build_checkers()
    ck.index_path("synth_200.mfur")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, FG_ROOT=ck.ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29731", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "MULTI_OK" in out.stdout


def test_shard_by_kmers_edge_cases():
    from fulgor_b200 import replicate

    assert replicate.shard_by_kmers(np.array([0], dtype=np.uint64), 31, 4) == [0, 0, 0, 0, 0]
    off = np.array([0, 10, 20, 30], dtype=np.uint64)  # all reads shorter than k
    cuts = replicate.shard_by_kmers(off, 31, 2)
    assert cuts[0] == 0 and cuts[-1] == 3
