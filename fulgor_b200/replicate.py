"""Multi-GPU replication of the index image: one process per GPU (torchrun), rank 0 parses + flattens the .fur once, ONE
broadcast of the bytes (NCCL over NVLink on GPUs; gloo in the CPU tests) replicates it, every rank adopts its copy. The
pseudoalignment path itself has no collective: reads shard trivially."""
import numpy as np

from . import index as _index


def broadcast_image(index_path, device, src=0):
    """returns a uint8 torch tensor on `device` holding the flattened image on every rank of the default process group"""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    size = torch.zeros(1, dtype=torch.int64, device=device)
    host = None
    if rank == src:
        host = _index.build_image(index_path)
        size[0] = host.size
    if world > 1:
        dist.broadcast(size, src)
    image = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    if rank == src:
        image.copy_(torch.from_numpy(host))
    if world > 1:
        dist.broadcast(image, src)
    return image


def open_replica(index_path, local_device):
    """Index handle on cuda:`local_device` backed by the broadcast image (the tensor is kept alive by the handle)"""
    import torch

    image = broadcast_image(index_path, torch.device("cuda", local_device))
    torch.cuda.synchronize(local_device)
    return _index.Index.adopt_device_image(image.data_ptr(), image.numel(), local_device, keepalive=image), image


def shard_range(n_total, rank, world):
    """contiguous, balanced [lo, hi) of a batch of n_total reads for this rank"""
    return n_total * rank // world, n_total * (rank + 1) // world


def shard_by_kmers(read_off, k, world):
    """cut points (world + 1) splitting a batch into contiguous ranges balanced by k-mer count (mixed-length reads)"""
    L = np.diff(np.asarray(read_off, dtype=np.int64))
    work = np.concatenate([[0], np.cumsum(np.maximum(L - k + 1, 0) + 8)])
    cuts = [int(np.searchsorted(work, work[-1] * g // world, side="left")) for g in range(world)] + [len(L)]
    return cuts
