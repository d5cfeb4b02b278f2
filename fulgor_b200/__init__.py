"""fulgor_b200 -- B200 (sm_100a) implementation of Fulgor's pseudoalignment hot path.

The product is the C-ABI library `libfulgor_gpu.so` (include/fulgor_gpu.h; CUDA sources under
fulgor_b200/csrc/). This package is the thin Python mirror of the reference's query interface used by
the tests and bench.py."""
from .index import (FULL_INTERSECTION, THRESHOLD_UNION, FulgorGpuError, Index, PinnedBuffer, bind_host_thread, build_image, image_info, lib, pack_reads,
                    unpack_bitmaps)  # noqa: F401

__all__ = ["FULL_INTERSECTION", "THRESHOLD_UNION", "FulgorGpuError", "Index", "PinnedBuffer", "bind_host_thread", "build_image", "image_info", "lib", "pack_reads", "unpack_bitmaps"]
