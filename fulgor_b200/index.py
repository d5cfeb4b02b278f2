"""Host-side mirror of the reference's `index<ColorSets>` query interface (reference include/index.hpp:39-46)
over the C ABI of libfulgor_gpu.so (include/fulgor_gpu.h). Same method names and argument meaning as the
reference; batched, because one GPU call serves many reads:

    fetch_color_set_ids            <- index::fetch_color_set_ids            src/ps_full_intersection.cpp:335-374
    pseudoalign_full_intersection  <- fetch + index::pseudoalign_full_intersection  src/ps_full_intersection.cpp:377-400
    pseudoalign_threshold_union    <- index::pseudoalign_threshold_union    src/ps_threshold_union.cpp:321-402

Reads are passed as (bases uint8[total], read_off uint64[n+1]); results come back in CSR form
(off uint64[n+1], values uint32[off[n]]). There is no CPU fallback: without the CUDA library or
without a GPU every compute call raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FULGOR_GPU_LIB") or os.path.join(_HERE, "libfulgor_gpu.so")  # the override is for kernel-variant A/B runs

FULL_INTERSECTION = 0
THRESHOLD_UNION = 1
E2BIG = -7
ENODEV = -19

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


class FulgorGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfulgor_gpu error {code}: {msg}")
        self.code = code


class Info(C.Structure):
    _fields_ = [("k", C.c_uint32), ("m", C.c_uint32), ("num_kmers", C.c_uint64), ("num_unitigs", C.c_uint64),
                ("num_color_sets", C.c_uint64), ("num_colors", C.c_uint32), ("type", C.c_uint32),
                ("image_bytes", C.c_uint64), ("device", C.c_int32), ("pad", C.c_uint32)]


_lib = None


def lib():
    """The C-ABI library; raises (loudly) when it has not been built: there is no Python/CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m fulgor_b200.build` (needs nvcc); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.fulgor_gpu_last_error.restype = C.c_char_p
        L.fulgor_gpu_version.restype = C.c_char_p
        L.fulgor_gpu_device_count.restype = C.c_int
        L.fulgor_gpu_host_alloc.restype = C.c_void_p
        L.fulgor_gpu_host_alloc.argtypes = [C.c_uint64]
        L.fulgor_gpu_host_free.argtypes = [C.c_void_p]
        L.fulgor_gpu_bind_host_thread.argtypes = [C.c_int]
        L.fulgor_gpu_image_build.argtypes = [C.c_char_p, C.POINTER(_u8p), _u64p]
        L.fulgor_gpu_image_free.argtypes = [_u8p]
        L.fulgor_gpu_image_info.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Info)]
        L.fulgor_gpu_index_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        L.fulgor_gpu_index_open_image.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
        L.fulgor_gpu_index_adopt_device_image.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
        L.fulgor_gpu_index_close.argtypes = [C.c_void_p]
        L.fulgor_gpu_index_close.restype = None
        L.fulgor_gpu_index_info.argtypes = [C.c_void_p, C.POINTER(Info)]
        L.fulgor_gpu_fetch_color_set_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.fulgor_gpu_pseudoalign.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64]
        L.fulgor_gpu_pseudoalign_dedup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.fulgor_gpu_kmer_conservation.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64]
        L.fulgor_gpu_kmer_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.fulgor_gpu_pseudoalign_device.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64,
                                                    C.c_void_p, C.c_void_p, C.c_uint64, _u64p]
        L.fulgor_gpu_last_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_float * 3)]
        L.fulgor_gpu_pack_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, _u64p, _u64p, C.c_int]
        L.fulgor_gpu_pseudoalign_packed.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                                    C.c_void_p, C.c_void_p, C.c_uint64]
        L.fulgor_gpu_pseudoalign_bitmaps.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.fulgor_gpu_pseudoalign_packed_bitmaps.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                                            C.c_void_p]
        L.fulgor_gpu_pseudoalign_packed_device.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                                           C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, _u64p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise FulgorGpuError(rc, lib().fulgor_gpu_last_error().decode(errors="replace"))


def build_image(index_path):
    """Reference-built .fur/.mfur -> flattened device image (numpy uint8, host). No GPU needed.
    Replaces essentials::load(index, path) (reference tools/pseudoalign.cpp:340)."""
    L = lib()
    p = _u8p()
    n = C.c_uint64(0)
    _check(L.fulgor_gpu_image_build(os.fsencode(index_path), C.byref(p), C.byref(n)))
    try:
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()
    finally:
        L.fulgor_gpu_image_free(p)


def image_info(image):
    info = Info()
    _check(lib().fulgor_gpu_image_info(image.ctypes.data, image.size, C.byref(info)))
    return info


def pack_reads(reads, threads=0):
    """(bases, read_off) -> (words uint32, read_len uint32 with FULGOR_GPU_READ_HAS_INVALID flags, invalid_pos uint64): the packed
    read form of include/fulgor_gpu.h, produced on host threads (no GPU needed)"""
    bases, off = reads
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = off.size - 1
    L = lib()
    lens = np.zeros(max(1, n), dtype=np.uint32)
    nw, ni = C.c_uint64(0), C.c_uint64(0)
    wcap = int((int(off[n]) - int(off[0])) // 16 + n + 1) if n else 1
    words = np.zeros(wcap, dtype=np.uint32)
    inv = np.zeros(64, dtype=np.uint64)
    while True:
        rc = L.fulgor_gpu_pack_reads(bases.ctypes.data, off.ctypes.data, n, words.ctypes.data, words.size, lens.ctypes.data, inv.ctypes.data, inv.size,
                                     C.byref(nw), C.byref(ni), int(threads))
        if rc == E2BIG:
            if nw.value > words.size:
                words = np.zeros(nw.value, dtype=np.uint32)
            if ni.value > inv.size:
                inv = np.zeros(ni.value, dtype=np.uint64)
            continue
        _check(rc)
        return words[: nw.value], lens[:n], inv[: ni.value]


def unpack_bitmaps(bitmaps, num_colors):
    """bitmap rows (n x ceil(num_colors / 32) uint32) -> CSR (off uint64[n+1], colors uint32): the lists fulgor_gpu_pseudoalign returns"""
    rows = np.ascontiguousarray(bitmaps, dtype=np.uint32)
    bits = np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little")[:, :num_colors]
    off = np.zeros(rows.shape[0] + 1, dtype=np.uint64)
    off[1:] = np.cumsum(bits.sum(axis=1, dtype=np.uint64))
    return off, np.nonzero(bits)[1].astype(np.uint32)


def bind_host_thread(device):
    """CPU affinity of the calling thread -> the CPUs next to `device` (NUMA-local pinned buffers); returns the CPUs bound or 0"""
    return int(lib().fulgor_gpu_bind_host_thread(int(device)))


class PinnedBuffer:
    """Pinned host memory from the library (cudaHostAlloc) exposed as a numpy array."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = lib().fulgor_gpu_host_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError(f"cudaHostAlloc({self.nbytes}) failed")
        self._raw = (C.c_uint8 * max(1, self.nbytes)).from_address(self.ptr)

    def view(self, dtype, count=None, offset=0):
        a = np.frombuffer(self._raw, dtype=np.uint8, count=self.nbytes)[offset:]
        a = a[: (a.size // np.dtype(dtype).itemsize) * np.dtype(dtype).itemsize].view(dtype)
        return a if count is None else a[:count]

    def free(self):
        if self.ptr:
            self._raw = None
            lib().fulgor_gpu_host_free(self.ptr)
            self.ptr = None


class Index:
    """A Fulgor index resident on one GPU."""

    def __init__(self, handle):
        self._h = handle
        info = Info()
        _check(lib().fulgor_gpu_index_info(self._h, C.byref(info)))
        self.info = info
        self.k, self.m = info.k, info.m
        self.num_colors = info.num_colors
        self.num_color_sets = info.num_color_sets
        self.num_unitigs = info.num_unitigs
        self.num_kmers = info.num_kmers
        self.type = info.type
        self.device = info.device
        self._keepalive = None

    # -- constructors
    @classmethod
    def open(cls, index_path, device=0):
        h = C.c_void_p()
        _check(lib().fulgor_gpu_index_open(os.fsencode(index_path), device, C.byref(h)))
        return cls(h)

    @classmethod
    def from_image(cls, image, device=0):
        h = C.c_void_p()
        _check(lib().fulgor_gpu_index_open_image(image.ctypes.data, image.size, device, C.byref(h)))
        return cls(h)

    @classmethod
    def adopt_device_image(cls, device_ptr, nbytes, device, keepalive=None):
        """Wrap an image already in device memory (e.g. a torch uint8 CUDA tensor filled by an NCCL broadcast)."""
        h = C.c_void_p()
        _check(lib().fulgor_gpu_index_adopt_device_image(device_ptr, nbytes, device, C.byref(h)))
        x = cls(h)
        x._keepalive = keepalive
        return x

    def close(self):
        if self._h:
            lib().fulgor_gpu_index_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers
    @staticmethod
    def _reads(reads):
        bases, off = reads
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        return bases, off, off.size - 1

    def _csr_call(self, fn, n, cap):
        """call fn(off, vals, cap) -> rc, retrying once with the exact capacity on E2BIG"""
        off = np.zeros(n + 1, dtype=np.uint64)
        while True:
            vals = np.empty(max(1, cap), dtype=np.uint32)
            rc = fn(off, vals, cap)
            if rc == E2BIG:
                cap = int(off[n])
                continue
            _check(rc)
            return off, vals[: int(off[n])]

    # -- the reference interface, batched
    def fetch_color_set_ids(self, reads, want_positive=False, cap=None):
        bases, off, n = self._reads(reads)
        npos = np.zeros(n, dtype=np.uint32) if want_positive else None
        L = lib()

        def fn(o, v, c):
            return L.fulgor_gpu_fetch_color_set_ids(self._h, bases.ctypes.data, off.ctypes.data, n, o.ctypes.data, v.ctypes.data, c,
                                                    npos.ctypes.data if want_positive else None)

        res = self._csr_call(fn, n, cap if cap is not None else 4 * n + 64)
        return res + (npos,) if want_positive else res

    def pseudoalign(self, reads, algo=FULL_INTERSECTION, threshold=1.0, cap=None):
        bases, off, n = self._reads(reads)
        L = lib()

        def fn(o, v, c):
            return L.fulgor_gpu_pseudoalign(self._h, algo, float(threshold), bases.ctypes.data, off.ctypes.data, n, o.ctypes.data, v.ctypes.data, c)

        return self._csr_call(fn, n, cap if cap is not None else 8 * n + 64)

    def pseudoalign_dedup(self, reads, cap=None):
        """full intersection computed once per distinct color-set-id list (the reference's --deduplicate):
        returns (rep_of_read, color_off, colors); colors of read i = colors[color_off[rep[i]] : color_off[rep[i] + 1]]"""
        bases, off, n = self._reads(reads)
        rep = np.zeros(n, dtype=np.uint32)
        L = lib()

        def fn(o, v, c):
            return L.fulgor_gpu_pseudoalign_dedup(self._h, bases.ctypes.data, off.ctypes.data, n, rep.ctypes.data, o.ctypes.data, v.ctypes.data, c)

        o, v = self._csr_call(fn, n, cap if cap is not None else 8 * n + 64)
        return rep, o, v

    def kmer_conservation(self, reads, cap=None):
        """index::kmer_conservation for a batch: (triple_off[n+1], triples[t, 3] = start_pos_in_query, num_kmers, color_set_id)"""
        bases, off, n = self._reads(reads)
        L = lib()
        toff = np.zeros(n + 1, dtype=np.uint64)
        cap = cap if cap is not None else 4 * n + 64
        while True:
            tr = np.empty(3 * max(1, cap), dtype=np.uint32)
            rc = L.fulgor_gpu_kmer_conservation(self._h, bases.ctypes.data, off.ctypes.data, n, toff.ctypes.data, tr.ctypes.data, cap)
            if rc == E2BIG:
                cap = int(toff[n])
                continue
            _check(rc)
            return toff, tr[: 3 * int(toff[n])].reshape(-1, 3)

    def kmer_matches(self, reads):
        """index::kmer_matches for a batch: (word_off[n+1], positive_words uint32, counts[n, num_colors])"""
        bases, off, n = self._reads(reads)
        L = lib()
        woff = np.zeros(n + 1, dtype=np.uint64)
        counts = np.zeros((n, self.num_colors), dtype=np.uint32)
        cap = int(off[n] - off[0]) // 32 + n + 1
        words = np.zeros(max(1, cap), dtype=np.uint32)
        _check(L.fulgor_gpu_kmer_matches(self._h, bases.ctypes.data, off.ctypes.data, n, woff.ctypes.data, words.ctypes.data, cap, counts.ctypes.data))
        return woff, words[: int(woff[n])], counts

    def pseudoalign_packed(self, packed, algo=FULL_INTERSECTION, threshold=1.0, cap=None):
        """fulgor_gpu_pseudoalign_packed: packed = (words, read_len, invalid_pos) from pack_reads"""
        words, lens, inv = packed
        n = lens.size
        L = lib()

        def fn(o, v, c):
            return L.fulgor_gpu_pseudoalign_packed(self._h, algo, float(threshold), words.ctypes.data, lens.ctypes.data, n, inv.ctypes.data if inv.size else None,
                                                   inv.size, o.ctypes.data, v.ctypes.data, c)

        return self._csr_call(fn, n, cap if cap is not None else 8 * n + 64)

    def pseudoalign_bitmaps(self, reads, algo=FULL_INTERSECTION, threshold=1.0, packed=False):
        """bitmap rows instead of lists: uint32[n, ceil(num_colors / 32)]; reads = (bases, read_off), or pack_reads' triple with packed=True"""
        wpr = (self.num_colors + 31) // 32
        L = lib()
        if packed:
            words, lens, inv = reads
            n = lens.size
            out = np.zeros((n, wpr), dtype=np.uint32)
            _check(L.fulgor_gpu_pseudoalign_packed_bitmaps(self._h, algo, float(threshold), words.ctypes.data, lens.ctypes.data, n,
                                                           inv.ctypes.data if inv.size else None, inv.size, out.ctypes.data))
            return out
        bases, off, n = self._reads(reads)
        out = np.zeros((n, wpr), dtype=np.uint32)
        _check(L.fulgor_gpu_pseudoalign_bitmaps(self._h, algo, float(threshold), bases.ctypes.data, off.ctypes.data, n, out.ctypes.data))
        return out

    def pseudoalign_full_intersection(self, reads, cap=None):
        return self.pseudoalign(reads, FULL_INTERSECTION, 1.0, cap)

    def pseudoalign_threshold_union(self, reads, threshold, cap=None):
        return self.pseudoalign(reads, THRESHOLD_UNION, threshold, cap)

    # -- raw pointer forms (bench.py: pinned host buffers / device-resident inputs)
    def pseudoalign_raw(self, algo, threshold, bases_ptr, read_off_ptr, n, color_off_ptr, colors_ptr, cap):
        return lib().fulgor_gpu_pseudoalign(self._h, algo, float(threshold), bases_ptr, read_off_ptr, n, color_off_ptr, colors_ptr, cap)

    def pseudoalign_device(self, algo, threshold, d_bases, d_read_off, n, read_off_base, d_color_off, d_colors, cap):
        total = C.c_uint64(0)
        rc = lib().fulgor_gpu_pseudoalign_device(self._h, algo, float(threshold), d_bases, d_read_off, n, read_off_base, d_color_off, d_colors, cap,
                                                 C.byref(total))
        _check(rc)
        return total.value

    def last_kernel_times(self):
        ms = (C.c_float * 3)()
        launches = lib().fulgor_gpu_last_kernel_times(self._h, C.byref(ms))
        return launches, [float(v) for v in ms]
