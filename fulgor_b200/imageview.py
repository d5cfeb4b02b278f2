"""Read-only numpy views into a flattened index image (layout: fulgor_b200/csrc/image.h). Used by bench.py to
compute the ALGORITHMIC bytes of a batch (compressed size of every hit color set) and by tests."""
import ctypes as C

import numpy as np

FGI_MAX_SKEW = 16


class Header(C.Structure):
    _fields_ = [
        ("magic", C.c_uint64), ("total_bytes", C.c_uint64), ("k", C.c_uint32), ("m", C.c_uint32), ("num_kmers", C.c_uint64),
        ("hash_magic", C.c_uint64), ("bucketer_T", C.c_uint64), ("num_minimizers", C.c_uint64), ("num_super_kmers", C.c_uint64),
        ("num_unitigs", C.c_uint64), ("num_string_words", C.c_uint64),
        ("skew_min_log2", C.c_uint32), ("skew_max_log2", C.c_uint32), ("skew_log2_max_bucket", C.c_uint32), ("num_skew", C.c_uint32),
        ("skew_phf", C.c_uint32 * FGI_MAX_SKEW), ("skew_pos_base", C.c_uint64 * FGI_MAX_SKEW),
        ("type", C.c_uint32), ("num_colors", C.c_uint32), ("num_color_sets", C.c_uint64),
        ("num_partitions", C.c_uint32), ("num_phfs", C.c_uint32), ("num_phf_parts", C.c_uint32), ("pad0", C.c_uint32),
        ("off_phfs", C.c_uint64), ("off_phf_parts", C.c_uint64), ("off_hashed_pilots", C.c_uint64), ("off_free_slots", C.c_uint64),
        ("off_bucket_begin", C.c_uint64), ("off_sk_records", C.c_uint64), ("off_strings", C.c_uint64), ("off_skew_positions", C.c_uint64),
        ("off_hybrids", C.c_uint64), ("off_set_bit_off", C.c_uint64), ("off_color_words", C.c_uint64), ("off_meta_off", C.c_uint64),
        ("off_meta_vals", C.c_uint64), ("off_part_min_color", C.c_uint64), ("off_part_sets_before", C.c_uint64),
        ("off_sk_cid", C.c_uint64), ("num_unpinned", C.c_uint64), ("guard_max_hash", C.c_uint64), ("reserved", C.c_uint64 * 5),
    ]


HYBRID_DTYPE = np.dtype([("num_colors", "<u4"), ("sparse_thr", "<u4"), ("very_dense_thr", "<u4"), ("kind", "<u4"),
                         ("num_sets", "<u8"), ("set_off_base", "<u8"), ("word_base", "<u8")])


def header(image):
    h = Header.from_buffer_copy(image[: C.sizeof(Header)].tobytes())
    assert h.magic == 0x3330474D49475546 and h.total_bytes == image.size, "not a fulgor-b200 image"
    return h


def section(image, off, dtype, count):
    return np.frombuffer(image, dtype=dtype, count=count, offset=off)


def hybrids(image):
    h = header(image)
    return section(image, h.off_hybrids, HYBRID_DTYPE, h.num_partitions)


def bucket_sizes(image):
    h = header(image)
    return np.diff(section(image, h.off_bucket_begin, "<u4", h.num_minimizers + 1).astype(np.int64))


def _container_set_bits(image, h, hy):
    """compressed bits one query of each set of a container reads: the set itself (hybrid), or its difference list plus the
    representative it is coded against (differential; the representative of a cluster sits right before the cluster's first set)"""
    n = int(hy["num_sets"])
    base = h.off_set_bit_off + 8 * int(hy["set_off_base"])
    off = section(image, base, "<u8", n + 1).astype(np.int64)
    bits = np.diff(off)
    if int(hy["kind"]) == 1:
        rep = section(image, base + 8 * (n + 1), "<u8", n).astype(np.int64)
        first = np.ones(n, dtype=bool)
        first[1:] = rep[1:] != rep[:-1]
        rep_end = off[:-1][first][np.cumsum(first) - 1]  # offset of the first set of each set's cluster
        bits = bits + (rep_end - rep)
    return bits


def color_set_bits(image):
    """compressed size in bits of every color set: m_offsets[c+1] - m_offsets[c] for a hybrid index (reference
    include/color_sets/hybrid.hpp:309-313); for a meta index the bits of the partial sets it visits plus 64 per meta list
    (SURVEY.md 8(d)); differential containers add the representative's bits."""
    h = header(image)
    hy = hybrids(image)
    if h.type in (0, 2):
        return _container_set_bits(image, h, hy[0])
    meta_off = section(image, h.off_meta_off, "<u8", h.num_color_sets + 1).astype(np.int64)
    meta_vals = section(image, h.off_meta_vals, "<u4", int(meta_off[-1]))
    before = section(image, h.off_part_sets_before, "<u4", h.num_partitions + 1).astype(np.int64)
    partial_bits = np.concatenate([_container_set_bits(image, h, hy[p]) for p in range(h.num_partitions)])
    # indexed by meta color (partitions are consecutive ranges of meta colors)
    out = np.zeros(h.num_color_sets, dtype=np.int64)
    for c in range(h.num_color_sets):
        b = int(meta_off[c])
        n = int(meta_vals[b])
        out[c] = 64 + int(partial_bits[meta_vals[b + 1: b + 1 + n]].sum())
    assert before[-1] == partial_bits.size
    return out
