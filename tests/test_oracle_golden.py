"""The oracle against the committed golden vectors (tests/golden/*.npz, produced by the unmodified reference via
tests/golden/make_golden.py). Runs everywhere, including boxes where /root/reference does not exist."""
import glob
import os

import numpy as np
import pytest

import _checkers as ck

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def load_case(path):
    g = np.load(path)
    reads = ck.gen_reads(int(g["n"]), int(g["min_len"]), int(g["max_len"]), seed=int(g["seed"]), genomes=str(g["index"]).split(".")[0])
    assert int(reads[0].astype(np.uint64).sum()) == int(g["reads_checksum"]), "read generator drifted from the golden inputs"
    return g, reads


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(path):
    g, reads = load_case(path)
    o = ck.Oracle(ck.index_path(str(g["index"])))
    off, vals = o.fetch_color_set_ids(reads)
    assert np.array_equal(off, g["cid_off"]) and np.array_equal(vals, g["cids"])
    off, vals = o.pseudoalign(reads, 0)
    assert np.array_equal(off, g["fi_off"]) and np.array_equal(vals, g["fi"])
    for j, t in enumerate(g["thresholds"]):
        off, vals = o.pseudoalign(reads, 1, float(t))
        assert np.array_equal(off, g[f"tu{j}_off"]) and np.array_equal(vals, g[f"tu{j}"])
    o.close()


def test_golden_present():
    assert len(GOLDEN) >= 4


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_gpu_matches_golden(path, built_lib):
    """the CUDA path against the reference's own outputs (no oracle in between)"""
    import fulgor_b200 as fg

    g, reads = load_case(path)
    with fg.Index.open(ck.index_path(str(g["index"])), 0) as gpu:
        off, vals = gpu.fetch_color_set_ids(reads)
        assert np.array_equal(off, g["cid_off"]) and np.array_equal(vals, g["cids"])
        off, vals = gpu.pseudoalign(reads, 0)
        assert np.array_equal(off, g["fi_off"]) and np.array_equal(vals, g["fi"])
        for j, t in enumerate(g["thresholds"]):
            off, vals = gpu.pseudoalign(reads, 1, float(t))
            assert np.array_equal(off, g[f"tu{j}_off"]) and np.array_equal(vals, g[f"tu{j}"])
