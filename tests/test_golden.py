"""Golden fixtures produced by the reference itself (tests/golden/make_golden.py): the oracle must reproduce them
on CPU without oracle/_ref, the CUDA path must reproduce them through the C-ABI on the GPU."""
import glob
import os

import numpy as np
import pytest

from avxwindowfmindex_b200 import read_awfmi
from oracle import harness

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load(case):
    return read_awfmi(os.path.join(GOLDEN, case + ".awfmi")), np.load(os.path.join(GOLDEN, case + ".npz"))


def test_fixtures_present():
    assert set(CASES) >= {"nuc_k4_r4", "amino_k2_r3", "nuc_k3_r1", "nuc_k3_r251", "four_records"}


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_golden(case):
    arrays, g = load(case)
    oracle = harness.Oracle(arrays)
    counts, _, _ = oracle.count(g["letters"], g["offsets"])
    assert np.array_equal(counts, g["counts"])
    hit_offsets, positions, _ = oracle.locate(g["letters"], g["offsets"])
    assert np.array_equal(hit_offsets, g["hit_offsets"])
    assert np.array_equal(positions, g["positions"])


def test_known_answers_four_records():
    """test/multiSequenceIndexTest/AwFmMultiSequenceTest.c:627-754: acdef -> contig 0 offset 0, g -> 1/0,
    hikl -> 2/0, m -> 3/0; fg, gh, lm span record boundaries -> no hit; header letters t, v, w, y absent."""
    arrays, g = load("four_records")
    counts = dict(zip([b"acdef", b"g", b"hikl", b"m", b"fg", b"gh", b"lm", b"t", b"v", b"w", b"y"], g["counts"]))
    assert [counts[q] for q in (b"acdef", b"g", b"hikl", b"m")] == [1, 1, 1, 1]
    assert all(counts[q] == 0 for q in (b"fg", b"gh", b"lm", b"t", b"v", b"w", b"y"))
    assert g["contig_of_hit"][:4].tolist() == [[0, 0], [1, 0], [2, 0], [3, 0]]
    # the oracle's contig mapping (FastaVector.c:338-381 semantics) on the metadata stored in the file
    oracle = harness.Oracle(arrays)
    import ctypes as C
    ends = np.ascontiguousarray(arrays.fasta_metadata[:, 1])
    for p, (s, l) in zip(g["positions"], g["contig_of_hit"]):
        si, li = C.c_uint64(), C.c_uint64()
        assert oracle.lib.awfm_oracle_contig_of(ends.ctypes.data, len(ends), int(p), C.byref(si), C.byref(li)) == 0
        assert (si.value, li.value) == (int(s), int(l))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_reproduces_golden(case):
    from avxwindowfmindex_b200 import GpuIndex
    arrays, g = load(case)
    gpu = GpuIndex(arrays)
    for lpq in (8, 4, 2, 1):
        for variant in (0, 1):
            gpu.set_tuning(count_lpq=lpq, locate_lpq=lpq, count_variant=variant, locate_variant=variant)
            counts = gpu.count(g["letters"], g["offsets"])
            assert np.array_equal(counts, g["counts"]), (case, lpq, variant)
            hit_offsets, positions = gpu.locate(g["letters"], g["offsets"])
            assert np.array_equal(hit_offsets, g["hit_offsets"]), (case, lpq, variant)
            assert np.array_equal(positions, g["positions"]), (case, lpq, variant)
    if "contig_of_hit" in g.files:  # the FASTA case: every hit mapped to (record, offset) as the reference does
        seq, loc, bad = gpu.map_positions(g["positions"])
        assert bad == 0 and np.array_equal(np.stack([seq, loc], axis=1), g["contig_of_hit"])
    gpu.close()
