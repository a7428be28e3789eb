"""The `.awfmi` v8 reader/writer against files written by the reference (src/AwFmFile.c:20-193)."""
import os

import numpy as np
import pytest

from avxwindowfmindex_b200 import read_awfmi, write_awfmi
from avxwindowfmindex_b200.index import sa_bit_width, sa_byte_length


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r3", "amino_r8", "amino_r1"])
def test_reader_matches_reference_memory(small_indexes, name):
    b = small_indexes[name]
    mine = read_awfmi(b.path)
    ref = b.arrays
    assert (mine.alphabet, mine.seed_k, mine.sa_ratio, mine.bwt_length) == (ref.alphabet, ref.seed_k, ref.sa_ratio, ref.bwt_length)
    assert np.array_equal(mine.blocks, ref.blocks)
    assert np.array_equal(mine.prefix_sums, ref.prefix_sums)
    assert np.array_equal(mine.seed_table, ref.seed_table)
    assert np.array_equal(mine.sa_bytes, ref.sa_bytes)
    assert mine.blocks.ctypes.data % 32 == 0


@pytest.mark.parametrize("name", ["nuc_r8", "amino_r8"])
def test_writer_round_trip_is_byte_identical(small_indexes, tmp_path, name):
    b = small_indexes[name]
    out = str(tmp_path / "copy.awfmi")
    write_awfmi(read_awfmi(b.path), out)
    assert open(out, "rb").read() == open(b.path, "rb").read()


def test_reference_reads_our_file(small_indexes, reference, tmp_path):
    """awFmReadIndexFromFile accepts a file written by write_awfmi and searches it identically."""
    b = small_indexes["nuc_r8"]
    out = str(tmp_path / "ours.awfmi")
    write_awfmi(b.arrays, out)
    ptr = reference.read_index(out, keep_sa=True)
    q = np.frombuffer(bytes(b.text[100:100 + 4000]), np.uint8)
    assert np.array_equal(reference.count(ptr, q, fixed_len=10), reference.count(b.ptr, q, fixed_len=10))
    reference.dealloc_index(ptr)


def test_sa_width_and_size_formulas():
    assert sa_bit_width(3_100_000_001) == 32 and sa_bit_width(1_000_000_001) == 30 and sa_bit_width(1_000_001) == 20
    assert sa_byte_length(1_000_001, 8) == (125_001 * 20 + 7) // 8 + 8
