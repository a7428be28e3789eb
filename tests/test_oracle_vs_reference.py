"""Pins oracle/awfm_oracle.c against the UNMODIFIED reference (oracle/_ref/libawfm_ref.so), function by function
and end to end.  CPU only.  Reference tests this mirrors: test/parallelSearch, test/inMemorySaTest,
test/backtraceTest, test/suffixArrayCompressionTests, test/occurrenceTests."""
import ctypes as C

import numpy as np
import pytest

from avxwindowfmindex_b200 import abi
from avxwindowfmindex_b200.index import aligned_empty
from avxwindowfmindex_b200.search import pack_queries
from oracle import harness
from conftest import make_queries


def brute_force_positions(text, query, amino):
    """All text positions where query matches after the reference's sanitisation (case fold for nucleotides,
    ambiguity letters -> one class)."""
    def norm(b):
        a = np.frombuffer(bytes(b), dtype=np.uint8)
        return np.array([harness_letter(amino, c) for c in a], dtype=np.uint8)
    t, q = norm(text), norm(query)
    if len(q) == 0 or len(q) > len(t):
        return np.zeros(0, np.uint64)
    windows = np.lib.stride_tricks.sliding_window_view(t, len(q))
    return np.nonzero((windows == q).all(axis=1))[0].astype(np.uint64)


_AMINO_TABLE = [20, 0, 20, 1, 2, 3, 4, 5, 6, 7, 20, 8, 9, 10, 11, 20, 12, 13, 14, 15, 16, 20, 17, 18, 20, 19, 20, 20, 20, 20, 20, 20]


def harness_letter(amino, c):
    if amino:
        return 21 if c == ord("$") else _AMINO_TABLE[c & 31]
    c |= 0x20
    return {ord("a"): 0, ord("c"): 1, ord("g"): 2, ord("t"): 3, ord("u"): 3, ord("$"): 5}.get(c, 4)


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r1", "nuc_r3", "nuc_r16", "nuc_r200", "nuc_r255", "amino_r8",
                                  "amino_r2", "amino_r1"])
def test_count_and_locate_match_reference(small_indexes, reference, name):
    b = small_indexes[name]
    k = b.arrays.seed_k
    queries = make_queries(b.text, b.amino, seed=11, num=600, min_len=1, max_len=max(12, k + 9), seed_k=k)
    letters, offsets = pack_queries(queries)
    oracle = harness.Oracle(b.arrays)
    for threads in (1, 4):
        ref_counts = reference.count(b.ptr, letters, offsets, threads=threads)
        o_counts, o_ranges, work = oracle.count(letters, offsets, threads=threads)
        assert np.array_equal(ref_counts, o_counts)
    rc, ref_counts2, ref_pos = reference.locate(b.ptr, letters, offsets, threads=3)
    assert rc == abi.AwFmSuccess
    hit_offsets, positions, work = oracle.locate(letters, offsets, threads=2)
    assert np.array_equal(np.diff(hit_offsets).astype(np.uint32), ref_counts2)
    for i, p in enumerate(ref_pos):  # element-wise, SA order (not sorted)
        assert np.array_equal(p, positions[int(hit_offsets[i]):int(hit_offsets[i + 1])]), (name, i, queries[i])
    # independent second check, as the reference's own tests do: brute force over the text
    for i in range(0, len(queries), 7):
        bf = brute_force_positions(b.text, queries[i], b.amino)
        assert np.array_equal(np.sort(ref_pos[i]), bf), (name, queries[i])
    assert work["hits"] == int(hit_offsets[-1])


@pytest.mark.parametrize("name", ["nuc_r8", "amino_r8"])
def test_ranges_match_reference_primitives(small_indexes, reference, name):
    """The batched API does not expose ranges, so they are rebuilt from the reference's exported primitives in the
    order parallelSearchFindKmerSeedsForBlock / ExtendKmersInBlock apply them (src/AwFmParallelSearch.c:222-313)."""
    b = small_indexes[name]
    lib = reference.lib
    amino = b.amino
    k = b.arrays.seed_k
    seed_fn = lib.awFmAminoKmerSeedRangeFromTable if amino else lib.awFmNucleotideKmerSeedRangeFromTable
    seed_fn.restype = abi.AwFmSearchRange
    seed_fn.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    nonseeded = lib.awFmAminoNonSeededSearch if amino else lib.awFmNucleotideNonSeededSearch
    nonseeded.restype = None
    nonseeded.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.POINTER(abi.AwFmSearchRange)]
    step = lib.awFmAminoIterativeStepBackwardSearch if amino else lib.awFmNucleotideIterativeStepBackwardSearch
    step.restype = None
    step.argtypes = [C.c_void_p, C.POINTER(abi.AwFmSearchRange), C.c_uint8]
    can_use = lib.awFmQueryCanUseKmerTable
    can_use.restype = C.c_bool
    can_use.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    to_index = lib.awFmAsciiAminoAcidToLetterIndex if amino else lib.awFmAsciiNucleotideToLetterIndex
    to_index.restype = C.c_uint8
    to_index.argtypes = [C.c_uint8]

    queries = make_queries(b.text, amino, seed=5, num=400, min_len=1, max_len=k + 8, seed_k=k)
    letters, offsets = pack_queries(queries)
    _, o_ranges, _ = harness.Oracle(b.arrays).count(letters, offsets)
    for i, q in enumerate(queries):
        rng = abi.AwFmSearchRange()
        n = len(q)
        if can_use(b.ptr, q, n):
            rng = seed_fn(b.ptr, q, n)
        else:
            start = 0 if n < k else n - k
            nonseeded(b.ptr, q[start:], min(n, k), C.byref(rng))
        j = k
        while True:
            j += 1
            if not (n >= j and rng.startPtr <= rng.endPtr):
                break
            step(b.ptr, C.byref(rng), to_index(q[n - j]))
        assert (rng.startPtr, rng.endPtr) == (int(o_ranges[i, 0]), int(o_ranges[i, 1])), (name, q)


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r3", "amino_r8", "amino_r1"])
def test_backtrace_and_letter_at_every_position(small_indexes, reference, name):
    b = small_indexes[name]
    fn = reference.lib.awFmAminoBacktraceBwtPosition if b.amino else reference.lib.awFmNucleotideBacktraceBwtPosition
    oracle = harness.Oracle(b.arrays)
    for p in range(b.arrays.bwt_length):
        assert fn(b.ptr, p) == oracle.lib.awfm_oracle_backtrace_step(oracle.ixp, p), p


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r1", "nuc_r200", "amino_r2"])
def test_sampled_sa_values(small_indexes, reference, name):
    b = small_indexes[name]
    oracle = harness.Oracle(b.arrays)
    sa_struct = C.addressof(reference.struct(b.ptr)) + abi.AwFmIndex.suffixArray.offset
    n = (b.arrays.bwt_length + b.arrays.sa_ratio - 1) // b.arrays.sa_ratio
    expect = b.arrays.sa_values()
    for j in range(n):
        r = reference.lib.awFmGetValueFromCompressedSuffixArray(sa_struct, j)
        assert r == oracle.lib.awfm_oracle_sa_value(oracle.ixp, j) == int(expect[j])


@pytest.mark.parametrize("amino", [False, True])
def test_rank_selectors_on_garbage_blocks(reference, amino):
    """The reference's letter selectors test only some code bits (src/AwFmOccurrence.c:18-35, 65-134); the oracle's
    (code, care) tables must agree on ARBITRARY block contents, for every letter and every masked position.  Driven
    through the reference's LF step on a fake index whose prefix sums and base occurrences are zero, so the step
    returns the masked popcounts themselves (src/AwFmSearch.c:42-159)."""
    rng = np.random.default_rng(3)
    bbytes = abi.AMINO_BLOCK_BYTES if amino else abi.NUC_BLOCK_BYTES
    nvec = 5 if amino else 3
    nblocks = 64
    blocks = aligned_empty(nblocks * bbytes)
    blocks[:] = 0
    view = blocks.reshape(nblocks, bbytes)
    view[:, : 32 * nvec] = rng.integers(0, 256, (nblocks, 32 * nvec), dtype=np.uint8)
    prefix = np.zeros(24, np.uint64)
    fake = abi.AwFmIndex()
    fake.bwtLength = nblocks * 256
    fake.bwtBlockList = blocks.ctypes.data
    fake.prefixSums = prefix.ctypes.data
    fake.config.alphabetType = abi.AwFmAlphabetAmino if amino else abi.AwFmAlphabetDna
    step = reference.lib.awFmAminoIterativeStepBackwardSearch if amino else reference.lib.awFmNucleotideIterativeStepBackwardSearch
    step.restype = None
    step.argtypes = [C.c_void_p, C.POINTER(abi.AwFmSearchRange), C.c_uint8]
    oracle_lib = harness.Oracle.__new__(harness.Oracle)
    lib = C.CDLL(harness.ORACLE_LIB)
    lib.awfm_oracle_block_popcount.restype = C.c_uint32
    lib.awfm_oracle_block_popcount.argtypes = [C.c_void_p, C.c_uint8, C.c_uint8, C.c_uint8]
    del oracle_lib
    alphabet = 1 if amino else 2
    for letter in range(21 if amino else 5):
        for blk in range(nblocks):
            for local in (0, 1, 31, 32, 63, 64, 100, 127, 128, 191, 192, 254, 255, int(rng.integers(0, 256))):
                p = blk * 256 + local
                r = abi.AwFmSearchRange(p + 1, p)  # sp-1 == ep == p
                step(C.addressof(fake), C.byref(r), letter)
                expect = lib.awfm_oracle_block_popcount(blocks.ctypes.data + blk * bbytes, alphabet, letter, local)
                assert r.startPtr == expect and r.endPtr == expect - 1 + (1 << 64) * (expect == 0), (letter, blk, local)


def test_letter_tables(reference):
    lib = C.CDLL(harness.ORACLE_LIB)
    lib.awfm_oracle_letter_index.restype = C.c_uint8
    lib.awfm_oracle_letter_index.argtypes = [C.c_uint8, C.c_uint8]
    lib.awfm_oracle_letter_is_ambiguous.argtypes = [C.c_uint8, C.c_uint8]
    r = reference.lib
    r.awFmAsciiNucleotideToLetterIndex.restype = C.c_uint8
    r.awFmAsciiNucleotideToLetterIndex.argtypes = [C.c_uint8]
    r.awFmAsciiAminoAcidToLetterIndex.restype = C.c_uint8
    r.awFmAsciiAminoAcidToLetterIndex.argtypes = [C.c_uint8]
    r.awFmLetterIsAmbiguous.restype = C.c_bool
    r.awFmLetterIsAmbiguous.argtypes = [C.c_char, C.c_int]
    for c in range(256):
        assert lib.awfm_oracle_letter_index(2, c) == r.awFmAsciiNucleotideToLetterIndex(c)
        assert lib.awfm_oracle_letter_index(1, c) == r.awFmAsciiAminoAcidToLetterIndex(c)
        if c < 128:  # tolower() on negative chars is locale/UB territory in the reference
            assert bool(lib.awfm_oracle_letter_is_ambiguous(2, c)) == r.awFmLetterIsAmbiguous(bytes([c]), abi.AwFmAlphabetDna)
            assert bool(lib.awfm_oracle_letter_is_ambiguous(1, c)) == r.awFmLetterIsAmbiguous(bytes([c]), abi.AwFmAlphabetAmino)
