"""Index contents the kernels never see from a well-formed index, and array shapes the small fixtures do not reach:
  * ARBITRARY letter bit-vectors (codes the index builder never emits: 0b111, 0b000 inside the text, the 11 unused
    amino codes) under an arbitrary seed table: the device's (code, care) selectors must rank exactly like the
    reference's boolean forms (src/AwFmOccurrence.c:18-35, 65-134).  The oracle is pinned against the reference on the
    same kind of blocks (tests/test_oracle_vs_reference.py::test_rank_selectors_on_garbage_blocks);
  * sampled-SA fields wider than 32 bits (an index above 4 Gbp has w = 33..; w > 57 needs a ninth byte in the
    reference, src/AwFmSuffixArray.c:131-139): the same small index re-packed at w = 33, 40, 57, 58, 64.
Run with -m gpu on a B200."""
import numpy as np
import pytest

from avxwindowfmindex_b200 import GpuGroup, GpuIndex, IndexArrays, abi, pack_queries_bits
from avxwindowfmindex_b200.search import QUERY_2BIT
from avxwindowfmindex_b200.index import aligned_empty, sa_num_samples
from oracle import harness

pytestmark = pytest.mark.gpu

NUC_CODE_CARE = [(6, 6), (5, 5), (3, 3), (1, 7), (2, 7)]  # A C G T X: match when (stored ^ code) & care == 0
AMINO_CODE_CARE = [(c & 0xFF, c >> 8) for c in
                   (0x1C0C, 0x0F17, 0x1303, 0x1606, 0x0F1E, 0x151A, 0x0F1B, 0x1619, 0x1A15, 0x131C, 0x0F1D,
                    0x0F08, 0x1909, 0x0F04, 0x1C13, 0x1A0A, 0x1505, 0x1916, 0x0F01, 0x0F02, 0x0F1F)]


def garbage_index(amino, num_blocks, seed_k, rng):
    """Random letter bit-vectors; base occurrences CONSISTENT with them under the reference's selectors (the device
    layout stores them as differences, which only exist for consistent counts); prefix sums zero; a seed table of
    arbitrary in-range positions."""
    nvec = 5 if amino else 3
    table = AMINO_CODE_CARE if amino else NUC_CODE_CARE
    bbytes = abi.AMINO_BLOCK_BYTES if amino else abi.NUC_BLOCK_BYTES
    n = num_blocks * 256
    bits = rng.integers(0, 2, (nvec, n), dtype=np.uint8)       # bit v of the code stored at every position
    codes = sum(bits[v].astype(np.uint32) << v for v in range(nvec))
    blocks = aligned_empty(num_blocks * bbytes)
    blocks[:] = 0
    view = blocks.reshape(num_blocks, bbytes)
    for v in range(nvec):  # vector v of block b: 256 bits, bit t <-> position 256*b + t, little-endian
        view[:, 32 * v:32 * (v + 1)] = np.packbits(bits[v].reshape(num_blocks, 256), axis=1, bitorder="little")
    base = view[:, 32 * nvec:].view("<u8")
    for letter, (code, care) in enumerate(table):
        match = ((codes ^ code) & care) == 0
        per_block = match.reshape(num_blocks, 256).sum(axis=1)
        base[:, letter] = np.concatenate([[0], np.cumsum(per_block)[:-1]])
    card = 20 if amino else 4
    prefix = np.zeros(card + 2, np.uint64)
    prefix[-1] = n
    num_seeds = card ** seed_k
    sp = rng.integers(1, n - 1, num_seeds).astype(np.uint64)
    ep = np.minimum(sp + rng.integers(0, 3000, num_seeds).astype(np.uint64), np.uint64(n - 1))
    seeds = np.stack([sp, ep], axis=1)
    seeds[::17] = seeds[::17][:, ::-1] + np.array([1, 0], np.uint64)  # some stored invalid pairs (sp > ep)
    return IndexArrays(abi.AwFmAlphabetAmino if amino else abi.AwFmAlphabetDna, seed_k, 8, n, blocks, prefix, seeds, None)


@pytest.mark.parametrize("amino", [False, True])
def test_selectors_on_arbitrary_block_contents(amino):
    rng = np.random.default_rng(11 + amino)
    k = 2 if amino else 5
    arrays = garbage_index(amino, 300, k, rng)
    alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY" if amino else b"ACGT", np.uint8)
    card = len(alphabet)
    # every seed entry under every letter: one LF step on two arbitrary positions of arbitrary bit-vectors
    num_seeds = card ** k
    idx = np.arange(num_seeds)
    tails = np.stack([alphabet[(idx // card ** (k - 1 - j)) % card] for j in range(k)], axis=1)
    rows = [np.concatenate([np.full((num_seeds, 1), ch, np.uint8), tails], axis=1) for ch in alphabet]
    rows.append(np.concatenate([np.full((num_seeds, 1), ord("X" if amino else "N"), np.uint8), tails], axis=1))
    letters = np.ascontiguousarray(np.concatenate(rows).reshape(-1))
    length = k + 1
    oracle = harness.Oracle(arrays)
    o_counts, o_ranges, _ = oracle.count(letters, fixed_len=length)
    gpu = GpuIndex(arrays)
    for variant, lpq in ((1, 2), (1, 1), (0, 4), (0, 1)):
        gpu.set_tuning(sweep_min_queries=-1, count_variant=variant, count_lpq=lpq)
        counts, ranges = gpu.count(letters, fixed_len=length, want_ranges=True)
        assert np.array_equal(ranges, o_ranges), (amino, variant, lpq)
        assert np.array_equal(counts, o_counts), (amino, variant, lpq)
    gpu.set_tuning(sweep_min_queries=1, sweep_profile=1, count_variant=1)
    counts, ranges = gpu.count(letters, fixed_len=length, want_ranges=True)
    assert gpu.sweep_stage_ms(), "the batch did not take the sweep path"
    assert np.array_equal(counts, o_counts) and np.array_equal(ranges, o_ranges), (amino, "sweep")
    gpu.close()


class WideSaIndex(IndexArrays):
    """The same index with its sampled SA re-packed at an arbitrary field width."""
    forced_width = 32

    @property
    def sa_width(self):
        return self.forced_width


def repack(values, width):
    out = np.zeros((len(values) * width + 7) // 8 + 8, np.uint8)
    bits = ((values[:, None] >> np.arange(width, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1)
    packed = np.packbits(bits, bitorder="little")
    out[: len(packed)] = packed
    return out


@pytest.mark.parametrize("width", [33, 40, 57, 58, 63, 64])
def test_suffix_array_fields_wider_than_32_bits(small_indexes, width):
    b = small_indexes["nuc_r3"]
    a = b.arrays
    values = a.sa_values()
    assert len(values) == sa_num_samples(a.bwt_length, a.sa_ratio)
    wide = WideSaIndex(a.alphabet, a.seed_k, a.sa_ratio, a.bwt_length, a.blocks, a.prefix_sums, a.seed_table,
                       repack(values, width))
    wide.forced_width = width
    assert np.array_equal(wide.sa_values(), values)
    rng = np.random.default_rng(width)
    starts = rng.integers(0, len(b.text) - 9, 3000)
    letters = np.ascontiguousarray(np.concatenate([b.text[s:s + 9] for s in starts]).astype(np.uint8))
    o_hit, o_pos, _ = harness.Oracle(a).locate(letters, fixed_len=9)       # the index as the reference built it
    w_hit, w_pos, _ = harness.Oracle(wide).locate(letters, fixed_len=9)    # oracle on the wide fields
    assert np.array_equal(o_hit, w_hit) and np.array_equal(o_pos, w_pos)
    gpu = GpuIndex(wide)
    for variant in (1, 0):
        gpu.set_tuning(locate_variant=variant)
        hit, pos = gpu.locate(letters, fixed_len=9)
        assert np.array_equal(hit, o_hit) and np.array_equal(pos, o_pos), (width, variant)
    gpu.close()


def big_garbage_index(num_blocks, pattern_blocks, seed_k, rng):
    """A nucleotide 'index' of num_blocks * 256 positions (more than 2^32) without 13 GB of random bits: a random
    pattern of pattern_blocks blocks repeated, base occurrences consistent with it, REALISTIC prefix sums (C[c] =
    matches of the letters before c) so that LF steps with G/T land above 2^32 again, and a seed table whose ranges
    lie anywhere in the BWT — below, above and across position 2^32."""
    bbytes = abi.NUC_BLOCK_BYTES
    n = num_blocks * 256
    pattern = rng.integers(0, 256, (pattern_blocks, 96), dtype=np.uint8)
    words = pattern.view("<u8").reshape(pattern_blocks, 3, 4)
    per_block = []
    for code, care in NUC_CODE_CARE:
        match = np.full((pattern_blocks, 4), np.uint64(0xFFFFFFFFFFFFFFFF))
        for v in range(3):
            if (care >> v) & 1:
                match &= words[:, v, :] if (code >> v) & 1 else ~words[:, v, :]
        per_block.append(np.bitwise_count(match).sum(axis=1).astype(np.uint64))
    blocks = aligned_empty(num_blocks * bbytes)
    view = blocks.reshape(num_blocks, bbytes)
    for lo in range(0, num_blocks, pattern_blocks):
        hi = min(num_blocks, lo + pattern_blocks)
        view[lo:hi, :96] = pattern[: hi - lo]
    base = view[:, 96:].view("<u8")
    base[:] = 0
    index = np.arange(num_blocks, dtype=np.uint64)
    within, repeat = (index % np.uint64(pattern_blocks)).astype(np.int64), index // np.uint64(pattern_blocks)
    totals = []
    for letter, counts in enumerate(per_block):
        before = np.concatenate([[np.uint64(0)], np.cumsum(counts, dtype=np.uint64)[:-1]])
        base[:, letter] = before[within] + repeat * np.uint64(counts.sum())
        full, rest = divmod(num_blocks, pattern_blocks)
        totals.append(int(counts.sum()) * full + int(counts[:rest].sum()))
    prefix = np.zeros(6, np.uint64)
    prefix[0] = 1  # the sentinel sorts first: no range of a real index starts at position 0 (sp - 1 is ranked)
    for c in range(1, 5):
        prefix[c] = prefix[c - 1] + np.uint64(totals[c - 1])
    prefix[5] = n
    assert int(prefix[4]) < n
    num_seeds = 4 ** seed_k
    sp = rng.integers(1, n - 4000, num_seeds).astype(np.uint64)
    sp[::4] = np.uint64(1 << 32) - rng.integers(0, 1500, len(sp[::4])).astype(np.uint64)   # ranges across 2^32
    sp[1::4] = np.uint64(1 << 32) + rng.integers(0, n - (1 << 32) - 4000, len(sp[1::4])).astype(np.uint64)  # above it
    ep = sp + rng.integers(0, 3000, num_seeds).astype(np.uint64)
    ep[5::64] = np.minimum(sp[5::64] + np.uint64(20_000_000), np.uint64(n - 1))  # wider than a wide record's 24-bit field
    sp[6::64] = np.minimum(sp[6::64], np.uint64(n - (1 << 25)))
    ep[6::64] = sp[6::64] + np.uint64((1 << 24) - 2)                              # the widest range that still fits
    seeds = np.stack([sp, ep], axis=1)
    seeds[::17] = seeds[::17][:, ::-1] + np.array([1, 0], np.uint64)  # some stored invalid pairs (sp > ep)
    return IndexArrays(abi.AwFmAlphabetDna, seed_k, 8, n, blocks, prefix, seeds, None)


def test_positions_beyond_two_to_the_32():
    """bwtLength = 1.25 * 2^32 (a two-strand human genome is 6.2 G positions): every 64-bit path of the upload-time
    re-layout (sector counts relative to 2^16-position superblocks, 64-bit superblock rows), of the tile kernels and of
    the sweep's wide passes (40-bit positions + 24-bit widths in 16-byte records, seed ranges too wide for that taking the
    per-query search), on ranges below, above and across position 2^32, bit-exact against the oracle; fixed-length ASCII,
    2-bit packed and variable-length batches."""
    rng = np.random.default_rng(2032)
    k = 5
    num_blocks = (5 << 30) // 256 + 777
    arrays = big_garbage_index(num_blocks, 1 << 15, k, rng)
    alphabet = np.frombuffer(b"ACGT", np.uint8)
    num_seeds = 4 ** k
    idx = np.arange(num_seeds)
    tails = np.stack([alphabet[(idx // 4 ** (k - 1 - j)) % 4] for j in range(k)], axis=1)
    batches = []
    rows = [np.concatenate([np.full((num_seeds, 1), ch, np.uint8), tails], axis=1) for ch in b"ACGTN"]
    batches.append((np.ascontiguousarray(np.concatenate(rows).reshape(-1)), k + 1))      # every seed under every letter
    for length, weights in ((k + 3, (1, 1, 1, 1)), (k + 7, (1, 1, 3, 6)), (k + 12, (0, 1, 4, 12))):  # G/T-heavy: stays high
        p = np.array(weights, float) / sum(weights)
        batches.append((np.ascontiguousarray(alphabet[rng.choice(4, (20000, length), p=p)].reshape(-1)), length))
    oracle = harness.Oracle(arrays)
    gpu = GpuIndex(arrays)
    group = GpuGroup(indexes=[gpu])
    seen_high = 0
    for letters, length in batches:
        o_counts, o_ranges, _ = oracle.count(letters, fixed_len=length)
        seen_high += int((o_ranges[:, 0] >= np.uint64(1 << 32)).sum())
        for variant, lpq in ((1, 2), (1, 1), (0, 2)):
            gpu.set_tuning(sweep_min_queries=-1, count_variant=variant, count_lpq=lpq)
            counts, ranges = gpu.count(letters, fixed_len=length, want_ranges=True)
            assert np.array_equal(ranges, o_ranges), (length, variant, lpq)
            assert np.array_equal(counts, o_counts), (length, variant, lpq)
        gpu.set_tuning(sweep_min_queries=1, sweep_profile=1, count_variant=1)
        counts, ranges = gpu.count(letters, fixed_len=length, want_ranges=True)
        assert gpu.sweep_stage_ms(), "the batch did not take the sweep path"
        assert np.array_equal(ranges, o_ranges), (length, "sweep")
        assert np.array_equal(counts, o_counts), (length, "sweep")
        if b"N" not in letters.tobytes():
            packed = pack_queries_bits(letters, length)
            assert np.array_equal(group.count(packed, QUERY_2BIT, fixed_len=length), o_counts), (length, "2bit")
            assert gpu.sweep_stage_ms()
    assert seen_high > 1000, "the batches never reached positions above 2^32"
    # variable lengths through the wide passes: the three random batches cut to lengths k-1 .. their own
    for letters, length in batches[1:]:
        rows = letters.reshape(-1, length)
        lengths = rng.integers(k - 1, length + 1, len(rows))
        offsets = np.zeros(len(rows) + 1, np.uint64)
        offsets[1:] = np.cumsum(lengths)
        var = np.concatenate([rows[i, length - lengths[i]:] for i in range(len(rows))])
        o_counts, o_ranges, _ = oracle.count(var, offsets)
        counts, ranges = gpu.count(var, offsets, want_ranges=True)
        assert gpu.sweep_stage_ms(), "the variable-length batch did not take the sweep path"
        assert np.array_equal(counts, o_counts) and np.array_equal(ranges, o_ranges), (length, "variable")
    group.close()
    gpu.close()
