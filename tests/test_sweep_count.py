"""Sweep count path (csrc/awfm_sweep.cuh): pack -> radix sort on the seed index -> one pass per LF step over records
bucketed by the next letter.  Forced here on small reference-built indexes (sweep_min_queries=1) and compared bit for
bit with the oracle and with the tile kernels: counts only depend on the query, not on the order it is processed in."""
import numpy as np
import pytest

from avxwindowfmindex_b200 import GpuIndex
from oracle import harness

pytestmark = pytest.mark.gpu


def fixed_batch(b, length, num, seed, irregular=True):
    if b.amino:
        return fixed_batch_amino(b, length, num, seed, irregular)
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    text = b.text
    rows = np.empty((num, length), dtype=np.uint8)
    for i in range(num):
        if rng.random() < 0.6:
            s = int(rng.integers(0, len(text) - length))
            rows[i] = text[s:s + length]
        else:
            rows[i] = alphabet[rng.integers(0, 4, length)]
    if irregular and num >= 40:
        pick = rng.choice(num, num // 20, replace=False)
        for j, i in enumerate(pick):
            col = int(rng.integers(0, length))
            rows[i, col] = (ord("N"), ord("n"), ord("$"), ord("x"), ord("-"))[j % 5]
        lower = rng.choice(num, num // 10, replace=False)
        rows[lower] |= 0x20
        rows[rng.choice(num, num // 25, replace=False)] = np.frombuffer(b"U", dtype=np.uint8)[0]
    return rows.reshape(-1)


def fixed_batch_amino(b, length, num, seed, irregular=True):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    text = b.text
    rows = np.empty((num, length), dtype=np.uint8)
    for i in range(num):
        if rng.random() < 0.6:
            s = int(rng.integers(0, len(text) - length))
            rows[i] = text[s:s + length]
        else:
            rows[i] = alphabet[rng.integers(0, 20, length)]
    if irregular and num >= 40:
        pick = rng.choice(num, num // 20, replace=False)
        for j, i in enumerate(pick):
            col = int(rng.integers(0, length))
            rows[i, col] = (ord("X"), ord("b"), ord("$"), ord("z"), ord("-"), ord("j"), ord("O"))[j % 7]
        lower = rng.choice(num, num // 10, replace=False)
        rows[lower] |= 0x20
    return rows.reshape(-1)


@pytest.mark.parametrize("name", ["amino_r8", "amino_r2", "amino_r1"])
def test_sweep_amino_matches_oracle(small_indexes, name):
    """20 buckets in 10 double-ended arrays, 5 bits per remaining letter, mixed-radix seed index as the sort key."""
    b = small_indexes[name]
    k = b.arrays.seed_k
    oracle = harness.Oracle(b.arrays)
    gpu = GpuIndex(b.arrays)
    for length, num in ((k, 700), (k + 1, 513), (k + 2, 1), (k + 3, 5000), (k + 6, 1031), (k + 2, 20000)):
        letters = fixed_batch(b, length, num, seed=length * 17 + num)
        o_counts, o_ranges, _ = oracle.count(letters, fixed_len=length)
        for bits, local, items in ((32, 8, 4), (16, 0, 2), (0, 8, 1), (3, 5, 8)):
            gpu.set_tuning(sweep_min_queries=1, sweep_sort_bits=bits, sweep_local_bits=local, sweep_items=items,
                           sweep_first_items=items, sweep_profile=1)
            counts = gpu.count(letters, fixed_len=length)
            assert np.array_equal(counts, o_counts), (name, length, num, bits, local, items)
            assert len(gpu.sweep_stage_ms()) == 3 + max(length - k, 1), "the batch did not take the sweep path"
        gpu.set_tuning(sweep_sort_bits=32, sweep_local_bits=-1, sweep_items=4, sweep_first_items=4, sweep_profile=1)
        counts, ranges = gpu.count(letters, fixed_len=length, want_ranges=True)
        assert gpu.sweep_stage_ms(), "range output did not take the sweep path"
        assert np.array_equal(counts, o_counts) and np.array_equal(ranges, o_ranges), (name, length, num)
        gpu.set_tuning(sweep_min_queries=-1)
        assert np.array_equal(gpu.count(letters, fixed_len=length), o_counts)
    # 7 letters left of the seed do not fit the 32-bit payload: the tile kernel answers
    gpu.set_tuning(sweep_min_queries=1, sweep_profile=1)
    letters = fixed_batch(b, k + 7, 300, seed=3)
    o_counts, _, _ = oracle.count(letters, fixed_len=k + 7)
    assert np.array_equal(gpu.count(letters, fixed_len=k + 7), o_counts)
    assert not gpu.sweep_stage_ms()
    gpu.close()


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r3", "nuc_r16", "nuc_r1"])
def test_sweep_matches_oracle(small_indexes, name):
    b = small_indexes[name]
    k = b.arrays.seed_k
    oracle = harness.Oracle(b.arrays)
    gpu = GpuIndex(b.arrays)
    for length, num in ((k, 700), (k + 1, 513), (k + 2, 1), (k + 5, 255), (k + 8, 5000), (k + 16, 1031), (k + 3, 20000),
                        (k + 17, 900), (k + 20, 3001), (k + 24, 1500)):  # 17..24 letters left of the seed: sweepRefill
        letters = fixed_batch(b, length, num, seed=length * 31 + num)
        o_counts, o_ranges, _ = oracle.count(letters, fixed_len=length)
        for bits, local, items in ((32, 8, 4), (16, 0, 2), (0, 8, 1), (3, 5, 8), (32, 0, 4)):
            gpu.set_tuning(sweep_min_queries=1, sweep_sort_bits=bits, sweep_local_bits=local, sweep_items=items,
                           sweep_profile=1)
            counts = gpu.count(letters, fixed_len=length)
            assert np.array_equal(counts, o_counts), (name, length, num, bits, local, items)
            assert len(gpu.sweep_stage_ms()) == 3 + max(length - k, 1), "the batch did not take the sweep path"
        # range output: every query's final (sp, ep) as the reference leaves it, incl. the pair a dying search stops at;
        # with 32-bit positions and with the 64-bit-position passes an index beyond 2^32 positions takes
        # (sweep_ordered_emit: the last pass's survivors leave through sweepEmit, bucketed by slices of the id space, or straight from the pass)
        for wide, emit, items in ((0, 1, 4), (0, 0, 4), (1, 1, 4), (1, 0, 4), (0, 1, 1), (0, 1, 8)):
            gpu.set_tuning(sweep_sort_bits=32, sweep_local_bits=-1, sweep_items=items, sweep_first_items=items, sweep_profile=1,
                           sweep_wide=wide, sweep_ordered_emit=emit)
            counts, ranges = gpu.count(letters, fixed_len=length, want_ranges=True)
            assert gpu.sweep_stage_ms(), "range output did not take the sweep path"
            assert np.array_equal(counts, o_counts) and np.array_equal(ranges, o_ranges), (name, length, num, wide, emit, items)
        gpu.set_tuning(sweep_items=4, sweep_first_items=4, sweep_wide=0, sweep_ordered_emit=1)
        gpu.set_tuning(sweep_min_queries=-1)
        assert np.array_equal(gpu.count(letters, fixed_len=length), o_counts)
    gpu.close()


def variable_batch(b, num, seed, lo, hi, irregular=True):
    """Queries of lengths lo..hi (uniform), 60 % substrings of the text, some with irregular letters / lower case."""
    rng = np.random.default_rng(seed)
    lengths = rng.integers(lo, hi + 1, num)
    offsets = np.zeros(num + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lengths)
    alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY" if b.amino else b"ACGT", dtype=np.uint8)
    odd = (b"X", b"b", b"$", b"z", b"-", b"j", b"O") if b.amino else (b"N", b"n", b"$", b"x", b"-")
    text = b.text
    letters = np.empty(int(offsets[-1]), dtype=np.uint8)
    for i in range(num):
        o, n = int(offsets[i]), int(lengths[i])
        if n == 0:
            continue
        if rng.random() < 0.6:
            s = int(rng.integers(0, len(text) - n))
            letters[o:o + n] = text[s:s + n]
        else:
            letters[o:o + n] = alphabet[rng.integers(0, len(alphabet), n)]
        if irregular:
            r = rng.random()
            if r < 0.05:
                letters[o + int(rng.integers(0, n))] = odd[i % len(odd)][0]
            elif r < 0.15:
                letters[o:o + n] |= 0x20
            elif r < 0.18 and not b.amino:
                letters[o:o + n] = ord("U")
    return letters, offsets


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r16", "amino_r8", "amino_r1"])
def test_sweep_variable_lengths(small_indexes, name):
    """Variable-length batches (letters + offsets): every record carries its own length as a marker bit above its
    remaining letters.  Lengths from 0 to past what the payload holds — shorter than the seed k-mer, exactly k, up to
    k + 15 (amino: k + 6) letters, and longer — in one batch; counts and ranges against the oracle, with every ordering
    configuration, sliced scratch, and the tile kernel as cross-check."""
    b = small_indexes[name]
    k = b.arrays.seed_k
    room = 6 if b.amino else 15
    oracle = harness.Oracle(b.arrays)
    gpu = GpuIndex(b.arrays)
    for num, lo, hi, seed in ((9000, 0, k + room + 3, 1), (5000, k, k + room, 2), (1, k + 2, k + 2, 3),
                              (300, 1, max(k - 1, 1), 4), (2500, k + room, k + room, 5), (20000, k, k + 4, 6)):
        letters, offsets = variable_batch(b, num, seed=seed * 131 + num, lo=lo, hi=hi)
        o_counts, o_ranges, _ = oracle.count(letters, offsets)
        for bits, local, own, max_batch, wide in ((32, -1, 1, 1 << 27, 0), (16, 0, 0, 1 << 27, 0), (3, 5, 1, 1024, 0),
                                                  (32, 8, 1, 4096, 0), (32, -1, 1, 1 << 27, 1)):
            gpu.set_tuning(sweep_min_queries=1, sweep_variable=1, sweep_sort_bits=bits, sweep_local_bits=local,
                           sweep_own_sort=own, sweep_max_batch=max_batch, sweep_profile=1, sweep_wide=wide)
            counts, ranges = gpu.count(letters, offsets, want_ranges=True)
            assert len(gpu.sweep_stage_ms()) == 3 + room, "the batch did not take the sweep path"
            assert np.array_equal(counts, o_counts), (name, num, lo, hi, bits, local, own, max_batch)
            assert np.array_equal(ranges, o_ranges), (name, num, lo, hi, bits, local, own, max_batch)
            assert np.array_equal(gpu.count(letters, offsets), o_counts)
        gpu.set_tuning(sweep_variable=0, sweep_max_batch=1 << 27, sweep_wide=0)
        counts, ranges = gpu.count(letters, offsets, want_ranges=True)
        assert not gpu.sweep_stage_ms(), "sweep_variable=0 still took the sweep path"
        assert np.array_equal(counts, o_counts) and np.array_equal(ranges, o_ranges)
    gpu.close()


def test_sweep_variable_lengths_on_device_buffers(small_indexes):
    """awfm_gpu_count_device with an offsets array whose letters end exactly at the end of the allocation's payload
    (the pack kernel reads aligned words but nothing past offsets[n]) and a letter base that is only 16-B aligned."""
    import torch
    b = small_indexes["nuc_r8"]
    k = b.arrays.seed_k
    oracle = harness.Oracle(b.arrays)
    gpu = GpuIndex(b.arrays)
    gpu.set_tuning(sweep_min_queries=1, sweep_profile=1)
    for num, seed in ((4097, 11), (513, 12), (7, 13)):
        letters, offsets = variable_batch(b, num, seed=seed, lo=k - 2, hi=k + 17)
        o_counts, o_ranges, _ = oracle.count(letters, offsets)
        pad = torch.zeros(len(letters) + 48, dtype=torch.uint8, device="cuda")
        d_letters = pad[16:16 + len(letters)]
        d_letters.copy_(torch.from_numpy(letters))
        d_offsets = torch.from_numpy(offsets.view(np.int64)).cuda()
        d_counts = torch.full((num,), 7, dtype=torch.int32, device="cuda")
        d_ranges = torch.zeros((num, 2), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        gpu.count_device(d_letters.data_ptr(), d_offsets.data_ptr(), 0, num, d_counts.data_ptr(), d_ranges.data_ptr())
        torch.cuda.synchronize()
        assert gpu.sweep_stage_ms(), "the batch did not take the sweep path"
        assert np.array_equal(d_counts.cpu().numpy().astype(np.uint32), o_counts)
        assert np.array_equal(d_ranges.cpu().numpy().view(np.uint64), o_ranges)
    gpu.close()


def test_sweep_falls_back_outside_its_domain(small_indexes):
    """Too many letters left of the seed in a fixed-length batch: the tile kernels answer, same results."""
    b = small_indexes["nuc_r8"]
    k = b.arrays.seed_k
    oracle = harness.Oracle(b.arrays)
    gpu = GpuIndex(b.arrays)
    gpu.set_tuning(sweep_min_queries=1)
    for length in (k + 25, 33, 40):  # more than 24 letters left of the seed / more than 32 letters in all
        letters = fixed_batch(b, length, 300, seed=5)
        o_counts, o_ranges, _ = oracle.count(letters, fixed_len=length)
        gpu.set_tuning(sweep_profile=1)
        assert np.array_equal(gpu.count(letters, fixed_len=length), o_counts)
        assert not gpu.sweep_stage_ms(), "a batch outside the sweep's domain took it"
    letters = fixed_batch(b, k + 4, 300, seed=6)
    o_counts, o_ranges, _ = oracle.count(letters, fixed_len=k + 4)
    counts, ranges = gpu.count(letters, fixed_len=k + 4, want_ranges=True)
    assert np.array_equal(counts, o_counts) and np.array_equal(ranges, o_ranges)
    offsets = np.arange(0, len(letters) + 1, k + 4, dtype=np.uint64)
    assert np.array_equal(gpu.count(letters, offsets), o_counts)
    gpu.close()


def test_sweep_slices_large_batches(small_indexes):
    """A batch larger than sweep_max_batch goes through the scratch in slices; slice starts stay 16-B aligned."""
    b = small_indexes["nuc_r8"]
    k = b.arrays.seed_k
    oracle = harness.Oracle(b.arrays)
    gpu = GpuIndex(b.arrays)
    for length in (k + 3, k + 6):
        letters = fixed_batch(b, length, 5000, seed=77 + length)
        o_counts, _, _ = oracle.count(letters, fixed_len=length)
        for max_batch in (256, 1000, 4999, 5000):
            gpu.set_tuning(sweep_min_queries=1, sweep_max_batch=max_batch, sweep_profile=1)
            assert np.array_equal(gpu.count(letters, fixed_len=length), o_counts), (length, max_batch)
            assert gpu.sweep_stage_ms()
    gpu.close()


def test_sweep_with_derived_seed_table(small_indexes):
    b = small_indexes["nuc_r16"]
    k = b.arrays.seed_k
    oracle = harness.Oracle(b.arrays)
    gpu = GpuIndex(b.arrays)
    gpu.set_tuning(sweep_min_queries=1)
    gpu.extend_seed_table(k + 3)
    for length in (k + 1, k + 3, k + 4, k + 11):
        letters = fixed_batch(b, length, 3000, seed=length)
        o_counts, _, _ = oracle.count(letters, fixed_len=length)
        assert np.array_equal(gpu.count(letters, fixed_len=length), o_counts), length
    gpu.close()


def test_sweep_repeated_calls_and_streams(small_indexes):
    """Scratch is reused across calls, grown when a larger batch arrives, and serialised across streams."""
    import torch
    b = small_indexes["nuc_r8"]
    k = b.arrays.seed_k
    oracle = harness.Oracle(b.arrays)
    gpu = GpuIndex(b.arrays)
    gpu.set_tuning(sweep_min_queries=1)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    for i, num in enumerate((4000, 100, 9000, 9000)):
        letters = fixed_batch(b, k + 6, num, seed=100 + i)
        d_letters = torch.from_numpy(letters).cuda()
        d_counts = torch.full((num,), 7, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        gpu.count_device(d_letters.data_ptr(), None, k + 6, num, d_counts.data_ptr(), None, streams[i % 2].cuda_stream)
        outs.append((letters, d_letters, d_counts))
    torch.cuda.synchronize()
    for letters, _, d_counts in outs:
        o_counts, _, _ = oracle.count(letters, fixed_len=k + 6)
        assert np.array_equal(d_counts.cpu().numpy().astype(np.uint32), o_counts)
    gpu.close()


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r16", "amino_r8"])
def test_own_bucket_passes_and_cub_order_alike(small_indexes, name):
    """The ordering step (csrc/awfm_sort.cuh: two unstable most-significant-digit-first bucket passes) against CUB's
    radix sort on the same batches: counts and ranges identical to the oracle either way, for one- and two-digit splits
    and for batches smaller and larger than one sort tile."""
    b = small_indexes[name]
    k = b.arrays.seed_k
    oracle = harness.Oracle(b.arrays)
    gpu = GpuIndex(b.arrays)
    for length, num in ((k + 1, 3), (k + 3, 3071), (k + 2, 3073), (k + 4, 50000), (k, 9000)):
        letters = fixed_batch(b, length, num, seed=num + length)
        o_counts, o_ranges, _ = oracle.count(letters, fixed_len=length)
        for own in (1, 0):
            for bits, local in ((32, -1), (32, 0), (5, 3), (9, 2), (16, 0), (1, 8)):
                gpu.set_tuning(sweep_min_queries=1, sweep_own_sort=own, sweep_sort_bits=bits, sweep_local_bits=local,
                               sweep_profile=1)
                counts, ranges = gpu.count(letters, fixed_len=length, want_ranges=True)
                assert gpu.sweep_stage_ms(), "the batch did not take the sweep path"
                assert np.array_equal(counts, o_counts), (name, length, num, own, bits, local)
                assert np.array_equal(ranges, o_ranges), (name, length, num, own, bits, local)
    gpu.close()


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r16", "amino_r8"])
def test_compact_pairs_through_the_bucket_passes(small_indexes, name):
    """csrc/awfm_sort.cuh, compact pairs: one 8-byte word per pair between the pack kernel and the second bucket pass
    (query id implicit in the pack kernel's output, explicit once the first digit has left the key).  Two-digit
    orderings (sweep_local_bits = 0 leaves every key bit to the bucket passes) with the compact words on and off, for
    fixed-length ASCII and 2-bit batches, variable-length batches, batches with irregular letters, batches smaller and
    larger than one sort tile: counts and ranges equal to the oracle's."""
    from avxwindowfmindex_b200 import GpuGroup, pack_queries_bits
    from avxwindowfmindex_b200.search import QUERY_2BIT
    b = small_indexes[name]
    k = b.arrays.seed_k
    room = 6 if b.amino else 15
    oracle = harness.Oracle(b.arrays)
    gpu = GpuIndex(b.arrays)
    group = GpuGroup(indexes=[gpu])
    for compact in (1, 0):
        for local in (0, 2):
            gpu.set_tuning(sweep_min_queries=1, sweep_own_sort=1, sweep_sort_bits=32, sweep_local_bits=local,
                           sweep_compact_pairs=compact, sweep_profile=1)
            for length, num in ((k + 1, 3), (k + 3, 3071), (k + 2, 3073), (k + 4, 50000), (k, 9000), (k + room // 2, 7001)):
                letters = fixed_batch(b, length, num, seed=num + 3 * length)
                o_counts, o_ranges, _ = oracle.count(letters, fixed_len=length)
                counts, ranges = gpu.count(letters, fixed_len=length, want_ranges=True)
                assert gpu.sweep_stage_ms(), "the batch did not take the sweep path"
                assert np.array_equal(counts, o_counts), (name, length, num, compact, local)
                assert np.array_equal(ranges, o_ranges), (name, length, num, compact, local)
                if not b.amino:
                    clean = fixed_batch(b, length, num, seed=num + 5 * length, irregular=False) & 0xDF  # upper case
                    clean[~np.isin(clean, np.frombuffer(b"ACGT", dtype=np.uint8))] = ord("A")  # (the text holds N's)
                    c_counts, _, _ = oracle.count(clean, fixed_len=length)
                    packed = pack_queries_bits(clean, length)
                    assert np.array_equal(group.count(packed, QUERY_2BIT, fixed_len=length), c_counts), (name, length, num, compact, local, "2bit")
            for num, lo, hi, seed in ((9000, 0, k + room + 3, 11), (20000, k, k + 4, 12), (2, k + 1, k + 2, 13)):
                letters, offsets = variable_batch(b, num, seed=seed * 17 + num, lo=lo, hi=hi)
                o_counts, o_ranges, _ = oracle.count(letters, offsets)
                counts, ranges = gpu.count(letters, offsets, want_ranges=True)
                assert len(gpu.sweep_stage_ms()) == 3 + room, "the batch did not take the sweep path"
                assert np.array_equal(counts, o_counts), (name, num, lo, hi, compact, local)
                assert np.array_equal(ranges, o_ranges), (name, num, lo, hi, compact, local)
    group.close()
    gpu.close()


def test_twelve_byte_records_and_the_wide_range_escape(reference, tmp_path):
    """Nucleotide batches with at most 8 letters left of the seed travel as 12-byte records with a 16-bit range width;
    a query whose SEED range is wider than 65534 positions leaves the sweep for the generic per-query search.  With a
    2-letter seed table over 1.2 Mbp every seed range is ~75 k wide, so every query takes the escape; with k = 6 only
    some do.  ASCII and 2-bit packed input, against the oracle, with the 16-byte records as cross-check."""
    from avxwindowfmindex_b200 import GpuGroup, abi, pack_queries_bits
    from avxwindowfmindex_b200.search import QUERY_2BIT
    rng = np.random.default_rng(5)
    text = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 1_200_000)]
    text[:400_000] = ord("A")  # a long run: the seed ranges of AAAAAA... are very wide
    for k in (2, 6):
        ptr = reference.create_index(text.tobytes(), str(tmp_path / f"wide{k}.awfmi"), abi.AwFmAlphabetDna, k, 16)
        arrays = reference.arrays(ptr)
        oracle = harness.Oracle(arrays)
        gpu = GpuIndex(arrays)
        group = GpuGroup(indexes=[gpu])
        for length in (k, k + 1, k + 5, k + 8):
            starts = rng.integers(0, len(text) - length, 6000)
            letters = np.ascontiguousarray(text[starts[:, None] + np.arange(length)[None, :]].reshape(-1))
            letters.reshape(-1, length)[::3] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (2000, length))]
            o_counts, o_ranges, _ = oracle.count(letters, fixed_len=length)
            for rec12 in (1, 0):
                gpu.set_tuning(sweep_min_queries=1, sweep_record12=rec12, sweep_profile=1)
                counts, ranges = gpu.count(letters, fixed_len=length, want_ranges=True)
                assert gpu.sweep_stage_ms(), "the batch did not take the sweep path"
                assert np.array_equal(counts, o_counts) and np.array_equal(ranges, o_ranges), (k, length, rec12)
                packed = pack_queries_bits(letters, length)
                assert np.array_equal(group.count(packed, QUERY_2BIT, fixed_len=length), o_counts), (k, length, rec12, "2bit")
        group.close()
        gpu.close()
        reference.dealloc_index(ptr)
