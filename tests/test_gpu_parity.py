"""Parity of the CUDA path (through the C-ABI and through the drop-in entry points) with the oracle and with the
compiled reference, bit for bit: ranges as left by the search (including the stored invalid pair), u32 counts,
position lists element-wise in SA order, return codes, capacity semantics.  Run with -m gpu on a B200."""
import ctypes as C

import numpy as np
import pytest

from avxwindowfmindex_b200 import (GpuIndex, KmerSearchList, abi, capi, parallel_search_count,
                                   parallel_search_locate)
from avxwindowfmindex_b200.search import pack_queries
from oracle import harness
from conftest import make_queries, make_text

pytestmark = pytest.mark.gpu

ALL = ["nuc_r8", "nuc_r1", "nuc_r3", "nuc_r16", "nuc_r200", "nuc_r255", "amino_r8", "amino_r2", "amino_r1"]


def queries_for(b, seed=21, num=700):
    k = b.arrays.seed_k
    return make_queries(b.text, b.amino, seed=seed, num=num, min_len=1, max_len=max(12, k + 9), seed_k=k)


@pytest.mark.parametrize("name", ALL)
def test_packed_batch_matches_oracle_and_reference(small_indexes, reference, name):
    b = small_indexes[name]
    letters, offsets = pack_queries(queries_for(b))
    oracle = harness.Oracle(b.arrays)
    o_counts, o_ranges, _ = oracle.count(letters, offsets)
    o_hit, o_pos, _ = oracle.locate(letters, offsets)
    r_counts = reference.count(b.ptr, letters, offsets, threads=2)
    assert np.array_equal(o_counts, r_counts)
    gpu = GpuIndex(b.arrays)
    for lpq in (8, 4, 2, 1):
        for variant in (1, 0):
            gpu.set_tuning(count_lpq=lpq, locate_lpq=lpq, count_variant=variant, locate_variant=variant)
            counts, ranges = gpu.count(letters, offsets, want_ranges=True)
            assert np.array_equal(counts, r_counts), (name, lpq, variant)
            assert np.array_equal(ranges, o_ranges), (name, lpq, variant)
            hit, pos, ranges2 = gpu.locate(letters, offsets, want_ranges=True)
            assert np.array_equal(hit, o_hit) and np.array_equal(pos, o_pos), (name, lpq, variant)
            assert np.array_equal(ranges2, o_ranges)
    gpu.close()


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r16", "amino_r8"])
def test_fixed_length_batches(small_indexes, name):
    b = small_indexes[name]
    oracle = harness.Oracle(b.arrays)
    gpu = GpuIndex(b.arrays)
    rng = np.random.default_rng(4)
    for length in (1, b.arrays.seed_k - 1, b.arrays.seed_k, b.arrays.seed_k + 1, 12, 33):
        if length < 1:
            continue
        starts = rng.integers(0, len(b.text) - length, 500)
        letters = np.concatenate([b.text[s:s + length] for s in starts]).astype(np.uint8)
        letters[::97] = ord("A")  # perturb some queries so a few miss
        o_counts, o_ranges, _ = oracle.count(letters, fixed_len=length)
        for variant in (0, 1):
            gpu.set_tuning(count_variant=variant)
            counts, ranges = gpu.count(letters, fixed_len=length, want_ranges=True)
            assert np.array_equal(counts, o_counts) and np.array_equal(ranges, o_ranges), (name, length, variant)
        o_hit, o_pos, _ = oracle.locate(letters, fixed_len=length)
        hit, pos = gpu.locate(letters, fixed_len=length)
        assert np.array_equal(hit, o_hit) and np.array_equal(pos, o_pos)
    gpu.close()


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r3", "amino_r8", "amino_r2"])
def test_drop_in_entry_points_match_reference(small_indexes, reference, name):
    """Same search list driven through the reference's awFmParallelSearchCount/Locate and through ours."""
    b = small_indexes[name]
    lib = capi.load()
    letters, offsets = pack_queries(queries_for(b, seed=33, num=900))
    index_struct = b.arrays.as_awfm_index()  # what a C caller holds after awFmReadIndexFromFile(..., true)
    ip = C.addressof(index_struct)
    r_counts = reference.count(b.ptr, letters, offsets, threads=4)
    rc, r_counts2, r_pos = reference.locate(b.ptr, letters, offsets, threads=4)
    for threads in (1, 4):
        sl = KmerSearchList(lib, len(offsets) - 1).fill(letters, offsets)
        parallel_search_count(lib, ip, sl, threads)
        assert lib.awFmGpuLastCountStatus() == abi.AwFmSuccess
        assert np.array_equal(sl.counts(), r_counts)
        # the list is reusable across calls (src/AwFmIndex.h:344-345): locate on the same list
        assert parallel_search_locate(lib, ip, sl, threads) == abi.AwFmSuccess == rc
        assert np.array_equal(sl.counts(), r_counts2)
        mine = sl.positions()
        for i, p in enumerate(r_pos):
            assert np.array_equal(p, mine[i]), (name, i)
        e = sl.entries()
        n = sl.count
        # capacity = max(old capacity, count): grown by realloc to exactly count (src/AwFmParallelSearch.c:367-387)
        assert np.array_equal(e["capacity"][:n], np.maximum(4, e["count"][:n]))
        # second locate on the same list: capacities already fit, results identical
        assert parallel_search_locate(lib, ip, sl, threads) == abi.AwFmSuccess
        assert all(np.array_equal(p, q) for p, q in zip(r_pos, sl.positions()))
        sl.close()
    lib.awFmGpuReleaseIndex(ip)


def test_drop_in_with_suffix_array_left_on_disk(small_indexes, reference):
    """keepSuffixArrayInMemory=false: the reference preads one value per hit (src/AwFmFile.c:484-522); the drop-in
    reads the SA section once from index->fileDescriptor.  Driven on the reference's own struct AwFmIndex."""
    b = small_indexes["nuc_r8"]
    lib = capi.load()
    disk = reference.read_index(b.path, keep_sa=False)
    assert not reference.struct(disk).suffixArray.values
    letters, offsets = pack_queries(queries_for(b, seed=8, num=300))
    rc, r_counts, r_pos = reference.locate(disk, letters, offsets, threads=2)
    sl = KmerSearchList(lib, len(offsets) - 1).fill(letters, offsets)
    assert parallel_search_locate(lib, disk, sl, 2) == rc == abi.AwFmSuccess
    assert np.array_equal(sl.counts(), r_counts)
    assert all(np.array_equal(p, q) for p, q in zip(r_pos, sl.positions()))
    sl.close()
    lib.awFmGpuReleaseIndex(disk)
    reference.dealloc_index(disk)


def test_empty_and_degenerate_inputs(small_indexes):
    b = small_indexes["nuc_r8"]
    lib = capi.load()
    gpu = GpuIndex(b.arrays)
    assert len(gpu.count(np.zeros(0, np.uint8), np.zeros(1, np.uint64))) == 0
    hit, pos = gpu.locate(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert hit.tolist() == [0] and len(pos) == 0
    # zero-length query and '$' inside a query: undefined in the reference, defined as "no match" here and in the oracle
    letters, offsets = pack_queries([b"", b"AC$T", b"$", b"ACGT"])
    oracle = harness.Oracle(b.arrays)
    o_counts, o_ranges, _ = oracle.count(letters, offsets)
    counts, ranges = gpu.count(letters, offsets, want_ranges=True)
    assert np.array_equal(counts, o_counts) and counts[:3].tolist() == [0, 0, 0]
    assert np.array_equal(ranges, o_ranges)
    gpu.close()
    # empty search list through the drop-in
    ix = b.arrays.as_awfm_index()
    sl = KmerSearchList(lib, 0)
    parallel_search_count(lib, C.addressof(ix), sl, 4)
    assert parallel_search_locate(lib, C.addressof(ix), sl, 4) == abi.AwFmSuccess
    sl.close()
    lib.awFmGpuReleaseIndex(C.addressof(ix))


def test_long_queries_and_unaligned_batches(small_indexes):
    """Queries longer than the shared-memory staging window and a letters pointer that is not 16-B aligned."""
    b = small_indexes["nuc_r16"]
    oracle = harness.Oracle(b.arrays)
    gpu = GpuIndex(b.arrays)
    rng = np.random.default_rng(9)
    qs = []
    for _ in range(40):
        length = int(rng.integers(2000, 12000))
        s = int(rng.integers(0, len(b.text) - length))
        qs.append(bytes(b.text[s:s + length]))
    qs += [bytes(b.text[10:30])] * 300
    letters, offsets = pack_queries(qs)
    o_counts, o_ranges, _ = oracle.count(letters, offsets)
    assert (o_counts[:40] >= 1).all()
    for variant in (0, 1):
        gpu.set_tuning(count_variant=variant)
        counts, ranges = gpu.count(letters, offsets, want_ranges=True)
        assert np.array_equal(counts, o_counts) and np.array_equal(ranges, o_ranges)
    # device API with a deliberately misaligned letters pointer
    import torch
    d_letters = torch.zeros(len(letters) + 64, dtype=torch.uint8, device="cuda")
    d_letters[3:3 + len(letters)] = torch.from_numpy(letters).cuda()
    d_offsets = torch.from_numpy(offsets.astype(np.int64)).cuda()
    n = len(offsets) - 1
    d_counts = torch.zeros(n, dtype=torch.int32, device="cuda")
    gpu.set_tuning(count_variant=1)
    gpu.count_device(d_letters.data_ptr() + 3, d_offsets.data_ptr(), 0, n, d_counts.data_ptr(), None,
                     torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_counts.cpu().numpy().astype(np.uint32), o_counts)
    gpu.close()


def test_device_buffer_pipeline_with_torch(small_indexes):
    """count_device -> scan_ranges_device -> locate_device on torch-owned device memory and torch's stream."""
    import torch
    b = small_indexes["amino_r8"]
    letters, offsets = pack_queries(queries_for(b, seed=77, num=1500))
    oracle = harness.Oracle(b.arrays)
    o_hit, o_pos, _ = oracle.locate(letters, offsets)
    gpu = GpuIndex(b.arrays)
    n = len(offsets) - 1
    stream = torch.cuda.current_stream().cuda_stream
    d_letters = torch.from_numpy(letters).cuda()
    d_offsets = torch.from_numpy(offsets.astype(np.int64)).cuda()
    d_counts = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_ranges = torch.zeros((n, 2), dtype=torch.int64, device="cuda")
    d_hit = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    gpu.count_device(d_letters.data_ptr(), d_offsets.data_ptr(), 0, n, d_counts.data_ptr(), d_ranges.data_ptr(), stream)
    gpu.scan_ranges_device(d_ranges.data_ptr(), n, d_hit.data_ptr(), stream)
    total = int(d_hit[-1].item())
    assert total == int(o_hit[-1])
    d_pos = torch.zeros(total, dtype=torch.int64, device="cuda")
    half = total // 2  # two launches over flat hit sub-ranges
    gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), n, 0, half, d_pos.data_ptr(), stream)
    gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), n, half, total, d_pos.data_ptr() + 8 * half, stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_hit.cpu().numpy().astype(np.uint64), o_hit)
    assert np.array_equal(d_pos.cpu().numpy().astype(np.uint64), o_pos)
    assert gpu.stats()["launches"] >= 1
    gpu.close()


def test_medium_index_many_queries(reference, tmp_path):
    """2 Mbp reference-built index, 200k random + sampled 14-mers, chunked through the drop-in pipeline."""
    text = make_text(2_000_003, False, seed=12)
    path = str(tmp_path / "medium.awfmi")
    ptr = reference.create_index(text.tobytes(), path, abi.AwFmAlphabetDna, 8, 8)
    arrays = reference.arrays(ptr)
    rng = np.random.default_rng(2)
    n, length = 200_000, 14
    starts = rng.integers(0, len(text) - length, n)
    letters = text[starts[:, None] + np.arange(length)[None, :]].reshape(-1).copy()
    rand = rng.integers(0, n, n // 2)
    letters.reshape(n, length)[rand, rng.integers(0, length, n // 2)] = ord("C")
    oracle = harness.Oracle(arrays)
    o_counts, o_ranges, work = oracle.count(letters, fixed_len=length, threads=8)
    r_counts = reference.count(ptr, letters, fixed_len=length, threads=8)
    assert np.array_equal(o_counts, r_counts)
    lib = capi.load()
    ix = arrays.as_awfm_index()
    sl = KmerSearchList(lib, n).fill(letters, fixed_len=length)
    import os
    os.environ["AWFM_GPU_CHUNK_QUERIES"] = "30000"  # force several pipeline chunks
    os.environ["AWFM_GPU_LOCATE_CHUNK_QUERIES"] = "17000"
    try:
        parallel_search_count(lib, C.addressof(ix), sl, 8)
        assert np.array_equal(sl.counts(), r_counts)
        assert parallel_search_locate(lib, C.addressof(ix), sl, 8) == abi.AwFmSuccess
    finally:
        del os.environ["AWFM_GPU_CHUNK_QUERIES"]
        del os.environ["AWFM_GPU_LOCATE_CHUNK_QUERIES"]
    o_hit, o_pos, _ = oracle.locate(letters, fixed_len=length, threads=8)
    mine = sl.positions()
    flat = np.concatenate(mine)
    assert np.array_equal(flat, o_pos)
    # size-independent property: every reported position really matches the query text
    for i in range(0, n, 997):
        for p in mine[i]:
            assert bytes(text[int(p):int(p) + length]) == bytes(letters[i * length:(i + 1) * length])
    sl.close()
    lib.awFmGpuReleaseIndex(C.addressof(ix))
    reference.dealloc_index(ptr)
