"""Device groups and the packed-batch engine (include/awfm_gpu.h: awfm_gpu_group_*; drop-in: awFmGpuCountPacked /
awFmGpuLocatePacked): every query format, pageable and page-locked buffers, one and several contexts per group, the
sweep and the tile path, against the oracle and the compiled reference, bit for bit.  A box with one GPU exercises the
fan-out with two contexts on the same device.  Run with -m gpu on a B200."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from avxwindowfmindex_b200 import (GpuGroup, GpuIndex, KmerSearchList, PinnedArray, abi, capi, pack_queries_bits,
                                   parallel_search_count, parallel_search_locate)
from avxwindowfmindex_b200.search import QUERY_2BIT, QUERY_5BIT, QUERY_ASCII, pack_queries
from oracle import harness
from conftest import make_queries

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fixed_queries(b, length, num, seed):
    """num fixed-length queries: two thirds cut from the text (hits), the rest random; plain letters only."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY" if b.amino else b"ACGT", dtype=np.uint8)
    text = b.text.copy()
    bad = ~np.isin(text & 0xDF, alphabet)
    text[bad] = alphabet[0]
    text = text & 0xDF  # upper case
    starts = rng.integers(0, len(text) - length, num)
    q = text[starts[:, None] + np.arange(length)[None, :]]
    rnd = rng.random(num) < 0.33
    q[rnd] = alphabet[rng.integers(0, len(alphabet), (int(rnd.sum()), length))]
    return np.ascontiguousarray(q.reshape(-1))


def two_contexts(arrays):
    """Two independent contexts on device 0 (plus a second device when the box has one)."""
    lib = capi.load()
    devs = [0, 1 % max(1, lib.awfm_gpu_device_count())]
    return [GpuIndex(arrays, device=d) for d in devs]


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r16", "amino_r8"])
@pytest.mark.parametrize("sweep", [False, True])
def test_group_count_every_format(small_indexes, name, sweep):
    b = small_indexes[name]
    k = b.arrays.seed_k
    oracle = harness.Oracle(b.arrays)
    indexes = two_contexts(b.arrays)
    group = GpuGroup(indexes=indexes)
    # small chunks and shards so that the pipeline has several chunks per device and both devices get a shard
    group.set_tuning(packed_chunk_queries=1024, packed_min_shard=512, sweep_min_queries=1 if sweep else -1)
    bits_fmt = QUERY_5BIT if b.amino else QUERY_2BIT
    for length in (k, k + 1, k + 4, (k + 6) if b.amino else 20, 31) + (() if b.amino else (28, 29, 32)):
        letters = fixed_queries(b, length, 5000 + length, seed=length)
        o_counts, _, _ = oracle.count(letters, fixed_len=length)
        counts = group.count(letters, QUERY_ASCII, fixed_len=length)
        assert np.array_equal(counts, o_counts), (name, length, "ascii")
        packed = pack_queries_bits(letters, length, amino=b.amino)
        counts = group.count(packed, bits_fmt, fixed_len=length)
        assert np.array_equal(counts, o_counts), (name, length, "bits")
        # page-locked input and output: the DMA engines work on the caller's memory in place
        pin, pout = PinnedArray(len(packed), np.uint8), PinnedArray(len(o_counts), np.uint32)
        pin.array[:] = packed
        pout.array[:] = 0xFFFFFFFF
        group.count(pin.array, bits_fmt, fixed_len=length, out=pout.array)
        assert np.array_equal(pout.array, o_counts), (name, length, "pinned")
        pin.close(), pout.close()
    group.close()
    for ix in indexes:
        ix.close()


def test_group_count_variable_length_and_irregular(small_indexes, reference):
    """Variable-length ASCII batches (offsets) with ambiguity letters, lower case, short queries: chunk boundaries fall
    at arbitrary letter offsets."""
    for name in ("nuc_r8", "amino_r8"):
        b = small_indexes[name]
        letters, offsets = pack_queries(make_queries(b.text, b.amino, seed=77, num=3000, min_len=1, max_len=40,
                                                     seed_k=b.arrays.seed_k))
        r_counts = reference.count(b.ptr, letters, offsets, threads=2)
        indexes = two_contexts(b.arrays)
        group = GpuGroup(indexes=indexes)
        o_hit, o_pos, _ = harness.Oracle(b.arrays).locate(letters, offsets)
        for sweep in (-1, 1):  # tile kernel | sweep with marker-bit payloads (chunks start at unaligned letter offsets)
            group.set_tuning(packed_chunk_queries=256, packed_min_shard=700, sweep_min_queries=sweep)
            counts = group.count(letters, QUERY_ASCII, offsets=offsets)
            assert np.array_equal(counts, r_counts), (name, sweep)
            hit, pos = group.locate(letters, QUERY_ASCII, offsets=offsets)
            assert np.array_equal(hit, o_hit) and np.array_equal(pos, o_pos), (name, sweep)
        group.close()
        for ix in indexes:
            ix.close()


def test_five_bit_ambiguity_codes(small_indexes):
    """5-bit codes >= 20 are the ambiguity letter: same result as an ASCII query holding 'X' there."""
    b = small_indexes["amino_r8"]
    length = b.arrays.seed_k + 3
    letters = fixed_queries(b, length, 4000, seed=5)
    letters.reshape(-1, length)[::7, 1] = ord("X")
    letters.reshape(-1, length)[::11, length - 1] = ord("X")  # inside the seed window: non-seeded start
    o_counts, _, _ = harness.Oracle(b.arrays).count(letters, fixed_len=length)
    packed = pack_queries_bits(letters, length, amino=True)
    gpu = GpuIndex(b.arrays)
    group = GpuGroup(indexes=[gpu])
    for sweep in (-1, 1):
        group.set_tuning(sweep_min_queries=sweep)
        assert np.array_equal(group.count(packed, QUERY_5BIT, fixed_len=length), o_counts), sweep
    group.close()
    gpu.close()


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r3", "amino_r2"])
def test_group_locate_csr(small_indexes, name):
    b = small_indexes[name]
    length = b.arrays.seed_k + 2
    letters = fixed_queries(b, length, 6000, seed=9)
    o_hit, o_pos, _ = harness.Oracle(b.arrays).locate(letters, fixed_len=length)
    indexes = two_contexts(b.arrays)
    group = GpuGroup(indexes=indexes)
    group.set_tuning(packed_chunk_queries=1024, packed_min_shard=512, packed_window_hits=500)
    packed = pack_queries_bits(letters, length, amino=b.amino)
    for q, fmt in ((letters, QUERY_ASCII), (packed, QUERY_5BIT if b.amino else QUERY_2BIT)):
        hit, pos = group.locate(q, fmt, fixed_len=length)
        assert np.array_equal(hit, o_hit) and np.array_equal(pos, o_pos), (name, fmt)
    # too small a positions buffer: offsets and the total only, nothing written past the buffer
    total = C.c_uint64()
    hit = np.zeros(len(o_hit), np.uint64)
    small = np.full(3, 0xABCD, np.uint64)
    capi.check(group.lib.awfm_gpu_group_locate(group.handle, letters.ctypes.data, QUERY_ASCII, None, length,
                                               len(o_hit) - 1, hit.ctypes.data, small.ctypes.data, 3, None, None,
                                               C.byref(total)))
    assert total.value == o_hit[-1] and np.array_equal(hit, o_hit) and (small == 0xABCD).all()
    # page-locked outputs
    ph, pp = PinnedArray(len(o_hit), np.uint64), PinnedArray(max(1, len(o_pos)), np.uint64)
    group.locate(letters, QUERY_ASCII, fixed_len=length, out=(ph.array, pp.array))
    assert np.array_equal(ph.array, o_hit) and np.array_equal(pp.array[: len(o_pos)], o_pos)
    ph.close(), pp.close()
    group.close()
    for ix in indexes:
        ix.close()


def test_group_search_list_fan_out(small_indexes, reference):
    """The reference's own list layout over two contexts: chunk r on device r mod 2; same results as the reference."""
    for name in ("nuc_r8", "amino_r8"):
        b = small_indexes[name]
        letters, offsets = pack_queries(make_queries(b.text, b.amino, seed=3, num=5000, min_len=1, max_len=30,
                                                     seed_k=b.arrays.seed_k))
        r_counts = reference.count(b.ptr, letters, offsets, threads=2)
        rc, r_counts2, r_pos = reference.locate(b.ptr, letters, offsets, threads=2)
        indexes = two_contexts(b.arrays)
        group = GpuGroup(indexes=indexes)
        group.set_tuning(chunk_queries=256, locate_chunk_queries=256)
        lib = group.lib
        for threads in (1, 5):
            sl = KmerSearchList(lib, len(offsets) - 1).fill(letters, offsets)
            capi.check(lib.awfm_gpu_group_search_list_count(group.handle, sl.ptr.contents.kmerSearchData, sl.count, threads))
            assert np.array_equal(sl.counts(), r_counts), (name, threads)
            capi.check(lib.awfm_gpu_group_search_list_locate(group.handle, sl.ptr.contents.kmerSearchData, sl.count, threads))
            assert np.array_equal(sl.counts(), r_counts2)
            assert all(np.array_equal(p, q) for p, q in zip(r_pos, sl.positions())), (name, threads)
            sl.close()
        group.close()
        for ix in indexes:
            ix.close()


def test_drop_in_packed_api_and_device_list(small_indexes, reference):
    """awFmGpuCountPacked / awFmGpuLocatePacked on the reference's own struct AwFmIndex, with AWFM_GPU_DEVICES naming
    device 0 twice... is not allowed (a replica per device), so the list is "0" here and "all" on a multi-GPU box."""
    b = small_indexes["nuc_r8"]
    lib = capi.load()
    length = 14
    letters = fixed_queries(b, length, 3000, seed=12)
    r_counts = reference.count(b.ptr, letters, fixed_len=length, threads=2)
    rc, _, r_pos = reference.locate(b.ptr, letters, fixed_len=length, threads=2)
    os.environ["AWFM_GPU_DEVICES"] = "all"
    try:
        assert lib.awFmGpuNumDevices(b.ptr) == lib.awfm_gpu_device_count()
        counts = np.zeros(len(r_counts), np.uint32)
        packed = pack_queries_bits(letters, length)
        assert lib.awFmGpuCountPacked(b.ptr, packed.ctypes.data, abi.AwFmGpuKmer2Bit, None, length, len(counts),
                                      counts.ctypes.data) == abi.AwFmSuccess
        assert np.array_equal(counts, r_counts)
        total = C.c_uint64()
        hit = np.zeros(len(counts) + 1, np.uint64)
        assert lib.awFmGpuLocatePacked(b.ptr, letters.ctypes.data, abi.AwFmGpuKmerAscii, None, length, len(counts),
                                       hit.ctypes.data, None, 0, None, None, C.byref(total)) == abi.AwFmSuccess
        pos = np.zeros(max(1, total.value), np.uint64)
        assert lib.awFmGpuLocatePacked(b.ptr, letters.ctypes.data, abi.AwFmGpuKmerAscii, None, length, len(counts),
                                       hit.ctypes.data, pos.ctypes.data, len(pos), None, None,
                                       C.byref(total)) == abi.AwFmSuccess
        assert np.array_equal(np.diff(hit).astype(np.uint32), r_counts)
        assert np.array_equal(pos[: total.value], np.concatenate(r_pos) if total.value else pos[:0])
        # the unchanged entry points run on the same device list
        sl = KmerSearchList(lib, len(counts)).fill(letters, fixed_len=length)
        parallel_search_count(lib, b.ptr, sl, 3)
        assert np.array_equal(sl.counts(), r_counts)
        assert parallel_search_locate(lib, b.ptr, sl, 3) == abi.AwFmSuccess
        assert all(np.array_equal(p, q) for p, q in zip(r_pos, sl.positions()))
        sl.close()
    finally:
        del os.environ["AWFM_GPU_DEVICES"]
        lib.awFmGpuReleaseIndex(b.ptr)


def test_stale_index_at_a_recycled_address_is_detected(reference, tmp_path):
    """awFmDeallocIndex + awFmCreateIndex of a same-length text commonly reuses every address; the reference API has no
    dealloc hook, so the drop-in must notice by content that its device copy is stale."""
    lib = capi.load()
    rng = np.random.default_rng(1)
    seen = set()
    for round_ in range(4):
        text = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 40000)]
        ptr = reference.create_index(text.tobytes(), str(tmp_path / "recycled.awfmi"), abi.AwFmAlphabetDna, 6, 4)
        seen.add(ptr.value)
        starts = rng.integers(0, len(text) - 12, 2000)
        letters = np.ascontiguousarray(text[starts[:, None] + np.arange(12)[None, :]].reshape(-1))
        r_counts = reference.count(ptr, letters, fixed_len=12, threads=1)
        assert (r_counts >= 1).all()
        sl = KmerSearchList(lib, len(r_counts)).fill(letters, fixed_len=12)
        parallel_search_count(lib, ptr, sl, 2)  # no awFmGpuReleaseIndex between the rounds, on purpose
        assert np.array_equal(sl.counts(), r_counts), round_
        sl.close()
        reference.dealloc_index(ptr)


def test_two_host_threads_search_one_index_concurrently(small_indexes, reference):
    """The reference's entry points are re-entrant for distinct lists on a shared const index
    (src/AwFmParallelSearch.c:95-220 keeps no global state); so are the drop-in's (one lane per call)."""
    import threading
    b = small_indexes["nuc_r8"]
    lib = capi.load()
    index_struct = b.arrays.as_awfm_index()
    ip = C.addressof(index_struct)
    assert lib.awFmGpuPrepareIndex(ip) == abi.AwFmSuccess
    jobs = []
    for t in range(4):
        letters, offsets = pack_queries(make_queries(b.text, False, seed=100 + t, num=4000, min_len=1, max_len=25,
                                                     seed_k=b.arrays.seed_k))
        rc, r_counts, r_pos = reference.locate(b.ptr, letters, offsets, threads=2)
        jobs.append((letters, offsets, r_counts, r_pos))
    errors = []

    def work(job):
        letters, offsets, r_counts, r_pos = job
        for _ in range(5):
            sl = KmerSearchList(lib, len(offsets) - 1).fill(letters, offsets)
            parallel_search_count(lib, ip, sl, 2)
            if not np.array_equal(sl.counts(), r_counts):
                errors.append("count")
            if parallel_search_locate(lib, ip, sl, 2) != abi.AwFmSuccess:
                errors.append("rc")
            if not all(np.array_equal(p, q) for p, q in zip(r_pos, sl.positions())):
                errors.append("positions")
            sl.close()

    threads = [threading.Thread(target=work, args=(j,)) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    lib.awFmGpuReleaseIndex(ip)


def test_short_openmp_team_is_handled(small_indexes, tmp_path):
    """OMP_THREAD_LIMIT=1: the runtime grants one thread whatever numThreads asks for; the engines must notice and
    still fill every count and position (the reference's `omp parallel for` degrades gracefully too)."""
    b = small_indexes["nuc_r8"]
    script = f"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, {ROOT!r})
from avxwindowfmindex_b200 import KmerSearchList, abi, capi, read_awfmi, parallel_search_count, parallel_search_locate
from oracle import harness
arrays = read_awfmi({b.path!r})
lib = capi.load()
ix = arrays.as_awfm_index()
rng = np.random.default_rng(0)
letters = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 12 * 300000)]
o_counts, _, _ = harness.Oracle(arrays).count(letters, fixed_len=12, threads=1)
sl = KmerSearchList(lib, 300000).fill(letters, fixed_len=12)
parallel_search_count(lib, C.addressof(ix), sl, 8)
assert lib.awFmGpuLastCountStatus() == abi.AwFmSuccess
assert np.array_equal(sl.counts(), o_counts), "counts differ under OMP_THREAD_LIMIT=1"
assert parallel_search_locate(lib, C.addressof(ix), sl, 8) == abi.AwFmSuccess
assert np.array_equal(sl.counts(), o_counts)
print("ok")
"""
    env = dict(os.environ, OMP_THREAD_LIMIT="1")
    out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("name", ["nuc_r8", "amino_r8"])
def test_locate_prepare_device_then_walk(small_indexes, name):
    """awfm_gpu_locate_prepare_device: search + ranges of the queries with hits only + hit offsets scanned from the
    counts, in one call on device buffers; followed by awfm_gpu_locate_device.  Sweep and tile path."""
    import torch
    b = small_indexes[name]
    length = b.arrays.seed_k + 3
    letters = fixed_queries(b, length, 7000, seed=31)
    n = len(letters) // length
    oracle = harness.Oracle(b.arrays)
    o_counts, _, _ = oracle.count(letters, fixed_len=length)
    o_hit, o_pos, _ = oracle.locate(letters, fixed_len=length)
    gpu = GpuIndex(b.arrays)
    stream = torch.cuda.current_stream().cuda_stream
    d_q = torch.from_numpy(letters).cuda()
    for sweep in (-1, 1):
        gpu.set_tuning(sweep_min_queries=sweep)
        d_c = torch.full((n,), -1, dtype=torch.int32, device="cuda")
        d_r = torch.full((n, 2), -7, dtype=torch.int64, device="cuda")  # entries of queries without hits stay untouched
        d_h = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
        gpu.locate_prepare_device(d_q.data_ptr(), QUERY_ASCII, length, n, d_c.data_ptr(), d_r.data_ptr(), d_h.data_ptr(), stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_c.cpu().numpy().astype(np.uint32), o_counts), (name, sweep)
        assert np.array_equal(d_h.cpu().numpy().astype(np.uint64), o_hit), (name, sweep)
        untouched = (d_r.cpu().numpy() == -7).all(axis=1)
        assert np.array_equal(untouched, o_counts == 0), "ranges must be written exactly for the queries with hits"
        total = int(o_hit[-1])
        d_p = torch.zeros(max(total, 1), dtype=torch.int64, device="cuda")
        gpu.locate_device(d_r.data_ptr(), d_h.data_ptr(), n, 0, total, d_p.data_ptr(), stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_p[:total].cpu().numpy().astype(np.uint64), o_pos), (name, sweep)
    gpu.close()
