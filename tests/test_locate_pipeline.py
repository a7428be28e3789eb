"""The drop-in awFmParallelSearchLocate is a four-station pipeline over chunks of the list (pack -> ship -> walk ->
finish, csrc/awfm_b200.cu).  These tests push many small chunks, every packing path (mixed lengths, equal lengths
copied, equal lengths read in place from page-locked memory, a list that only looks contiguous) and the windowed
path of chunks with very many hits through it, with 1..8 host threads, and compare element-wise with the compiled
reference (src/AwFmParallelSearch.c:95-157, 315-387).  Run with -m gpu on a B200."""
import contextlib
import ctypes as C
import os

import numpy as np
import pytest

from avxwindowfmindex_b200 import KmerSearchList, abi, capi, parallel_search_count, parallel_search_locate
from avxwindowfmindex_b200.search import pack_queries
from conftest import make_queries

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def engine_env(**kv):
    """The shim reads its tuning from the environment when it uploads an index (awfm_dropin.c)."""
    old = {k: os.environ.get(k) for k in kv}
    os.environ.update({k: str(v) for k, v in kv.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def check_against_reference(sl, r_counts, r_pos):
    assert np.array_equal(sl.counts(), r_counts)
    mine = sl.positions()
    for i, p in enumerate(r_pos):
        assert np.array_equal(p, mine[i]), i
    e = sl.entries()
    n = sl.count
    # capacity = max(old capacity, count), grown by realloc to exactly count (src/AwFmParallelSearch.c:367-387)
    assert np.array_equal(e["capacity"][:n], np.maximum(4, e["count"][:n]))


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r16", "amino_r8"])
def test_many_small_chunks_mixed_lengths(small_indexes, reference, name):
    b = small_indexes[name]
    lib = capi.load()
    k = b.arrays.seed_k
    queries = make_queries(b.text, b.amino, seed=77, num=3000, min_len=1, max_len=k + 9, seed_k=k)
    letters, offsets = pack_queries(queries)
    rc, r_counts, r_pos = reference.locate(b.ptr, letters, offsets, threads=4)
    ix = b.arrays.as_awfm_index()
    ip = C.addressof(ix)
    for chunk, threads in ((64, 1), (64, 8), (200, 3), (1 << 18, 5)):
        with engine_env(AWFM_GPU_LOCATE_CHUNK_QUERIES=chunk, AWFM_GPU_CHUNK_QUERIES=chunk):
            sl = KmerSearchList(lib, len(queries)).fill(letters, offsets)
            assert parallel_search_locate(lib, ip, sl, threads) == rc == abi.AwFmSuccess
            check_against_reference(sl, r_counts, r_pos)
            # the list is reusable (src/AwFmIndex.h:344-345): count, then locate again with the grown capacities
            parallel_search_count(lib, ip, sl, threads)
            assert lib.awFmGpuLastCountStatus() == abi.AwFmSuccess
            assert np.array_equal(sl.counts(), r_counts)
            assert parallel_search_locate(lib, ip, sl, threads) == abi.AwFmSuccess
            check_against_reference(sl, r_counts, r_pos)
            sl.close()
            lib.awFmGpuReleaseIndex(ip)


@pytest.mark.parametrize("name", ["nuc_r8", "amino_r2"])
def test_windowed_path_for_chunks_with_many_hits(small_indexes, reference, name):
    """Short queries have thousands of hits each; with the inline limit at 0 every chunk that has hits is finished
    through windows of 37 flat hit indices, so position lists straddle many windows."""
    b = small_indexes[name]
    lib = capi.load()
    k = b.arrays.seed_k
    queries = make_queries(b.text, b.amino, seed=5, num=400, min_len=1, max_len=k + 2, seed_k=k)
    letters, offsets = pack_queries(queries)
    rc, r_counts, r_pos = reference.locate(b.ptr, letters, offsets, threads=4)
    assert int(r_counts.max()) > 100
    ix = b.arrays.as_awfm_index()
    ip = C.addressof(ix)
    for threads in (1, 6):
        with engine_env(AWFM_GPU_LOCATE_CHUNK_QUERIES=96, AWFM_GPU_LOCATE_INLINE_HITS=0,
                        AWFM_GPU_LOCATE_WINDOW_HITS=37):
            sl = KmerSearchList(lib, len(queries)).fill(letters, offsets)
            assert parallel_search_locate(lib, ip, sl, threads) == rc
            check_against_reference(sl, r_counts, r_pos)
            sl.close()
            lib.awFmGpuReleaseIndex(ip)
    # mixed: only the chunks above 2000 hits take the windows
    with engine_env(AWFM_GPU_LOCATE_CHUNK_QUERIES=64, AWFM_GPU_LOCATE_INLINE_HITS=2000,
                    AWFM_GPU_LOCATE_WINDOW_HITS=1000):
        sl = KmerSearchList(lib, len(queries)).fill(letters, offsets)
        assert parallel_search_locate(lib, ip, sl, 4) == rc
        check_against_reference(sl, r_counts, r_pos)
        sl.close()
        lib.awFmGpuReleaseIndex(ip)


def test_equal_length_queries_every_packing_path(small_indexes, reference):
    """Equal lengths: copied from pageable memory, read in place from page-locked memory, and a page-locked list
    whose probe entries (first / middle / last of a chunk) look contiguous while others are not."""
    import torch
    b = small_indexes["nuc_r16"]
    lib = capi.load()
    n, length = 5000, 11
    rng = np.random.default_rng(3)
    starts = rng.integers(0, len(b.text) - length, n)
    letters = b.text[starts[:, None] + np.arange(length)[None, :]].reshape(-1).copy()
    letters.reshape(n, length)[rng.integers(0, n, n // 3), rng.integers(0, length, n // 3)] = ord("G")
    rc, r_counts, r_pos = reference.locate(b.ptr, letters, fixed_len=length, threads=4)
    ix = b.arrays.as_awfm_index()
    ip = C.addressof(ix)
    pinned = torch.from_numpy(letters.copy()).pin_memory()
    with engine_env(AWFM_GPU_LOCATE_CHUNK_QUERIES=512, AWFM_GPU_CHUNK_QUERIES=512):
        for source in ("pageable", "pinned", "pinned_shuffled"):
            buf = letters if source == "pageable" else pinned.numpy()
            sl = KmerSearchList(lib, n).fill(buf, fixed_len=length)
            expect_counts, expect_pos = r_counts, r_pos
            if source == "pinned_shuffled":
                # swap two interior entries of every chunk: same strings, the list is no longer back to back
                e = sl.entries()
                perm = np.arange(n)
                for first in range(0, n - 512, 512):
                    perm[first + 10], perm[first + 20] = perm[first + 20], perm[first + 10]
                e["kmerString"][:n] = e["kmerString"][:n][perm]
                expect_counts = r_counts[perm]
                expect_pos = [r_pos[i] for i in perm]
            for threads in (1, 7):
                assert parallel_search_locate(lib, ip, sl, threads) == rc
                check_against_reference(sl, expect_counts, expect_pos)
                parallel_search_count(lib, ip, sl, threads)
                assert np.array_equal(sl.counts(), expect_counts)
            sl.close()
    lib.awFmGpuReleaseIndex(ip)


def test_caller_capacity_is_kept_and_grown_exactly(small_indexes, reference):
    """A caller may hand over lists with larger position lists; they are kept, smaller ones are grown to exactly
    `count` (src/AwFmParallelSearch.c:367-387)."""
    b = small_indexes["nuc_r8"]
    lib = capi.load()
    libc = C.CDLL(None)
    libc.realloc.restype = C.c_void_p
    libc.realloc.argtypes = [C.c_void_p, C.c_size_t]
    queries = make_queries(b.text, False, seed=9, num=300, min_len=2, max_len=9, seed_k=b.arrays.seed_k)
    letters, offsets = pack_queries(queries)
    rc, r_counts, r_pos = reference.locate(b.ptr, letters, offsets, threads=2)
    ix = b.arrays.as_awfm_index()
    ip = C.addressof(ix)
    sl = KmerSearchList(lib, len(queries)).fill(letters, offsets)
    e = sl.entries()
    roomy = np.arange(0, len(queries), 3)
    for i in roomy:  # 5000 slots: more than any of these queries' hit counts
        e["positionList"][i] = libc.realloc(int(e["positionList"][i]), 5000 * 8)
        e["capacity"][i] = 5000
    with engine_env(AWFM_GPU_LOCATE_CHUNK_QUERIES=64):
        assert parallel_search_locate(lib, ip, sl, 4) == rc
    assert np.array_equal(sl.counts(), r_counts)
    mine = sl.positions()
    assert all(np.array_equal(p, q) for p, q in zip(r_pos, mine))
    expected_capacity = np.maximum(4, r_counts)
    expected_capacity[roomy] = np.maximum(5000, r_counts[roomy])
    assert np.array_equal(sl.entries()["capacity"][: len(queries)], expected_capacity)
    sl.close()
    lib.awFmGpuReleaseIndex(ip)


@pytest.mark.parametrize("name", ["nuc_r8", "amino_r8"])
def test_count_pipeline_rounds_and_driver_thread(small_indexes, reference, name):
    """awFmParallelSearchCount in rounds (csrc/awfm_b200.cu: ship r-1 / pack r / scatter r-4): many chunks, teams small
    enough for the calling thread to pack as well (<= 4) and large enough for a dedicated driver thread, mixed lengths,
    equal lengths copied, equal lengths read in place from page-locked memory; the list is reused across calls."""
    import torch
    b = small_indexes[name]
    lib = capi.load()
    k = b.arrays.seed_k
    ix = b.arrays.as_awfm_index()
    ip = C.addressof(ix)
    mixed = pack_queries(make_queries(b.text, b.amino, seed=5, num=6000, min_len=1, max_len=k + 9, seed_k=k))
    length = k + 4
    rng = np.random.default_rng(9)
    starts = rng.integers(0, len(b.text) - length, 7001)
    fixed = np.concatenate([b.text[s:s + length] for s in starts]).astype(np.uint8)
    fixed[::53] = b.text[0]
    fixed_offsets = np.arange(0, len(fixed) + 1, length, dtype=np.uint64)
    pinned = torch.from_numpy(fixed.copy()).pin_memory()
    cases = [("mixed", mixed[0], mixed[1], {}), ("equal, copied", fixed, fixed_offsets, {}),
             ("equal, in place", pinned.numpy(), None, {"fixed_len": length})]
    for label, letters, offsets, kw in cases:
        r_counts = reference.count(b.ptr, np.asarray(letters), offsets if offsets is not None else fixed_offsets, threads=4)
        n = len(r_counts)
        for chunk, threads in ((512, 1), (512, 3), (512, 6), (640, 9), (1 << 16, 7)):
            with engine_env(AWFM_GPU_CHUNK_QUERIES=chunk):
                lib.awFmGpuReleaseIndex(ip)
                sl = KmerSearchList(lib, n).fill(letters, offsets, **kw)
                for rep in range(3):  # stale counts twice, then counts that are already right (left unwritten by the engine)
                    if rep < 2:
                        sl.entries()["count"][:n] = 0xDEAD
                    parallel_search_count(lib, ip, sl, threads)
                    assert lib.awFmGpuLastCountStatus() == abi.AwFmSuccess
                    assert np.array_equal(sl.counts(), r_counts), (name, label, chunk, threads)
                sl.close()
    lib.awFmGpuReleaseIndex(ip)


@pytest.mark.parametrize("n", [0, 1, 2047, 2048, 2049, 524_288, 524_289, 1_300_003])
def test_hit_offset_scan_against_numpy(small_indexes, n):
    """scanTileSums / scanTileBases / scanTileOffsets (csrc/awfm_kernels.cuh): exclusive scan of the u32-truncated range
    lengths (src/AwFmParallelSearch.c:328,367).  Sizes around one tile (2048 queries), around 256 tiles (where a thread of
    the base kernel starts owning more than one tile) and far beyond; empty ranges, ranges longer than 2^32 (truncated),
    ranges at the top of the 64-bit position space; from ranges (awfm_gpu_scan_ranges_device) and, through a search, from
    counts (awfm_gpu_locate_prepare_device)."""
    import torch
    from avxwindowfmindex_b200 import GpuIndex
    b = small_indexes["nuc_r8"]
    gpu = GpuIndex(b.arrays)
    rng = np.random.default_rng(n + 5)
    sp = rng.integers(1, 1 << 40, n, dtype=np.uint64)
    length = rng.integers(0, 9, n, dtype=np.uint64)               # 0 = empty range
    length[rng.random(n) < 0.01] = np.uint64((1 << 32) + 7)        # u32 truncation: counts as 7
    ep = sp + length - np.uint64(1)                                # empty: ep = sp - 1
    if n:
        sp[-1], ep[-1] = np.uint64(0xFFFFFFFFFFFFFFF0), np.uint64(0xFFFFFFFFFFFFFFF4)
        length[-1] = 5
    want = np.zeros(n + 1, dtype=np.uint64)
    want[1:] = np.cumsum(length & np.uint64(0xFFFFFFFF))
    ranges = np.stack([sp, ep], axis=1).astype(np.uint64) if n else np.zeros((0, 2), np.uint64)
    d_ranges = torch.zeros((max(n, 1), 2), dtype=torch.int64, device="cuda")
    d_ranges[:n] = torch.from_numpy(ranges.view(np.int64))
    d_hit = torch.full((n + 1,), -1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    gpu.scan_ranges_device(d_ranges.data_ptr(), n, d_hit.data_ptr(), st)
    torch.cuda.synchronize()
    assert np.array_equal(d_hit.cpu().numpy().view(np.uint64), want), n
    if n >= 2048:  # from counts: a real search over sampled 8-mers, offsets against the cumulative counts
        L = 8
        starts = rng.integers(0, len(b.text) - L, n)
        letters = b.text[starts[:, None] + np.arange(L)[None, :]].reshape(-1).copy()
        d_letters = torch.from_numpy(letters).cuda()
        d_counts = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_r = torch.zeros((n, 2), dtype=torch.int64, device="cuda")
        gpu.locate_prepare_device(d_letters.data_ptr(), 0, L, n, d_counts.data_ptr(), d_r.data_ptr(), d_hit.data_ptr(), st)
        torch.cuda.synchronize()
        counts = d_counts.cpu().numpy().view(np.uint32).astype(np.uint64)
        want = np.zeros(n + 1, dtype=np.uint64)
        want[1:] = np.cumsum(counts)
        assert np.array_equal(d_hit.cpu().numpy().view(np.uint64), want), n
    gpu.close()
