"""Derived structures (awfm_gpu_ctx_extend_seed_table / awfm_gpu_ctx_densify_suffix_array): they change how many
dependent DRAM round trips a query costs, never a result.  Every output — ranges as the reference leaves them
(including the stored invalid pair), counts, positions in SA order — must stay bit-exact against the oracle for
every depth / ratio, every kernel variant, and queries that cannot use the deep table (short, ambiguous)."""
import numpy as np
import pytest

from avxwindowfmindex_b200 import GpuIndex, capi
from avxwindowfmindex_b200.search import pack_queries
from conftest import make_queries
from oracle import harness

pytestmark = pytest.mark.gpu


def queries_for(b, seed, num=700):
    k = b.arrays.seed_k
    qs = make_queries(b.text, b.amino, seed, num, 1, k + 9, k)
    # ambiguity letter at every distance from the end, around the deep-table window
    amb = b"X" if b.amino else b"N"
    base = bytes(b.text[11:11 + k + 8])
    for d in range(0, k + 6):
        q = bytearray(base)
        q[len(q) - 1 - d] = amb[0]
        qs.append(bytes(q))
    return qs


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r3", "nuc_r255", "amino_r8", "amino_r2"])
def test_deep_seed_table_is_invisible_in_results(small_indexes, name):
    b = small_indexes[name]
    letters, offsets = pack_queries(queries_for(b, seed=5))
    oracle = harness.Oracle(b.arrays)
    o_counts, o_ranges, _ = oracle.count(letters, offsets)
    o_hit, o_pos, _ = oracle.locate(letters, offsets)
    gpu = GpuIndex(b.arrays)
    base_bytes = gpu.device_bytes()
    k = b.arrays.seed_k
    card = 20 if b.amino else 4
    for depth in (k + 1, k + 2, k + 4 if not b.amino else k + 2):
        ms = gpu.extend_seed_table(depth)
        assert ms > 0 and gpu.device_bytes() == base_bytes + card ** depth * 8
        for variant in (0, 1):
            for lpq in (1, 2, 4):
                gpu.set_tuning(count_variant=variant, count_lpq=lpq)
                counts, ranges = gpu.count(letters, offsets, want_ranges=True)
                assert np.array_equal(counts, o_counts), (name, depth, variant, lpq)
                assert np.array_equal(ranges, o_ranges), (name, depth, variant, lpq)
        hit, pos = gpu.locate(letters, offsets)
        assert np.array_equal(hit, o_hit) and np.array_equal(pos, o_pos)
        # A/B switch keeps the table but routes around it
        gpu.set_tuning(use_deep_seed_table=0)
        assert np.array_equal(gpu.count(letters, offsets), o_counts)
        gpu.set_tuning(use_deep_seed_table=1)
    gpu.extend_seed_table(0)
    assert gpu.device_bytes() == base_bytes
    assert np.array_equal(gpu.count(letters, offsets), o_counts)
    with pytest.raises(capi.AwfmGpuError):
        gpu.extend_seed_table(40)
    gpu.close()


def test_deep_seed_table_entries_equal_stepped_ranges(small_indexes):
    """every entry of the depth-(k+2) table, read back through fixed-length queries that enumerate all (k+2)-mers,
    equals the oracle's range for that k-mer (the reference's seed + 2 steps with stop-on-invalid)"""
    b = small_indexes["nuc_r8"]  # seed k = 5
    k2 = b.arrays.seed_k + 2
    n = 4 ** k2
    idx = np.arange(n, dtype=np.int64)
    digits = (idx[:, None] // (4 ** np.arange(k2 - 1, -1, -1))[None, :]) % 4
    letters = np.frombuffer(b"ACGT", np.uint8)[digits].reshape(-1).copy()
    oracle = harness.Oracle(b.arrays)
    o_counts, o_ranges, _ = oracle.count(letters, fixed_len=k2, threads=4)
    gpu = GpuIndex(b.arrays)
    gpu.extend_seed_table(k2)
    counts, ranges = gpu.count(letters, fixed_len=k2, want_ranges=True)
    assert np.array_equal(ranges, o_ranges) and np.array_equal(counts, o_counts)
    assert (o_ranges[:, 0] > o_ranges[:, 1]).any()  # the stored invalid pairs are part of the comparison
    gpu.close()


@pytest.mark.parametrize("name", ["nuc_r8", "nuc_r16", "nuc_r200", "nuc_r255", "nuc_r3", "amino_r8", "amino_r2"])
def test_dense_suffix_array_is_invisible_in_results(small_indexes, name):
    b = small_indexes[name]
    letters, offsets = pack_queries(queries_for(b, seed=6))
    oracle = harness.Oracle(b.arrays)
    o_hit, o_pos, _ = oracle.locate(letters, offsets)
    gpu = GpuIndex(b.arrays)
    base_bytes = gpu.device_bytes()
    for new_ratio in (1, 2, 5):
        if new_ratio >= b.arrays.sa_ratio:
            continue
        ms = gpu.densify_suffix_array(new_ratio)
        assert ms > 0 and gpu.device_bytes() > base_bytes
        for variant in (0, 1):
            for lpq in (1, 2, 4):
                gpu.set_tuning(locate_variant=variant, locate_lpq=lpq)
                hit, pos = gpu.locate(letters, offsets)
                assert np.array_equal(hit, o_hit) and np.array_equal(pos, o_pos), (name, new_ratio, variant, lpq)
    gpu.densify_suffix_array(0)
    assert gpu.device_bytes() == base_bytes
    hit, pos = gpu.locate(letters, offsets)
    assert np.array_equal(pos, o_pos)
    gpu.close()


def test_dense_suffix_array_holds_the_true_suffix_array(small_indexes):
    """ratio 1: locating every single-position range [p, p] returns SA[p]; compare with the suffix array of the text"""
    b = small_indexes["nuc_r16"]
    text = bytes(b.text).lower().replace(b"n", b"x") + b"$"
    order = {c: i for i, c in enumerate(b"$acgtx")}
    ranks = np.array([order[c] for c in text], dtype=np.int64)
    n = len(text)
    sa = np.array(sorted(range(n), key=lambda i: ranks[i:].tobytes()), dtype=np.uint64)  # byte order == rank order
    import torch
    gpu = GpuIndex(b.arrays)
    gpu.densify_suffix_array(1)
    d_ranges = torch.stack([torch.arange(n), torch.arange(n)], dim=1).to(torch.int64).cuda().contiguous()
    d_hit = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    d_pos = torch.zeros(n, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    gpu.scan_ranges_device(d_ranges.data_ptr(), n, d_hit.data_ptr(), st)
    gpu.locate_device(d_ranges.data_ptr(), d_hit.data_ptr(), n, 0, n, d_pos.data_ptr(), st)
    torch.cuda.synchronize()
    assert np.array_equal(d_pos.cpu().numpy().astype(np.uint64), sa)
    gpu.close()


def test_both_together_on_the_drop_in_env(small_indexes, reference, monkeypatch):
    """AWFM_GPU_SEED_DEPTH / AWFM_GPU_SA_RATIO make the drop-in derive both at upload; the reference's own outputs
    are still reproduced"""
    import ctypes as C
    from avxwindowfmindex_b200 import KmerSearchList, abi, parallel_search_locate
    b = small_indexes["nuc_r8"]
    lib = capi.load()
    letters, offsets = pack_queries(queries_for(b, seed=7))
    rc, r_counts, r_pos = reference.locate(b.ptr, letters, offsets, threads=2)
    monkeypatch.setenv("AWFM_GPU_SEED_DEPTH", str(b.arrays.seed_k + 3))
    monkeypatch.setenv("AWFM_GPU_SA_RATIO", "1")
    ix = b.arrays.as_awfm_index()
    sl = KmerSearchList(lib, len(offsets) - 1).fill(letters, offsets)
    assert parallel_search_locate(lib, C.addressof(ix), sl, 2) == abi.AwFmSuccess
    assert np.array_equal(sl.counts(), r_counts)
    assert all(np.array_equal(p, q) for p, q in zip(r_pos, sl.positions()))
    sl.close()
    lib.awFmGpuReleaseIndex(C.addressof(ix))
