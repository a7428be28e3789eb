"""The C-ABI library loads on a CPU-only box and exports every function include/*.h declares (no compute calls)."""
import ctypes as C
import os
import re

from avxwindowfmindex_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(awfm_gpu_[a-z0-9_]+|awFm[A-Za-z]+)\s*\(", text)
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(capi.LIB_PATH)
    gpu = declared_functions("awfm_gpu.h")
    dropin = declared_functions("awfm_abi.h")
    assert set(gpu) == set(capi.GPU_SYMBOLS)
    assert set(dropin) == set(capi.DROPIN_SYMBOLS)
    for name in gpu + dropin:
        assert hasattr(lib, name), name


def test_no_device_means_loud_failure():
    """Without a GPU the product must fail, not fall back (this test only runs its assertion on GPU-less boxes)."""
    import numpy as np
    import pytest
    from avxwindowfmindex_b200 import GpuIndex, read_awfmi
    lib = capi.load()
    if lib.awfm_gpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    arrays = read_awfmi(os.path.join(ROOT, "tests", "golden", "nuc_k4_r4.awfmi"))
    with pytest.raises(capi.AwfmGpuError):
        GpuIndex(arrays)
    # the reference-facing entry point reports failure too: Locate returns a failure code
    from avxwindowfmindex_b200 import KmerSearchList, parallel_search_locate
    ix = arrays.as_awfm_index()
    sl = KmerSearchList(lib, 2).fill(np.frombuffer(b"acgtacgt", np.uint8), fixed_len=4)
    assert parallel_search_locate(lib, C.addressof(ix), sl, 1) < 0
    sl.close()


def test_search_list_allocation_semantics():
    """awFmCreateKmerSearchList / awFmDeallocKmerSearchList (src/AwFmParallelSearch.c:36-93): capacity, count 0,
    one 4-slot position list per entry."""
    from avxwindowfmindex_b200 import KmerSearchList
    sl = KmerSearchList(capi.load(), 17)
    e = sl.entries()
    assert sl.ptr.contents.capacity == 17 and sl.ptr.contents.count == 0
    assert (e["capacity"] == 4).all() and (e["count"] == 0).all() and (e["positionList"] != 0).all()
    assert (e["kmerString"] == 0).all() and (e["kmerLength"] == 0).all()
    sl.close()


def test_every_tuning_key_is_documented_in_the_header():
    """awfm_gpu_ctx_set_tuning's keys (csrc/awfm_b200.cu) and the list in include/awfm_gpu.h must not drift apart."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "avxwindowfmindex_b200", "csrc", "awfm_b200.cu")).read()
    body = src[src.index('extern "C" int awfm_gpu_ctx_set_tuning'):]
    body = body[:body.index("unknown tuning key")]
    keys = set(re.findall(r'k == "([a-z0-9_]+)"', body))
    assert len(keys) >= 15
    header = open(os.path.join(root, "include", "awfm_gpu.h")).read()
    missing = sorted(k for k in keys if f'"{k}"' not in header)
    assert not missing, f"tuning keys not documented in include/awfm_gpu.h: {missing}"
