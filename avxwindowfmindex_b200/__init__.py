"""avxwindowfmindex_b200 — B200-native (sm_100a) batched exact-match k-mer search for AwFmIndex.

Only the hot path of TravisWheelerLab/AvxWindowFmIndex is here: awFmParallelSearchCount / awFmParallelSearchLocate
over an AwFmKmerSearchList, on an unchanged `.awfmi` index.  The product is csrc/libawfm_b200.so (hand-written
CUDA + C-ABI, include/awfm_gpu.h, include/awfm_abi.h); the Python modules are a ctypes veneer for tests and bench.
"""
from . import abi, build_index, capi, index, search, synth  # noqa: F401
from .build_index import DeviceBuiltIndex  # noqa: F401
from .index import IndexArrays, read_awfmi, write_awfmi  # noqa: F401
from .search import (GpuGroup, GpuIndex, KmerSearchList, PinnedArray, pack_queries_bits, parallel_search_count,  # noqa: F401
                     parallel_search_locate)
