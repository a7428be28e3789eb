"""Loader for csrc/libawfm_b200.so and prototypes of every symbol include/awfm_gpu.h and include/awfm_abi.h declare.

The library is built in-tree by `__graft_entry__.build()` (or `make -C avxwindowfmindex_b200/csrc`).  There is no
fallback of any kind: a missing library raises ImportError-like RuntimeError, a missing GPU makes calls fail.
"""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# AWFM_B200_LIB: development only — a variant build of the same library (tools/build_variant.sh) for A/B probes
LIB_PATH = os.environ.get("AWFM_B200_LIB") or os.path.join(_HERE, "csrc", "libawfm_b200.so")

GPU_SYMBOLS = [
    "awfm_gpu_last_error", "awfm_gpu_device_count", "awfm_gpu_ctx_create", "awfm_gpu_ctx_create_from_device",
    "awfm_gpu_ctx_destroy", "awfm_gpu_ctx_device_bytes", "awfm_gpu_ctx_get_stats", "awfm_gpu_ctx_set_tuning",
    "awfm_gpu_count_host", "awfm_gpu_locate_host", "awfm_gpu_count_device", "awfm_gpu_scan_ranges_device",
    "awfm_gpu_locate_device", "awfm_gpu_search_list_count", "awfm_gpu_search_list_locate",
    "awfm_gpu_gather_bandwidth", "awfm_gpu_build_index", "awfm_gpu_build_index_host", "awfm_gpu_built_view",
    "awfm_gpu_built_download", "awfm_gpu_built_destroy", "awfm_gpu_built_tie_rounds", "awfm_gpu_synth_letters", "awfm_gpu_set_l2_fetch_granularity",
    "awfm_gpu_ctx_create_from_file", "awfm_gpu_ctx_set_sequences", "awfm_gpu_ctx_extend_seed_table",
    "awfm_gpu_ctx_densify_suffix_array", "awfm_gpu_map_positions_device", "awfm_gpu_map_positions_host",
    "awfm_gpu_ctx_sweep_stage_ms", "awfm_gpu_ctx_sweep_live", "awfm_gpu_count_device_format", "awfm_gpu_locate_prepare_device",
    "awfm_gpu_group_create", "awfm_gpu_group_create_from_contexts", "awfm_gpu_group_destroy", "awfm_gpu_group_size",
    "awfm_gpu_group_context", "awfm_gpu_group_set_sequences", "awfm_gpu_group_set_tuning", "awfm_gpu_group_get_stats",
    "awfm_gpu_group_count", "awfm_gpu_group_locate", "awfm_gpu_group_search_list_count",
    "awfm_gpu_group_search_list_locate", "awfm_gpu_host_alloc", "awfm_gpu_host_free", "awfm_gpu_host_register",
    "awfm_gpu_host_unregister", "awfm_gpu_device_malloc", "awfm_gpu_device_free", "awfm_gpu_ipc_export",
    "awfm_gpu_ipc_open", "awfm_gpu_ipc_close", "awfm_gpu_peer_copy_async",
]
DROPIN_SYMBOLS = [
    "awFmCreateKmerSearchList", "awFmDeallocKmerSearchList", "awFmParallelSearchCount", "awFmParallelSearchLocate",
    "awFmGpuReleaseIndex", "awFmGpuPrepareIndex", "awFmGpuLastCountStatus", "awFmGpuGetLocalSequencePositions",
    "awFmGpuNumDevices", "awFmGpuCountPacked", "awFmGpuLocatePacked", "awFmGpuHostAlloc", "awFmGpuHostFree",
]

_lib = None


class AwfmGpuError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"awfm_gpu error {code}: {message}")
        self.code = code


def declare_search_list_api(lib):
    """Prototypes of the four reference entry points (src/AwFmIndex.h:308,326-327,364-367,400-403); works for the
    drop-in library and for the compiled reference alike."""
    lib.awFmCreateKmerSearchList.restype = C.POINTER(abi.AwFmKmerSearchList)
    lib.awFmCreateKmerSearchList.argtypes = [C.c_size_t]
    lib.awFmDeallocKmerSearchList.restype = None
    lib.awFmDeallocKmerSearchList.argtypes = [C.POINTER(abi.AwFmKmerSearchList)]
    lib.awFmParallelSearchCount.restype = None
    lib.awFmParallelSearchCount.argtypes = [C.c_void_p, C.POINTER(abi.AwFmKmerSearchList), C.c_uint32]
    lib.awFmParallelSearchLocate.restype = C.c_int
    lib.awFmParallelSearchLocate.argtypes = [C.c_void_p, C.POINTER(abi.AwFmKmerSearchList), C.c_uint32]
    return lib


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the search path.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    vp, u64, u32, i64 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64
    lib.awfm_gpu_last_error.restype = C.c_char_p
    lib.awfm_gpu_device_count.restype = C.c_int
    lib.awfm_gpu_ctx_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(abi.awfm_index_view)]
    lib.awfm_gpu_ctx_create_from_device.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(abi.awfm_index_view)]
    lib.awfm_gpu_ctx_destroy.argtypes = [vp]
    lib.awfm_gpu_ctx_destroy.restype = None
    lib.awfm_gpu_ctx_device_bytes.argtypes = [vp]
    lib.awfm_gpu_ctx_device_bytes.restype = u64
    lib.awfm_gpu_ctx_get_stats.argtypes = [vp, C.POINTER(abi.awfm_gpu_stats)]
    lib.awfm_gpu_ctx_set_tuning.argtypes = [vp, C.c_char_p, i64]
    lib.awfm_gpu_ctx_sweep_stage_ms.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
    lib.awfm_gpu_count_host.argtypes = [vp, vp, vp, u32, u64, vp, vp]
    lib.awfm_gpu_locate_host.argtypes = [vp, vp, vp, u32, u64, vp, vp, u64, vp]
    lib.awfm_gpu_count_device.argtypes = [vp, vp, vp, u32, u64, vp, vp, vp]
    lib.awfm_gpu_scan_ranges_device.argtypes = [vp, vp, u64, vp, vp]
    lib.awfm_gpu_locate_device.argtypes = [vp, vp, vp, u64, u64, u64, vp, vp]
    lib.awfm_gpu_search_list_count.argtypes = [vp, vp, u64, u32]
    lib.awfm_gpu_search_list_locate.argtypes = [vp, vp, u64, u32]
    lib.awfm_gpu_gather_bandwidth.argtypes = [C.c_int, u64, u32, u64, C.c_int, C.POINTER(C.c_double)]
    u8 = C.c_uint8
    lib.awfm_gpu_build_index.argtypes = [C.POINTER(vp), C.c_int, vp, u64, u8, u8, u8]
    lib.awfm_gpu_build_index_host.argtypes = [C.POINTER(vp), C.c_int, vp, u64, u8, u8, u8]
    lib.awfm_gpu_built_view.argtypes = [vp, C.POINTER(abi.awfm_index_view), C.POINTER(u64), C.POINTER(C.c_double)]
    lib.awfm_gpu_built_download.argtypes = [vp, vp, vp, vp, vp]
    lib.awfm_gpu_built_destroy.argtypes = [vp]
    lib.awfm_gpu_built_destroy.restype = None
    lib.awfm_gpu_built_tie_rounds.argtypes = [vp]
    lib.awfm_gpu_built_tie_rounds.restype = C.c_uint32
    lib.awfm_gpu_synth_letters.argtypes = [C.c_int, vp, u64, u64, u64, C.c_int]
    lib.awfm_gpu_set_l2_fetch_granularity.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.awfm_gpu_ctx_create_from_file.argtypes = [C.POINTER(vp), C.c_int, C.c_char_p, C.c_int, C.POINTER(abi.awfm_file_info)]
    lib.awfm_gpu_ctx_extend_seed_table.argtypes = [vp, u32, C.POINTER(C.c_double)]
    lib.awfm_gpu_ctx_densify_suffix_array.argtypes = [vp, u32, C.POINTER(C.c_double)]
    lib.awfm_gpu_ctx_set_sequences.argtypes = [vp, vp, u64]
    lib.awfm_gpu_map_positions_device.argtypes = [vp, vp, u64, vp, vp, vp]
    lib.awfm_gpu_map_positions_host.argtypes = [vp, vp, u64, vp, vp, C.POINTER(u64)]
    lib.awfm_gpu_ctx_sweep_live.argtypes = [vp, C.POINTER(u64), C.c_int, C.POINTER(u64)]
    lib.awfm_gpu_count_device_format.argtypes = [vp, vp, u32, vp, u32, u64, vp, vp, vp]
    lib.awfm_gpu_locate_prepare_device.argtypes = [vp, vp, u32, vp, u32, u64, vp, vp, vp, vp]
    lib.awfm_gpu_group_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.POINTER(abi.awfm_index_view)]
    lib.awfm_gpu_group_create_from_contexts.argtypes = [C.POINTER(vp), C.POINTER(vp), C.c_int]
    lib.awfm_gpu_group_destroy.argtypes = [vp]
    lib.awfm_gpu_group_destroy.restype = None
    lib.awfm_gpu_group_size.argtypes = [vp]
    lib.awfm_gpu_group_context.argtypes = [vp, C.c_int]
    lib.awfm_gpu_group_context.restype = vp
    lib.awfm_gpu_group_set_sequences.argtypes = [vp, vp, u64]
    lib.awfm_gpu_group_set_tuning.argtypes = [vp, C.c_char_p, i64]
    lib.awfm_gpu_group_get_stats.argtypes = [vp, C.POINTER(abi.awfm_gpu_stats)]
    lib.awfm_gpu_group_count.argtypes = [vp, vp, u32, vp, u32, u64, vp]
    lib.awfm_gpu_group_locate.argtypes = [vp, vp, u32, vp, u32, u64, vp, vp, u64, vp, vp, C.POINTER(u64)]
    lib.awfm_gpu_group_search_list_count.argtypes = [vp, vp, u64, u32]
    lib.awfm_gpu_group_search_list_locate.argtypes = [vp, vp, u64, u32]
    lib.awfm_gpu_host_alloc.argtypes = [C.POINTER(vp), u64]
    lib.awfm_gpu_host_free.argtypes = [vp]
    lib.awfm_gpu_host_free.restype = None
    lib.awfm_gpu_host_register.argtypes = [vp, u64]
    lib.awfm_gpu_host_unregister.argtypes = [vp]
    lib.awfm_gpu_device_malloc.argtypes = [C.c_int, C.POINTER(vp), u64]
    lib.awfm_gpu_device_free.argtypes = [C.c_int, vp]
    lib.awfm_gpu_ipc_export.argtypes = [C.c_int, vp, vp]
    lib.awfm_gpu_ipc_open.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.awfm_gpu_ipc_close.argtypes = [C.c_int, vp]
    lib.awfm_gpu_peer_copy_async.argtypes = [C.c_int, vp, vp, u64, vp]
    declare_search_list_api(lib)
    lib.awFmGpuNumDevices.argtypes = [vp]
    lib.awFmGpuNumDevices.restype = C.c_int
    lib.awFmGpuCountPacked.argtypes = [vp, vp, C.c_int, vp, u32, u64, vp]
    lib.awFmGpuCountPacked.restype = C.c_int
    lib.awFmGpuLocatePacked.argtypes = [vp, vp, C.c_int, vp, u32, u64, vp, vp, u64, vp, vp, C.POINTER(u64)]
    lib.awFmGpuLocatePacked.restype = C.c_int
    lib.awFmGpuHostAlloc.argtypes = [C.c_size_t]
    lib.awFmGpuHostAlloc.restype = vp
    lib.awFmGpuHostFree.argtypes = [vp]
    lib.awFmGpuHostFree.restype = None
    lib.awFmGpuGetLocalSequencePositions.argtypes = [vp, vp, C.c_size_t, vp, vp]
    lib.awFmGpuGetLocalSequencePositions.restype = C.c_int
    lib.awFmGpuReleaseIndex.argtypes = [vp]
    lib.awFmGpuReleaseIndex.restype = None
    lib.awFmGpuPrepareIndex.argtypes = [vp]
    lib.awFmGpuPrepareIndex.restype = C.c_int
    lib.awFmGpuLastCountStatus.restype = C.c_int
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise AwfmGpuError(code, load().awfm_gpu_last_error().decode("utf-8", "replace"))
