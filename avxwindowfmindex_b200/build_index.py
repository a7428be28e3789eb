"""Device-side index construction: wraps awfm_gpu_build_index* of include/awfm_gpu.h.

Produces, on the GPU, the same arrays awFmCreateIndex builds on the CPU (src/AwFmCreate.c:31-450): raw BWT blocks,
prefix sums, seed table, bit-packed sampled SA.  `to_host()` gives an IndexArrays that write_awfmi() turns into an
`.awfmi` file the reference reads; `gpu_index()` makes the arrays searchable without leaving the device.
"""
import ctypes as C

import numpy as np

from . import abi, capi
from .index import IndexArrays, aligned_empty
from .search import GpuIndex


class DeviceBuiltIndex:
    def __init__(self, handle, device):
        self.lib = capi.load()
        self.handle = handle
        self.device = device
        self._view = abi.awfm_index_view()
        ties, ms = C.c_uint64(), C.c_double()
        capi.check(self.lib.awfm_gpu_built_view(handle, C.byref(self._view), C.byref(ties), C.byref(ms)))
        self.tie_suffixes = int(ties.value)
        self.tie_rounds = int(self.lib.awfm_gpu_built_tie_rounds(handle))
        self.build_ms = float(ms.value)

    @classmethod
    def from_host_text(cls, text, alphabet=abi.AwFmAlphabetDna, seed_k=12, sa_ratio=8, device=0):
        lib = capi.load()
        text = np.ascontiguousarray(np.frombuffer(bytes(text), dtype=np.uint8) if not isinstance(text, np.ndarray) else text)
        h = C.c_void_p()
        capi.check(lib.awfm_gpu_build_index_host(C.byref(h), device, text.ctypes.data, len(text), alphabet, seed_k, sa_ratio))
        return cls(h, device)

    @classmethod
    def from_device_text(cls, d_text_ptr, n, alphabet=abi.AwFmAlphabetDna, seed_k=12, sa_ratio=8, device=0):
        lib = capi.load()
        h = C.c_void_p()
        capi.check(lib.awfm_gpu_build_index(C.byref(h), device, d_text_ptr, n, alphabet, seed_k, sa_ratio))
        return cls(h, device)

    @property
    def view(self):
        return self._view

    def to_host(self) -> IndexArrays:
        v = self._view
        amino = v.alphabet == abi.AwFmAlphabetAmino
        card = 20 if amino else 4
        blocks = aligned_empty(v.numBlocks * (abi.AMINO_BLOCK_BYTES if amino else abi.NUC_BLOCK_BYTES))
        prefix = np.zeros(card + 2, dtype=np.uint64)
        seeds = np.zeros((card ** v.seedK, 2), dtype=np.uint64)
        sa = np.zeros(v.saByteLength, dtype=np.uint8)
        capi.check(self.lib.awfm_gpu_built_download(self.handle, blocks.ctypes.data, prefix.ctypes.data,
                                                    seeds.ctypes.data, sa.ctypes.data))
        return IndexArrays(int(v.alphabet), int(v.seedK), int(v.saRatio), int(v.bwtLength), blocks, prefix, seeds, sa)

    def gpu_index(self) -> GpuIndex:
        return GpuIndex.from_device_view(self._view, self.device)

    def close(self):
        if self.handle:
            self.lib.awfm_gpu_built_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
