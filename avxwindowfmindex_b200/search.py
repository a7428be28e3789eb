"""Host-side mirror of the reference's batched search interface over the B200 C-ABI.

Two levels, both thin:
  * `GpuIndex`       — an index resident in HBM (awfm_gpu_ctx) with the packed-batch calls of include/awfm_gpu.h.
  * `KmerSearchList` + `parallel_search_count` / `parallel_search_locate` — the reference's own calling
    convention (src/AwFmIndex.h:308-403): a list of {kmerString, kmerLength, positionList, count, capacity}
    entries handed to awFmParallelSearchCount / awFmParallelSearchLocate.  `lib` may be the drop-in
    (capi.load()) or the compiled reference, so parity tests run the same code against both.
"""
import ctypes as C

import numpy as np

from . import abi, capi
from .index import IndexArrays


def pack_queries(queries):
    """list of bytes -> (letters uint8, offsets uint64[n+1])"""
    lengths = np.fromiter((len(q) for q in queries), dtype=np.uint64, count=len(queries))
    offsets = np.zeros(len(queries) + 1, dtype=np.uint64)
    np.cumsum(lengths, out=offsets[1:])
    letters = np.frombuffer(b"".join(queries), dtype=np.uint8).copy() if len(queries) else np.zeros(0, np.uint8)
    return letters, offsets


def _ptr(a):
    return None if a is None else a.ctypes.data


class GpuIndex:
    """Device-resident index.  Construction uploads and re-lays-out the arrays (see csrc/awfm_device.cuh)."""

    def __init__(self, arrays: IndexArrays, device=0):
        self.lib = capi.load()
        self.arrays = arrays
        self.device = device
        self._ctx = C.c_void_p()
        view = arrays.view()
        capi.check(self.lib.awfm_gpu_ctx_create(C.byref(self._ctx), device, C.byref(view)))
        if arrays.fasta_metadata is not None and len(arrays.fasta_metadata):
            self.set_sequences(arrays.fasta_metadata)

    @classmethod
    def from_device_view(cls, view, device=0):
        """Index whose arrays already live on `device` (awfm_index_view of DEVICE pointers, e.g. from build_index)."""
        self = cls.__new__(cls)
        self.lib = capi.load()
        self.arrays = None
        self.device = device
        self._ctx = C.c_void_p()
        capi.check(self.lib.awfm_gpu_ctx_create_from_device(C.byref(self._ctx), device, C.byref(view)))
        return self

    @classmethod
    def from_file(cls, path, device=0, want_suffix_array=True):
        """Index loaded from an `.awfmi` file straight into HBM (awfm_gpu_ctx_create_from_file); `.info` holds the
        header fields."""
        self = cls.__new__(cls)
        self.lib = capi.load()
        self.arrays = None
        self.device = device
        self._ctx = C.c_void_p()
        self.info = abi.awfm_file_info()
        capi.check(self.lib.awfm_gpu_ctx_create_from_file(C.byref(self._ctx), device, str(path).encode(),
                                                          int(want_suffix_array), C.byref(self.info)))
        return self

    @property
    def ctx(self):
        return self._ctx

    def close(self):
        if self._ctx:
            self.lib.awfm_gpu_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_sequences(self, fasta_metadata):
        """Record table of a multi-sequence index: uint64 (numSequences, 2) = (headerEnd, sequenceEnd) per record, the
        reference's struct FastaVectorMetadata array as stored in the `.awfmi` file (IndexArrays.fasta_metadata)."""
        meta = np.ascontiguousarray(fasta_metadata, dtype=np.uint64).reshape(-1, 2)
        capi.check(self.lib.awfm_gpu_ctx_set_sequences(self._ctx, _ptr(meta), len(meta)))

    def map_positions(self, positions):
        """Batched awFmGetLocalSequencePositionFromIndexPosition: (sequence_index, local_position, num_illegal)."""
        positions = np.ascontiguousarray(positions, dtype=np.uint64)
        seq = np.zeros(len(positions), np.uint64)
        loc = np.zeros(len(positions), np.uint64)
        bad = C.c_uint64()
        capi.check(self.lib.awfm_gpu_map_positions_host(self._ctx, _ptr(positions), len(positions), _ptr(seq), _ptr(loc),
                                                        C.byref(bad)))
        return seq, loc, int(bad.value)

    def map_positions_device(self, d_positions, n, d_sequence_index, d_local_position, stream=0):
        capi.check(self.lib.awfm_gpu_map_positions_device(self._ctx, d_positions, n, d_sequence_index, d_local_position,
                                                          stream or None))

    def extend_seed_table(self, depth):
        """Derive a deeper seed table on the device (0 or <= seed k drops it); returns the build time in ms."""
        ms = C.c_double()
        capi.check(self.lib.awfm_gpu_ctx_extend_seed_table(self._ctx, int(depth), C.byref(ms)))
        return ms.value

    def densify_suffix_array(self, new_ratio):
        """Derive SA samples at every new_ratio-th BWT position (0 drops them); returns the build time in ms."""
        ms = C.c_double()
        capi.check(self.lib.awfm_gpu_ctx_densify_suffix_array(self._ctx, int(new_ratio), C.byref(ms)))
        return ms.value

    def set_tuning(self, **kv):
        for k, v in kv.items():
            capi.check(self.lib.awfm_gpu_ctx_set_tuning(self._ctx, k.encode(), int(v)))

    def sweep_stage_ms(self):
        """Device time (ms) of every stage of the most recent sweep count call made under sweep_profile=1."""
        ms = (C.c_double * 32)()
        n = self.lib.awfm_gpu_ctx_sweep_stage_ms(self._ctx, ms, 32)
        return [ms[i] for i in range(max(n, 0))]

    def sweep_live(self):
        """(live, irregular): live[0] = queries, live[p] = queries still searching after LF step p, of the most recent
        sweep count call ([] when it took the tile kernel)."""
        live = (C.c_uint64 * 32)()
        irregular = C.c_uint64()
        n = self.lib.awfm_gpu_ctx_sweep_live(self._ctx, live, 32, C.byref(irregular))
        return [int(live[i]) for i in range(max(n, 0))], int(irregular.value)

    def count_device_format(self, d_queries, fmt, fixed_len, n, d_counts, d_ranges=None, stream=0):
        capi.check(self.lib.awfm_gpu_count_device_format(self._ctx, d_queries, fmt, None, fixed_len, n, d_counts,
                                                         d_ranges or None, stream or None))

    def device_bytes(self):
        return int(self.lib.awfm_gpu_ctx_device_bytes(self._ctx))

    def stats(self):
        s = abi.awfm_gpu_stats()
        capi.check(self.lib.awfm_gpu_ctx_get_stats(self._ctx, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    # ---- packed batch, host buffers ----
    def count(self, letters, offsets=None, fixed_len=0, want_ranges=False):
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        n = (len(offsets) - 1) if offsets is not None else (len(letters) // fixed_len if fixed_len else 0)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        counts = np.zeros(n, dtype=np.uint32)
        ranges = np.zeros((n, 2), dtype=np.uint64) if want_ranges else None
        capi.check(self.lib.awfm_gpu_count_host(self._ctx, _ptr(letters), _ptr(offsets), fixed_len, n,
                                                _ptr(counts), _ptr(ranges)))
        return (counts, ranges) if want_ranges else counts

    def locate(self, letters, offsets=None, fixed_len=0, want_ranges=False):
        """CSR result: (hit_offsets uint64[n+1], positions uint64[total]) in SA order within each query."""
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        n = (len(offsets) - 1) if offsets is not None else (len(letters) // fixed_len if fixed_len else 0)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        hit_offsets = np.zeros(n + 1, dtype=np.uint64)
        ranges = np.zeros((n, 2), dtype=np.uint64) if want_ranges else None
        # one call when the hits fit a generous first guess (the C-ABI fills hitOffsets before it checks the capacity),
        # a second one with the exact size otherwise
        guess = max(4 * n, 1 << 12)
        positions = np.zeros(guess, dtype=np.uint64)
        rc = self.lib.awfm_gpu_locate_host(self._ctx, _ptr(letters), _ptr(offsets), fixed_len, n, _ptr(hit_offsets),
                                           _ptr(positions), guess, _ptr(ranges))
        total = int(hit_offsets[n]) if n else 0
        if rc != 0 and total > guess:
            positions = np.zeros(total, dtype=np.uint64)
            rc = self.lib.awfm_gpu_locate_host(self._ctx, _ptr(letters), _ptr(offsets), fixed_len, n,
                                               _ptr(hit_offsets), _ptr(positions), total, _ptr(ranges))
        capi.check(rc)
        positions = positions[:total].copy() if total < len(positions) else positions
        return (hit_offsets, positions, ranges) if want_ranges else (hit_offsets, positions)

    # ---- packed batch, device buffers (raw device pointers, e.g. torch tensors' data_ptr()) ----
    def count_device(self, d_letters, d_offsets, fixed_len, n, d_counts, d_ranges=None, stream=0):
        capi.check(self.lib.awfm_gpu_count_device(self._ctx, d_letters, d_offsets or None, fixed_len, n, d_counts,
                                                  d_ranges or None, stream or None))

    def locate_prepare_device(self, d_queries, fmt, fixed_len, n, d_counts, d_ranges, d_hit_offsets, stream=0):
        """search + hit offsets in one call; d_ranges is only written for queries with hits"""
        capi.check(self.lib.awfm_gpu_locate_prepare_device(self._ctx, d_queries, fmt, None, fixed_len, n, d_counts,
                                                           d_ranges, d_hit_offsets, stream or None))

    def scan_ranges_device(self, d_ranges, n, d_hit_offsets, stream=0):
        capi.check(self.lib.awfm_gpu_scan_ranges_device(self._ctx, d_ranges, n, d_hit_offsets, stream or None))

    def locate_device(self, d_ranges, d_hit_offsets, n, hit_begin, hit_end, d_positions, stream=0):
        capi.check(self.lib.awfm_gpu_locate_device(self._ctx, d_ranges, d_hit_offsets, n, hit_begin, hit_end,
                                                   d_positions, stream or None))


QUERY_ASCII, QUERY_2BIT, QUERY_5BIT = 0, 2, 5  # include/awfm_gpu.h: awfm_query_format


def pack_queries_bits(letters, length, amino=False):
    """ASCII fixed-length queries -> the 2-bit (nucleotide) / 5-bit (amino) packed format of include/awfm_gpu.h:
    query i occupies bytes [i*B, (i+1)*B), B = ceil(length*bits/8), letter j in bits [j*bits, (j+1)*bits) of that
    little-endian byte string.  Nucleotide letters must be A/C/G/T/U (either case); amino letters outside the 20
    standard ones become code 20 (ambiguity).  Test/bench helper (numpy), not on the product path."""
    letters = np.ascontiguousarray(letters, dtype=np.uint8).reshape(-1, length)
    n = letters.shape[0]
    if amino:
        bits = 5
        table = np.full(256, 20, dtype=np.uint8)
        for i, ch in enumerate(b"ACDEFGHIKLMNPQRSTVWY"):
            table[ch] = i
            table[ch | 0x20] = i
    else:
        bits = 2
        table = np.full(256, 255, dtype=np.uint8)
        for i, chs in enumerate((b"Aa", b"Cc", b"Gg", b"TtUu")):
            for ch in chs:
                table[ch] = i
    codes = table[letters]
    if not amino and (codes == 255).any():
        raise ValueError("the 2-bit format holds A/C/G/T(U) only")
    nbytes = (length * bits + 7) // 8
    out = np.zeros((n, nbytes), dtype=np.uint8)
    step = max(1, (1 << 22) // max(length, 1))
    for a in range(0, n, step):  # bounded temporaries
        c = codes[a:a + step].astype(np.uint64)
        acc = np.zeros((c.shape[0], (length * bits + 63) // 64), dtype=np.uint64)
        for j in range(length):
            bit = j * bits
            w, sh = bit // 64, bit % 64
            acc[:, w] |= c[:, j] << np.uint64(sh)
            if sh + bits > 64:
                acc[:, w + 1] |= c[:, j] >> np.uint64(64 - sh)
        out[a:a + step] = acc.view(np.uint8).reshape(c.shape[0], -1)[:, :nbytes]
    return out.reshape(-1)


class PinnedArray:
    """numpy view of page-locked host memory from awfm_gpu_host_alloc (every GPU of the box can DMA from / to it)."""

    def __init__(self, shape, dtype):
        self.lib = capi.load()
        self.dtype = np.dtype(dtype)
        self.shape = (shape,) if np.isscalar(shape) else tuple(shape)
        nbytes = max(16, int(np.prod(self.shape)) * self.dtype.itemsize)
        p = C.c_void_p()
        capi.check(self.lib.awfm_gpu_host_alloc(C.byref(p), nbytes))
        self.ptr = p.value
        buf = (C.c_uint8 * nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def close(self):
        if self.ptr:
            self.array = None
            self.lib.awfm_gpu_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuGroup:
    """A device group (awfm_gpu_group): one call fanned out over several GPUs from this process, and the pipelined
    packed-batch engine.  Built from host arrays (replicated onto `devices`) or from existing GpuIndex objects."""

    def __init__(self, arrays: IndexArrays = None, devices=None, indexes=None):
        self.lib = capi.load()
        self._g = C.c_void_p()
        self._keep = indexes
        if indexes is not None:
            ctxs = (C.c_void_p * len(indexes))(*[ix.ctx for ix in indexes])
            capi.check(self.lib.awfm_gpu_group_create_from_contexts(C.byref(self._g), ctxs, len(indexes)))
        else:
            view = arrays.view()
            self._arrays = arrays
            if devices is None:
                capi.check(self.lib.awfm_gpu_group_create(C.byref(self._g), None, 0, C.byref(view)))
            else:
                devs = (C.c_int * len(devices))(*devices)
                capi.check(self.lib.awfm_gpu_group_create(C.byref(self._g), devs, len(devices), C.byref(view)))
            if arrays.fasta_metadata is not None and len(arrays.fasta_metadata):
                self.set_sequences(arrays.fasta_metadata)

    @property
    def handle(self):
        return self._g

    @property
    def size(self):
        return int(self.lib.awfm_gpu_group_size(self._g))

    def close(self):
        if self._g:
            self.lib.awfm_gpu_group_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_sequences(self, fasta_metadata):
        meta = np.ascontiguousarray(fasta_metadata, dtype=np.uint64).reshape(-1, 2)
        capi.check(self.lib.awfm_gpu_group_set_sequences(self._g, _ptr(meta), len(meta)))

    def set_tuning(self, **kv):
        for k, v in kv.items():
            capi.check(self.lib.awfm_gpu_group_set_tuning(self._g, k.encode(), int(v)))

    def stats(self):
        s = abi.awfm_gpu_stats()
        capi.check(self.lib.awfm_gpu_group_get_stats(self._g, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    @staticmethod
    def _n(queries, fmt, offsets, fixed_len):
        if offsets is not None:
            return len(offsets) - 1
        bits = {QUERY_ASCII: 8, QUERY_2BIT: 2, QUERY_5BIT: 5}[fmt]
        qb = (fixed_len * bits + 7) // 8
        return len(queries) // qb if qb else 0

    def count(self, queries, fmt=QUERY_ASCII, offsets=None, fixed_len=0, out=None):
        """awfm_gpu_group_count: `queries` uint8 (numpy, page-locked or not), returns uint32 counts (`out` if given)."""
        n = self._n(queries, fmt, offsets, fixed_len)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        counts = out if out is not None else np.zeros(n, dtype=np.uint32)
        capi.check(self.lib.awfm_gpu_group_count(self._g, queries.ctypes.data, fmt, _ptr(offsets), fixed_len, n,
                                                 counts.ctypes.data))
        return counts

    def locate(self, queries, fmt=QUERY_ASCII, offsets=None, fixed_len=0, mapped=False, out=None):
        """awfm_gpu_group_locate: (hit_offsets, positions[, sequence_index, local_position]).  `out` = (hit_offsets,
        positions[, seq, loc]) preallocated arrays (e.g. page-locked) to be filled instead of fresh ones."""
        n = self._n(queries, fmt, offsets, fixed_len)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        total = C.c_uint64()
        if out is None:
            hit = np.zeros(n + 1, dtype=np.uint64)
            capi.check(self.lib.awfm_gpu_group_locate(self._g, queries.ctypes.data, fmt, _ptr(offsets), fixed_len, n,
                                                      hit.ctypes.data, None, 0, None, None, C.byref(total)))
            pos = np.zeros(int(total.value), dtype=np.uint64)
            seq = np.zeros(int(total.value), dtype=np.uint64) if mapped else None
            loc = np.zeros(int(total.value), dtype=np.uint64) if mapped else None
            if total.value == 0:
                return (hit, pos, seq, loc) if mapped else (hit, pos)
        else:
            hit, pos = out[0], out[1]
            seq, loc = (out[2], out[3]) if mapped else (None, None)
        capi.check(self.lib.awfm_gpu_group_locate(self._g, queries.ctypes.data, fmt, _ptr(offsets), fixed_len, n,
                                                  hit.ctypes.data, pos.ctypes.data, len(pos), _ptr(seq), _ptr(loc),
                                                  C.byref(total)))
        self.last_total = int(total.value)
        return (hit, pos, seq, loc) if mapped else (hit, pos)


class KmerSearchList:
    """The reference's AwFmKmerSearchList, allocated and freed by `lib` itself (awFmCreateKmerSearchList /
    awFmDeallocKmerSearchList), filled the way README.md's example does: set kmerString/kmerLength per entry and
    `count` on the list.  Query bytes live in one numpy buffer owned by this object (the library never frees them,
    src/AwFmIndex.h:316-321)."""

    _DTYPE = np.dtype([("kmerString", "<u8"), ("kmerLength", "<u8"), ("positionList", "<u8"), ("count", "<u4"),
                       ("capacity", "<u4")])

    def __init__(self, lib, capacity):
        self.lib = lib
        self.capacity = capacity
        self.ptr = lib.awFmCreateKmerSearchList(capacity)
        if not self.ptr:
            raise MemoryError("awFmCreateKmerSearchList returned NULL")
        self._letters = None

    def entries(self):
        """numpy structured view (no copy) of the 32-B AwFmKmerSearchData array."""
        n = self.capacity
        if n == 0:
            return np.zeros(0, dtype=self._DTYPE)
        addr = C.addressof(self.ptr.contents.kmerSearchData.contents)
        buf = (C.c_uint8 * (32 * n)).from_address(addr)
        return np.frombuffer(buf, dtype=self._DTYPE)

    def fill(self, letters, offsets=None, fixed_len=0):
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        if offsets is None:
            n = len(letters) // fixed_len if fixed_len else 0
            starts = np.arange(n, dtype=np.uint64) * np.uint64(fixed_len)
            lengths = np.full(n, fixed_len, dtype=np.uint64)
        else:
            offsets = np.asarray(offsets, dtype=np.uint64)
            n = len(offsets) - 1
            starts, lengths = offsets[:-1], offsets[1:] - offsets[:-1]
        assert n <= self.capacity
        self._letters = letters
        e = self.entries()
        e["kmerString"][:n] = np.uint64(letters.ctypes.data) + starts
        e["kmerLength"][:n] = lengths
        self.ptr.contents.count = n
        return self

    @property
    def count(self):
        return int(self.ptr.contents.count)

    def counts(self):
        return self.entries()["count"][: self.count].copy()

    def positions(self):
        """list of uint64 arrays, one per query, copied out of the malloc'd position lists."""
        e = self.entries()
        out = []
        for i in range(self.count):
            c = int(e["count"][i])
            if c == 0:
                out.append(np.zeros(0, np.uint64))
            else:
                buf = (C.c_uint64 * c).from_address(int(e["positionList"][i]))
                out.append(np.frombuffer(buf, dtype=np.uint64).copy())
        return out

    def positions_flat(self, limit=None):
        """positions of the first `limit` queries, concatenated in query order (CSR values)."""
        e = self.entries()
        n = self.count if limit is None else min(limit, self.count)
        counts = e["count"][:n]
        lists = e["positionList"][:n]
        out = np.zeros(int(counts.sum(dtype=np.uint64)), np.uint64)
        o = 0
        for i in np.flatnonzero(counts):
            c = int(counts[i])
            out[o:o + c] = np.frombuffer((C.c_uint64 * c).from_address(int(lists[i])), dtype=np.uint64)
            o += c
        return out

    def close(self):
        if self.ptr:
            self.lib.awFmDeallocKmerSearchList(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def parallel_search_count(lib, index_ptr, search_list: KmerSearchList, num_threads=4):
    """awFmParallelSearchCount(index, searchList, numThreads) — src/AwFmIndex.h:400-403"""
    lib.awFmParallelSearchCount(index_ptr, search_list.ptr, num_threads)


def parallel_search_locate(lib, index_ptr, search_list: KmerSearchList, num_threads=4):
    """awFmParallelSearchLocate(index, searchList, numThreads) -> enum AwFmReturnCode — src/AwFmIndex.h:364-367"""
    return int(lib.awFmParallelSearchLocate(index_ptr, search_list.ptr, num_threads))
