"""TEST INFRASTRUCTURE — ctypes access to the two checkers:

  * `Reference` : the UNMODIFIED reference compiled by `make -C oracle ref` into oracle/_ref/libawfm_ref.so
                  (index construction with libdivsufsort, file I/O, and its own OpenMP/AVX2 batched search);
  * `Oracle`    : the scalar C restatement oracle/awfm_oracle.c (liboracle), which needs no AVX2/OpenMP and is
                  the on-box checker when _ref is unavailable.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from avxwindowfmindex_b200 import abi
from avxwindowfmindex_b200.capi import declare_search_list_api
from avxwindowfmindex_b200.index import IndexArrays, aligned_empty, sa_byte_length
from avxwindowfmindex_b200.search import KmerSearchList

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libawfm_ref.so")
ORACLE_LIB = os.path.join(HERE, "libawfm_oracle.so")
REFERENCE_TREE = os.environ.get("AWFM_REFERENCE_TREE", "/root/reference")


def build(ref=True):
    """Compile the checkers (idempotent).  _ref is only (re)built where the reference tree exists."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir(os.path.join(REFERENCE_TREE, "src")):
        subprocess.run(["make", "-s", "-C", HERE, "ref", f"REF={REFERENCE_TREE}"], check=True)
        dropin = os.path.join(HERE, "..", "avxwindowfmindex_b200", "csrc", "libawfm_b200.so")
        if os.path.exists(dropin):  # the C consumer of tests/test_c_consumer.py (needs the reference's header)
            subprocess.run(["make", "-s", "-C", HERE, "consumers", f"REF={REFERENCE_TREE}"], check=True)


def have_reference():
    return os.path.exists(REF_LIB)


# ----------------------------------------------------------------------------------------------- reference
class Reference:
    def __init__(self):
        if not have_reference():
            raise RuntimeError(f"{REF_LIB} missing: run `make -C oracle ref` where /root/reference exists")
        lib = C.CDLL(REF_LIB, mode=C.RTLD_LOCAL)  # RTLD_LOCAL + -Bsymbolic: never interposed by the drop-in
        declare_search_list_api(lib)
        vp = C.c_void_p
        lib.awFmCreateIndex.restype = C.c_int
        lib.awFmCreateIndex.argtypes = [C.POINTER(vp), C.POINTER(abi.AwFmIndexConfiguration), vp, C.c_size_t, C.c_char_p]
        lib.awFmCreateIndexFromFasta.restype = C.c_int
        lib.awFmCreateIndexFromFasta.argtypes = [C.POINTER(vp), C.POINTER(abi.AwFmIndexConfiguration), C.c_char_p, C.c_char_p]
        lib.awFmReadIndexFromFile.restype = C.c_int
        lib.awFmReadIndexFromFile.argtypes = [C.POINTER(vp), C.c_char_p, C.c_bool]
        lib.awFmDeallocIndex.restype = None
        lib.awFmDeallocIndex.argtypes = [vp]
        lib.awFmGetLocalSequencePositionFromIndexPosition.restype = C.c_int
        lib.awFmGetLocalSequencePositionFromIndexPosition.argtypes = [vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        lib.awFmNucleotideBacktraceBwtPosition.restype = C.c_size_t
        lib.awFmNucleotideBacktraceBwtPosition.argtypes = [vp, C.c_uint64]
        lib.awFmAminoBacktraceBwtPosition.restype = C.c_size_t
        lib.awFmAminoBacktraceBwtPosition.argtypes = [vp, C.c_uint64]
        lib.awFmGetValueFromCompressedSuffixArray.restype = C.c_size_t
        lib.awFmGetValueFromCompressedSuffixArray.argtypes = [vp, C.c_size_t]
        lib.AwFmMaskedVectorPopcount.restype = C.c_uint32
        self.lib = lib

    # -- index lifecycle (the reference's own code) --
    def create_index(self, sequence, path, alphabet=abi.AwFmAlphabetDna, seed_k=8, sa_ratio=8, keep_sa=True,
                     store_sequence=False):
        """awFmCreateIndex (src/AwFmCreate.c:31-137): builds with divsufsort64 and writes `path`."""
        if os.path.exists(path):
            os.remove(path)
        seq = np.ascontiguousarray(np.frombuffer(bytes(sequence), dtype=np.uint8))
        cfg = abi.AwFmIndexConfiguration(sa_ratio, seed_k, alphabet, keep_sa, store_sequence)
        out = C.c_void_p()
        rc = self.lib.awFmCreateIndex(C.byref(out), C.byref(cfg), seq.ctypes.data, len(seq), path.encode())
        if rc < 0:
            raise RuntimeError(f"awFmCreateIndex failed: {rc}")
        return out

    def create_index_from_fasta(self, fasta_path, path, alphabet=abi.AwFmAlphabetDna, seed_k=8, sa_ratio=8,
                                keep_sa=True):
        if os.path.exists(path):
            os.remove(path)
        cfg = abi.AwFmIndexConfiguration(sa_ratio, seed_k, alphabet, keep_sa, True)
        out = C.c_void_p()
        rc = self.lib.awFmCreateIndexFromFasta(C.byref(out), C.byref(cfg), fasta_path.encode(), path.encode())
        if rc < 0:
            raise RuntimeError(f"awFmCreateIndexFromFasta failed: {rc}")
        return out

    def read_index(self, path, keep_sa=True):
        out = C.c_void_p()
        rc = self.lib.awFmReadIndexFromFile(C.byref(out), path.encode(), keep_sa)
        if rc < 0:
            raise RuntimeError(f"awFmReadIndexFromFile failed: {rc}")
        return out

    def dealloc_index(self, index_ptr):
        self.lib.awFmDeallocIndex(index_ptr)

    @staticmethod
    def struct(index_ptr):
        return C.cast(index_ptr, C.POINTER(abi.AwFmIndex)).contents

    def arrays(self, index_ptr, copy=True):
        """IndexArrays view of a reference-owned struct AwFmIndex."""
        s = self.struct(index_ptr)
        amino = s.config.alphabetType == abi.AwFmAlphabetAmino
        card = 20 if amino else 4
        nblocks = 1 + (s.bwtLength - 1) // 256
        bbytes = abi.AMINO_BLOCK_BYTES if amino else abi.NUC_BLOCK_BYTES

        def grab(addr, nbytes, dtype):
            buf = (C.c_uint8 * nbytes).from_address(addr)
            a = np.frombuffer(buf, dtype=dtype)
            return a.copy() if copy else a

        raw = grab(s.bwtBlockList, nblocks * bbytes, np.uint8)
        if copy:
            blocks = aligned_empty(len(raw))
            blocks[:] = raw
        else:
            blocks = raw
        prefix = grab(s.prefixSums, (card + 2) * 8, "<u8")
        seeds = grab(s.kmerSeedTable, (card ** s.config.kmerLengthInSeedTable) * 16, "<u8").reshape(-1, 2)
        sa = None
        if s.suffixArray.values:
            sa = grab(s.suffixArray.values, s.suffixArray.compressedByteLength, np.uint8)
            assert s.suffixArray.compressedByteLength == sa_byte_length(s.bwtLength, s.config.suffixArrayCompressionRatio)
        meta = header = None
        if s.fastaVector:
            fv = C.cast(s.fastaVector, C.POINTER(abi.FastaVector)).contents
            meta = grab(fv.metadata.data, fv.metadata.count * 16, "<u8").reshape(-1, 2)
            header = bytes(grab(fv.header.charData, fv.header.count, np.uint8))
        return IndexArrays(int(s.config.alphabetType), int(s.config.kmerLengthInSeedTable),
                           int(s.config.suffixArrayCompressionRatio), int(s.bwtLength), blocks, prefix, seeds, sa,
                           int(s.featureFlags), False, None, header, meta)

    # -- the batched search, reference implementation --
    def search_list(self, letters, offsets=None, fixed_len=0):
        n = (len(offsets) - 1) if offsets is not None else (len(letters) // fixed_len if fixed_len else 0)
        return KmerSearchList(self.lib, n).fill(letters, offsets, fixed_len)

    def count(self, index_ptr, letters, offsets=None, fixed_len=0, threads=1):
        sl = self.search_list(letters, offsets, fixed_len)
        self.lib.awFmParallelSearchCount(index_ptr, sl.ptr, threads)
        out = sl.counts()
        sl.close()
        return out

    def locate(self, index_ptr, letters, offsets=None, fixed_len=0, threads=1):
        sl = self.search_list(letters, offsets, fixed_len)
        rc = int(self.lib.awFmParallelSearchLocate(index_ptr, sl.ptr, threads))
        counts, pos = sl.counts(), sl.positions()
        sl.close()
        return rc, counts, pos

    def contig_of(self, index_ptr, global_position):
        seq, loc = C.c_size_t(), C.c_size_t()
        rc = self.lib.awFmGetLocalSequencePositionFromIndexPosition(index_ptr, global_position, C.byref(seq), C.byref(loc))
        return rc, seq.value, loc.value


# ----------------------------------------------------------------------------------------------- oracle
class _OracleIndex(C.Structure):
    _fields_ = [("blocks", C.c_void_p), ("numBlocks", C.c_uint64), ("prefixSums", C.c_void_p),
                ("seedTable", C.c_void_p), ("saBytes", C.c_void_p), ("saByteLength", C.c_uint64),
                ("bwtLength", C.c_uint64), ("saBitWidth", C.c_uint8), ("saRatio", C.c_uint8),
                ("seedK", C.c_uint8), ("alphabet", C.c_uint8)]


class OracleWork(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("queries", "seeded", "lfSteps", "lfBlockReads", "queryLetters", "hits",
                                          "backtraceSteps", "countBytes", "locateBytes")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Oracle:
    def __init__(self, arrays: IndexArrays):
        if not os.path.exists(ORACLE_LIB):
            build(ref=False)
        lib = C.CDLL(ORACLE_LIB, mode=C.RTLD_LOCAL)
        vp, u64, u32, u8 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint8
        lib.awfm_oracle_occ.restype = u64
        lib.awfm_oracle_occ.argtypes = [vp, u8, u64]
        lib.awfm_oracle_block_popcount.restype = u32
        lib.awfm_oracle_block_popcount.argtypes = [vp, u8, u8, u8]
        lib.awfm_oracle_letter_at.restype = u8
        lib.awfm_oracle_letter_at.argtypes = [vp, u64]
        lib.awfm_oracle_letter_index.restype = u8
        lib.awfm_oracle_letter_index.argtypes = [u8, u8]
        lib.awfm_oracle_letter_is_ambiguous.restype = C.c_int
        lib.awfm_oracle_letter_is_ambiguous.argtypes = [u8, u8]
        lib.awfm_oracle_backtrace_step.restype = u64
        lib.awfm_oracle_backtrace_step.argtypes = [vp, u64]
        lib.awfm_oracle_sa_value.restype = u64
        lib.awfm_oracle_sa_value.argtypes = [vp, u64]
        lib.awfm_oracle_step.restype = None
        lib.awfm_oracle_step.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), u8]
        lib.awfm_oracle_count.restype = None
        lib.awfm_oracle_count.argtypes = [vp, vp, vp, u32, u64, vp, vp, vp, C.c_int]
        lib.awfm_oracle_locate.restype = u64
        lib.awfm_oracle_locate.argtypes = [vp, vp, vp, u32, u64, vp, vp, vp, C.c_int]
        lib.awfm_oracle_contig_of.restype = C.c_int
        lib.awfm_oracle_contig_of.argtypes = [vp, u64, u64, C.POINTER(u64), C.POINTER(u64)]
        self.lib = lib
        self.arrays = arrays
        ix = _OracleIndex()
        ix.blocks = arrays.blocks.ctypes.data
        ix.numBlocks = arrays.num_blocks
        ix.prefixSums = arrays.prefix_sums.ctypes.data
        ix.seedTable = arrays.seed_table.ctypes.data
        ix.saBytes = arrays.sa_bytes.ctypes.data if arrays.sa_bytes is not None else None
        ix.saByteLength = len(arrays.sa_bytes) if arrays.sa_bytes is not None else 0
        ix.bwtLength = arrays.bwt_length
        ix.saBitWidth = arrays.sa_width
        ix.saRatio = arrays.sa_ratio
        ix.seedK = arrays.seed_k
        ix.alphabet = arrays.alphabet
        self.ix = ix
        self.ixp = C.addressof(ix)

    @staticmethod
    def _n(letters, offsets, fixed_len):
        return (len(offsets) - 1) if offsets is not None else (len(letters) // fixed_len if fixed_len else 0)

    def count(self, letters, offsets=None, fixed_len=0, threads=1):
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        offsets = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.uint64)
        n = self._n(letters, offsets, fixed_len)
        counts = np.zeros(n, np.uint32)
        ranges = np.zeros((n, 2), np.uint64)
        work = OracleWork()
        self.lib.awfm_oracle_count(self.ixp, letters.ctypes.data, None if offsets is None else offsets.ctypes.data,
                                   fixed_len, n, counts.ctypes.data, ranges.ctypes.data, C.addressof(work), threads)
        return counts, ranges, work.as_dict()

    def locate(self, letters, offsets=None, fixed_len=0, threads=1):
        letters = np.ascontiguousarray(letters, dtype=np.uint8)
        offsets = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.uint64)
        n = self._n(letters, offsets, fixed_len)
        hit_offsets = np.zeros(n + 1, np.uint64)
        op = None if offsets is None else offsets.ctypes.data
        total = self.lib.awfm_oracle_locate(self.ixp, letters.ctypes.data, op, fixed_len, n,
                                            hit_offsets.ctypes.data, None, None, threads)
        positions = np.zeros(int(total), np.uint64)
        work = OracleWork()
        self.lib.awfm_oracle_locate(self.ixp, letters.ctypes.data, op, fixed_len, n, hit_offsets.ctypes.data,
                                    positions.ctypes.data if total else None, C.addressof(work), threads)
        return hit_offsets, positions, work.as_dict()
