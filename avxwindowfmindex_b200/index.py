"""Host-side container for the arrays of an AwFm index and the unchanged `.awfmi` version-8 file format.

`IndexArrays` holds exactly what the search path reads from the reference's `struct AwFmIndex`
(src/AwFmIndex.h:94-109): the BWT block array, prefix sums, k-mer seed table and bit-packed sampled suffix array.
`read_awfmi` / `write_awfmi` follow the layout written by awFmWriteIndexToFile (src/AwFmFile.c:20-193) and read by
awFmReadIndexFromFile (src/AwFmFile.c:195-449); section offsets as in src/AwFmFile.c:524-558.
"""
import ctypes as C
import dataclasses
import struct

import numpy as np

from . import abi

MAGIC = b"AwFmIndex\n"  # src/AwFmFile.c:17-18
VERSION = 8             # src/AwFmIndexStruct.h:9


def aligned_empty(nbytes, alignment=64):
    raw = np.empty(nbytes + alignment, dtype=np.uint8)
    shift = (-raw.ctypes.data) % alignment
    return raw[shift:shift + nbytes]


def sa_bit_width(bwt_length):  # src/AwFmSuffixArray.c:12-18
    return max(1, int(bwt_length - 1).bit_length())


def sa_num_samples(bwt_length, ratio):  # src/AwFmSuffixArray.c:144-147
    return (bwt_length + ratio - 1) // ratio


def sa_byte_length(bwt_length, ratio):  # src/AwFmSuffixArray.c:41-53
    bits = sa_num_samples(bwt_length, ratio) * sa_bit_width(bwt_length)
    return (bits + 7) // 8 + 8


@dataclasses.dataclass
class IndexArrays:
    alphabet: int          # 1 amino, 2 DNA, 3 RNA
    seed_k: int
    sa_ratio: int
    bwt_length: int
    blocks: np.ndarray      # uint8, numBlocks * 160|352, 32-B aligned
    prefix_sums: np.ndarray  # uint64, |A|+2
    seed_table: np.ndarray  # uint64, (|A|^k, 2)
    sa_bytes: np.ndarray    # uint8 or None
    feature_flags: int = 0
    store_sequence: bool = False
    sequence: bytes = None
    fasta_header: bytes = None
    fasta_metadata: np.ndarray = None  # uint64 (numSequences, 2): headerEnd, sequenceEnd

    @property
    def amino(self):
        return self.alphabet == abi.AwFmAlphabetAmino

    @property
    def cardinality(self):
        return 20 if self.amino else 4

    @property
    def block_bytes(self):
        return abi.AMINO_BLOCK_BYTES if self.amino else abi.NUC_BLOCK_BYTES

    @property
    def num_blocks(self):
        return 1 + (self.bwt_length - 1) // 256

    @property
    def sa_width(self):
        return sa_bit_width(self.bwt_length)

    def view(self):
        """awfm_index_view over the numpy buffers (keep `self` alive while it is in use)."""
        v = abi.awfm_index_view()
        v.blocks = self.blocks.ctypes.data
        v.numBlocks = self.num_blocks
        v.prefixSums = self.prefix_sums.ctypes.data
        v.seedTable = self.seed_table.ctypes.data
        v.saBytes = self.sa_bytes.ctypes.data if self.sa_bytes is not None else None
        v.saByteLength = len(self.sa_bytes) if self.sa_bytes is not None else 0
        v.bwtLength = self.bwt_length
        v.saBitWidth = self.sa_width
        v.saRatio = self.sa_ratio
        v.seedK = self.seed_k
        v.alphabet = self.alphabet
        return v

    def as_awfm_index(self):
        """A `struct AwFmIndex` (reference layout) whose pointers alias these arrays — what a C caller of the
        reference API holds after awFmReadIndexFromFile(..., keepSuffixArrayInMemory=true)."""
        assert self.blocks.ctypes.data % 32 == 0
        ix = abi.AwFmIndex()
        ix.versionNumber = VERSION
        ix.featureFlags = self.feature_flags
        ix.bwtLength = self.bwt_length
        ix.bwtBlockList = self.blocks.ctypes.data
        ix.prefixSums = self.prefix_sums.ctypes.data
        ix.kmerSeedTable = self.seed_table.ctypes.data
        ix.fileHandle = None
        ix.config.suffixArrayCompressionRatio = self.sa_ratio
        ix.config.kmerLengthInSeedTable = self.seed_k
        ix.config.alphabetType = self.alphabet
        ix.config.keepSuffixArrayInMemory = self.sa_bytes is not None
        ix.config.storeOriginalSequence = False
        ix.fileDescriptor = -1
        ix.fastaVector = None
        ix.suffixArray.valueBitWidth = self.sa_width
        ix.suffixArray.values = self.sa_bytes.ctypes.data if self.sa_bytes is not None else None
        ix.suffixArray.compressedByteLength = len(self.sa_bytes) if self.sa_bytes is not None else 0
        return ix

    def sa_values(self):
        """Unpacked sampled SA (for tests): field j at bit j*w, little-endian."""
        w, n = self.sa_width, sa_num_samples(self.bwt_length, self.sa_ratio)
        bits = np.unpackbits(self.sa_bytes, bitorder="little")[: n * w].reshape(n, w).astype(np.uint64)
        return (bits << np.arange(w, dtype=np.uint64)).sum(axis=1, dtype=np.uint64)


def section_digests(ix: "IndexArrays", chunk_bytes=1 << 26, with_chunks=True):
    """SHA-256 of every section of an index as the `.awfmi` file stores it (blocks, prefix sums, seed table, packed
    sampled SA), plus a short digest per 64-MiB chunk so that a mismatch can be localised.  Used to prove that a
    device-built index equals the one the reference's awFmCreateIndex builds (tools/ref_index_hashes.py)."""
    import hashlib

    def one(a):
        b = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        whole = hashlib.sha256()
        chunks = []
        for o in range(0, len(b), chunk_bytes):
            part = b[o:o + chunk_bytes]
            whole.update(part)
            if with_chunks:
                chunks.append(hashlib.sha256(part).hexdigest()[:16])
        out = {"bytes": int(len(b)), "sha256": whole.hexdigest()}
        if with_chunks:
            out.update({"chunk_bytes": chunk_bytes, "chunks": chunks})
        return out

    out = {"blocks": one(ix.blocks), "prefix_sums": one(ix.prefix_sums.astype("<u8")),
           "seed_table": one(ix.seed_table.astype("<u8"))}
    if ix.sa_bytes is not None:
        out["suffix_array"] = one(ix.sa_bytes)
    return out


def read_awfmi(path, keep_suffix_array=True):
    with open(path, "rb") as f:
        data = f.read()
    if data[:10] != MAGIC:
        raise ValueError("not an .awfmi file (bad magic)")
    version, flags = struct.unpack_from("<II", data, 10)
    if version != VERSION:
        raise ValueError(f"unsupported .awfmi version {version}")
    ratio, seed_k, alphabet, store_seq = struct.unpack_from("<BBBB", data, 18)
    (bwt_length,) = struct.unpack_from("<Q", data, 22)
    amino = alphabet == abi.AwFmAlphabetAmino
    block_bytes = abi.AMINO_BLOCK_BYTES if amino else abi.NUC_BLOCK_BYTES
    card = 20 if amino else 4
    num_blocks = 1 + (bwt_length - 1) // 256
    off = 30
    blocks = aligned_empty(num_blocks * block_bytes)
    blocks[:] = np.frombuffer(data, np.uint8, num_blocks * block_bytes, off)
    off += num_blocks * block_bytes
    prefix_sums = np.frombuffer(data, "<u8", card + 2, off).copy()
    off += (card + 2) * 8
    num_seeds = card ** seed_k
    seed_table = np.frombuffer(data, "<u8", num_seeds * 2, off).reshape(num_seeds, 2).copy()
    off += num_seeds * 16
    sequence = None
    if store_seq:
        sequence = data[off: off + bwt_length - 1]
        off += bwt_length - 1
    sa_len = sa_byte_length(bwt_length, ratio)
    sa_bytes = np.frombuffer(data, np.uint8, sa_len, off).copy() if keep_suffix_array else None
    off += sa_len
    header = metadata = None
    if flags & 1:  # src/AwFmIndexStruct.h:10, src/AwFmFile.c:360-440
        header_len, meta_count = struct.unpack_from("<QQ", data, off)
        off += 16
        header = data[off: off + header_len]
        off += header_len
        metadata = np.frombuffer(data, "<u8", meta_count * 2, off).reshape(meta_count, 2).copy()
    return IndexArrays(alphabet, seed_k, ratio, bwt_length, blocks, prefix_sums, seed_table, sa_bytes, flags,
                       bool(store_seq), sequence, header, metadata)


def write_awfmi(ix: IndexArrays, path):
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", VERSION, ix.feature_flags))
        f.write(struct.pack("<BBBB", ix.sa_ratio, ix.seed_k, ix.alphabet, 1 if ix.store_sequence else 0))
        f.write(struct.pack("<Q", ix.bwt_length))
        f.write(ix.blocks.tobytes())
        f.write(ix.prefix_sums.astype("<u8").tobytes())
        f.write(ix.seed_table.astype("<u8").tobytes())
        if ix.store_sequence:
            f.write(ix.sequence)
        f.write(ix.sa_bytes.tobytes())
        if ix.feature_flags & 1:
            f.write(struct.pack("<QQ", len(ix.fasta_header), len(ix.fasta_metadata)))
            f.write(ix.fasta_header)
            f.write(ix.fasta_metadata.astype("<u8").tobytes())
