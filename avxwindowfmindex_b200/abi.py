"""ctypes mirrors of include/awfm_abi.h and include/awfm_gpu.h (same field order, sizes and offsets).

The reference's public structs are plain data (src/AwFmIndex.h:55-123); these classes let Python tests and the
bench build / inspect them exactly as a C caller would.
"""
import ctypes as C

AwFmAlphabetAmino, AwFmAlphabetDna, AwFmAlphabetRna = 1, 2, 3
AwFmSuccess, AwFmFileReadOkay = 1, 2
AwFmGpuKmerAscii, AwFmGpuKmer2Bit, AwFmGpuKmer5Bit = 0, 2, 5  # enum AwFmGpuKmerFormat (include/awfm_abi.h)
AwFmGeneralFailure, AwFmAllocationFailure, AwFmFileReadFail = -1, -3, -11
AwFmUnsupportedVersionError, AwFmNullPtrError, AwFmIllegalPositionError = -2, -4, -6
NUC_BLOCK_BYTES, AMINO_BLOCK_BYTES = 160, 352


class AwFmIndexConfiguration(C.Structure):  # src/AwFmIndex.h:74-80
    _fields_ = [
        ("suffixArrayCompressionRatio", C.c_uint8),
        ("kmerLengthInSeedTable", C.c_uint8),
        ("alphabetType", C.c_int),
        ("keepSuffixArrayInMemory", C.c_bool),
        ("storeOriginalSequence", C.c_bool),
    ]


class AwFmCompressedSuffixArray(C.Structure):  # src/AwFmIndex.h:82-86
    _fields_ = [
        ("valueBitWidth", C.c_uint8),
        ("values", C.c_void_p),
        ("compressedByteLength", C.c_uint64),
    ]


class AwFmSearchRange(C.Structure):  # src/AwFmIndex.h:88-91
    _fields_ = [("startPtr", C.c_uint64), ("endPtr", C.c_uint64)]


class AwFmIndex(C.Structure):  # src/AwFmIndex.h:94-109
    _fields_ = [
        ("versionNumber", C.c_uint32),
        ("featureFlags", C.c_uint32),
        ("bwtLength", C.c_uint64),
        ("bwtBlockList", C.c_void_p),
        ("prefixSums", C.c_void_p),
        ("kmerSeedTable", C.c_void_p),
        ("fileHandle", C.c_void_p),
        ("config", AwFmIndexConfiguration),
        ("fileDescriptor", C.c_int),
        ("suffixArrayFileOffset", C.c_size_t),
        ("sequenceFileOffset", C.c_size_t),
        ("fastaVector", C.c_void_p),
        ("suffixArray", AwFmCompressedSuffixArray),
    ]


class AwFmKmerSearchData(C.Structure):  # src/AwFmIndex.h:111-117
    _fields_ = [
        ("kmerString", C.c_void_p),
        ("kmerLength", C.c_uint64),
        ("positionList", C.c_void_p),
        ("count", C.c_uint32),
        ("capacity", C.c_uint32),
    ]


class AwFmKmerSearchList(C.Structure):  # src/AwFmIndex.h:119-123
    _fields_ = [
        ("capacity", C.c_size_t),
        ("count", C.c_size_t),
        ("kmerSearchData", C.POINTER(AwFmKmerSearchData)),
    ]


class FastaVectorString(C.Structure):  # lib/FastaVector/src/FastaVectorString.h
    _fields_ = [("charData", C.c_void_p), ("capacity", C.c_size_t), ("count", C.c_size_t)]


class FastaVectorMetadataVector(C.Structure):  # lib/FastaVector/src/FastaVectorMetadataVector.h:15-19
    _fields_ = [("data", C.c_void_p), ("capacity", C.c_size_t), ("count", C.c_size_t)]


class FastaVector(C.Structure):  # lib/FastaVector/src/FastaVector.h
    _fields_ = [("sequence", FastaVectorString), ("header", FastaVectorString), ("metadata", FastaVectorMetadataVector)]


class awfm_index_view(C.Structure):  # include/awfm_gpu.h
    _fields_ = [
        ("blocks", C.c_void_p),
        ("numBlocks", C.c_uint64),
        ("prefixSums", C.c_void_p),
        ("seedTable", C.c_void_p),
        ("saBytes", C.c_void_p),
        ("saByteLength", C.c_uint64),
        ("bwtLength", C.c_uint64),
        ("saBitWidth", C.c_uint8),
        ("saRatio", C.c_uint8),
        ("seedK", C.c_uint8),
        ("alphabet", C.c_uint8),
    ]


class awfm_gpu_stats(C.Structure):  # include/awfm_gpu.h
    _fields_ = [
        ("kernelMs", C.c_double),
        ("h2dMs", C.c_double),
        ("d2hMs", C.c_double),
        ("launches", C.c_uint64),
        ("queries", C.c_uint64),
        ("hits", C.c_uint64),
        ("h2dBytes", C.c_uint64),
        ("d2hBytes", C.c_uint64),
    ]


assert C.sizeof(AwFmIndex) == 112 and AwFmIndex.config.offset == 48 and AwFmIndex.suffixArray.offset == 88
assert C.sizeof(AwFmKmerSearchData) == 32 and C.sizeof(AwFmKmerSearchList) == 24
assert C.sizeof(AwFmIndexConfiguration) == 12 and C.sizeof(AwFmCompressedSuffixArray) == 24


class awfm_file_info(C.Structure):  # include/awfm_gpu.h
    _fields_ = [("bwtLength", C.c_uint64), ("numSequences", C.c_uint64), ("suffixArrayByteLength", C.c_uint64),
                ("versionNumber", C.c_uint32), ("featureFlags", C.c_uint32),
                ("suffixArrayCompressionRatio", C.c_uint8), ("kmerLengthInSeedTable", C.c_uint8),
                ("alphabetType", C.c_uint8), ("storeOriginalSequence", C.c_uint8)]
