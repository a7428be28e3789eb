/*
 * awfm_abi.h — layout-compatible restatement of the PUBLIC data structures of the reference library
 * (TravisWheelerLab/AvxWindowFmIndex, src/AwFmIndex.h) that the batched k-mer search path reads and writes.
 *
 * The reference's header cannot be included from CUDA or from SIMD-free C (it pulls in <immintrin.h> and embeds
 * __m256i in public structs, src/AwFmIndex.h:40-65), so the drop-in shim and the tests use these plain-C mirrors.
 * Every struct below has the same size, alignment and member offsets as the reference's; `tests/test_abi_layout.py`
 * proves it by compiling a probe against the real header when /root/reference is present, and the static
 * assertions at the bottom pin the numbers that probe produced.
 *
 * Names keep the reference's spelling so that code written against AwFmIndex.h reads the same; a translation unit
 * must include EITHER this file OR the reference's AwFmIndex.h, never both.
 */
#ifndef AWFM_ABI_H
#define AWFM_ABI_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/AwFmIndex.h:20-28 */
#define AW_FM_POSITIONS_PER_FM_BLOCK 256
#define AW_FM_NUCLEOTIDE_VECTORS_PER_WINDOW 3
#define AW_FM_NUCLEOTIDE_CARDINALITY 4
#define AW_FM_AMINO_VECTORS_PER_WINDOW 5
#define AW_FM_AMINO_CARDINALITY 20

/* src/AwFmIndex.h:30-34 */
enum AwFmAlphabetType { AwFmAlphabetAmino = 1, AwFmAlphabetDna = 2, AwFmAlphabetRna = 3 };

/* A 256-bit letter bit-vector: bit (p % 8) of byte (p / 8) belongs to block position p (src/AwFmCreate.c:296-335).
 * The reference types it __m256i / 2x uint8x16_t (src/AwFmIndex.h:40-52); only size (32) and alignment (32) matter. */
typedef struct AwFmBitVector256 {
  uint8_t bytes[32];
} __attribute__((aligned(32))) AwFmBitVector256;

/* src/AwFmIndex.h:55-65.  160 B and 352 B; baseOccurrences[c] = occurrences of letter c in BWT[0, 256*blockIndex). */
struct AwFmAminoBlock {
  AwFmBitVector256 letterBitVectors[AW_FM_AMINO_VECTORS_PER_WINDOW];
  uint64_t baseOccurrences[AW_FM_AMINO_CARDINALITY + 4];
};
struct AwFmNucleotideBlock {
  AwFmBitVector256 letterBitVectors[AW_FM_NUCLEOTIDE_VECTORS_PER_WINDOW];
  uint64_t baseOccurrences[AW_FM_NUCLEOTIDE_CARDINALITY + 4];
};

union AwFmBwtBlockList { /* src/AwFmIndex.h:67-70 */
  struct AwFmNucleotideBlock *asNucleotide;
  struct AwFmAminoBlock *asAmino;
};

struct AwFmIndexConfiguration { /* src/AwFmIndex.h:74-80 */
  uint8_t suffixArrayCompressionRatio;
  uint8_t kmerLengthInSeedTable;
  enum AwFmAlphabetType alphabetType;
  bool keepSuffixArrayInMemory;
  bool storeOriginalSequence;
};

struct AwFmCompressedSuffixArray { /* src/AwFmIndex.h:82-86 */
  uint8_t valueBitWidth;
  uint8_t *values;
  uint64_t compressedByteLength;
};

struct AwFmSearchRange { /* src/AwFmIndex.h:88-91; inclusive [startPtr, endPtr], valid iff startPtr <= endPtr */
  uint64_t startPtr;
  uint64_t endPtr;
};

/* lib/FastaVector (submodule @7ac0534): only the record table is read on this path (contig mapping of hits,
 * src/AwFmSearch.c:284-301).  FastaVectorString.h, FastaVectorMetadataVector.h:10-19, FastaVector.h:23-32. */
struct FastaVectorString {
  char *charData;
  size_t capacity;
  size_t count;
};
struct FastaVectorMetadata {
  size_t headerEndPosition;
  size_t sequenceEndPosition; /* cumulative end of the record in the concatenated text, separator included */
};
struct FastaVectorMetadataVector {
  struct FastaVectorMetadata *data;
  size_t capacity;
  size_t count;
};
struct FastaVector {
  struct FastaVectorString sequence;
  struct FastaVectorString header;
  struct FastaVectorMetadataVector metadata;
};

struct AwFmIndex { /* src/AwFmIndex.h:94-109 */
  uint32_t versionNumber;
  uint32_t featureFlags;
  uint64_t bwtLength;
  union AwFmBwtBlockList bwtBlockList;
  uint64_t *prefixSums;
  struct AwFmSearchRange *kmerSeedTable;
  FILE *fileHandle;
  struct AwFmIndexConfiguration config;
  int fileDescriptor;
  size_t suffixArrayFileOffset;
  size_t sequenceFileOffset;
  struct FastaVector *fastaVector;
  struct AwFmCompressedSuffixArray suffixArray;
};

struct AwFmKmerSearchData { /* src/AwFmIndex.h:111-117 */
  char *kmerString;
  uint64_t kmerLength;
  uint64_t *positionList;
  uint32_t count;
  uint32_t capacity;
};

struct AwFmKmerSearchList { /* src/AwFmIndex.h:119-123 */
  size_t capacity;
  size_t count;
  struct AwFmKmerSearchData *kmerSearchData;
};

/* src/AwFmIndex.h:132-138 */
enum AwFmReturnCode {
  AwFmSuccess = 1,
  AwFmFileReadOkay = 2,
  AwFmFileWriteOkay = 3,
  AwFmGeneralFailure = -1,
  AwFmUnsupportedVersionError = -2,
  AwFmAllocationFailure = -3,
  AwFmNullPtrError = -4,
  AwFmSuffixArrayCreationFailure = -5,
  AwFmIllegalPositionError = -6,
  AwFmNoFileSrcGiven = -7,
  AwFmNoDatabaseSequenceGiven = -8,
  AwFmFileFormatError = -9,
  AwFmFileOpenFail = -10,
  AwFmFileReadFail = -11,
  AwFmFileWriteFail = -12,
  AwFmErrorDbSequenceNull = -13,
  AwFmErrorSuffixArrayNull = -14,
  AwFmFileAlreadyExists = -15
};

/* ---- the four entry points of the hot path (the translation unit src/AwFmParallelSearch.c), unchanged ---- */

/* src/AwFmIndex.h:308, src/AwFmParallelSearch.c:36-84 */
struct AwFmKmerSearchList *awFmCreateKmerSearchList(const size_t capacity);
/* src/AwFmIndex.h:326-327, src/AwFmParallelSearch.c:86-93 */
void awFmDeallocKmerSearchList(struct AwFmKmerSearchList *restrict const searchList);
/* src/AwFmIndex.h:364-367, src/AwFmParallelSearch.c:95-157 */
enum AwFmReturnCode awFmParallelSearchLocate(const struct AwFmIndex *restrict const index,
                                             struct AwFmKmerSearchList *restrict const searchList,
                                             uint32_t numThreads);
/* src/AwFmIndex.h:400-403, src/AwFmParallelSearch.c:159-220 */
void awFmParallelSearchCount(const struct AwFmIndex *restrict const index,
                             struct AwFmKmerSearchList *restrict const searchList, uint32_t numThreads);

/* ---- additive entry points of the B200 drop-in (not in the reference) ---- */

/* Drops the device-resident copy of `index` (call before awFmDeallocIndex; the reference API has no hook). */
void awFmGpuReleaseIndex(const struct AwFmIndex *index);
/* Uploads `index` to the device now instead of lazily on the first batched call. AwFmSuccess or a failure code. */
enum AwFmReturnCode awFmGpuPrepareIndex(const struct AwFmIndex *index);
/* Return code of the most recent awFmParallelSearchCount on this thread (the reference's is void). */
enum AwFmReturnCode awFmGpuLastCountStatus(void);
/* Batched awFmGetLocalSequencePositionFromIndexPosition (src/AwFmIndex.h, src/AwFmSearch.c:284-301) on the device:
 * for each of `count` global positions, the record number and the offset inside that record.  Returns
 * AwFmUnsupportedVersionError when the index holds no FastaVector (as the reference does), AwFmIllegalPositionError
 * when at least one position lies beyond the last record (those entries are set to SIZE_MAX, the others are valid),
 * else AwFmSuccess. */
enum AwFmReturnCode awFmGpuGetLocalSequencePositions(const struct AwFmIndex *index, const size_t *globalPositions,
                                                     size_t count, size_t *sequenceNumbers,
                                                     size_t *localSequencePositions);

/* Number of GPUs the batched calls on `index` are fanned out over (AWFM_GPU_DEVICES = "all" | "0,1,..." selects them;
 * default one GPU, AWFM_GPU_DEVICE or 0); uploads the index if it is not resident yet.  0 on failure. */
int awFmGpuNumDevices(const struct AwFmIndex *index);

/* ---- the packed-batch API: the same two searches without the per-query 32-B structs and
 *      pointers of src/AwFmIndex.h:111-123 — contiguous k-mers in, flat arrays out.  With page-locked buffers
 *      (awFmGpuHostAlloc) the GPUs' copy engines read and write the caller's memory in place, chunk-pipelined with the
 *      search, every selected GPU working on its own contiguous shard of the batch. ---- */
enum AwFmGpuKmerFormat {
  AwFmGpuKmerAscii = 0, /* one byte per letter, what kmerString holds in the reference API                          */
  AwFmGpuKmer2Bit = 2,  /* nucleotide: letter j of k-mer i in bits [2j,2j+2) of bytes [i*B,(i+1)*B), B = ceil(L/4);
                           codes 0..3 = A,C,G,T — the reference's letter indices (src/AwFmLetter.c:4-22)            */
  AwFmGpuKmer5Bit = 5   /* amino: 5 bits per letter, B = ceil(5L/8); codes 0..19 in the reference's letter-index order
                           (src/AwFmLetter.c:55-67), >= 20 = ambiguity                                              */
};
/* counts[i] = what awFmParallelSearchCount stores in kmerSearchData[i].count.  kmerOffsets (numKmers+1 entries) is for
 * variable-length ASCII batches; NULL = every k-mer has kmerLength letters. */
enum AwFmReturnCode awFmGpuCountPacked(const struct AwFmIndex *index, const void *kmers, enum AwFmGpuKmerFormat format,
                                       const uint64_t *kmerOffsets, uint32_t kmerLength, uint64_t numKmers,
                                       uint32_t *counts);
/* CSR form of awFmParallelSearchLocate: positions[hitOffsets[i] .. hitOffsets[i+1]) = kmerSearchData[i].positionList
 * (same order).  *totalHits = hitOffsets[numKmers].  With positions == NULL or positionsCapacity < *totalHits only
 * hitOffsets and *totalHits are produced (size the buffer, call again).  sequenceNumbers / localSequencePositions (both
 * or neither) receive awFmGetLocalSequencePositionFromIndexPosition's answer for every hit. */
enum AwFmReturnCode awFmGpuLocatePacked(const struct AwFmIndex *index, const void *kmers, enum AwFmGpuKmerFormat format,
                                        const uint64_t *kmerOffsets, uint32_t kmerLength, uint64_t numKmers,
                                        uint64_t *hitOffsets, uint64_t *positions, uint64_t positionsCapacity,
                                        uint64_t *sequenceNumbers, uint64_t *localSequencePositions,
                                        uint64_t *totalHits);
void *awFmGpuHostAlloc(size_t bytes); /* page-locked, visible to every GPU; NULL on failure */
void awFmGpuHostFree(void *p);

#ifndef __cplusplus
_Static_assert(sizeof(struct AwFmNucleotideBlock) == 160, "nucleotide block is 160 B");
_Static_assert(sizeof(struct AwFmAminoBlock) == 352, "amino block is 352 B");
_Static_assert(offsetof(struct AwFmNucleotideBlock, baseOccurrences) == 96, "nuc base occurrences at 96");
_Static_assert(offsetof(struct AwFmAminoBlock, baseOccurrences) == 160, "amino base occurrences at 160");
_Static_assert(sizeof(struct AwFmSearchRange) == 16, "range is 16 B");
_Static_assert(sizeof(struct AwFmKmerSearchData) == 32, "search data is 32 B");
_Static_assert(offsetof(struct AwFmKmerSearchData, count) == 24, "count at 24");
_Static_assert(offsetof(struct AwFmKmerSearchData, capacity) == 28, "capacity at 28");
_Static_assert(sizeof(struct AwFmKmerSearchList) == 24, "search list is 24 B");
_Static_assert(sizeof(struct AwFmIndex) == 112, "index struct is 112 B");
_Static_assert(offsetof(struct AwFmIndex, config) == 48, "config at 48");
_Static_assert(offsetof(struct AwFmIndex, suffixArray) == 88, "suffixArray at 88");
_Static_assert(sizeof(struct FastaVector) == 72, "FastaVector is 72 B");
_Static_assert(offsetof(struct FastaVector, metadata) == 48, "metadata vector at 48");
_Static_assert(sizeof(struct FastaVectorMetadata) == 16, "record metadata is 16 B");
#endif

#ifdef __cplusplus
}
#endif
#endif /* AWFM_ABI_H */
