/*
 * awfm_gpu.h — thin C-ABI between host code (C, Python/ctypes, anything with an FFI) and the hand-written sm_100a
 * CUDA implementation of AwFmIndex's batched exact-match k-mer search.  Plain pointers and sizes only; no CUDA,
 * torch or reference types appear in any signature.  Library: avxwindowfmindex_b200/csrc/libawfm_b200.so.
 *
 * What each entry point replaces in the reference (TravisWheelerLab/AvxWindowFmIndex @92b849f):
 *
 *   awfm_gpu_ctx_create / _destroy   the read-only arrays of struct AwFmIndex (src/AwFmIndex.h:94-109) made resident
 *                                    in HBM: bwtBlockList, prefixSums, kmerSeedTable, suffixArray.values.
 *   awfm_gpu_count_*                 awFmParallelSearchCount (src/AwFmParallelSearch.c:159-220) =
 *                                    parallelSearchFindKmerSeedsForBlock (:222-271, seed table src/AwFmKmerTable.c:4-51,
 *                                    non-seeded start src/AwFmSearch.c:485-520) + parallelSearchExtendKmersInBlock
 *                                    (:273-313; LF step src/AwFmSearch.c:42-159; rank src/AwFmOccurrence.c:8-135 and
 *                                    src/AwFmSimdConfig.c:89-114) + awFmSearchRangeLength (src/AwFmIndexStruct.c:126-130).
 *   awfm_gpu_locate_*                awFmParallelSearchLocate (src/AwFmParallelSearch.c:95-157): the above, then
 *                                    parallelSearchTracebackPositionLists (:315-365; backtrace step src/AwFmSearch.c:369-427,
 *                                    letter-at-position src/AwFmOccurrence.c:170-217, sampled-SA read
 *                                    src/AwFmSuffixArray.c:114-142,179-203).
 *   awfm_gpu_search_list_*           the same two calls over the reference's own AwFmKmerSearchList memory layout
 *                                    (src/AwFmIndex.h:111-123): pipelined host marshalling + kernels + scatter.
 *
 * Query batch format ("packed batch"): the ASCII letters of all queries concatenated in `letters`; query i is
 * letters[offsets[i] .. offsets[i+1]).  If `offsets` is NULL every query has exactly `fixedLen` letters and
 * query i starts at i*fixedLen.  Letters are the raw bytes the reference API would be given (any case, ambiguity
 * codes allowed); translation to letter indices follows src/AwFmLetter.c:4-22,55-67 on the device.
 *
 * All functions return AWFM_GPU_OK (0) or a negative awfm_gpu_status; awfm_gpu_last_error() returns a
 * thread-local human-readable message.  There is NO CPU fallback: without a usable CUDA device every call fails.
 */
#ifndef AWFM_GPU_H
#define AWFM_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum awfm_gpu_status {
  AWFM_GPU_OK = 0,
  AWFM_GPU_ERR_ARG = -1,     /* bad argument (NULL, unsupported alphabet/width, ...)        */
  AWFM_GPU_ERR_NO_DEVICE = -2, /* no CUDA device / driver                                   */
  AWFM_GPU_ERR_ALLOC = -3,   /* device or pinned-host allocation failed                     */
  AWFM_GPU_ERR_CUDA = -4,    /* any other CUDA runtime / launch error                       */
  AWFM_GPU_ERR_NO_SA = -5    /* locate requested but the context holds no sampled SA        */
} awfm_gpu_status;

/* Plain view of the index arrays (all HOST pointers; copied, never retained).  Field meaning = the reference's. */
typedef struct awfm_index_view {
  const void *blocks;         /* struct AwFm{Nucleotide,Amino}Block[numBlocks] (160 / 352 B each)               */
  uint64_t numBlocks;         /* 1 + (bwtLength-1)/256                    (src/AwFmIndexStruct.c:104-106)       */
  const uint64_t *prefixSums; /* |alphabet|+2 entries                     (src/AwFmCreate.c:338-343)            */
  const void *seedTable;      /* struct AwFmSearchRange[|alphabet|^seedK] (src/AwFmCreate.c:407-450)            */
  const uint8_t *saBytes;     /* bit-packed sampled SA, or NULL (count-only context)  (src/AwFmSuffixArray.c)   */
  uint64_t saByteLength;      /* suffixArray.compressedByteLength                                                */
  uint64_t bwtLength;
  uint8_t saBitWidth;         /* suffixArray.valueBitWidth, 1..64                                                */
  uint8_t saRatio;            /* config.suffixArrayCompressionRatio, 1..255                                      */
  uint8_t seedK;              /* config.kmerLengthInSeedTable                                                    */
  uint8_t alphabet;           /* enum AwFmAlphabetType: 1 amino, 2 DNA, 3 RNA                                    */
} awfm_index_view;

typedef struct awfm_range { /* == struct AwFmSearchRange */
  uint64_t startPtr;
  uint64_t endPtr;
} awfm_range;

typedef struct awfm_kmer_search_data { /* == struct AwFmKmerSearchData, 32 B (src/AwFmIndex.h:111-117) */
  char *kmerString;
  uint64_t kmerLength;
  uint64_t *positionList;
  uint32_t count;
  uint32_t capacity;
} awfm_kmer_search_data;

typedef struct awfm_gpu_ctx awfm_gpu_ctx;

/* How a batch of queries is stored (awfm_gpu_count_device_format, awfm_gpu_group_*):
 *   AWFM_QUERY_ASCII  one byte per letter, the raw bytes the reference API would be given (any case, ambiguity codes);
 *                     fixed length, or variable length through an offsets array;
 *   AWFM_QUERY_2BIT   nucleotide indexes, fixed length L: query i occupies bytes [i*B, (i+1)*B) with B = ceil(L/4),
 *                     letter j in bits [2j, 2j+2) of that little-endian byte string; codes 0,1,2,3 = A,C,G,T(U) — the
 *                     reference's own letter indices (src/AwFmLetter.c:4-22).  No ambiguity codes.
 *   AWFM_QUERY_5BIT   amino indexes, fixed length L: B = ceil(5L/8), letter j in bits [5j, 5j+5); codes 0..19 = the
 *                     letters A C D E F G H I K L M N P Q R S T V W Y in the reference's index order
 *                     (src/AwFmLetter.c:55-67), any code >= 20 = the ambiguity letter.
 * A 20-mer costs 5 bytes on the bus instead of 20; an amino 8-mer 5 instead of 8. */
typedef enum awfm_query_format { AWFM_QUERY_ASCII = 0, AWFM_QUERY_2BIT = 2, AWFM_QUERY_5BIT = 5 } awfm_query_format;

/* Kernel-only timing and work counters of the most recent call on a context (device time from CUDA events). */
typedef struct awfm_gpu_stats {
  double kernelMs;        /* sum over launches of the search / backtrace kernels          */
  double h2dMs, d2hMs;    /* only filled by the *_host and search_list calls              */
  uint64_t launches;      /* kernels launched by the call (all of ours, incl. scans)      */
  uint64_t queries, hits;
  uint64_t h2dBytes, d2hBytes;
} awfm_gpu_stats;

const char *awfm_gpu_last_error(void);
int awfm_gpu_device_count(void);

/* ---- index residency ---- */
int awfm_gpu_ctx_create(awfm_gpu_ctx **ctx, int device, const awfm_index_view *view);
/* Same, but every array pointer in `view` is a DEVICE pointer on `device` (index built or loaded on the GPU). */
int awfm_gpu_ctx_create_from_device(awfm_gpu_ctx **ctx, int device, const awfm_index_view *view);
/* Index straight from an unchanged `.awfmi` version-8 file (replaces awFmReadIndexFromFile,
 * src/AwFmFile.c:195-449, + the upload): the file is mapped and its sections go from the page cache to the device,
 * no host copy of the index is built.  wantSuffixArray = 0 gives a count-only context (the reference's
 * keepSuffixArrayInMemory = false would pread per hit instead, src/AwFmFile.c:484-522).  A FastaVector record table
 * in the file is installed for awfm_gpu_map_positions_*.  Format errors are AWFM_GPU_ERR_ARG and are detected before
 * any CUDA call.  `info` (may be NULL) receives the header fields. */
typedef struct awfm_file_info {
  uint64_t bwtLength, numSequences, suffixArrayByteLength;
  uint32_t versionNumber, featureFlags;
  uint8_t suffixArrayCompressionRatio, kmerLengthInSeedTable, alphabetType, storeOriginalSequence;
} awfm_file_info;
int awfm_gpu_ctx_create_from_file(awfm_gpu_ctx **ctx, int device, const char *path, int wantSuffixArray,
                                  awfm_file_info *info);
void awfm_gpu_ctx_destroy(awfm_gpu_ctx *ctx);
uint64_t awfm_gpu_ctx_device_bytes(const awfm_gpu_ctx *ctx);
int awfm_gpu_ctx_get_stats(const awfm_gpu_ctx *ctx, awfm_gpu_stats *out);
/* Tuning knobs (kernel variant selection for measurement; defaults are the shipped configuration).
 * keys: "count_lpq" (lanes per query: 1,2,4,8; clamped to what the alphabet's line layout supports), "locate_lpq",
 * "count_variant" (0 group-per-query from global memory, 1 CTA tiles staged in shared memory), "locate_variant"
 * (0 group-per-hit, 1 group-per-hit with refill), "blocks_per_sm" (0 = occupancy query), "use_deep_seed_table" (0/1:
 * A/B switch for a table already derived with awfm_gpu_ctx_extend_seed_table); of the search-list engine:
 * "chunk_queries" / "locate_chunk_queries" (queries per pipeline chunk of count / locate; chunk_queries 0 = automatic,
 * a power of two between 2^16 and 2^19 that leaves every device about 48 chunks), "locate_inline_hits" (a
 * chunk with more hits is finished through windows), "locate_window_hits" (hits per such window).
 * Sweep count path (csrc/awfm_sweep.cuh; large batches of either alphabet, fixed or variable length, counts and ranges):
 * "sweep_min_queries" (0 = automatic: batches of at least max(2^22, one query per two 128-B lines of the index) queries — with range
 * output one per line, three per four lines when the ranges leave through the ordered emit — and no derived deep seed table; n > 0 = batches of at least n queries; -1 = never), "sweep_sort_bits" (top bits of the seed
 * index the radix sort orders, default 32 = all but the low "sweep_local_bits"), "sweep_local_bits" (0..8, or -1 =
 * automatic, the default: low bits ordered inside each tile of the first pass instead), "sweep_items" /
 * "sweep_first_items" (records per thread and tile of the later passes / of the first pass: 1, 2, 4, 8; default 4),
 * "sweep_max_batch" (queries per slice of the scratch — 96 B per query, 352 B for amino indexes —, default 2^27; when
 * even that does not fit the call is answered by the tile kernel), "sweep_profile" (0/1: record an event after every
 * stage of the next calls), "sweep_own_sort" (1, the default: the hand-written bucket passes of csrc/awfm_sort.cuh order
 * the pairs whenever at most 16 key bits have to be ordered globally; 0: CUB's radix sort, kept as cross-check and for
 * deeper seed tables), "sweep_compact_pairs" (1, the default: between the pack kernel and the second bucket pass a pair
 * travels as ONE 8-byte word — key, remaining letters and, once the first digit has left the key, the query id —
 * whenever those fit 64 bits, instead of a 4-byte key and an 8-byte payload; 0: always 4 + 8 bytes), "sweep_ordered_emit"
 * (1, the default: with range output and at most 2^24 queries the last pass hands its survivors, bucketed by the sixteenth
 * of the id space, to a kernel that writes counts and ranges slice by slice; 0: scattered straight from the pass),
 * "sweep_record12" (1: live records of nucleotide batches with at most 8 letters left of the seed
 * travel as 12 bytes — 16-bit range width, 16 bits of letters — and queries whose seed range is wider than 65534 take
 * the generic per-query search; 0, the default: 16-byte records — the passes are latency-bound, not byte-bound, and the
 * smaller records measured no faster, profiles/r02_sweep_probe.jsonl), "sweep_variable" (1, the default: variable-length
 * ASCII batches — an offsets array — take the sweep too, every record carrying its own length as a marker bit above its
 * remaining letters; queries shorter than the seed k-mer or with more than 15 (amino: 6) letters left of it are answered
 * by the generic per-query search inside the same call; 0: such batches always take the tile kernel), "sweep_wide" (0, the
 * default: 64-bit positions only for nucleotide indexes of 2^32 .. 2^40 positions, where they are needed; 1: always —
 * cross-check on small indexes.  Records stay 16 bytes: 40 bits of position, 24 bits of range width; a query whose seed
 * range is wider than 2^24 - 2 is answered by the generic per-query search). */
int awfm_gpu_ctx_set_tuning(awfm_gpu_ctx *ctx, const char *key, int64_t value);
/* Device time of every stage of the most recent sweep count call made with "sweep_profile" = 1, in launch order:
 * clear + pack, radix sort, first pass (seed entry + LF step 1), one entry per further pass, irregular queries.
 * Returns the number of entries written (<= capacity), 0 when the last count call did not take the sweep path. */
int awfm_gpu_ctx_sweep_stage_ms(awfm_gpu_ctx *ctx, double *ms, int capacity);
/* Live records of the most recent sweep count call: live[0] = queries of the batch, live[p] = queries still searching
 * after LF step p (the records pass p appended).  Returns the number of entries written, 0 when the last count call did
 * not take the sweep path.  The bench's compulsory-traffic model for the sweep is computed from these. */
int awfm_gpu_ctx_sweep_live(awfm_gpu_ctx *ctx, uint64_t *live, int capacity, uint64_t *irregular /* may be NULL */);

/* ---- derived structures: spend HBM (180 GB per B200) to remove dependent DRAM round trips.  Both are computed on
 *      the device from the unchanged index with the search kernels' own primitives, hold exactly the values the
 *      reference would compute at query time, and are optional (default: absent). ----
 * extend_seed_table: a deeper k-mer seed table.  Entry x of depth d = the range the reference holds after the last
 *   d letters x of a query: kmerSeedTable entry of the last k letters (src/AwFmKmerTable.c:21-51) pushed through
 *   d-k LF steps with the reference's stop-when-invalid rule (src/AwFmParallelSearch.c:279-311).  A query whose last
 *   d letters are all searchable letters then opens with ONE table read instead of 1 + (d-k) step pairs; any other
 *   query takes the normal path.  Size |alphabet|^d x 8 B (16 B when bwtLength > 2^32).  depth <= seedK drops it.
 * densify_suffix_array: SA samples at every newRatio-th BWT position (newRatio < the index's ratio), obtained by
 *   running the reference's own backtrace walk (src/AwFmParallelSearch.c:333-361) for each of them once.  Locate
 *   then walks newRatio-1 steps per hit on average instead of ratio-1 (none at newRatio = 1).  Size
 *   ceil(bwtLength/newRatio) x 4 B (8 B when bwtLength > 2^32).  newRatio = 0 or >= the index's ratio drops it. */
int awfm_gpu_ctx_extend_seed_table(awfm_gpu_ctx *ctx, uint32_t depth, double *buildMs /* may be NULL */);
int awfm_gpu_ctx_densify_suffix_array(awfm_gpu_ctx *ctx, uint32_t newRatio, double *buildMs /* may be NULL */);

/* ---- packed batch, HOST buffers (H2D, kernels, D2H inside the call) ---- */
int awfm_gpu_count_host(awfm_gpu_ctx *ctx, const uint8_t *letters, const uint64_t *offsets, uint32_t fixedLen,
                        uint64_t numQueries, uint32_t *counts /* out, numQueries */,
                        awfm_range *ranges /* out, numQueries, may be NULL */);
/* Locate, CSR output: hitOffsets[numQueries+1] (exclusive prefix sum of range lengths) and positions
 * [hitOffsets[numQueries]] in SA order within each query.  Two-step so the caller can size `positions`:
 * pass positions == NULL to get only hitOffsets; then call again with a buffer of positionsCapacity entries. */
int awfm_gpu_locate_host(awfm_gpu_ctx *ctx, const uint8_t *letters, const uint64_t *offsets, uint32_t fixedLen,
                         uint64_t numQueries, uint64_t *hitOffsets /* out, numQueries+1 */,
                         uint64_t *positions /* out or NULL */, uint64_t positionsCapacity,
                         awfm_range *ranges /* out, may be NULL */);

/* ---- packed batch, DEVICE buffers on the context's device, asynchronous on `stream` (a cudaStream_t) ---- */
int awfm_gpu_count_device(awfm_gpu_ctx *ctx, const uint8_t *dLetters, const uint64_t *dOffsets, uint32_t fixedLen,
                          uint64_t numQueries, uint32_t *dCounts, awfm_range *dRanges /* may be NULL */,
                          void *stream);
/* Same for any awfm_query_format (dOffsets must be NULL unless format is AWFM_QUERY_ASCII). */
int awfm_gpu_count_device_format(awfm_gpu_ctx *ctx, const void *dQueries, uint32_t format, const uint64_t *dOffsets,
                                 uint32_t fixedLen, uint64_t numQueries, uint32_t *dCounts,
                                 awfm_range *dRanges /* may be NULL */, void *stream);
/* dRanges (in) come from awfm_gpu_count_device; dHitOffsets (out, numQueries+1) is their exclusive scan. */
int awfm_gpu_scan_ranges_device(awfm_gpu_ctx *ctx, const awfm_range *dRanges, uint64_t numQueries,
                                uint64_t *dHitOffsets, void *stream);
/* The front end of a device-resident locate in ONE call (what the locate pipelines of this library use themselves):
 * the search, then dHitOffsets (numQueries+1) scanned from the u32 counts.  dRanges[q] is written ONLY for queries with
 * dCounts[q] > 0 — enough for awfm_gpu_locate_device, and it spares a scattered 16-B store per query without hits and
 * three quarters of the scan's reads. */
int awfm_gpu_locate_prepare_device(awfm_gpu_ctx *ctx, const void *dQueries, uint32_t format, const uint64_t *dOffsets,
                                   uint32_t fixedLen, uint64_t numQueries, uint32_t *dCounts, awfm_range *dRanges,
                                   uint64_t *dHitOffsets, void *stream);
/* Backtrace + sampled-SA read for flat hit indices [hitBegin, hitEnd) into dPositions[h - hitBegin]. */
int awfm_gpu_locate_device(awfm_gpu_ctx *ctx, const awfm_range *dRanges, const uint64_t *dHitOffsets,
                           uint64_t numQueries, uint64_t hitBegin, uint64_t hitEnd, uint64_t *dPositions,
                           void *stream);

/* ---- contig mapping for multi-sequence (FASTA) indexes — SURVEY.md §8 row f2.  Replaces, for whole batches of
 *      hits, awFmGetLocalSequencePositionFromIndexPosition (src/AwFmSearch.c:284-301), i.e.
 *      fastaVectorGetLocalSequencePositionFromGlobal (lib/FastaVector/src/FastaVector.c:338-381). ---- */
/* `metadata` = the reference's own struct FastaVectorMetadata[numSequences] (two size_t per record:
 * headerEndPosition, sequenceEndPosition; lib/FastaVector/src/FastaVectorMetadataVector.h:10-13), HOST memory, copied. */
int awfm_gpu_ctx_set_sequences(awfm_gpu_ctx *ctx, const void *metadata, uint64_t numSequences);
/* For every position: sequenceIndex = number of records ending at or before it, localPosition = offset inside that
 * record.  A position beyond the last record's end (the reference's AwFmIllegalPositionError) gives UINT64_MAX in
 * both outputs.  Device buffers, asynchronous on `stream`. */
int awfm_gpu_map_positions_device(awfm_gpu_ctx *ctx, const uint64_t *dPositions, uint64_t numPositions,
                                  uint64_t *dSequenceIndex, uint64_t *dLocalPosition, void *stream);
/* Same on HOST buffers; *numIllegal (may be NULL) receives the number of illegal positions. */
int awfm_gpu_map_positions_host(awfm_gpu_ctx *ctx, const uint64_t *positions, uint64_t numPositions,
                                uint64_t *sequenceIndex, uint64_t *localPosition, uint64_t *numIllegal);

/* ---- the reference's own search-list layout (used by the drop-in shim) ---- */
/* Fills data[i].count for i < numQueries (uint32 truncation as in src/AwFmParallelSearch.c:187-190). */
int awfm_gpu_search_list_count(awfm_gpu_ctx *ctx, awfm_kmer_search_data *data, uint64_t numQueries,
                               uint32_t numThreads);
/* Fills count + positionList with the reference's capacity semantics (realloc to exactly count when
 * count > capacity, src/AwFmParallelSearch.c:367-387).  Returns AWFM_GPU_ERR_ALLOC if a realloc failed.
 * Chunks of the list are pipelined: pack (host team) -> H2D + search + scan -> walk + D2H -> scatter (host team). */
int awfm_gpu_search_list_locate(awfm_gpu_ctx *ctx, awfm_kmer_search_data *data, uint64_t numQueries,
                                uint32_t numThreads);

/* ---- device groups: ONE call fanned out over several GPUs of the box (single process, index
 *      replicated in every GPU's HBM, query i -> GPU floor(i*G/N) in contiguous shards, one DMA stream set per GPU).
 *      Results land in the caller's HOST arrays directly, so no device-to-device gather exists on this path.
 *      The reference's parallel axis is inside the call as well (`omp parallel for`, src/AwFmParallelSearch.c:103-106,
 *      167-170).  A group of one device is the pipelined single-GPU engine. ---- */
typedef struct awfm_gpu_group awfm_gpu_group;
/* Replicates the index (HOST arrays in `view`) onto every listed device; numDevices = 0 / devices = NULL = all visible. */
int awfm_gpu_group_create(awfm_gpu_group **group, const int *devices, int numDevices, const awfm_index_view *view);
/* Groups contexts that already exist (distinct contexts holding the same index, e.g. one built on the device);
 * the group does not own them. */
int awfm_gpu_group_create_from_contexts(awfm_gpu_group **group, awfm_gpu_ctx *const *contexts, int numContexts);
void awfm_gpu_group_destroy(awfm_gpu_group *group);
int awfm_gpu_group_size(const awfm_gpu_group *group);
awfm_gpu_ctx *awfm_gpu_group_context(awfm_gpu_group *group, int i);
int awfm_gpu_group_set_sequences(awfm_gpu_group *group, const void *metadata, uint64_t numSequences);
/* keys: "packed_chunk_queries" (queries per pipeline chunk of awfm_gpu_group_count/_locate, multiple of 256; default
 * 3 * 2^23), "packed_min_shard" (a device is only given a shard of at least this many queries; default 2^16),
 * "packed_window_hits" (hits per walk window of awfm_gpu_group_locate; default 2^21); any other key is forwarded to
 * every context (awfm_gpu_ctx_set_tuning). */
int awfm_gpu_group_set_tuning(awfm_gpu_group *group, const char *key, int64_t value);
int awfm_gpu_group_get_stats(awfm_gpu_group *group, awfm_gpu_stats *out); /* summed over the devices, last call */

/* SURVEY.md §8 row f1 — the additive packed-batch API: contiguous queries in HOST memory (ASCII, 2-bit or 5-bit, see
 * awfm_query_format) in, flat arrays out; escapes the reference's per-query 32-B structs with pointers
 * (src/AwFmIndex.h:111-123; list setup src/AwFmParallelSearch.c:36-84).  Every device works through its shard in
 * chunks on three streams: the H2D copy of chunk i+1, the search of chunk i and the D2H copy of chunk i-1 overlap.
 * Buffers that are page-locked (awfm_gpu_host_alloc / awfm_gpu_host_register / cudaHostAlloc) are read and written by
 * the DMA engines in place; pageable ones go through page-locked staging.  `offsets` (numQueries+1 letter offsets) is
 * for variable-length ASCII batches, else NULL with every query `fixedLen` letters long. */
int awfm_gpu_group_count(awfm_gpu_group *group, const void *queries, uint32_t format, const uint64_t *offsets,
                         uint32_t fixedLen, uint64_t numQueries, uint32_t *counts /* out, numQueries */);
/* Locate, CSR output as awfm_gpu_locate_host: hitOffsets[numQueries+1], positions[hitOffsets[numQueries]] in SA order
 * within each query.  *totalHits (may be NULL) always receives hitOffsets[numQueries].  If positions is NULL or
 * positionsCapacity is smaller than that total, only hitOffsets are produced and the call returns AWFM_GPU_OK — size
 * the buffer and call again.  sequenceIndex / localPosition (both or neither; same length as positions) additionally
 * receive awFmGetLocalSequencePositionFromIndexPosition's answer for every hit (src/AwFmSearch.c:284-301), computed on
 * the device right after the walk (needs awfm_gpu_group_set_sequences). */
int awfm_gpu_group_locate(awfm_gpu_group *group, const void *queries, uint32_t format, const uint64_t *offsets,
                          uint32_t fixedLen, uint64_t numQueries, uint64_t *hitOffsets, uint64_t *positions,
                          uint64_t positionsCapacity, uint64_t *sequenceIndex, uint64_t *localPosition,
                          uint64_t *totalHits);
/* The reference's list layout over all devices of the group: chunk r of the list is searched on device r mod G. */
int awfm_gpu_group_search_list_count(awfm_gpu_group *group, awfm_kmer_search_data *data, uint64_t numQueries,
                                     uint32_t numThreads);
int awfm_gpu_group_search_list_locate(awfm_gpu_group *group, awfm_kmer_search_data *data, uint64_t numQueries,
                                      uint32_t numThreads);

/* page-locked host memory every device of the box can DMA from / to */
int awfm_gpu_host_alloc(void **p, uint64_t bytes);
void awfm_gpu_host_free(void *p);
int awfm_gpu_host_register(void *p, uint64_t bytes);
int awfm_gpu_host_unregister(void *p);

/* ---- result gather between PROCESSES over NVLink without a collective (one process per GPU): the root
 *      exports a device buffer, every other rank opens it and writes its shard of the results into its slot with a
 *      copy-engine peer copy on its own stream — no SM of either GPU is involved (an NCCL gather's channel CTAs share
 *      the SMs with the persistent search kernels).  handle = 64 bytes (cudaIpcMemHandle_t), to be passed between the
 *      processes by any means. ---- */
int awfm_gpu_device_malloc(int device, void **dPtr, uint64_t bytes);
int awfm_gpu_device_free(int device, void *dPtr);
int awfm_gpu_ipc_export(int device, const void *dPtr, void *handle64);
int awfm_gpu_ipc_open(int device, const void *handle64, void **dPtr);
int awfm_gpu_ipc_close(int device, void *dPtr);
/* dst/src: device pointers valid in this process (local, peer-mapped or IPC-opened); asynchronous on `stream` */
int awfm_gpu_peer_copy_async(int device, void *dst, const void *src, uint64_t bytes, void *stream);

/* ---- index construction on the device (replaces awFmCreateIndex, src/AwFmCreate.c:31-137,
 *      for texts with bwtLength < 2^32).  Output arrays are in the reference's own formats (raw 160/352-B blocks,
 *      prefix sums, seed table, bit-packed sampled SA) and stay on the device inside the handle.  The suffix order is
 *      a radix sort on the first 22 (amino: 13) symbols followed by prefix doubling over the suffixes that still tie,
 *      all on the device: repetitive texts cost O(log(longest repeat)) extra rounds. ---- */
typedef struct awfm_built_index awfm_built_index;
int awfm_gpu_build_index(awfm_built_index **built, int device, const uint8_t *dText /* DEVICE, ASCII */,
                         uint64_t textLength, uint8_t alphabet, uint8_t seedK, uint8_t saRatio);
int awfm_gpu_build_index_host(awfm_built_index **built, int device, const uint8_t *text /* HOST, ASCII */,
                              uint64_t textLength, uint8_t alphabet, uint8_t seedK, uint8_t saRatio);
/* `view` receives DEVICE pointers (feed it to awfm_gpu_ctx_create_from_device); valid until _destroy. */
int awfm_gpu_built_view(awfm_built_index *built, awfm_index_view *view, uint64_t *tieSuffixes, double *buildMs);
/* prefix-doubling rounds the suffix sort needed after its radix pass (0 = no suffixes tied) */
uint32_t awfm_gpu_built_tie_rounds(const awfm_built_index *built);
/* copies into caller-sized HOST buffers (sizes follow from the view); NULL pointers are skipped */
int awfm_gpu_built_download(awfm_built_index *built, void *blocks, uint64_t *prefixSums, void *seedTable,
                            uint8_t *saBytes);
void awfm_gpu_built_destroy(awfm_built_index *built);
/* splitmix64 letter stream (same as avxwindowfmindex_b200/synth.py): element i uses counter start + i */
int awfm_gpu_synth_letters(int device, uint8_t *dOut, uint64_t count, uint64_t seed, uint64_t start, int amino);

/* ---- measurement helper: random-gather bandwidth with this path's access shape ---- */
/* Reads `numReads` independent pseudo-random `bytesPerRead`-byte records (16, 32, 64 or 128, aligned to their
 * size) from a `arrayBytes` device buffer; returns achieved GB/s (bytes consumed / device time) in *gbps. */
int awfm_gpu_gather_bandwidth(int device, uint64_t arrayBytes, uint32_t bytesPerRead, uint64_t numReads,
                              int lanesPerRead, double *gbps);

/* Sets (bytes = 32, 64, 128) or just queries (bytes = 0) the device's L2->DRAM fetch granularity hint
 * (cudaLimitMaxL2FetchGranularity).  Random 64-B half-line reads waste half of every 128-B fetch otherwise. */
int awfm_gpu_set_l2_fetch_granularity(int device, int bytes, int *actual);

#ifdef __cplusplus
}
#endif
#endif /* AWFM_GPU_H */
