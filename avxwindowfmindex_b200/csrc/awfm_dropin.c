/*
 * awfm_dropin.c — the reference's translation unit src/AwFmParallelSearch.c re-provided over the B200 path.
 *
 * Exports exactly the four public symbols that TU defines (src/AwFmIndex.h:308, 326-327, 364-367, 400-403):
 *   awFmCreateKmerSearchList, awFmDeallocKmerSearchList, awFmParallelSearchCount, awFmParallelSearchLocate
 * with unchanged signatures, ownership rules and return codes, plus three additive awFmGpu* helpers.
 * Everything else of the reference library (index creation, file I/O, single-query helpers) is NOT here; link the
 * reference for those.  This file only unpacks struct AwFmIndex into the plain-C view of
 * include/awfm_gpu.h and forwards to the CUDA engine.  There is no CPU search path: if no CUDA device can be
 * used, Locate returns AwFmGeneralFailure and Count reports through awFmGpuLastCountStatus()/stderr.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/awfm_abi.h"
#include "../../include/awfm_gpu.h"

#define DEFAULT_POSITION_LIST_CAPACITY 4 /* src/AwFmParallelSearch.c:13 */
#define MAX_CACHED_INDEXES 16

struct CachedIndex {
  const struct AwFmIndex *index;
  const void *blocks, *seedTable; /* fingerprint: a recycled AwFmIndex* with other arrays is a different index */
  uint64_t bwtLength;
  awfm_gpu_ctx *ctx;
};
static struct CachedIndex gCache[MAX_CACHED_INDEXES];
static pthread_mutex_t gCacheLock = PTHREAD_MUTEX_INITIALIZER;
static __thread enum AwFmReturnCode gLastCountStatus = AwFmSuccess;

static int selectedDevice(void) {
  const char *e = getenv("AWFM_GPU_DEVICE");
  return e ? atoi(e) : 0;
}

static enum AwFmReturnCode mapStatus(int gpuStatus) {
  switch (gpuStatus) {
    case AWFM_GPU_OK: return AwFmSuccess;
    case AWFM_GPU_ERR_ALLOC: return AwFmAllocationFailure;
    case AWFM_GPU_ERR_ARG: return AwFmNullPtrError;
    case AWFM_GPU_ERR_NO_SA: return AwFmErrorSuffixArrayNull;
    default: return AwFmGeneralFailure;
  }
}

/* The reference leaves the sampled SA on disk when keepSuffixArrayInMemory is false and preads one value per hit
 * (src/AwFmFile.c:484-522).  The device needs it resident, so the whole section is read once; values are identical. */
static uint8_t *readSuffixArrayFromFile(const struct AwFmIndex *index) {
  const uint64_t bytes = index->suffixArray.compressedByteLength;
  uint8_t *buffer = malloc(bytes ? bytes : 1);
  if (!buffer) return NULL;
  uint64_t done = 0;
  while (done < bytes) {
    ssize_t got = pread(index->fileDescriptor, buffer + done, bytes - done, (off_t)(index->suffixArrayFileOffset + done));
    if (got <= 0) {
      free(buffer);
      return NULL;
    }
    done += (uint64_t)got;
  }
  return buffer;
}

static int contextFor(const struct AwFmIndex *index, awfm_gpu_ctx **out) {
  pthread_mutex_lock(&gCacheLock);
  int freeSlot = -1;
  for (int i = 0; i < MAX_CACHED_INDEXES; i++) {
    struct CachedIndex *c = &gCache[i];
    if (c->ctx && c->index == index) {
      if (c->blocks == index->bwtBlockList.asNucleotide && c->seedTable == index->kmerSeedTable &&
          c->bwtLength == index->bwtLength) {
        *out = c->ctx;
        pthread_mutex_unlock(&gCacheLock);
        return AWFM_GPU_OK;
      }
      awfm_gpu_ctx_destroy(c->ctx); /* same address, different index: stale entry */
      c->ctx = NULL;
    }
    if (!c->ctx && freeSlot < 0) freeSlot = i;
  }
  if (freeSlot < 0) { /* evict the oldest slot (round robin) */
    static int victim = 0;
    freeSlot = victim;
    victim = (victim + 1) % MAX_CACHED_INDEXES;
    awfm_gpu_ctx_destroy(gCache[freeSlot].ctx);
    gCache[freeSlot].ctx = NULL;
  }
  awfm_index_view view;
  memset(&view, 0, sizeof view);
  view.blocks = index->bwtBlockList.asNucleotide;
  view.bwtLength = index->bwtLength;
  view.numBlocks = 1 + (index->bwtLength - 1) / AW_FM_POSITIONS_PER_FM_BLOCK; /* src/AwFmIndexStruct.c:104-106 */
  view.prefixSums = index->prefixSums;
  view.seedTable = index->kmerSeedTable;
  view.saBitWidth = index->suffixArray.valueBitWidth;
  view.saRatio = index->config.suffixArrayCompressionRatio;
  view.seedK = index->config.kmerLengthInSeedTable;
  /* "!= AwFmAlphabetAmino" selects the nucleotide path in the reference (src/AwFmParallelSearch.c:250) */
  view.alphabet = index->config.alphabetType == AwFmAlphabetAmino ? 1 : 2;
  view.saByteLength = index->suffixArray.compressedByteLength;
  uint8_t *saFromFile = NULL;
  if (index->suffixArray.values) {
    view.saBytes = index->suffixArray.values;
  } else if (index->fileHandle && index->suffixArray.compressedByteLength) {
    saFromFile = readSuffixArrayFromFile(index);
    view.saBytes = saFromFile; /* NULL -> count-only context; locate then reports AwFmFileReadFail */
  }
  awfm_gpu_ctx *ctx = NULL;
  int rc = awfm_gpu_ctx_create(&ctx, selectedDevice(), &view);
  free(saFromFile);
  if (rc == AWFM_GPU_OK) {
    const char *e;
    if ((e = getenv("AWFM_GPU_COUNT_LPQ"))) awfm_gpu_ctx_set_tuning(ctx, "count_lpq", atoll(e));
    if ((e = getenv("AWFM_GPU_LOCATE_LPQ"))) awfm_gpu_ctx_set_tuning(ctx, "locate_lpq", atoll(e));
    if ((e = getenv("AWFM_GPU_COUNT_VARIANT"))) awfm_gpu_ctx_set_tuning(ctx, "count_variant", atoll(e));
    if ((e = getenv("AWFM_GPU_CHUNK_QUERIES"))) awfm_gpu_ctx_set_tuning(ctx, "chunk_queries", atoll(e));
    if ((e = getenv("AWFM_GPU_LOCATE_CHUNK_QUERIES"))) awfm_gpu_ctx_set_tuning(ctx, "locate_chunk_queries", atoll(e));
    if ((e = getenv("AWFM_GPU_LOCATE_INLINE_HITS"))) awfm_gpu_ctx_set_tuning(ctx, "locate_inline_hits", atoll(e));
    if ((e = getenv("AWFM_GPU_LOCATE_WINDOW_HITS"))) awfm_gpu_ctx_set_tuning(ctx, "locate_window_hits", atoll(e));
    if ((e = getenv("AWFM_GPU_LOCATE_VARIANT"))) awfm_gpu_ctx_set_tuning(ctx, "locate_variant", atoll(e));
    /* opt-in derived structures (include/awfm_gpu.h): deeper seed table, denser SA samples */
    if ((e = getenv("AWFM_GPU_SEED_DEPTH")) && atoi(e) > 0) rc = awfm_gpu_ctx_extend_seed_table(ctx, (uint32_t)atoi(e), NULL);
    if (rc == AWFM_GPU_OK && view.saBytes && (e = getenv("AWFM_GPU_SA_RATIO")) && atoi(e) > 0)
      rc = awfm_gpu_ctx_densify_suffix_array(ctx, (uint32_t)atoi(e), NULL);
    /* multi-sequence index: the record table rides along for awFmGpuGetLocalSequencePositions */
    if (rc == AWFM_GPU_OK && index->fastaVector && index->fastaVector->metadata.count)
      rc = awfm_gpu_ctx_set_sequences(ctx, index->fastaVector->metadata.data, index->fastaVector->metadata.count);
    if (rc != AWFM_GPU_OK) {
      awfm_gpu_ctx_destroy(ctx);
      pthread_mutex_unlock(&gCacheLock);
      return rc;
    }
    struct CachedIndex *c = &gCache[freeSlot];
    c->index = index;
    c->blocks = index->bwtBlockList.asNucleotide;
    c->seedTable = index->kmerSeedTable;
    c->bwtLength = index->bwtLength;
    c->ctx = ctx;
    *out = ctx;
  }
  pthread_mutex_unlock(&gCacheLock);
  return rc;
}

void awFmGpuReleaseIndex(const struct AwFmIndex *index) {
  pthread_mutex_lock(&gCacheLock);
  for (int i = 0; i < MAX_CACHED_INDEXES; i++) {
    if (gCache[i].ctx && gCache[i].index == index) {
      awfm_gpu_ctx_destroy(gCache[i].ctx);
      memset(&gCache[i], 0, sizeof gCache[i]);
    }
  }
  pthread_mutex_unlock(&gCacheLock);
}

enum AwFmReturnCode awFmGpuPrepareIndex(const struct AwFmIndex *index) {
  if (!index) return AwFmNullPtrError;
  awfm_gpu_ctx *ctx = NULL;
  return mapStatus(contextFor(index, &ctx));
}

enum AwFmReturnCode awFmGpuLastCountStatus(void) { return gLastCountStatus; }

enum AwFmReturnCode awFmGpuGetLocalSequencePositions(const struct AwFmIndex *index, const size_t *globalPositions,
                                                     size_t count, size_t *sequenceNumbers,
                                                     size_t *localSequencePositions) {
  if (!index || (count && (!globalPositions || !sequenceNumbers || !localSequencePositions))) return AwFmNullPtrError;
  if (!index->fastaVector) return AwFmUnsupportedVersionError; /* src/AwFmSearch.c:287-289 */
  awfm_gpu_ctx *ctx = NULL;
  int rc = contextFor(index, &ctx);
  uint64_t illegal = 0;
  if (rc == AWFM_GPU_OK)
    rc = awfm_gpu_map_positions_host(ctx, (const uint64_t *)globalPositions, count, (uint64_t *)sequenceNumbers,
                                     (uint64_t *)localSequencePositions, &illegal);
  if (rc != AWFM_GPU_OK) {
    fprintf(stderr, "awFmGpuGetLocalSequencePositions (B200): %s\n", awfm_gpu_last_error());
    return mapStatus(rc);
  }
  return illegal ? AwFmIllegalPositionError : AwFmSuccess;
}

/* src/AwFmParallelSearch.c:36-84: the list, its 32-B entries, and one 4-slot position list per entry */
struct AwFmKmerSearchList *awFmCreateKmerSearchList(const size_t capacity) {
  struct AwFmKmerSearchList *list = malloc(sizeof *list);
  if (!list) return NULL;
  list->capacity = capacity;
  list->count = 0;
  list->kmerSearchData = malloc(capacity * sizeof(struct AwFmKmerSearchData));
  if (!list->kmerSearchData) {
    free(list);
    return NULL;
  }
  size_t made = 0;
  for (; made < capacity; made++) {
    struct AwFmKmerSearchData *d = &list->kmerSearchData[made];
    d->kmerString = NULL;
    d->kmerLength = 0;
    d->count = 0;
    d->capacity = DEFAULT_POSITION_LIST_CAPACITY;
    d->positionList = malloc(DEFAULT_POSITION_LIST_CAPACITY * sizeof(uint64_t));
    if (!d->positionList) break;
  }
  if (made != capacity) { /* any allocation failure undoes everything and returns NULL, like the reference */
    for (size_t i = 0; i < made; i++) free(list->kmerSearchData[i].positionList);
    free(list->kmerSearchData);
    free(list);
    return NULL;
  }
  return list;
}

/* src/AwFmParallelSearch.c:86-93: frees `capacity` position lists (not `count`) */
void awFmDeallocKmerSearchList(struct AwFmKmerSearchList *restrict const searchList) {
  for (size_t i = 0; i < searchList->capacity; i++) free(searchList->kmerSearchData[i].positionList);
  free(searchList->kmerSearchData);
  free(searchList);
}

/* src/AwFmParallelSearch.c:159-220.  numThreads = host marshalling threads (the device ignores it). */
void awFmParallelSearchCount(const struct AwFmIndex *restrict const index,
                             struct AwFmKmerSearchList *restrict const searchList, uint32_t numThreads) {
  const uint32_t searchListCount = (uint32_t)searchList->count; /* uint32 truncation, :164 */
  awfm_gpu_ctx *ctx = NULL;
  int rc = contextFor(index, &ctx);
  if (rc == AWFM_GPU_OK)
    rc = awfm_gpu_search_list_count(ctx, (awfm_kmer_search_data *)searchList->kmerSearchData, searchListCount,
                                    numThreads);
  gLastCountStatus = mapStatus(rc);
  if (rc != AWFM_GPU_OK)
    fprintf(stderr, "awFmParallelSearchCount (B200): %s — counts were NOT computed (no CPU fallback)\n",
            awfm_gpu_last_error());
}

/* src/AwFmParallelSearch.c:95-157.  Returns AwFmSuccess, AwFmFileReadFail when the sampled SA cannot be obtained
 * (the reference's only failure), or AwFmAllocationFailure / AwFmGeneralFailure for device-side failures. */
enum AwFmReturnCode awFmParallelSearchLocate(const struct AwFmIndex *restrict const index,
                                             struct AwFmKmerSearchList *restrict const searchList,
                                             uint32_t numThreads) {
  const uint32_t searchListCount = (uint32_t)searchList->count; /* :100 */
  awfm_gpu_ctx *ctx = NULL;
  int rc = contextFor(index, &ctx);
  if (rc == AWFM_GPU_OK)
    rc = awfm_gpu_search_list_locate(ctx, (awfm_kmer_search_data *)searchList->kmerSearchData, searchListCount,
                                     numThreads);
  if (rc == AWFM_GPU_ERR_NO_SA) return AwFmFileReadFail;
  if (rc != AWFM_GPU_OK) fprintf(stderr, "awFmParallelSearchLocate (B200): %s\n", awfm_gpu_last_error());
  return mapStatus(rc);
}
