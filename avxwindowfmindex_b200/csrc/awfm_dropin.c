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
#define MAX_DEVICES 64

/* One device-resident replica set (a group of one or more GPUs) per struct AwFmIndex the caller searches with.
 * The reference API has no hook on awFmDeallocIndex, and malloc commonly hands a new index the addresses of a freed
 * one, so an entry is recognised by a FINGERPRINT of the index contents (header fields, array addresses, the prefix
 * sums and samples of the block list, seed table and suffix array), recomputed on every call — not by addresses alone.
 * Entries are reference-counted: a stale or evicted entry that another thread is still searching with is only
 * unlinked, and destroyed by its last user. */
struct CachedIndex {
  const struct AwFmIndex *index;
  uint64_t fingerprint;
  awfm_gpu_group *group;
  int users; /* calls in flight */
  int dead;  /* unlinked from the table: the last user destroys it */
};
static struct CachedIndex *gCache[MAX_CACHED_INDEXES];
static pthread_mutex_t gCacheLock = PTHREAD_MUTEX_INITIALIZER;
static __thread enum AwFmReturnCode gLastCountStatus = AwFmSuccess;

/* AWFM_GPU_DEVICES = "all" or a comma-separated list of CUDA ordinals: every batched call is fanned out over those
 * GPUs from inside the library (index replicated on each).  Otherwise one GPU: AWFM_GPU_DEVICE (default 0). */
static int selectedDevices(int *devices) {
  const char *list = getenv("AWFM_GPU_DEVICES");
  if (list && *list) {
    if (strcmp(list, "all") == 0) {
      int n = awfm_gpu_device_count();
      if (n > MAX_DEVICES) n = MAX_DEVICES;
      for (int i = 0; i < n; i++) devices[i] = i;
      return n > 0 ? n : 1; /* no device: let group creation report it */
    }
    int n = 0;
    const char *p = list;
    while (*p && n < MAX_DEVICES) {
      char *end = NULL;
      long v = strtol(p, &end, 10);
      if (end == p) break;
      devices[n++] = (int)v;
      p = (*end == ',') ? end + 1 : end;
      if (*end != ',') break;
    }
    if (n > 0) return n;
  }
  const char *e = getenv("AWFM_GPU_DEVICE");
  devices[0] = e ? atoi(e) : 0;
  return 1;
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

static inline uint64_t mix(uint64_t h, uint64_t v) { /* FNV-1a over 64-bit words, then a finaliser per word */
  h = (h ^ v) * 0x100000001B3ull;
  return h ^ (h >> 29);
}
static uint64_t hashBytes(uint64_t h, const void *p, size_t bytes) {
  const uint8_t *b = p;
  for (size_t i = 0; i + 8 <= bytes; i += 8) {
    uint64_t w;
    memcpy(&w, b + i, 8);
    h = mix(h, w);
  }
  return h;
}
static uint64_t numSeedsOf(const struct AwFmIndex *index) {
  const uint64_t card = index->config.alphabetType == AwFmAlphabetAmino ? 20 : 4;
  uint64_t n = 1;
  for (int i = 0; i < index->config.kmerLengthInSeedTable; i++) n *= card;
  return n;
}
static uint64_t fingerprintOf(const struct AwFmIndex *index) {
  const int amino = index->config.alphabetType == AwFmAlphabetAmino;
  const uint64_t blockBytes = amino ? sizeof(struct AwFmAminoBlock) : sizeof(struct AwFmNucleotideBlock);
  const uint64_t numBlocks = 1 + (index->bwtLength - 1) / AW_FM_POSITIONS_PER_FM_BLOCK;
  uint64_t h = 0xCBF29CE484222325ull;
  h = mix(h, index->bwtLength);
  h = mix(h, (uint64_t)(uintptr_t)index->bwtBlockList.asNucleotide);
  h = mix(h, (uint64_t)(uintptr_t)index->kmerSeedTable);
  h = mix(h, (uint64_t)(uintptr_t)index->prefixSums);
  h = mix(h, (uint64_t)(uintptr_t)index->suffixArray.values);
  h = mix(h, index->suffixArray.compressedByteLength);
  h = mix(h, ((uint64_t)index->config.suffixArrayCompressionRatio << 16) | ((uint64_t)index->config.kmerLengthInSeedTable << 8) |
                 (uint64_t)index->config.alphabetType);
  h = mix(h, (uint64_t)(uintptr_t)index->fastaVector);
  h = hashBytes(h, index->prefixSums, (amino ? 22 : 6) * 8);
  const uint8_t *blocks = (const uint8_t *)index->bwtBlockList.asNucleotide;
  for (uint64_t i = 0; i < 48; i++) { /* whole blocks, evenly spread, always including the first and the last */
    const uint64_t b = numBlocks <= 48 ? i : i * (numBlocks - 1) / 47;
    if (b >= numBlocks) break;
    h = hashBytes(h, blocks + b * blockBytes, blockBytes);
  }
  const uint64_t numSeeds = numSeedsOf(index);
  for (uint64_t i = 0; i < 64; i++) {
    const uint64_t e = numSeeds <= 64 ? i : i * (numSeeds - 1) / 63;
    if (e >= numSeeds) break;
    h = hashBytes(h, &index->kmerSeedTable[e], sizeof(struct AwFmSearchRange));
  }
  if (index->suffixArray.values && index->suffixArray.compressedByteLength >= 8) {
    const uint64_t words = index->suffixArray.compressedByteLength / 8;
    for (uint64_t i = 0; i < 64; i++) {
      const uint64_t w = words <= 64 ? i : i * (words - 1) / 63;
      if (w >= words) break;
      h = hashBytes(h, index->suffixArray.values + 8 * w, 8);
    }
  }
  return h;
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

static void destroyEntry(struct CachedIndex *c) {
  awfm_gpu_group_destroy(c->group);
  free(c);
}

static int createGroup(const struct AwFmIndex *index, awfm_gpu_group **out) {
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
  int devices[MAX_DEVICES];
  const int numDevices = selectedDevices(devices);
  awfm_gpu_group *group = NULL;
  int rc = awfm_gpu_group_create(&group, devices, numDevices, &view);
  free(saFromFile);
  if (rc != AWFM_GPU_OK) return rc;
  static const char *const knobs[][2] = {
      {"AWFM_GPU_COUNT_LPQ", "count_lpq"}, {"AWFM_GPU_LOCATE_LPQ", "locate_lpq"},
      {"AWFM_GPU_COUNT_VARIANT", "count_variant"}, {"AWFM_GPU_LOCATE_VARIANT", "locate_variant"},
      {"AWFM_GPU_CHUNK_QUERIES", "chunk_queries"}, {"AWFM_GPU_LOCATE_CHUNK_QUERIES", "locate_chunk_queries"},
      {"AWFM_GPU_LOCATE_INLINE_HITS", "locate_inline_hits"}, {"AWFM_GPU_LOCATE_WINDOW_HITS", "locate_window_hits"},
      {"AWFM_GPU_PACKED_CHUNK_QUERIES", "packed_chunk_queries"}, {"AWFM_GPU_PACKED_MIN_SHARD", "packed_min_shard"},
      {"AWFM_GPU_SWEEP_MIN_QUERIES", "sweep_min_queries"}};
  for (size_t i = 0; i < sizeof knobs / sizeof knobs[0]; i++) {
    const char *e = getenv(knobs[i][0]);
    if (e) awfm_gpu_group_set_tuning(group, knobs[i][1], atoll(e));
  }
  /* opt-in derived structures (include/awfm_gpu.h): deeper seed table, denser SA samples */
  const char *e;
  for (int d = 0; d < awfm_gpu_group_size(group) && rc == AWFM_GPU_OK; d++) {
    awfm_gpu_ctx *ctx = awfm_gpu_group_context(group, d);
    if ((e = getenv("AWFM_GPU_SEED_DEPTH")) && atoi(e) > 0) rc = awfm_gpu_ctx_extend_seed_table(ctx, (uint32_t)atoi(e), NULL);
    if (rc == AWFM_GPU_OK && view.saBytes && (e = getenv("AWFM_GPU_SA_RATIO")) && atoi(e) > 0)
      rc = awfm_gpu_ctx_densify_suffix_array(ctx, (uint32_t)atoi(e), NULL);
  }
  /* multi-sequence index: the record table rides along for the contig mapping */
  if (rc == AWFM_GPU_OK && index->fastaVector && index->fastaVector->metadata.count)
    rc = awfm_gpu_group_set_sequences(group, index->fastaVector->metadata.data, index->fastaVector->metadata.count);
  if (rc != AWFM_GPU_OK) {
    awfm_gpu_group_destroy(group);
    return rc;
  }
  *out = group;
  return AWFM_GPU_OK;
}

/* Returns the entry for `index` with its user count raised; pair with releaseEntry(). */
static int acquireEntry(const struct AwFmIndex *index, struct CachedIndex **out) {
  const uint64_t fp = fingerprintOf(index);
  pthread_mutex_lock(&gCacheLock);
  int freeSlot = -1;
  for (int i = 0; i < MAX_CACHED_INDEXES; i++) {
    struct CachedIndex *c = gCache[i];
    if (c && c->index == index) {
      if (c->fingerprint == fp) {
        c->users++;
        *out = c;
        pthread_mutex_unlock(&gCacheLock);
        return AWFM_GPU_OK;
      }
      gCache[i] = NULL; /* same address, different contents: the index was deallocated and another one built */
      c->dead = 1;
      if (c->users == 0) destroyEntry(c);
      c = NULL;
    }
    if (!gCache[i] && freeSlot < 0) freeSlot = i;
  }
  if (freeSlot < 0) { /* evict the oldest slot (round robin) */
    static int victim = 0;
    freeSlot = victim;
    victim = (victim + 1) % MAX_CACHED_INDEXES;
    struct CachedIndex *c = gCache[freeSlot];
    gCache[freeSlot] = NULL;
    c->dead = 1;
    if (c->users == 0) destroyEntry(c);
  }
  /* the upload happens under the table lock: concurrent first calls on the same index must not upload it twice */
  awfm_gpu_group *group = NULL;
  int rc = createGroup(index, &group);
  if (rc == AWFM_GPU_OK) {
    struct CachedIndex *c = calloc(1, sizeof *c);
    if (!c) {
      awfm_gpu_group_destroy(group);
      pthread_mutex_unlock(&gCacheLock);
      return AWFM_GPU_ERR_ALLOC;
    }
    c->index = index;
    c->fingerprint = fp;
    c->group = group;
    c->users = 1;
    gCache[freeSlot] = c;
    *out = c;
  }
  pthread_mutex_unlock(&gCacheLock);
  return rc;
}

static void releaseEntry(struct CachedIndex *c) {
  pthread_mutex_lock(&gCacheLock);
  const int last = --c->users == 0 && c->dead;
  pthread_mutex_unlock(&gCacheLock);
  if (last) destroyEntry(c);
}

void awFmGpuReleaseIndex(const struct AwFmIndex *index) {
  pthread_mutex_lock(&gCacheLock);
  for (int i = 0; i < MAX_CACHED_INDEXES; i++) {
    struct CachedIndex *c = gCache[i];
    if (c && c->index == index) {
      gCache[i] = NULL;
      c->dead = 1;
      if (c->users == 0) destroyEntry(c);
    }
  }
  pthread_mutex_unlock(&gCacheLock);
}

enum AwFmReturnCode awFmGpuPrepareIndex(const struct AwFmIndex *index) {
  if (!index) return AwFmNullPtrError;
  struct CachedIndex *entry = NULL;
  const int rc = acquireEntry(index, &entry);
  if (rc == AWFM_GPU_OK) releaseEntry(entry);
  return mapStatus(rc);
}

int awFmGpuNumDevices(const struct AwFmIndex *index) {
  if (!index) return 0;
  struct CachedIndex *entry = NULL;
  if (acquireEntry(index, &entry) != AWFM_GPU_OK) return 0;
  const int n = awfm_gpu_group_size(entry->group);
  releaseEntry(entry);
  return n;
}

enum AwFmReturnCode awFmGpuLastCountStatus(void) { return gLastCountStatus; }

enum AwFmReturnCode awFmGpuGetLocalSequencePositions(const struct AwFmIndex *index, const size_t *globalPositions,
                                                     size_t count, size_t *sequenceNumbers,
                                                     size_t *localSequencePositions) {
  if (!index || (count && (!globalPositions || !sequenceNumbers || !localSequencePositions))) return AwFmNullPtrError;
  if (!index->fastaVector) return AwFmUnsupportedVersionError; /* src/AwFmSearch.c:287-289 */
  struct CachedIndex *entry = NULL;
  int rc = acquireEntry(index, &entry);
  uint64_t illegal = 0;
  if (rc == AWFM_GPU_OK) {
    rc = awfm_gpu_map_positions_host(awfm_gpu_group_context(entry->group, 0), (const uint64_t *)globalPositions, count,
                                     (uint64_t *)sequenceNumbers, (uint64_t *)localSequencePositions, &illegal);
    releaseEntry(entry);
  }
  if (rc != AWFM_GPU_OK) {
    fprintf(stderr, "awFmGpuGetLocalSequencePositions (B200): %s\n", awfm_gpu_last_error());
    return mapStatus(rc);
  }
  return illegal ? AwFmIllegalPositionError : AwFmSuccess;
}

/* ---- the additive packed-batch API (include/awfm_abi.h) ---- */
enum AwFmReturnCode awFmGpuCountPacked(const struct AwFmIndex *index, const void *kmers, enum AwFmGpuKmerFormat format,
                                       const uint64_t *kmerOffsets, uint32_t kmerLength, uint64_t numKmers,
                                       uint32_t *counts) {
  if (!index) return AwFmNullPtrError;
  struct CachedIndex *entry = NULL;
  int rc = acquireEntry(index, &entry);
  if (rc == AWFM_GPU_OK) {
    rc = awfm_gpu_group_count(entry->group, kmers, (uint32_t)format, kmerOffsets, kmerLength, numKmers, counts);
    releaseEntry(entry);
  }
  if (rc != AWFM_GPU_OK) fprintf(stderr, "awFmGpuCountPacked (B200): %s\n", awfm_gpu_last_error());
  return mapStatus(rc);
}

enum AwFmReturnCode awFmGpuLocatePacked(const struct AwFmIndex *index, const void *kmers, enum AwFmGpuKmerFormat format,
                                        const uint64_t *kmerOffsets, uint32_t kmerLength, uint64_t numKmers,
                                        uint64_t *hitOffsets, uint64_t *positions, uint64_t positionsCapacity,
                                        uint64_t *sequenceNumbers, uint64_t *localSequencePositions,
                                        uint64_t *totalHits) {
  if (!index) return AwFmNullPtrError;
  struct CachedIndex *entry = NULL;
  int rc = acquireEntry(index, &entry);
  if (rc == AWFM_GPU_OK) {
    rc = awfm_gpu_group_locate(entry->group, kmers, (uint32_t)format, kmerOffsets, kmerLength, numKmers, hitOffsets,
                               positions, positionsCapacity, sequenceNumbers, localSequencePositions, totalHits);
    releaseEntry(entry);
  }
  if (rc == AWFM_GPU_ERR_NO_SA) return AwFmFileReadFail;
  if (rc != AWFM_GPU_OK) fprintf(stderr, "awFmGpuLocatePacked (B200): %s\n", awfm_gpu_last_error());
  return mapStatus(rc);
}

void *awFmGpuHostAlloc(size_t bytes) {
  void *p = NULL;
  return awfm_gpu_host_alloc(&p, bytes) == AWFM_GPU_OK ? p : NULL;
}
void awFmGpuHostFree(void *p) { awfm_gpu_host_free(p); }

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
  struct CachedIndex *entry = NULL;
  int rc = acquireEntry(index, &entry);
  if (rc == AWFM_GPU_OK) {
    rc = awfm_gpu_group_search_list_count(entry->group, (awfm_kmer_search_data *)searchList->kmerSearchData,
                                          searchListCount, numThreads);
    releaseEntry(entry);
  }
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
  struct CachedIndex *entry = NULL;
  int rc = acquireEntry(index, &entry);
  if (rc == AWFM_GPU_OK) {
    rc = awfm_gpu_group_search_list_locate(entry->group, (awfm_kmer_search_data *)searchList->kmerSearchData,
                                           searchListCount, numThreads);
    releaseEntry(entry);
  }
  if (rc == AWFM_GPU_ERR_NO_SA) return AwFmFileReadFail;
  if (rc != AWFM_GPU_OK) fprintf(stderr, "awFmParallelSearchLocate (B200): %s\n", awfm_gpu_last_error());
  return mapStatus(rc);
}
