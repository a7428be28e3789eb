// awfm_internal.cuh — state shared by the translation units behind include/awfm_gpu.h:
//   awfm_b200.cu   index residency, launch dispatch, single-device entry points, the search-list engine
//   awfm_multi.cu  device groups: one call fanned out over several GPUs, the pipelined packed-batch engine
// Nothing here is part of the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/awfm_gpu.h"
#include "awfm_kernels.cuh"
#include "awfm_sweep.cuh"

// ------------------------------------------------------------------------------------------------ errors
int awfm_fail(int code, const char *what, const char *detail = nullptr);  // sets the thread's last-error text
#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      cudaGetLastError();                                                                            \
      return awfm_fail(e_ == cudaErrorMemoryAllocation ? AWFM_GPU_ERR_ALLOC                          \
                       : (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? AWFM_GPU_ERR_NO_DEVICE \
                                                                                        : AWFM_GPU_ERR_CUDA, \
                       #call, cudaGetErrorString(e_));                                               \
    }                                                                                                \
  } while (0)

// ------------------------------------------------------------------------------------------------ context
struct EventPair {
  cudaEvent_t a, b;
};

struct GrowBuf {  // device (or pinned host) scratch that only ever grows; owned by a lane
  void *p = nullptr;
  size_t cap = 0;
  bool host = false;
  int ensure(size_t bytes);
  void release();
};

struct LocateScratch {  // per-stream scratch of scan + walk (two streams may not share one)
  void *scanTemp = nullptr;  // block sums of the range-length scan
  size_t scanTempBytes = 0;
  unsigned long long *dWorkCounter = nullptr;  // locateKernelRefill's chunk dispenser
};

struct SweepScratch {  // buffers of the sweep count path (awfm_sweep.cuh), grown on demand, one call at a time
  uint64_t cap = 0;                        // queries the buffers hold
  void *arena = nullptr;                   // one allocation carved into the buffers below
  uint32_t *keys[2] = {nullptr, nullptr};  // seed-table index per query, radix-sort double buffer
  uint64_t *vals[2] = {nullptr, nullptr};  // (remaining letters << 32) | query id
  uint4 *recs[2][awfm::kSweepMaxArrays] = {};  // two generations x (2 | 10) double-ended arrays
  int arrays = 0;                          // arrays per generation the arena was carved for
  uint32_t *ctrl = nullptr;                // [kSweepMaxPasses][stride] bucket counters, then the irregular-query counter
  uint32_t *irregularIds = nullptr;
  uint32_t *more = nullptr;                // [cap] letters 17.. left of the seed k-mer (sweepRefill)
  void *sortCtrl = nullptr;                // awfm_sort.cuh: SortCtrl + group counts + group cursors
  void *sortTemp = nullptr;                // CUB's temporary storage (seed tables deeper than 2^24 entries)
  size_t sortTempBytes = 0;
  cudaEvent_t done = nullptr;              // end of the last sweep: the next one (possibly on another stream) waits
  cudaEvent_t stage[awfm::kSweepMaxPasses + 4];  // stage boundaries of the most recent call ("sweep_profile")
  int numStages = 0, stagesRecorded = 0;
  uint64_t bytes = 0;
  uint32_t lastSteps = 0, lastBuckets = 0;  // passes and buckets of the most recent sweep (for awfm_gpu_ctx_sweep_live)
  uint64_t lastQueries = 0;
};

struct PipeSlot {  // one in-flight chunk of the search-list engine
  uint8_t *hLetters = nullptr, *dLetters = nullptr;
  uint64_t lettersCap = 0, dLettersCap = 0;
  uint64_t *hOffsets = nullptr, *dOffsets = nullptr;
  uint32_t *hCounts = nullptr, *dCounts = nullptr;
  uint32_t *hOld = nullptr;  // count engine: the counts the chunk's entries held when they were packed (pageable)
  uint4 *dRanges = nullptr;
  uint64_t queryCap = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  uint64_t first = 0, n = 0;
  bool busy = false;
  bool ready = false;  // count engine: D2H complete, counts not scattered yet
  // locate pipeline: hit offsets and positions of the chunk, both sides of the bus
  LocateScratch sc;
  uint64_t *hHit = nullptr, *dHit = nullptr, hitCap = 0;
  uint64_t *hPos = nullptr, *dPos = nullptr, posCap = 0;
  cudaEvent_t offsetsDone = nullptr;
  uint64_t total = 0;
  bool shipped = false, walked = false, big = false;
};

// Working state of ONE call.  A host-buffer / search-list call holds exactly one lane from entry to return, so calls
// from different host threads on the same index run concurrently on different lanes (the reference's entry points are
// re-entrant for distinct lists on a shared const index, src/AwFmParallelSearch.c:95-220).  Lane 0 also serves the
// asynchronous device-buffer entry points, which do not lock (the caller orders them through its stream).
struct Lane {
  std::mutex mu;
  bool ready = false;  // streams / events / counters created (lazily, on the context's device)
  LocateScratch sc;    // scratch of the device-/host-buffer calls (the list engine's slots have their own)
  SweepScratch sweep;
  uint64_t *hBigPos = nullptr, *dBigPos = nullptr, bigPosCap = 0;  // windowed positions of a chunk with very many hits
  GrowBuf dLetters, dOffsets, dCounts, dRanges, dHits, dPositions;  // persistent buffers of the *_host calls
  GrowBuf dUnpacked;  // ASCII letters of a 2-/5-bit batch that takes the tile kernels
  cudaEvent_t unpackDone = nullptr;  // last reader of dUnpacked: the next unpack (possibly on another stream) waits
  std::vector<EventPair> kernelEvents;  // of the most recent call
  size_t eventsUsed = 0;
  awfm_gpu_stats stats{};
  static constexpr int kSlots = 6;
  PipeSlot slots[kSlots];
};

struct awfm_gpu_ctx {
  int device = 0, numSMs = 0;
  awfm::DevIndex ix{};
  void *dLines = nullptr, *dXRel16 = nullptr, *dSuperC = nullptr, *dSeed = nullptr, *dSa = nullptr;
  uint64_t *dSequenceEnds = nullptr;
  void *dDeepSeed = nullptr, *dDenseSa = nullptr;  // derived structures (extend_seed_table / densify_suffix_array)
  uint64_t deepSeedBytes = 0, denseSaBytes = 0;
  uint32_t deepSeedKBuilt = 0;
  const void *origSa = nullptr;  // the index's own sampled SA, restored when the dense one is dropped
  uint32_t origSaBitWidth = 0, origSaRatio = 0, origSaRatioShift = 0;
  std::atomic<uint64_t> deviceBytes{0};
  bool hasSa = false;
  // tuning
  int countLpq = 2, locateLpq = 2, countVariant = 1, locateVariant = 1, ctaThreads = 256, blocksPerSm = 0 /* 0 = occupancy */;
  int64_t chunkQueries = 0;  // list count engine: queries per chunk; 0 = automatic (awfm_list_engine.inc)
  int64_t locateChunkQueries = 1 << 18;
  int64_t locateInlineHits = 1 << 22;  // a chunk with more hits than this is finished through windows of ...
  int64_t locateWindowHits = 1 << 26;  // ... this many flat hit indices
  int64_t sweepMinQueries = 0;  // 0 = automatic (see sweepEligible); 1 = whenever the batch qualifies; < 0 = never
  int64_t sweepMaxBatch = 1ll << 27;
  int sweepSortBits = 32, sweepLocalBits = -1 /* automatic */, sweepProfile = 0, sweepItems = 4, sweepFirstItems = 4;
  int sweepOrderedEmit = 1;  // range output of batches of up to 2^24 queries: survivors leave through sweepEmit (awfm_sweep.cuh)
  int sweepRecord12 = 0;  // 12-byte live records when the batch allows it (nucleotide, <= 8 letters left of the seed)
  int sweepWide = 0;      // 1: the 64-bit-position passes even on an index below 2^32 positions (cross-check)
  int sweepVariable = 1;  // variable-length batches may take the sweep (marker-bit payloads, sweepPackVar)
  int sweepCompactPairs = 1;  // 8-byte pairs through the bucket passes when key, payload and id fit (awfm_sort.cuh)
  int sweepOwnSort = 1;  // 1 = the hand-written stable radix passes (awfm_sort.cuh), 0 = CUB (cross-check)
  static constexpr int kLanes = 3;
  Lane lanes[kLanes];
  std::atomic<int> lastLane{0};  // lane of the most recent call: what get_stats / sweep_stage_ms report
  std::mutex mu;                 // structural changes (derived structures, record table): taken with every lane
};

// RAII: a lane held for the duration of one call.
struct LaneHold {
  awfm_gpu_ctx *c = nullptr;
  Lane *lane = nullptr;
  int rc = AWFM_GPU_OK;
  explicit LaneHold(awfm_gpu_ctx *ctx);  // sets the device, takes a free lane of 1..kLanes-1 (or waits), prepares it
  ~LaneHold();
  Lane &operator*() { return *lane; }
};
struct AllLanesHold {  // exclusive access: every lane + the context mutex
  awfm_gpu_ctx *c;
  explicit AllLanesHold(awfm_gpu_ctx *ctx);
  ~AllLanesHold();
};

int awfm_set_device(const awfm_gpu_ctx *c);
int awfm_lane_prepare(awfm_gpu_ctx *c, Lane &L);  // lazily creates the lane's streams / events (device must be current)
void awfm_begin_call(Lane &L);

// what the packed batch holds per letter (include/awfm_gpu.h: awfm_query_format)
struct PackedBatch {
  const uint8_t *data = nullptr;    // DEVICE pointer, 16-B aligned
  const uint64_t *offsets = nullptr;  // ASCII only: numQueries+1 letter offsets, or nullptr (fixed length)
  uint32_t format = AWFM_QUERY_ASCII;
  uint32_t length = 0;              // letters per query (fixed-length batches)
  uint32_t maxLength = 0;           // variable-length batches: longest query if the caller knows it, else 0
  uint64_t numQueries = 0;
  bool rangesOfHitsOnly = false;    // locate pipelines: dRanges is only written for queries with hits (the hit offsets
                                    // are then scanned from dCounts, awfm_scan_impl with fromCounts)
};
static inline uint32_t awfm_format_bits(uint32_t format) {
  return format == AWFM_QUERY_2BIT ? 2u : format == AWFM_QUERY_5BIT ? 5u : 8u;
}
static inline uint64_t awfm_query_bytes(uint32_t format, uint32_t length) {
  return ((uint64_t)length * awfm_format_bits(format) + 7) / 8;
}

// ---- device-side building blocks used by both translation units (all asynchronous on `st`) ----
int awfm_count_device_impl(awfm_gpu_ctx *c, Lane &L, const PackedBatch &batch, uint32_t *dCounts, awfm_range *dRanges,
                           cudaStream_t st);
// hit offsets = exclusive scan of the hit-list lengths, taken from the final ranges (lengthSource = awfm_range[n],
// fromCounts = false) or from the u32 counts of the same search (uint32_t[n], fromCounts = true)
int awfm_scan_impl(awfm_gpu_ctx *c, Lane &L, LocateScratch &sc, const void *lengthSource, bool fromCounts, uint64_t n,
                   uint64_t *dHitOffsets, uint64_t base, cudaStream_t st);
int awfm_locate_device_impl(awfm_gpu_ctx *c, Lane &L, LocateScratch &sc, const awfm_range *dRanges,
                            const uint64_t *dHitOffsets, uint64_t n, uint64_t hb, uint64_t he, uint64_t *dPos,
                            cudaStream_t st);
int awfm_map_device_impl(awfm_gpu_ctx *c, Lane &L, const uint64_t *dPos, uint64_t n, uint64_t *dSeq, uint64_t *dLocal,
                         cudaStream_t st);
bool awfm_is_pinned_host(const void *p);
int awfm_search_list_run(awfm_gpu_ctx *const *ctxs, int numContexts, awfm_kmer_search_data *data, uint64_t n,
                         uint32_t numThreads, bool locate);
