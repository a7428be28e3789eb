// awfm_sort.cuh — the ordering step of the sweep count path (awfm_sweep.cuh), hand-written for sm_100a.
//
// What the sweep needs is weaker than a sort: the (key, payload) pairs GROUPED by the top S bits of the key, groups in
// ascending order, any order inside a group (the first sweep pass finishes the low bits inside each of its tiles, and a
// query's result does not depend on the processing order).  No stability is required, so there is no look-back chain
// and no spinning anywhere: two most-significant-digit-first bucket passes of up to 8 bits each,
//
//   pack kernel        also counts the top digit A (shared-memory histogram per CTA, 2^dA atomics per CTA at the end)
//   sortBases<0>       exclusive scan of the 2^dA bucket counts -> bucket bases, append cursors, tile map of pass B
//   sortPass<false>    tiles of kSortTile pairs in input order: counting sort by digit A in shared memory (rank = one
//                      shared atomic per pair), ONE global atomicAdd per tile and bucket claims room behind the
//                      bucket's cursor, pairs leave the tile bucket by bucket (runs of ~kSortTile/2^dA neighbours)
//   sortDigitCounts    tiles inside the bucket regions: shared histogram of digit B, 2^dB atomics per tile
//   sortBases<1>       exclusive scan of the 2^(dA+dB) group counts (one CTA per bucket)
//   sortPass<true>     the same tile body on digit B, tiles cut so that none straddles two buckets
//
// against CUB's onesweep (round 1: 1.68 ms for 100 M pairs, 3.0 TB/s, profiles/r01_ncu_sweep_cub_onesweep.json) the pairs
// are moved the same two times, but the passes carry no decoupled look-back and no per-tile status polling.  CUB stays
// as the path for S > 16 (seed tables deeper than 2^24 entries) and as a cross-check ("sweep_own_sort" = 0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace awfm {

constexpr int kSortThreads = 256;
#ifndef AWFM_SORT_ITEMS
#define AWFM_SORT_ITEMS 12
#endif
constexpr int kSortItems = AWFM_SORT_ITEMS;
constexpr int kSortTile = kSortThreads * kSortItems;
#ifndef AWFM_SORT_MIN_CTAS
#define AWFM_SORT_MIN_CTAS 4
#endif
constexpr int kSortMaxDigitBits = 8;
constexpr int kSortBins = 1 << kSortMaxDigitBits;

// control block of one sort (device memory, zeroed before the pack kernel)
struct SortCtrl {
  uint32_t countA[kSortBins];       // pairs per top-digit bucket (pack kernel)
  uint32_t baseA[kSortBins + 1];    // first slot of every bucket (+ total)
  uint32_t cursorA[kSortBins];      // append cursor of pass A
  uint32_t tilesBefore[kSortBins + 1];  // pass B: tiles in the buckets before a (+ total)
  uint32_t ticket[4];               // tile tickets of pass A, the digit count, pass B
};
// followed in the same allocation by countB[2^(dA+dB)] and cursorB[2^(dA+dB)]

template <int LEVEL>  // 0: bucket bases from countA; 1: group bases from countB (grid = 2^dA CTAs)
__global__ void __launch_bounds__(kSortThreads)
    sortBases(SortCtrl *__restrict__ ctrl, uint32_t *__restrict__ countB, uint32_t *__restrict__ cursorB, uint32_t dA,
              uint32_t dB, uint32_t tilePairs /* pairs per tile of pass B */) {
  __shared__ uint32_t warpSums[kSortThreads / 32];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  auto blockExclusive = [&](uint32_t v, uint32_t &total) {  // exclusive scan over the CTA's 256 threads
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= (unsigned)d) incl += up;
    }
    __syncthreads();
    if (lane == 31u) warpSums[warp] = incl;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (unsigned w = 0; w < kSortThreads / 32; w++) {
      const uint32_t s = warpSums[w];
      before += w < warp ? s : 0u;
      all += s;
    }
    total = all;
    return before + incl - v;
  };
  if (LEVEL == 0) {
    const uint32_t bins = 1u << dA;
    const uint32_t c = threadIdx.x < bins ? ctrl->countA[threadIdx.x] : 0u;
    uint32_t total, tilesTotal;
    const uint32_t base = blockExclusive(c, total);
    const uint32_t tiles = (c + tilePairs - 1) / tilePairs;
    const uint32_t tb = blockExclusive(tiles, tilesTotal);
    if (threadIdx.x < bins) {
      ctrl->baseA[threadIdx.x] = base;
      ctrl->cursorA[threadIdx.x] = base;
      ctrl->tilesBefore[threadIdx.x] = tb;
    }
    if (threadIdx.x == 0) {
      ctrl->baseA[bins] = total;
      ctrl->tilesBefore[bins] = tilesTotal;
    }
  } else {
    const uint32_t a = blockIdx.x, bins = 1u << dB;
    const uint32_t c = threadIdx.x < bins ? countB[(a << dB) + threadIdx.x] : 0u;
    uint32_t total;
    const uint32_t local = blockExclusive(c, total);
    if (threadIdx.x < bins) cursorB[(a << dB) + threadIdx.x] = ctrl->baseA[a] + local;
  }
}

// pass B's tile t -> (bucket a, first pair, number of pairs); tiles never straddle two buckets
__device__ __forceinline__ void sortTileOfBucket(const SortCtrl *__restrict__ ctrl, uint32_t numBuckets, uint32_t tile,
                                                 uint32_t tilePairs, uint32_t &a, uint32_t &first, uint32_t &count) {
  uint32_t lo = 0, hi = numBuckets;  // last a with tilesBefore[a] <= tile
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (ctrl->tilesBefore[mid] <= tile) lo = mid;
    else hi = mid;
  }
  a = lo;
  const uint32_t j = tile - ctrl->tilesBefore[a];
  const uint32_t begin = ctrl->baseA[a], end = ctrl->baseA[a + 1];
  first = begin + j * tilePairs;
  count = min(tilePairs, end - first);
}

// Compact pairs (8 bytes instead of 4 + 8 between the pack kernel and the second bucket pass).  The pack kernel writes
// pair i at index i, so the query id is implicit there; pass A turns the pack word into a word that carries the id but
// no longer the digit it has just bucketed by (implied by the bucket the word lies in); pass B, whose tiles never
// straddle two buckets, restores the full key and writes the (key, payload | id) pairs the first sweep pass reads.
//   pack word:   key in bits [0, keyBits), payload in [keyBits, keyBits + restBits), bit 63 = irregular query
//   pass A word: key bits below digit A in [0, lowBits), payload in [lowBits, lowBits + restBits), id above them
//                (idBits wide; all ones = irregular query)
// Usable when keyBits + restBits <= 63 and lowBits + restBits + idBits <= 64 (100 M 20-mers on a k = 12 table:
// 16 + 16 + 27 = 59).  A pass then moves one 64-bit word per pair through registers, shared memory and the LSU instead
// of a 32-bit and a 64-bit one: the passes are bound by exactly that pipe (profiles/r02_ncu_sweep_kernels.json).
struct SortCompact {
  uint32_t keyBits, lowBits, restBits, idBits;
};
constexpr uint32_t kSortNoId = 0xFFFFFFFFu;  // (= kSweepNoId)

// histogram of digit B inside the bucket regions (keys only: 4 B per pair; COMPACT: the low half of pass A's words)
template <bool COMPACT, int ITEMS>  // ITEMS * kSortThreads = pairs per tile of pass B
__global__ void __launch_bounds__(kSortThreads)
    sortDigitCounts(const void *__restrict__ keysOrWords, SortCtrl *__restrict__ ctrl, uint32_t *__restrict__ countB,
                    uint32_t dA, uint32_t dB, uint32_t shiftB) {
  __shared__ uint32_t bins[kSortBins];
  __shared__ uint32_t tileShared;
  const uint32_t numBuckets = 1u << dA, maskB = (1u << dB) - 1u;
  const uint32_t totalTiles = ctrl->tilesBefore[numBuckets];
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) tileShared = atomicAdd(&ctrl->ticket[1], 1u);
    bins[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t tile = tileShared;
    if (tile >= totalTiles) break;
    uint32_t a, first, count;
    sortTileOfBucket(ctrl, numBuckets, tile, ITEMS * kSortThreads, a, first, count);
    uint32_t key[ITEMS];  // every load is issued before the first shared atomic waits on one
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      const uint32_t i = it * kSortThreads + threadIdx.x;
      const uint32_t src = first + min(i, count - 1u);
      if (COMPACT) key[it] = __ldg(reinterpret_cast<const uint2 *>(keysOrWords) + src).x;
      else key[it] = __ldg(reinterpret_cast<const uint32_t *>(keysOrWords) + src);
    }
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
      const uint32_t i = it * kSortThreads + threadIdx.x;
      if (i < count) atomicAdd(&bins[(key[it] >> shiftB) & maskB], 1u);
    }
    __syncthreads();
    const uint32_t c = bins[threadIdx.x];
    if (c) atomicAdd(countB + (a << dB) + threadIdx.x, c);
  }
}

// One bucket pass.  SECOND = false: tiles in input order, digit A, cursors ctrl->cursorA.  SECOND = true: tiles inside
// the bucket regions of pass A's output, digit B, cursors cursorB[a << dB | digit].
template <bool SECOND>
__global__ void __launch_bounds__(kSortThreads, AWFM_SORT_MIN_CTAS)
    sortPass(const uint32_t *__restrict__ keysIn, const uint64_t *__restrict__ valsIn, uint32_t *__restrict__ keysOut,
             uint64_t *__restrict__ valsOut, uint32_t numPairs, SortCtrl *__restrict__ ctrl, uint32_t *__restrict__ cursorB,
             uint32_t dA, uint32_t dB, uint32_t shift) {
  extern __shared__ __align__(16) uint8_t sortSmem[];
  uint64_t *sVal = reinterpret_cast<uint64_t *>(sortSmem);                        // kSortTile
  uint32_t *sKey = reinterpret_cast<uint32_t *>(sortSmem + 8 * (size_t)kSortTile);  // kSortTile
  __shared__ uint32_t bins[kSortBins], localBase[kSortBins], globalBase[kSortBins];
  __shared__ uint32_t warpSums[kSortThreads / 32];
  __shared__ uint32_t tileShared;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t digitBits = SECOND ? dB : dA, mask = (1u << digitBits) - 1u;
  const uint32_t numBuckets = 1u << dA;
  const uint32_t totalTiles = SECOND ? ctrl->tilesBefore[numBuckets] : (numPairs + kSortTile - 1) / kSortTile;
  for (;;) {
    __syncthreads();  // previous tile's staging consumed
    if (threadIdx.x == 0) tileShared = atomicAdd(&ctrl->ticket[SECOND ? 2 : 0], 1u);
    bins[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t tile = tileShared;
    if (tile >= totalTiles) break;
    uint32_t a = 0, first, count;
    if (SECOND) sortTileOfBucket(ctrl, numBuckets, tile, kSortTile, a, first, count);
    else first = tile * kSortTile, count = min((uint32_t)kSortTile, numPairs - first);
    uint32_t key[kSortItems], rank[kSortItems];
    uint64_t val[kSortItems];
#pragma unroll
    for (int it = 0; it < kSortItems; it++) {
      const uint32_t i = it * kSortThreads + threadIdx.x;
      const uint32_t src = first + min(i, count - 1u);
      key[it] = __ldg(keysIn + src);
      val[it] = __ldg(valsIn + src);
    }
#pragma unroll
    for (int it = 0; it < kSortItems; it++) {
      const uint32_t i = it * kSortThreads + threadIdx.x;
      if (i < count) rank[it] = atomicAdd(&bins[(key[it] >> shift) & mask], 1u);
    }
    __syncthreads();
    {  // exclusive scan of the tile's bin counts; room behind the global cursors
      const uint32_t c = bins[threadIdx.x];
      uint32_t incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (unsigned)d) incl += up;
      }
      if (lane == 31u) warpSums[warp] = incl;
      __syncthreads();
      uint32_t before = 0;
#pragma unroll
      for (unsigned w = 0; w < kSortThreads / 32; w++) before += w < warp ? warpSums[w] : 0u;
      localBase[threadIdx.x] = before + incl - c;
      uint32_t *cursor = SECOND ? cursorB + (a << dB) + threadIdx.x : ctrl->cursorA + threadIdx.x;
      globalBase[threadIdx.x] = c ? atomicAdd(cursor, c) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kSortItems; it++) {
      const uint32_t i = it * kSortThreads + threadIdx.x;
      if (i < count) {
        const uint32_t dst = localBase[(key[it] >> shift) & mask] + rank[it];
        sKey[dst] = key[it];
        sVal[dst] = val[it];
      }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kSortItems; it++) {
      const uint32_t i = it * kSortThreads + threadIdx.x;
      if (i < count) {
        const uint32_t k = sKey[i];
        const uint32_t d = (k >> shift) & mask;
        const uint32_t dst = globalBase[d] + (i - localBase[d]);
        keysOut[dst] = k;
        valsOut[dst] = sVal[i];
      }
    }
  }
}

// The same bucket pass on compact pairs.  SECOND = false: pack words in (id = index), pass A words out; digit A leaves
// the word before it is staged, so the tile keeps one byte per staged pair to know the bucket it is written to.
// SECOND = true: pass A words in, (key, payload | id) pairs out for the first sweep pass.  `shift` = position of the
// digit inside the low 32 bits of the input word (the key occupies the word's low bits in both formats).
#ifndef AWFM_SORT_COMPACT_ITEMS
#define AWFM_SORT_COMPACT_ITEMS 16  // (a word per pair leaves registers for more pairs per thread: 1.15 -> 1.01 ms per 100 M)
#endif
#ifndef AWFM_SORT_COMPACT_MIN_CTAS
#define AWFM_SORT_COMPACT_MIN_CTAS 4
#endif
constexpr int kSortCompactItems = AWFM_SORT_COMPACT_ITEMS;
constexpr int kSortCompactTile = kSortThreads * kSortCompactItems;
template <bool SECOND>
__global__ void __launch_bounds__(kSortThreads, AWFM_SORT_COMPACT_MIN_CTAS)
    sortPassCompact(const uint64_t *__restrict__ wordsIn, uint64_t *__restrict__ wordsOut, uint32_t *__restrict__ keysOut,
                    uint64_t *__restrict__ valsOut, uint32_t numPairs, SortCtrl *__restrict__ ctrl,
                    uint32_t *__restrict__ cursorB, uint32_t dA, uint32_t dB, uint32_t shift, const SortCompact f) {
  extern __shared__ __align__(16) uint8_t sortSmem[];
  uint64_t *sWord = reinterpret_cast<uint64_t *>(sortSmem);       // kSortCompactTile
  uint8_t *sDigit = sortSmem + 8 * (size_t)kSortCompactTile;             // kSortCompactTile (pass A only)
  __shared__ uint32_t bins[kSortBins], localBase[kSortBins], globalDelta[kSortBins];
  __shared__ uint32_t warpSums[kSortThreads / 32];
  __shared__ uint32_t tileShared;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t digitBits = SECOND ? dB : dA, mask = (1u << digitBits) - 1u;
  const uint32_t numBuckets = 1u << dA;
  const uint32_t totalTiles = SECOND ? ctrl->tilesBefore[numBuckets] : (numPairs + kSortCompactTile - 1) / kSortCompactTile;
  const uint64_t lowMask = (1ull << f.lowBits) - 1ull, restMask = (1ull << f.restBits) - 1ull;
  const uint32_t idShift = f.lowBits + f.restBits;  // <= 63
  const uint64_t idMask = (1ull << f.idBits) - 1ull;  // idBits <= 32
  for (;;) {
    __syncthreads();  // previous tile's staging consumed
    if (threadIdx.x == 0) tileShared = atomicAdd(&ctrl->ticket[SECOND ? 2 : 0], 1u);
    bins[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t tile = tileShared;
    if (tile >= totalTiles) break;
    uint32_t a = 0, first, count;
    if (SECOND) sortTileOfBucket(ctrl, numBuckets, tile, kSortCompactTile, a, first, count);
    else first = tile * kSortCompactTile, count = min((uint32_t)kSortCompactTile, numPairs - first);
    uint64_t word[kSortCompactItems];
    uint32_t rank[kSortCompactItems];
#pragma unroll
    for (int it = 0; it < kSortCompactItems; it++) {
      const uint32_t i = it * kSortThreads + threadIdx.x;
      word[it] = __ldg(wordsIn + first + min(i, count - 1u));
    }
#pragma unroll
    for (int it = 0; it < kSortCompactItems; it++) {
      const uint32_t i = it * kSortThreads + threadIdx.x;
      if (i < count) rank[it] = atomicAdd(&bins[((uint32_t)word[it] >> shift) & mask], 1u);
    }
    __syncthreads();
    {  // exclusive scan of the tile's bin counts; room behind the global cursors
      const uint32_t c = bins[threadIdx.x];
      uint32_t incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (unsigned)d) incl += up;
      }
      if (lane == 31u) warpSums[warp] = incl;
      __syncthreads();
      uint32_t before = 0;
#pragma unroll
      for (unsigned w = 0; w < kSortThreads / 32; w++) before += w < warp ? warpSums[w] : 0u;
      const uint32_t local = before + incl - c;
      localBase[threadIdx.x] = local;
      uint32_t *cursor = SECOND ? cursorB + (a << dB) + threadIdx.x : ctrl->cursorA + threadIdx.x;
      globalDelta[threadIdx.x] = (c ? atomicAdd(cursor, c) : 0u) - local;  // staged pair i of bucket d goes to globalDelta[d] + i
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kSortCompactItems; it++) {
      const uint32_t i = it * kSortThreads + threadIdx.x;
      if (i < count) {
        uint64_t w = word[it];
        const uint32_t d = ((uint32_t)w >> shift) & mask;
        const uint32_t dst = localBase[d] + rank[it];
        if (!SECOND) {  // pack word -> pass A word: digit A leaves, the id becomes explicit
          const uint64_t id = (w >> 63) ? idMask : (uint64_t)(first + i);
          w = (w & lowMask) | (((w >> f.keyBits) & restMask) << f.lowBits) | (id << idShift);
          sDigit[dst] = (uint8_t)d;
        }
        sWord[dst] = w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kSortCompactItems; it++) {
      const uint32_t i = it * kSortThreads + threadIdx.x;
      if (i < count) {
        const uint64_t w = sWord[i];
        if (!SECOND) {
          wordsOut[globalDelta[sDigit[i]] + i] = w;
        } else {
          const uint32_t dst = globalDelta[((uint32_t)w >> shift) & mask] + i;
          const uint32_t id = (uint32_t)((w >> idShift) & idMask);
          keysOut[dst] = (a << f.lowBits) | (uint32_t)(w & lowMask);
          valsOut[dst] = (((w >> f.lowBits) & restMask) << 32) | (id == (uint32_t)idMask ? kSortNoId : id);
        }
      }
    }
  }
}

constexpr size_t kSortSmemBytes = 12 * (size_t)kSortTile;
constexpr size_t kSortCompactSmemBytes = 9 * (size_t)kSortCompactTile;

}  // namespace awfm
