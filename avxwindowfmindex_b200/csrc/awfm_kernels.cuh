// awfm_kernels.cuh — the search kernels (sm_100a).  See awfm_device.cuh for the HBM layout.
#pragma once
#include "awfm_device.cuh"

namespace awfm {

struct QueryBatch {
  const uint8_t *letters;
  const uint64_t *offsets;  // numQueries+1 or nullptr (fixed length)
  uint64_t numQueries;
  uint32_t fixedLen;
  uint32_t rangesOfHitsOnly;  // != 0: `ranges` is only written for queries with a non-empty final range (locate)
};

__device__ __forceinline__ void queryExtent(const QueryBatch &qb, uint64_t q, uint64_t &off, uint64_t &len) {
  if (qb.offsets) {
    off = __ldg(qb.offsets + q);
    len = __ldg(qb.offsets + q + 1) - off;
  } else {
    off = q * (uint64_t)qb.fixedLen;
    len = qb.fixedLen;
  }
}

// Seed-table entry `index` of the original table (deep = false) or of the derived deeper one.
__device__ __forceinline__ void loadSeedEntry(const DevIndex &ix, bool deep, uint64_t index, uint64_t &sp, uint64_t &ep) {
  if (deep && !ix.deepSeedWide) {
    const uint2 r = __ldg(reinterpret_cast<const uint2 *>(ix.deepSeedTable) + index);
    sp = r.x;
    ep = r.y;
  } else {
    const uint4 r = __ldg((deep ? reinterpret_cast<const uint4 *>(ix.deepSeedTable) : ix.seedTable) + index);
    sp = (uint64_t)r.x | ((uint64_t)r.y << 32);
    ep = (uint64_t)r.z | ((uint64_t)r.w << 32);
  }
}

// Table index of the last `k` letters (leftmost most significant, src/AwFmKmerTable.c:21-51); false when one of them is
// not a searchable letter (src/AwFmKmerTable.c:4-19).  LETTERS are letter indices when TRANSLATED, else ASCII.
template <bool AMINO, bool TRANSLATED>
__device__ __forceinline__ bool seedIndexOf(const uint8_t *__restrict__ s, uint64_t len, uint32_t k, uint64_t &index) {
  constexpr uint32_t CARD = AMINO ? 20u : 4u;
  bool seedable = true;
  uint64_t t = 0;
  for (uint32_t i = 0; i < k; i++) {
    const uint32_t l = TRANSLATED ? (uint32_t)s[(len - k) + i] : letterIndex<AMINO>(__ldg(s + (len - k) + i));
    seedable &= (l < CARD);
    t = t * CARD + l;
  }
  index = t;
  return seedable;
}

// Opens the range for one query (src/AwFmParallelSearch.c:222-271): seed-table entry when the last k letters are
// all searchable letters (src/AwFmKmerTable.c:4-51), otherwise [C[c], C[c+1]-1] of the last letter
// (src/AwFmSearch.c:485-501).  Returns the number of leading letters still to be stepped through; sets an
// empty range (1,0) for inputs the reference leaves undefined (len == 0, '$' inside a query).
// With a derived deep table (awfm_gpu_ctx_extend_seed_table) a query whose last deepSeedK letters are all searchable
// starts from the range the reference would hold after stepping through those letters — same result, fewer steps.
template <bool AMINO, bool TRANSLATED = false>
__device__ __forceinline__ uint64_t openRange(const DevIndex &ix, const uint8_t *__restrict__ s, uint64_t len,
                                              uint64_t &sp, uint64_t &ep) {
  constexpr uint32_t CARD = AMINO ? 20u : 4u;
  const uint32_t k = ix.seedK;
  sp = 1;
  ep = 0;
  if (len == 0) return 0;
  uint64_t tableIndex;
  if (ix.deepSeedK && len >= ix.deepSeedK && seedIndexOf<AMINO, TRANSLATED>(s, len, ix.deepSeedK, tableIndex)) {
    loadSeedEntry(ix, true, tableIndex, sp, ep);
    return len - ix.deepSeedK;
  }
  if (len >= k && seedIndexOf<AMINO, TRANSLATED>(s, len, k, tableIndex)) {
    loadSeedEntry(ix, false, tableIndex, sp, ep);
    return len - k;
  }
  const uint32_t last = TRANSLATED ? (uint32_t)s[len - 1] : letterIndex<AMINO>(__ldg(s + len - 1));
  if (last > CARD) return 0;
  sp = ix.prefixSums[last];
  ep = ix.prefixSums[last + 1] - 1;
  return len - 1;
}

// ---------------------------------------------------------------------------------------------------------------
// count kernel, variant 0: one LPQ-lane group per query, grid-stride over queries.
// src/AwFmParallelSearch.c:159-220 (seed -> extend -> range length) for every query of the batch.
// ---------------------------------------------------------------------------------------------------------------
template <int LPQ, bool AMINO>
__global__ void __launch_bounds__(256)
    countKernelV0(const __grid_constant__ DevIndex ix, const __grid_constant__ QueryBatch qb,
                  uint32_t *__restrict__ counts, uint4 *__restrict__ ranges) {
  constexpr uint32_t CARD = AMINO ? 20u : 4u;
  const unsigned sub = threadIdx.x % LPQ;
  const unsigned mask = groupMaskOf<LPQ>();
  const uint64_t numGroups = (uint64_t)gridDim.x * blockDim.x / LPQ;
  for (uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPQ; q < qb.numQueries; q += numGroups) {
    uint64_t off, len, sp, ep;
    queryExtent(qb, q, off, len);
    const uint8_t *s = qb.letters + off;
    uint64_t next = openRange<AMINO>(ix, s, len, sp, ep);
    while (next > 0 && sp <= ep) {  // src/AwFmParallelSearch.c:279-311
      const uint32_t letter = letterIndex<AMINO>(__ldg(s + next - 1));
      if (letter > CARD) {  // '$' in a query: undefined in the reference, defined here as no match
        sp = 1;
        ep = 0;
        break;
      }
      lfStep<LPQ, AMINO>(ix, sp, ep, letter, sub, mask);
      next--;
    }
    if (sub == 0) {
      counts[q] = (uint32_t)(sp <= ep ? ep - sp + 1 : 0);  // src/AwFmIndexStruct.c:126-130, u32 store :187-190
      if (ranges && (sp <= ep || !qb.rangesOfHitsOnly))
        ranges[q] = make_uint4((uint32_t)sp, (uint32_t)(sp >> 32), (uint32_t)ep, (uint32_t)(ep >> 32));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// count kernel, variant 1: CTA tiles with shared-memory staging and refill.
//   phase A (thread per query): the tile's query letters are copied into shared memory with coalesced 128-bit
//           loads, translated to letter indices in place, the seed-table entry of every query is requested
//           (one independent 16-B gather per thread) and the open range parked in shared memory;
//   phase B (LPQ-lane group per query): groups pull queries from the tile through a shared counter and run the
//           LF loop; a group that finishes early immediately takes the next query, so lanes do not idle on the
//           step-count spread (1..len-k steps, range usually empties early on random queries).
// ---------------------------------------------------------------------------------------------------------------
template <int TILE>
struct TileSmem {
  uint64_t sp[TILE], ep[TILE];
  uint64_t next[TILE];   // letters still to step through
  uint32_t start[TILE];  // offset of the query's first letter inside `letters`
  uint32_t counter;
  uint32_t lettersBase;  // 16-B aligned-down global byte offset of letters[0] (low bits)
};

template <int LPQ, bool AMINO, int TILE, int LETTER_BYTES>
__global__ void __launch_bounds__(256)
    countKernelV1(const __grid_constant__ DevIndex ix, const __grid_constant__ QueryBatch qb,
                  uint32_t *__restrict__ counts, uint4 *__restrict__ ranges) {
  constexpr uint32_t CARD = AMINO ? 20u : 4u;
  __shared__ TileSmem<TILE> sm;
  __shared__ __align__(16) uint8_t letters[LETTER_BYTES];
  const unsigned sub = threadIdx.x % LPQ;
  const unsigned mask = groupMaskOf<LPQ>();
  const uint64_t numTiles = (qb.numQueries + TILE - 1) / TILE;

  for (uint64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
    const uint64_t q0 = tile * TILE;
    const uint32_t nq = (uint32_t)min((uint64_t)TILE, qb.numQueries - q0);
    uint64_t byte0, byte1, dummy;
    queryExtent(qb, q0, byte0, dummy);
    {
      uint64_t o, l;
      queryExtent(qb, q0 + nq - 1, o, l);
      byte1 = o + l;
    }
    const uint64_t aligned0 = byte0 & ~15ull;
    const uint64_t span = byte1 - aligned0;
    const bool staged = span <= LETTER_BYTES;  // uniform per CTA
    __syncthreads();                           // previous tile fully consumed
    if (staged) {
      const uint4 *src = reinterpret_cast<const uint4 *>(qb.letters + aligned0);
      const uint32_t n16 = (uint32_t)((span + 15) / 16);
      for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) {
        // translate ASCII -> letter index while staging (4 bytes at a time)
        uint32_t w[4];
        const uint64_t g = aligned0 + 16ull * i;
        if (g + 16 <= byte1) {
          const uint4 v = __ldg(src + i);
          w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
        } else {  // last chunk of the tile: never read past the batch's final letter
          w[0] = w[1] = w[2] = w[3] = 0;
          for (uint32_t b = 0; g + b < byte1; b++) w[b >> 2] |= (uint32_t)__ldg(qb.letters + g + b) << (8 * (b & 3));
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          uint32_t o = 0;
#pragma unroll
          for (int b = 0; b < 4; b++) o |= letterIndex<AMINO>((w[j] >> (8 * b)) & 0xFFu) << (8 * b);
          w[j] = o;
        }
        reinterpret_cast<uint4 *>(letters)[i] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    if (threadIdx.x == 0) sm.counter = 0;
    __syncthreads();

    // ---- phase A: open ranges, thread per query ----
    for (uint32_t t = threadIdx.x; t < nq; t += blockDim.x) {
      uint64_t off, len, sp, ep, next;
      queryExtent(qb, q0 + t, off, len);
      if (staged) next = openRange<AMINO, true>(ix, letters + (off - aligned0), len, sp, ep);
      else next = openRange<AMINO, false>(ix, qb.letters + off, len, sp, ep);
      sm.sp[t] = sp;
      sm.ep[t] = ep;
      sm.start[t] = (uint32_t)(off - aligned0);
      sm.next[t] = next;
    }
    __syncthreads();

    // ---- phase B: LF loop, group per query with refill ----
    for (;;) {
      uint32_t t = 0;
      if (sub == 0) t = atomicAdd(&sm.counter, 1u);
      t = __shfl_sync(mask, t, (threadIdx.x & 31u) / LPQ * LPQ);
      if (t >= nq) break;
      uint64_t sp = sm.sp[t], ep = sm.ep[t];
      uint64_t next = sm.next[t];
      const uint32_t start = sm.start[t];
      if (!staged) {  // tile too long for the staging window: letters straight from global memory
        uint64_t off, len;
        queryExtent(qb, q0 + t, off, len);
        const uint8_t *s = qb.letters + off;
        while (next > 0 && sp <= ep) {
          const uint32_t letter = letterIndex<AMINO>(__ldg(s + next - 1));
          if (letter > CARD) {
            sp = 1;
            ep = 0;
            break;
          }
          lfStep<LPQ, AMINO>(ix, sp, ep, letter, sub, mask);
          next--;
        }
      } else {
        const uint8_t *s = letters + start;
        while (next > 0 && sp <= ep) {
          const uint32_t letter = s[next - 1];
          if (letter > CARD) {
            sp = 1;
            ep = 0;
            break;
          }
          lfStep<LPQ, AMINO>(ix, sp, ep, letter, sub, mask);
          next--;
        }
      }
      if (sub == 0) {
        counts[q0 + t] = (uint32_t)(sp <= ep ? ep - sp + 1 : 0);
        if (ranges && (sp <= ep || !qb.rangesOfHitsOnly))
          ranges[q0 + t] = make_uint4((uint32_t)sp, (uint32_t)(sp >> 32), (uint32_t)ep, (uint32_t)(ep >> 32));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// locate kernel: one LPQ-lane group per hit, grid-stride with immediate refill.
// src/AwFmParallelSearch.c:315-365: p = sp + i; while p not sampled: p = LF(p), offset++;
// position = (SA[p / ratio] + offset) mod bwtLength (src/AwFmSuffixArray.c:179-203).
// Hit h of the flat CSR list belongs to the query q with hitOffsets[q] <= h < hitOffsets[q+1].
// ---------------------------------------------------------------------------------------------------------------
// Step 1, expandHits: positions[h - hitBegin] = sp(q) + (h - hitOffsets[q]) for every flat hit index h of the
// window, written by one warp per 32 queries (lanes stride over a query's hits, so a query with millions of hits
// is as cheap per hit as one with a single hit).
static __global__ void __launch_bounds__(256)
    expandHits(const uint4 *__restrict__ ranges, const uint64_t *__restrict__ hitOffsets, uint64_t numQueries,
               uint64_t hitBegin, uint64_t hitEnd, uint64_t *__restrict__ positions) {
  const unsigned lane = threadIdx.x & 31u;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t numWarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t q0 = warp * 32; q0 < numQueries; q0 += numWarps * 32) {
    const uint64_t q = q0 + lane;
    uint64_t a = 0, b = 0, sp = 0;
    if (q < numQueries) {
      a = __ldg(hitOffsets + q);
      b = __ldg(hitOffsets + q + 1);
      if (b > a) {  // (with ranges of hits only, the entry of a query without hits was never written)
        const uint4 r = __ldg(ranges + q);
        sp = (uint64_t)r.x | ((uint64_t)r.y << 32);
      }
    }
    const bool live = b > a && b > hitBegin && a < hitEnd;
    const uint64_t lo = max(a, hitBegin), hi = min(b, hitEnd);
    // short ranges (the common case on sparse hits): the lane writes its own few positions, neighbours in a warp
    // write neighbouring slots; long ranges are spread over the whole warp below
    const bool small = live && hi - lo <= 8;
    if (small)
      for (uint64_t h = lo; h < hi; h++) positions[h - hitBegin] = sp + (h - a);
    unsigned todo = __ballot_sync(0xFFFFFFFFu, live && !small);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const uint64_t qa = __shfl_sync(0xFFFFFFFFu, a, src);
      const uint64_t qlo = __shfl_sync(0xFFFFFFFFu, lo, src), qhi = __shfl_sync(0xFFFFFFFFu, hi, src);
      const uint64_t qsp = __shfl_sync(0xFFFFFFFFu, sp, src);
      for (uint64_t h = qlo + lane; h < qhi; h += 32) positions[h - hitBegin] = qsp + (h - qa);
    }
  }
}

// (value + offset) mod bwtLength (src/AwFmSuffixArray.c:179-203).  A walk passes BWT position 0 (always sampled)
// within bwtLength steps, so the sum is below 2*bwtLength on any well-formed index: one compare-and-subtract; the
// 64-bit division only runs on malformed input.
__device__ __forceinline__ uint64_t wrapPosition(uint64_t v, uint64_t n) {
  if (v < n) return v;
  v -= n;
  return v < n ? v : v % n;
}

// Step 2, locateKernel: in place, positions[i] holds a BWT position on entry and the text position on exit.
template <int LPQ, bool AMINO>
__global__ void __launch_bounds__(256)
    locateKernel(const __grid_constant__ DevIndex ix, uint64_t numHits, uint64_t *__restrict__ positions) {
  const unsigned sub = threadIdx.x % LPQ;
  const unsigned mask = groupMaskOf<LPQ>();
  const uint64_t numGroups = (uint64_t)gridDim.x * blockDim.x / LPQ;
  for (uint64_t h = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPQ; h < numHits; h += numGroups) {
    uint64_t p = positions[h];
    if (LPQ > 1) __syncwarp(mask);  // every lane of the group has read the slot before lane 0 overwrites it
    uint64_t offset = 0;
    while (!isSampled(ix, p)) {
      p = backtraceStep<LPQ, AMINO>(ix, p, sub, mask);
      offset++;
    }
    if (sub == 0) positions[h] = wrapPosition(saValue(ix, sampleIndexOf(ix, p)) + offset, ix.bwtLength);
  }
}

// Step 2, default variant: one LPQ-lane GROUP per hit with immediate refill.  The walk length of a hit is geometric
// (mean ratio-1, unbounded: sampling is by BWT position, src/AwFmIndexStruct.c:88-91), so a group bound to one hit
// per warp round idles for most of the round (the slowest of 32 geometric walks is ~4x the mean).  Here a group
// that reaches a sampled position finishes its hit (sampled-SA read, add, mod, store) and takes the next hit in the
// same round: every group keeps one independent DRAM request in flight.  Hits are handed out in 64-hit chunks from
// a global counter (one atomic per chunk per warp), so the tail of the launch is one chunk, not the slowest thread.
// All lanes of a group carry the same (h, p, offset); a nucleotide walk is owned by a single thread (LPQ = 1).
constexpr uint32_t kLocateChunk = 64;
template <int LPQ, bool AMINO>
__global__ void __launch_bounds__(256, 8)
    locateKernelRefill(const __grid_constant__ DevIndex ix, uint64_t numHits, uint64_t *__restrict__ positions,
                       unsigned long long *__restrict__ workCounter) {
  constexpr uint64_t kIdle = ~0ull;
  constexpr unsigned kLeaders = LPQ == 1 ? 0xFFFFFFFFu : LPQ == 2 ? 0x55555555u : LPQ == 4 ? 0x11111111u : 0x01010101u;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned sub = lane % LPQ;
  const unsigned mask = groupMaskOf<LPQ>();
  const unsigned leadersBelow = kLeaders & ((1u << (lane - sub)) - 1u);  // group leaders of lower-numbered groups
  uint64_t chunkNext = 0, chunkEnd = 0;  // warp-uniform: hits of the current chunk not handed out yet
  bool exhausted = false;                // warp-uniform: the global counter ran past numHits
  uint64_t h = kIdle, p = 0;
  uint32_t offset = 0;
  for (;;) {
    unsigned needMask = __ballot_sync(0xFFFFFFFFu, h == kIdle) & kLeaders;
    while (needMask) {
      if (chunkNext >= chunkEnd) {
        if (exhausted) break;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(workCounter, (unsigned long long)kLocateChunk);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= numHits) {
          exhausted = true;
          break;
        }
        chunkNext = base;
        chunkEnd = min((uint64_t)base + kLocateChunk, numHits);
      }
      const uint32_t avail = (uint32_t)(chunkEnd - chunkNext);
      const uint32_t rank = __popc(needMask & leadersBelow);
      if (h == kIdle && rank < avail) {
        h = chunkNext + rank;
        p = positions[h];
        offset = 0;
      }
      chunkNext += min(avail, (uint32_t)__popc(needMask));
      needMask = __ballot_sync(0xFFFFFFFFu, h == kIdle) & kLeaders;
    }
    if (needMask == kLeaders) break;  // no group holds a hit and none is left to hand out
    if (h != kIdle) {                 // uniform within a group
      if (isSampled(ix, p)) {
        const uint64_t v = wrapPosition(saValue(ix, sampleIndexOf(ix, p)) + offset, ix.bwtLength);
        if (LPQ > 1) __syncwarp(mask);  // every lane of the group has read positions[h] before it is overwritten
        if (sub == 0) positions[h] = v;
        h = kIdle;
      } else {
        p = backtraceStep<LPQ, AMINO>(ix, p, sub, mask);
        offset++;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Derived structures that trade HBM for fewer dependent DRAM round trips (built once per context, on request).
//
// extendSeedTable: level j -> level j+1 of the seed table.  Entry x of level j holds the range the reference has
// after the last j letters x of a query (src/AwFmParallelSearch.c:222-311: table entry for the last k letters, then
// one LF step per further letter while the range is valid — an invalid range is kept as it is).  Prepending letter
// c gives entry c*CARD^j + x of level j+1 = valid(entry) ? step(c, entry) : entry.  Exactly the reference's values,
// including the stored invalid pairs.
// ---------------------------------------------------------------------------------------------------------------
template <bool AMINO, bool SRC_WIDE, bool DST_WIDE>
__global__ void __launch_bounds__(256)
    extendSeedTable(const __grid_constant__ DevIndex ix, const void *__restrict__ src, uint64_t numSrc,
                    void *__restrict__ dst) {
  constexpr uint64_t CARD = AMINO ? 20 : 4;
  const uint64_t total = numSrc * CARD;
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t x = o % numSrc;
    const uint32_t c = (uint32_t)(o / numSrc);
    uint64_t sp, ep;
    if (SRC_WIDE) {
      const uint4 r = __ldg(reinterpret_cast<const uint4 *>(src) + x);
      sp = (uint64_t)r.x | ((uint64_t)r.y << 32);
      ep = (uint64_t)r.z | ((uint64_t)r.w << 32);
    } else {
      const uint2 r = __ldg(reinterpret_cast<const uint2 *>(src) + x);
      sp = r.x;
      ep = r.y;
    }
    if (sp <= ep) lfStep<1, AMINO>(ix, sp, ep, c, 0u, 0u);
    if (DST_WIDE)
      reinterpret_cast<uint4 *>(dst)[o] = make_uint4((uint32_t)sp, (uint32_t)(sp >> 32), (uint32_t)ep, (uint32_t)(ep >> 32));
    else
      reinterpret_cast<uint2 *>(dst)[o] = make_uint2((uint32_t)sp, (uint32_t)ep);
  }
}

// densify the sampled SA: out[j] = SA[j * newRatio] for j in [first, first + count), by the locate walk itself
// (src/AwFmParallelSearch.c:333-361).  `work` holds j*newRatio on entry (iota kernel) and text positions on exit.
static __global__ void saIota(uint64_t *__restrict__ work, uint64_t first, uint64_t count, uint64_t newRatio) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    work[i] = (first + i) * newRatio;
}
template <typename T>
__global__ void saNarrow(const uint64_t *__restrict__ work, uint64_t count, T *__restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = (T)work[i];
}

// ---------------------------------------------------------------------------------------------------------------
// contig mapping: global text position -> (sequence index, offset inside that sequence).
// awFmGetLocalSequencePositionFromIndexPosition (src/AwFmSearch.c:284-301) = fastaVectorGetLocalSequencePosition-
// FromGlobal (lib/FastaVector/src/FastaVector.c:338-381): E[s] = cumulative end of record s INCLUDING its one-byte
// separator; sequence = number of records with E[s] <= g (so g == E[last] yields numSequences, offset 0, exactly as
// the reference's binary search does); offset = g - E[sequence-1]; g > E[last] is the reference's
// AwFmIllegalPositionError and is reported as (UINT64_MAX, UINT64_MAX).
// The first levels of the search run on a 256-entry sample of E staged in shared memory; the record table itself
// (16 B x 10 k contigs for BASELINE cfg 5) stays L1/L2-resident.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMapSample = 256;
static __global__ void __launch_bounds__(256)
    mapPositionsKernel(const uint64_t *__restrict__ ends, uint64_t numSequences, const uint64_t *__restrict__ positions,
                       uint64_t n, uint64_t *__restrict__ sequenceIndex, uint64_t *__restrict__ localPosition) {
  __shared__ uint64_t sample[kMapSample];  // sample[j] = E[min((j+1)*stride, numSequences) - 1]
  const uint64_t stride = (numSequences + kMapSample - 1) / kMapSample;
  for (uint32_t j = threadIdx.x; j < kMapSample; j += blockDim.x) {
    const uint64_t last = min((uint64_t)(j + 1) * stride, numSequences);
    sample[j] = last ? __ldg(ends + last - 1) : 0;
  }
  __syncthreads();
  const uint64_t limit = numSequences ? __ldg(ends + numSequences - 1) : 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t g = positions[i];
    if (numSequences == 0 || g > limit) {
      sequenceIndex[i] = ~0ull;
      localPosition[i] = ~0ull;
      continue;
    }
    // bucket = number of sample entries <= g; records before bucket*stride all end at or before g
    uint32_t lo = 0, hi = kMapSample;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (g < sample[mid]) hi = mid;
      else lo = mid + 1;
    }
    uint64_t a = min((uint64_t)lo * stride, numSequences), b = min(a + stride, numSequences);
    while (a < b) {  // invariant: every record before a has E <= g, every record from b on has E > g
      const uint64_t mid = a + ((b - a) >> 1);
      if (g < __ldg(ends + mid)) b = mid;
      else a = mid + 1;
    }
    sequenceIndex[i] = a;
    localPosition[i] = a ? g - __ldg(ends + a - 1) : g;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// upload-time relayout: reference blocks -> lines (one thread per block)
// ---------------------------------------------------------------------------------------------------------------
// Nucleotide: first the superblock rows (one thread per 2^16-position superblock = 256 reference blocks), then one
// thread per reference block (256 positions) -> four 32-B sectors with 16-bit counts relative to the superblock row.
static __global__ void sectorSuperRows(const uint8_t *__restrict__ raw, uint64_t numBlocks, uint64_t firstBlock,
                                const uint64_t *__restrict__ prefixSums /* device copy, 6 entries */,
                                uint64_t *__restrict__ superCounts, uint64_t *__restrict__ superC) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // superblock inside this slab
  const uint64_t i = j * 256;                                          // its first block inside the slab
  if (i >= numBlocks) return;
  const uint64_t row = (firstBlock + i) >> 8;
  const uint64_t *base = reinterpret_cast<const uint64_t *>(raw + i * 160 + 96);
  for (int c = 0; c < 8; c++) {
    const uint64_t v = c < 5 ? base[c] : 0;
    superCounts[row * kSectorSuperStride + c] = v;
    superC[row * kSectorSuperStride + c] = c < 5 ? v + prefixSums[c] : 0;
  }
}
static __global__ void relayoutNucleotideSectors(const uint8_t *__restrict__ raw, uint64_t numBlocks, uint64_t firstBlock,
                                          const uint64_t *__restrict__ superCounts, uint4 *__restrict__ sectors,
                                          uint16_t *__restrict__ xRel16) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= numBlocks) return;
  const uint64_t b = firstBlock + i;
  const uint32_t *src = reinterpret_cast<const uint32_t *>(raw + i * 160);
  const uint64_t *base = reinterpret_cast<const uint64_t *>(raw + i * 160 + 96);
  const uint64_t *super = superCounts + (b >> 8) * kSectorSuperStride;
  uint32_t w[3][8];
#pragma unroll
  for (int v = 0; v < 3; v++)
#pragma unroll
    for (int j = 0; j < 8; j++) w[v][j] = src[8 * v + j];
  uint32_t rel[5];
#pragma unroll
  for (int c = 0; c < 5; c++) rel[c] = (uint32_t)(base[c] - super[c]);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    sectors[(4 * b + q) * kSectorU4] = make_uint4(w[0][2 * q], w[0][2 * q + 1], w[1][2 * q], w[1][2 * q + 1]);
    sectors[(4 * b + q) * kSectorU4 + 1] =
        make_uint4(w[2][2 * q], w[2][2 * q + 1], rel[0] | (rel[1] << 16), rel[2] | (rel[3] << 16));
    xRel16[4 * b + q] = (uint16_t)rel[4];
#pragma unroll
    for (int c = 0; c < 5; c++) {
      const uint32_t cc = nucCodeCare(c), code = cc & 0xFu, care = cc >> 4;
#pragma unroll
      for (int j = 2 * q; j < 2 * q + 2; j++) {
        uint32_t sel = 0xFFFFFFFFu;
#pragma unroll
        for (int v = 0; v < 3; v++)
          if ((care >> v) & 1u) sel &= ((code >> v) & 1u) ? w[v][j] : ~w[v][j];
        rel[c] += __popc(sel);
      }
    }
  }
}

// One thread per reference block (256 positions) -> four quarter-lines (see awfm_device.cuh).  The count of letter c
// at the start of quarter q = baseOccurrences[c] (block start) + popcount of c's selector over the 2q words before
// it, made relative to the enclosing superblock.
static __global__ void relayoutAmino(const uint8_t *__restrict__ raw, uint64_t numBlocks, uint64_t firstBlock,
                              const uint64_t *__restrict__ superCounts /* [numSuper][24] */,
                              uint4 *__restrict__ lines) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= numBlocks) return;
  const uint64_t b = firstBlock + i;
  const uint32_t *src = reinterpret_cast<const uint32_t *>(raw + i * 352);
  const uint64_t *base = reinterpret_cast<const uint64_t *>(raw + i * 352 + 160);
  const uint64_t *super = superCounts + ((b * 256) >> kSuperShift) * kAminoSuperStride;
  uint32_t w[5][8];
#pragma unroll
  for (int v = 0; v < 5; v++)
#pragma unroll
    for (int j = 0; j < 8; j++) w[v][j] = src[8 * v + j];
  uint32_t *dst = reinterpret_cast<uint32_t *>(lines + (4 * b) * kAminoLineU4);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    uint32_t *o = dst + 32 * q;
#pragma unroll
    for (int v = 0; v < 4; v++) {
      o[v] = w[v][2 * q];
      o[4 + v] = w[v][2 * q + 1];
    }
    o[8] = w[4][2 * q];
    o[9] = w[4][2 * q + 1];
    o[10] = 0;
  }
  for (int c = 0; c < 21; c++) {
    const uint32_t cc = kAminoCodeCare[c], code = cc & 0xFFu, care = cc >> 8;
    uint32_t rel = (uint32_t)(base[c] - super[c]);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      dst[32 * q + kAminoRelWord + c] = rel;
#pragma unroll
      for (int j = 2 * q; j < 2 * q + 2; j++) {
        uint32_t sel = 0xFFFFFFFFu;
#pragma unroll
        for (int v = 0; v < 5; v++)
          if ((care >> v) & 1u) sel &= ((code >> v) & 1u) ? w[v][j] : ~w[v][j];
        rel += __popc(sel);
      }
    }
  }
}

// Packed query formats of include/awfm_gpu.h -> the ASCII letters the search kernels take.  AWFM_QUERY_2BIT: codes 0..3
// = A,C,G,T; AWFM_QUERY_5BIT: codes 0..19 = the amino letters in the reference's index order (src/AwFmLetter.c:55-67),
// anything above = the ambiguity letter (written as 'X', which both the sanitizer and the ambiguity predicate treat as
// such, src/AwFmLetter.c:69-79,98-125).  Query i occupies bytes [i*B, (i+1)*B), B = ceil(len*bits/8), letter j in bits
// [j*bits, (j+1)*bits) of that little-endian byte string.  One thread per four output letters.
template <int BITS>
__global__ void __launch_bounds__(256)
    unpackQueries(const uint8_t *__restrict__ packed, uint64_t numQueries, uint32_t len, uint8_t *__restrict__ ascii) {
  const uint64_t totalLetters = numQueries * (uint64_t)len;
  const uint32_t B = (len * BITS + 7u) >> 3;
  const uint64_t totalBytes = numQueries * (uint64_t)B;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; 4 * w < totalLetters;
       w += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t out = 0;
#pragma unroll
    for (uint32_t t = 0; t < 4; t++) {
      const uint64_t i = 4 * w + t;
      if (i >= totalLetters) break;
      const uint64_t q = i / len;
      const uint32_t bit = (uint32_t)(i - q * len) * BITS;
      const uint64_t byte = q * B + (bit >> 3);
      uint32_t window = __ldg(packed + byte);
      if (byte + 1 < totalBytes) window |= (uint32_t)__ldg(packed + byte + 1) << 8;
      const uint32_t code = (window >> (bit & 7u)) & ((1u << BITS) - 1u);
      uint32_t ch;
      if (BITS == 2) ch = (0x54474341u >> (8u * code)) & 0xFFu;  // "ACGT"
      else ch = code < 20u ? (uint32_t)"ACDEFGHIKLMNPQRSTVWY"[code] : (uint32_t)'X';
      out |= ch << (8u * t);
    }
    if (4 * w + 4 <= totalLetters) reinterpret_cast<uint32_t *>(ascii)[w] = out;
    else
      for (uint32_t t = 0; 4 * w + t < totalLetters; t++) ascii[4 * w + t] = (uint8_t)(out >> (8u * t));
  }
}

// ---- range-length scan (src/AwFmParallelSearch.c:328,367: lengths are u32-truncated) in three launches of our own:
// per-tile sums, one CTA scanning the tile sums, per-tile exclusive scan + tile base.  hitOffsets[n] = base + total. ----
constexpr int kScanTile = 2048;  // queries per CTA: 256 threads x 8
__device__ __forceinline__ uint64_t rangeLengthOf(const uint4 r) {
  const uint64_t sp = (uint64_t)r.x | ((uint64_t)r.y << 32), ep = (uint64_t)r.z | ((uint64_t)r.w << 32);
  return (uint32_t)(sp <= ep ? ep - sp + 1 : 0);
}
// length of query q's hit list: from its final range, or from the u32 count the search stored (the same number, a
// quarter of the bytes; what the locate pipelines scan)
template <bool FROM_COUNTS>
__device__ __forceinline__ uint64_t hitsOfQuery(const void *__restrict__ src, uint64_t q) {
  if (FROM_COUNTS) return __ldg(reinterpret_cast<const uint32_t *>(src) + q);
  return rangeLengthOf(__ldg(reinterpret_cast<const uint4 *>(src) + q));
}
__device__ __forceinline__ uint64_t blockSum256(uint64_t v, uint64_t *warpSums /* [8] shared */) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
  if ((threadIdx.x & 31u) == 0) warpSums[threadIdx.x >> 5] = v;
  __syncthreads();
  uint64_t total = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) total += warpSums[w];
  __syncthreads();
  return total;
}
template <bool FROM_COUNTS>
__global__ void __launch_bounds__(256)
    scanTileSums(const void *__restrict__ ranges, uint64_t n, uint64_t *__restrict__ tileSums) {
  __shared__ uint64_t warpSums[8];
  const uint64_t q0 = (uint64_t)blockIdx.x * kScanTile;
  uint64_t mine = 0;
#pragma unroll
  for (int it = 0; it < kScanTile / 256; it++) {
    const uint64_t q = q0 + it * 256 + threadIdx.x;
    if (q < n) mine += hitsOfQuery<FROM_COUNTS>(ranges, q);
  }
  const uint64_t total = blockSum256(mine, warpSums);
  if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}
// one CTA: exclusive scan of the tile sums in place (numTiles = n / 2048: 24 k for 50 M queries), grand total to
// tileSums[numTiles].  Thread t owns the consecutive tiles [t*per, (t+1)*per): it sums them (independent loads, all in
// flight together), the 256 partial sums are scanned once, and the thread walks its tiles again writing the exclusive
// prefixes — two sweeps over an L2-resident array instead of numTiles/256 dependent rounds of load, scan and barrier
// (50 M queries: the whole scan 0.53 -> 0.44 ms from ranges, together with the 128-bit requests of scanTileOffsets).
static __global__ void __launch_bounds__(256) scanTileBases(uint64_t *__restrict__ tileSums, uint64_t numTiles, uint64_t base) {
  __shared__ uint64_t warpSums[8];
  const uint64_t per = (numTiles + 255) / 256;
  const uint64_t t0 = min(numTiles, (uint64_t)threadIdx.x * per), t1 = min(numTiles, t0 + per);
  uint64_t mine = 0;
#pragma unroll 8
  for (uint64_t t = t0; t < t1; t++) mine += tileSums[t];
  uint64_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if ((threadIdx.x & 31u) >= (unsigned)d) incl += up;
  }
  if ((threadIdx.x & 31u) == 31u) warpSums[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint64_t run = base + incl - mine;
  for (unsigned w = 0; w < (threadIdx.x >> 5); w++) run += warpSums[w];
  for (uint64_t t = t0; t < t1; t++) {
    const uint64_t v = tileSums[t];
    tileSums[t] = run;
    run += v;
  }
  if (threadIdx.x == 255) tileSums[numTiles] = run;  // (threads past the last tile carry the total through unchanged)
}
template <bool FROM_COUNTS>
__global__ void __launch_bounds__(256)
    scanTileOffsets(const void *__restrict__ ranges, uint64_t n, const uint64_t *__restrict__ tileBases,
                    uint64_t *__restrict__ hitOffsets) {
  __shared__ uint64_t warpSums[8];
  const uint64_t q0 = (uint64_t)blockIdx.x * kScanTile;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  // thread t owns the 8 consecutive queries q0 + 8t .. q0 + 8t + 7
  static_assert(kScanTile / 256 == 8, "vector paths below: 8 queries per thread");
  uint64_t len[kScanTile / 256], mine = 0;
  const uint64_t qFirst = q0 + (uint64_t)threadIdx.x * (kScanTile / 256);
  // whole tile inside the batch and 16-byte aligned arrays (q0 is a multiple of 2048): 128-bit loads and stores
  const bool vector = q0 + kScanTile <= n && (reinterpret_cast<uintptr_t>(ranges) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(hitOffsets) & 15u) == 0;
  if (FROM_COUNTS && vector) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(ranges) + qFirst));
    const uint4 b = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(ranges) + qFirst) + 1);
    len[0] = a.x, len[1] = a.y, len[2] = a.z, len[3] = a.w, len[4] = b.x, len[5] = b.y, len[6] = b.z, len[7] = b.w;
#pragma unroll
    for (int j = 0; j < 8; j++) mine += len[j];
  } else {
#pragma unroll
    for (int j = 0; j < kScanTile / 256; j++) {
      const uint64_t q = qFirst + j;
      len[j] = q < n ? hitsOfQuery<FROM_COUNTS>(ranges, q) : 0;
      mine += len[j];
    }
  }
  uint64_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= (unsigned)d) incl += up;
  }
  if (lane == 31u) warpSums[warp] = incl;
  __syncthreads();
  uint64_t run = __ldg(tileBases + blockIdx.x) + incl - mine;
  for (unsigned w = 0; w < warp; w++) run += warpSums[w];
  if (vector) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      reinterpret_cast<ulonglong2 *>(hitOffsets + qFirst)[j >> 1] = make_ulonglong2(run, run + len[j]);
      run += len[j] + len[j + 1];
    }
  } else {
#pragma unroll
    for (int j = 0; j < kScanTile / 256; j++) {
      const uint64_t q = qFirst + j;
      if (q < n) hitOffsets[q] = run;
      run += len[j];
    }
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) hitOffsets[n] = __ldg(tileBases + gridDim.x);
}
static __global__ void addHitBase(uint64_t *__restrict__ hitOffsets, uint64_t count, uint64_t base) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    hitOffsets[i] += base;
}

// range lengths (u32-truncated, src/AwFmParallelSearch.c:328,367) for the exclusive scan that builds hitOffsets
static __global__ void rangeLengths(const uint4 *__restrict__ ranges, uint64_t n, uint64_t *__restrict__ lengths) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 r = ranges[i];
  const uint64_t sp = (uint64_t)r.x | ((uint64_t)r.y << 32), ep = (uint64_t)r.z | ((uint64_t)r.w << 32);
  lengths[i] = (uint32_t)(sp <= ep ? ep - sp + 1 : 0);
}

// random-gather bandwidth probe with this path's access shape: `lanes` consecutive lanes read one record
template <int BYTES, int LANES>
__global__ void gatherProbe(const uint4 *__restrict__ data, uint64_t numRecords, uint64_t numReads,
                            uint64_t *__restrict__ sink) {
  constexpr int U4 = BYTES / 16;            // uint4 per record
  constexpr int PER_LANE = (U4 + LANES - 1) / LANES;
  const unsigned sub = threadIdx.x % LANES;
  const uint64_t numGroups = (uint64_t)gridDim.x * blockDim.x / LANES;
  uint64_t acc = 0;
  constexpr int UNROLL = 4;  // independent records in flight per group
  for (uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES; i0 < numReads;
       i0 += numGroups * UNROLL) {
    uint4 v[UNROLL][PER_LANE];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const uint64_t i = i0 + (uint64_t)u * numGroups;
      uint64_t z = i + 0x9E3779B97F4A7C15ull;  // splitmix64
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      const uint64_t rec = z % numRecords;
#pragma unroll
      for (int k = 0; k < PER_LANE; k++) {
        const int c = sub + LANES * k;
        v[u][k] = (c < U4 && i < numReads) ? __ldg(data + rec * U4 + c) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
#pragma unroll
      for (int k = 0; k < PER_LANE; k++) acc += v[u][k].x ^ v[u][k].y ^ v[u][k].z ^ v[u][k].w;
  }
  if (acc == 0x1234567ull) sink[0] = acc;  // keep the loads alive
}

}  // namespace awfm
