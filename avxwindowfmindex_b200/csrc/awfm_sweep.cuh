// awfm_sweep.cuh — the "sweep" count path for LARGE batches of queries (sm_100a).  Described for fixed-length
// batches of the nucleotide alphabet; amino indexes take the same path with 20 buckets and 5-bit letters
// (SweepAlphabet<true>), variable-length batches with a marker bit in every record (sweepPackVar, VARLEN), nucleotide
// indexes of 2^32 .. 2^40 positions with 64-bit positions in registers and 40 + 24 bits in the record (WIDE).
//
// The tile kernels of awfm_kernels.cuh pay one random DRAM line per rank: 1 + 5.64 lines per 20-mer at 3.1 Gbp, and
// the memory system delivers ~42 G random lines/s whatever their size (profiles/r01_granularity_probe.jsonl), so they
// stop at 6.4 G queries/s with the DRAM pipe 65 % busy.  A batch of 100 M queries, however, puts ~8 queries on every
// 128-B line of the index at every step — the misses are only random because the queries are processed in input order.
// Backward search keeps order: if the live queries are ordered by the start of their current range, the positions
// the next LF step ranks (sp-1, ep) are non-decreasing, and after the step the queries that prepended the same letter
// c are still ordered (sp' = C[c] + Occ(c, sp-1) is monotone in sp).  So:
//
//   sweepPack*   thread per query: seed-table index of the last k letters (src/AwFmKmerTable.c:21-51) as the sort key,
//                the remaining len-k letters packed 2 bits each (next letter to prepend in the low bits) + query id as
//                the payload.  Queries holding anything but A/C/G/T(/U) go to a side list (sweepIrregular).
//   ordering     two most-significant-digit-first bucket passes (awfm_sort.cuh; CUB's radix sort as cross-check) on
//                the top 16 key bits, the first pass's tiles order the 8 bits below them: the order has to be exact at
//                warp granularity — 32 consecutive queries should rank within a handful of neighbouring lines — or
//                the L1 data pipe pays one wavefront per lane (measured: DESIGN.md section 3, dropped variants).
//   sweepStep<FIRST>   finishes the order on those low bits inside each tile (shared-memory counting sort), reads the
//                seed entry (src/AwFmParallelSearch.c:222-271), does the first LF step (src/AwFmSearch.c:42-103) and
//                appends the still-valid query as a 16-B record {sp, ep-sp, id, letters} to the bucket of the letter
//                it has just prepended;
//   sweepStep<!FIRST>  one pass per further letter over the buckets in letter order A,C,G,T — which is ascending sp
//                order, because ranges of strings starting with A precede those starting with C, ... and inside a
//                bucket the append order is the (ascending) input order — one LF step per record, same append.  A
//                query whose range empties is dropped (its count stays at the zero the output was cleared to,
//                src/AwFmParallelSearch.c:279-311 stops there too); one that has prepended all its letters writes
//                count = ep-sp+1 (u32, :187-190).  With range output every query also stores its final (sp, ep) once:
//                the seed entry if that is empty, the pair its search stops at, or its last range.
//
// Buckets are filled tile by tile through one atomicAdd per bucket per tile (order inside a tile is kept, tiles land
// in roughly launch order), so the order of whole tiles is approximate: the ranks of the tiles in flight at any moment
// touch a window of ~1 % of the BWT, which the 126 MB L2 holds.  The index is then STREAMED once per pass instead of
// being gathered line by line, and all other traffic (keys, records) is sequential.  No spin-waits anywhere.
//
// Exactness: same seed entries, same ranks (sectorRank), same stop rule; only the processing order differs, and the
// result of a query does not depend on it.  Not covered here (the caller falls back to the tile kernels): fixed-length
// batches with len - k > 24 or len > 32 (amino: len - k > 6), k > 16 (amino: 7), amino indexes beyond 2^32 positions, nucleotide ones beyond
// 2^40; inside a variable-length batch, queries shorter than k or with more than 15 (amino: 6) letters left of the seed
// are answered by sweepIrregular within the same call.
#pragma once
#include "awfm_kernels.cuh"
#include "awfm_sort.cuh"

#include <type_traits>

namespace awfm {

#ifndef AWFM_SWEEP_THREADS
#define AWFM_SWEEP_THREADS 256
#endif
constexpr int kSweepThreads = AWFM_SWEEP_THREADS;  // 256 or 512 (measured: profiles/r01_sweep_probe.jsonl)
constexpr uint32_t kSweepNoId = 0xFFFFFFFFu;
constexpr int kSweepMaxPasses = 26;
constexpr uint32_t kSweepMaxRestNuc = 24;   // fixed-length nucleotide batches: letters left of the seed k-mer (16 in the record + a refill)

// One generation of live records: bucket b (= letter prepended last) lives in arr[b >> 1]; even buckets grow up
// from slot 0, odd ones down from slot cap-1, so two buckets of unknown sizes share cap slots (total <= cap).
// Nucleotide: 4 buckets in 2 arrays; amino: 20 buckets in 10 arrays.
constexpr int kSweepMaxArrays = 10;
constexpr int kSweepCtrlStride = 32;  // device counters reserved per generation (4 or 20 used)
struct SweepRecs {
  uint4 *arr[kSweepMaxArrays];
  uint32_t *count;  // [4 | 20] device counters of this generation
  uint64_t cap;
};
template <bool AMINO>
struct SweepAlphabet {
  static constexpr uint32_t kCard = AMINO ? 20u : 4u;       // plain letters = buckets
  static constexpr uint32_t kLetterBits = AMINO ? 5u : 2u;  // per remaining letter in the payload
};

// ---------------------------------------------------------------------------------------------------------------
// sweepPackWords: len % 4 == 0 and `letters` 16-B aligned.  One thread per query reads its len/4 words straight from
// global memory (a warp covers 32*len contiguous bytes; the repeats are L1 hits) and translates four letters at a time:
//   code  = ((w >> 1) ^ (w >> 2)) & 3 per byte maps A,C,G,T/U (either case) to 0,1,2,3 (src/AwFmLetter.c:4-22);
//   valid = the byte, lower-cased, equals "acgt"[code] (or 'u' for code 3) — anything else is an irregular query.
// The whole query becomes one number Q, two bits per letter, first letter most significant: the seed-table index of
// the last k letters (src/AwFmKmerTable.c:21-51) is Q mod 4^k, and Q >> 2k has the letter prepended next in its low
// bits.
// ---------------------------------------------------------------------------------------------------------------
// Top-digit histogram of the ordering step (awfm_sort.cuh), kept per CTA in shared memory by the pack kernels and
// flushed with one atomic per bin and CTA.  sortCtrl == nullptr: the caller orders the pairs some other way.
struct PackHistogram {
  uint32_t *bins;  // [kSortBins] shared
  __device__ __forceinline__ void begin(uint32_t *shared, const SortCtrl *ctrl) {
    bins = shared;
    if (ctrl) {
      for (uint32_t b = threadIdx.x; b < (uint32_t)kSortBins; b += blockDim.x) bins[b] = 0;
      __syncthreads();
    }
  }
  __device__ __forceinline__ void add(const SortCtrl *ctrl, uint32_t key, uint32_t shiftA) {
    if (ctrl) atomicAdd(&bins[key >> shiftA], 1u);
  }
  __device__ __forceinline__ void flush(SortCtrl *ctrl) {
    if (ctrl) {
      __syncthreads();
      for (uint32_t b = threadIdx.x; b < (uint32_t)kSortBins; b += blockDim.x)
        if (bins[b]) atomicAdd(&ctrl->countA[b], bins[b]);
    }
  }
};
// One packed query leaves the pack kernel either as a (key, payload | id) pair or — compactShift != 0, the bucket
// passes on compact pairs (awfm_sort.cuh) — as ONE word at index q: key | payload << compactShift, bit 63 = irregular.
// Nucleotide batches with 17..24 letters left of the seed k-mer (32-mers on a k = 12 table): the pair holds the first 16
// of them, letters 17.. go to more[q], from where sweepRefill hands them to the record once the 16 have been prepended.
__device__ __forceinline__ void sweepStorePair(uint32_t *__restrict__ keys, uint64_t *__restrict__ vals, uint64_t q, uint32_t key,
                                               uint64_t payload, bool irregular, uint32_t compactShift,
                                               uint32_t *__restrict__ more = nullptr) {
  if (compactShift) {
    vals[q] = (uint64_t)key | (payload << compactShift) | ((uint64_t)irregular << 63);
  } else {
    keys[q] = key;
    vals[q] = (payload << 32) | (irregular ? kSweepNoId : (uint32_t)q);
    if (more) more[q] = (uint32_t)(payload >> 32);
  }
}
__device__ __forceinline__ uint32_t packFourLetters(uint32_t w, uint32_t &bad) {
  const uint32_t code = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
  const uint32_t b0 = code & 0x01010101u, b1 = (code >> 1) & 0x01010101u, both = b0 & b1;
  const uint32_t expected = 0x61616161u + b0 * 2u + b1 * 6u + both * 11u;  // 'a', 'c' = +2, 'g' = +6, 't' = +19
  bad |= ((w | 0x20202020u) ^ expected) & ~both;                           // 't' ^ 'u' == 1, allowed where code == 3
  uint32_t t = __byte_perm(code, 0, 0x0123);                               // first letter into the top byte
  t = (t | (t >> 6)) & 0x000F000Fu;
  return (t | (t >> 12)) & 0xFFu;                                          // first letter in bits 7..6
}
template <int WORDS>  // words per query (len / 4), fully unrolled
__global__ void __launch_bounds__(256)
    sweepPackWords(const uint32_t *__restrict__ words, uint64_t numQueries, uint32_t k, uint32_t *__restrict__ keys,
                   uint64_t *__restrict__ vals, uint32_t *__restrict__ irregularIds,
                   uint32_t *__restrict__ irregularCount, SortCtrl *__restrict__ sortCtrl, uint32_t shiftA,
                   uint32_t compactShift, uint32_t *__restrict__ more /* or nullptr: at most 16 letters left of the seed */) {
  __shared__ uint32_t histShared[kSortBins];
  PackHistogram hist;
  hist.begin(histShared, sortCtrl);
  const uint64_t keyMask = (k >= 16) ? 0xFFFFFFFFull : ((1ull << (2 * k)) - 1ull);
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < numQueries;
       q += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t *src = words + q * WORDS;
    uint32_t w[WORDS];
#pragma unroll
    for (int i = 0; i < WORDS; i++) w[i] = __ldg(src + i);
    uint64_t Q = 0;
    uint32_t bad = 0;
#pragma unroll
    for (int i = 0; i < WORDS; i++) Q = (Q << 8) | packFourLetters(w[i], bad);
    if (bad) irregularIds[atomicAdd(irregularCount, 1u)] = (uint32_t)q;
    sweepStorePair(keys, vals, q, (uint32_t)(Q & keyMask), Q >> (2 * k), bad != 0, compactShift, more);
    hist.add(sortCtrl, (uint32_t)(Q & keyMask), shiftA);
  }
  hist.flush(sortCtrl);
}

// Amino counterpart (len % 4 == 0): the query's words straight from global memory, letters translated one by one
// (src/AwFmLetter.c:55-67), key = mixed-radix seed index of the last k letters (src/AwFmKmerTable.c:36-51), five bits
// per remaining letter with the letter prepended next in the low bits.
template <int WORDS>
__global__ void __launch_bounds__(256)
    sweepPackWordsAmino(const uint32_t *__restrict__ words, uint64_t numQueries, uint32_t k, uint32_t *__restrict__ keys,
                        uint64_t *__restrict__ vals, uint32_t *__restrict__ irregularIds,
                        uint32_t *__restrict__ irregularCount, SortCtrl *__restrict__ sortCtrl, uint32_t shiftA,
                        uint32_t compactShift) {
  __shared__ uint32_t histShared[kSortBins];
  PackHistogram hist;
  hist.begin(histShared, sortCtrl);
  constexpr uint32_t LEN = 4 * WORDS;
  const uint32_t rest = LEN - k;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < numQueries;
       q += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t *src = words + q * WORDS;
    uint32_t w[WORDS];
#pragma unroll
    for (int i = 0; i < WORDS; i++) w[i] = __ldg(src + i);
    uint32_t key = 0, packed = 0, bad = 0;
#pragma unroll
    for (uint32_t i = 0; i < LEN; i++) {
      const uint32_t l = aminoLetterIndex((w[i >> 2] >> (8u * (i & 3u))) & 0xFFu);
      bad |= l >= 20u;
      const uint32_t v = l < 20u ? l : 0u;
      if (i >= rest) key = key * 20u + v;              // leftmost of the last k letters most significant
      else packed |= v << (5u * (rest - 1u - i));      // letter prepended at step j+1 is s[rest-1-j]
    }
    if (bad) irregularIds[atomicAdd(irregularCount, 1u)] = (uint32_t)q;
    sweepStorePair(keys, vals, q, key, packed, bad != 0, compactShift);
    hist.add(sortCtrl, key, shiftA);
  }
  hist.flush(sortCtrl);
}

// ---------------------------------------------------------------------------------------------------------------
// sweepPack: 256 queries per tile staged through shared memory with coalesced 128-bit loads.
// ---------------------------------------------------------------------------------------------------------------
template <bool AMINO>  // amino: key = mixed-radix seed index (radix 20), 5 bits per remaining letter
__global__ void __launch_bounds__(256)
    sweepPack(const uint8_t *__restrict__ letters, uint64_t numQueries, uint32_t len, uint32_t k,
              uint32_t *__restrict__ keys, uint64_t *__restrict__ vals, uint32_t *__restrict__ irregularIds,
              uint32_t *__restrict__ irregularCount, SortCtrl *__restrict__ sortCtrl, uint32_t shiftA,
              uint32_t compactShift, uint32_t *__restrict__ more) {
  constexpr uint32_t CARD = SweepAlphabet<AMINO>::kCard, LB = SweepAlphabet<AMINO>::kLetterBits;
  extern __shared__ __align__(16) uint8_t sLetters[];  // 256 * len bytes, rounded up to 16
  __shared__ uint32_t histShared[kSortBins];
  PackHistogram hist;
  hist.begin(histShared, sortCtrl);
  const uint64_t numTiles = (numQueries + 255) / 256;
  const uint64_t totalBytes = numQueries * (uint64_t)len;
  const uint32_t rest = len - k;
  for (uint64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
    const uint64_t q0 = tile * 256;
    const uint32_t nq = (uint32_t)min((uint64_t)256, numQueries - q0);
    const uint64_t byte0 = q0 * (uint64_t)len;  // multiple of 256: 16-B aligned when `letters` is
    const uint32_t bytes = nq * len;
    __syncthreads();  // previous tile consumed
    for (uint32_t i = threadIdx.x; i < (bytes + 15) / 16; i += blockDim.x) {
      const uint64_t g = byte0 + 16ull * i;
      uint4 v;
      if (g + 16 <= totalBytes) {
        v = __ldg(reinterpret_cast<const uint4 *>(letters + g));
      } else {  // never read past the batch's final letter
        uint32_t w[4] = {0, 0, 0, 0};
        for (uint32_t b = 0; g + b < totalBytes; b++) w[b >> 2] |= (uint32_t)__ldg(letters + g + b) << (8 * (b & 3));
        v = make_uint4(w[0], w[1], w[2], w[3]);
      }
      reinterpret_cast<uint4 *>(sLetters)[i] = v;
    }
    __syncthreads();
    if (threadIdx.x < nq) {
      const uint8_t *s = sLetters + threadIdx.x * len;
      uint32_t key = 0, bad = 0;
      uint64_t packed = 0;
      for (uint32_t i = 0; i < k; i++) {  // leftmost of the last k letters most significant
        const uint32_t l = letterIndex<AMINO>(s[rest + i]);
        bad |= l >= CARD;
        key = key * CARD + (l < CARD ? l : 0u);
      }
      for (uint32_t j = 0; j < rest; j++) {  // letter prepended at step j+1 is s[rest-1-j]: low bits first
        const uint32_t l = letterIndex<AMINO>(s[rest - 1 - j]);
        bad |= l >= CARD;
        packed |= (uint64_t)(l < CARD ? l : 0u) << (LB * j);
      }
      if (bad) irregularIds[atomicAdd(irregularCount, 1u)] = (uint32_t)(q0 + threadIdx.x);
      sweepStorePair(keys, vals, q0 + threadIdx.x, key, packed, bad != 0, compactShift, more);
      hist.add(sortCtrl, key, shiftA);
    }
  }
  hist.flush(sortCtrl);
}

// ---------------------------------------------------------------------------------------------------------------
// sweepPackBits: the 2-bit packed query format of include/awfm_gpu.h (AWFM_QUERY_2BIT): query i occupies bytes
// [i*B, (i+1)*B), B = ceil(len/4), letter j in bits [2j, 2j+2) of that little-endian byte string, codes 0..3 = A,C,G,T
// (the reference's letter indices, src/AwFmLetter.c:4-22).  len <= 32.  A tile of 256 queries is staged through shared
// memory with coalesced 128-bit loads; one bit reversal + one swap of neighbouring bits turns the query into the same
// number Q sweepPackWords builds (first letter most significant).  The format has no ambiguity codes, so there are no
// irregular queries.
// ---------------------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
    sweepPackBits(const uint8_t *__restrict__ packed, uint64_t numQueries, uint32_t len, uint32_t k,
                  uint32_t *__restrict__ keys, uint64_t *__restrict__ vals, SortCtrl *__restrict__ sortCtrl,
                  uint32_t shiftA, uint32_t compactShift, uint32_t *__restrict__ more) {
  extern __shared__ __align__(16) uint8_t sPacked[];  // 256 * B bytes, rounded up to 16
  __shared__ uint32_t histShared[kSortBins];
  PackHistogram hist;
  hist.begin(histShared, sortCtrl);
  const uint32_t B = (len + 3u) >> 2;
  const uint64_t numTiles = (numQueries + 255) / 256;
  const uint64_t totalBytes = numQueries * (uint64_t)B;
  const uint64_t keyMask = (k >= 16) ? 0xFFFFFFFFull : ((1ull << (2 * k)) - 1ull);
  for (uint64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
    const uint64_t q0 = tile * 256;
    const uint32_t nq = (uint32_t)min((uint64_t)256, numQueries - q0);
    const uint64_t byte0 = q0 * (uint64_t)B;  // multiple of 256
    const uint32_t bytes = nq * B;
    __syncthreads();  // previous tile consumed
    for (uint32_t i = threadIdx.x; i < (bytes + 15) / 16; i += blockDim.x) {
      const uint64_t g = byte0 + 16ull * i;
      uint4 v;
      if (g + 16 <= totalBytes) {
        v = __ldg(reinterpret_cast<const uint4 *>(packed + g));
      } else {  // never read past the batch's final byte
        uint32_t w[4] = {0, 0, 0, 0};
        for (uint32_t b = 0; g + b < totalBytes; b++) w[b >> 2] |= (uint32_t)__ldg(packed + g + b) << (8 * (b & 3));
        v = make_uint4(w[0], w[1], w[2], w[3]);
      }
      reinterpret_cast<uint4 *>(sPacked)[i] = v;
    }
    __syncthreads();
    if (threadIdx.x < nq) {
      const uint8_t *s = sPacked + threadIdx.x * B;
      uint64_t raw = 0;
      for (uint32_t b = 0; b < B; b++) raw |= (uint64_t)s[b] << (8u * b);
      uint64_t r = __brevll(raw);  // letter j: bits (2j, 2j+1) -> (63-2j, 62-2j)
      r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);  // ... -> (62-2j, 63-2j)
      const uint64_t Q = r >> (64u - 2u * len);  // first letter most significant, 2 bits per letter
      sweepStorePair(keys, vals, q0 + threadIdx.x, (uint32_t)(Q & keyMask), Q >> (2 * k), false, compactShift, more);
      hist.add(sortCtrl, (uint32_t)(Q & keyMask), shiftA);
    }
  }
  hist.flush(sortCtrl);
}

// ---------------------------------------------------------------------------------------------------------------
// sweepPackVar: VARIABLE-length ASCII batches (letters + numQueries+1 letter offsets, the layout the search-list
// engine and awfm_gpu_count_device already use).  The payload's 32 bits hold the letters left of the seed k-mer as
// above PLUS a marker bit right above the last of them: a record has prepended all its letters when its payload has
// shrunk to 1, so every record carries its own length and the passes need no common one.  Queries shorter than k
// (the reference opens those from the last letter, src/AwFmParallelSearch.c:241-268), with more than 15 (amino: 6)
// letters left of the seed, or holding anything but plain letters go to the irregular list.
// `letters` is 16-byte aligned, nothing past offsets[numQueries] is read.
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kSweepVarMaxRestNuc = 15, kSweepVarMaxRestAmino = 6;
__device__ __forceinline__ uint32_t sweepSafeWord(const uint8_t *__restrict__ letters, uint64_t wordIndex, uint64_t totalBytes) {
  const uint64_t g = wordIndex * 4ull;
  if (g + 4 <= totalBytes) return __ldg(reinterpret_cast<const uint32_t *>(letters) + wordIndex);
  uint32_t w = 0;
  for (uint32_t b = 0; g + b < totalBytes; b++) w |= (uint32_t)__ldg(letters + g + b) << (8u * b);
  return w;
}
// one query straight from global memory (amino batches; nucleotide tiles too long for the staging buffer)
template <bool AMINO>
__device__ __forceinline__ void sweepPackVarDirect(const uint8_t *__restrict__ letters, uint64_t totalBytes, uint64_t o,
                                                   uint32_t len, uint32_t k, uint64_t keyMask, uint32_t &key,
                                                   uint32_t &payload, uint32_t &bad) {
  const uint32_t rest = len - k, shift = (uint32_t)(o & 3u) * 8u;
  const uint64_t w0 = o >> 2, wEnd = (o + len + 3) >> 2;  // words [w0, wEnd) cover the query's letters
  uint32_t a = sweepSafeWord(letters, w0, totalBytes);
  uint64_t Q = 0;
  uint32_t packed = 0;
  for (uint32_t i = 0; 4u * i < len; i++) {
    const uint32_t b = (w0 + i + 1 < wEnd) ? sweepSafeWord(letters, w0 + i + 1, totalBytes) : 0u;
    const uint32_t group = __funnelshift_r(a, b, shift);  // letters 4i .. 4i+3 of the query
    a = b;
    const uint32_t t = min(4u, len - 4u * i);
    if constexpr (AMINO) {
      for (uint32_t j = 0; j < t; j++) {
        const uint32_t at = 4u * i + j;
        const uint32_t l = aminoLetterIndex((group >> (8u * j)) & 0xFFu);
        bad |= l >= 20u;
        const uint32_t v = l < 20u ? l : 0u;
        if (at >= rest) key = key * 20u + v;            // leftmost of the last k letters most significant
        else packed |= v << (5u * (rest - 1u - at));    // letter prepended at step j+1 is s[rest-1-j]
      }
    } else {
      uint32_t bad4 = 0;
      uint32_t p = packFourLetters(group, bad4);  // first letter in bits 7..6
      if (t < 4u) p >>= 2u * (4u - t), bad4 &= (1u << (8u * t)) - 1u;
      Q = (Q << (2u * t)) | p;
      bad |= bad4;
    }
  }
  if constexpr (AMINO) {
    payload = packed | (1u << (5u * rest));
  } else {
    key = (uint32_t)(Q & keyMask);
    payload = (uint32_t)(Q >> (2 * k)) | (1u << (2u * rest));
  }
}
// Nucleotide tiles: the 256 queries' letters (at most 31 each if they are to stay in the sweep, so 8 KB hold them) are
// turned into a 2-bit stream cooperatively — thread i translates bytes [32i, 32i+32) of the tile with eight
// packFourLetters and notes which of its eight 4-letter groups held anything but A/C/G/T/U — and every query then
// cuts its own letters out of that stream with two funnel shifts.  A 4-letter group may straddle two queries: an
// irregular letter then sends its neighbour to the irregular list as well, which only costs that query the slower path.
constexpr uint32_t kSweepVarStageBytes = 8 * 1024;
template <bool AMINO>
__global__ void __launch_bounds__(256)
    sweepPackVar(const uint8_t *__restrict__ letters, const uint64_t *__restrict__ offsets, uint64_t numQueries, uint32_t k,
                 uint32_t *__restrict__ keys, uint64_t *__restrict__ vals, uint32_t *__restrict__ irregularIds,
                 uint32_t *__restrict__ irregularCount, SortCtrl *__restrict__ sortCtrl, uint32_t shiftA,
                 uint32_t compactShift) {
  constexpr uint32_t kMaxRest = AMINO ? kSweepVarMaxRestAmino : kSweepVarMaxRestNuc;
  __shared__ uint32_t histShared[kSortBins];
  __shared__ uint64_t sOff[257];
  __shared__ uint32_t sCodes[AMINO ? 1 : kSweepVarStageBytes / 16 + 2];  // word j: letters 16j..16j+15, first in the top bits
  __shared__ uint8_t sBadGroups[AMINO ? 1 : kSweepVarStageBytes / 32 + 4];  // byte i: groups 8i..8i+7, first in bit 7
  PackHistogram hist;
  hist.begin(histShared, sortCtrl);
  const uint64_t totalBytes = __ldg(offsets + numQueries);
  const uint64_t keyMask = (k >= 16) ? 0xFFFFFFFFull : ((1ull << (2 * k)) - 1ull);
  const uint64_t numTiles = (numQueries + 255) / 256;
  // this thread's offset of the tile after the current one (and, thread 0, that tile's closing offset): requested one
  // tile ahead so that only the letter loads sit between a tile's barriers
  auto fetchOffsets = [&](uint64_t tile, uint64_t &mine, uint64_t &closing) {
    if (tile >= numTiles) return;
    const uint64_t q0 = tile * 256;
    const uint32_t nq = (uint32_t)min((uint64_t)256, numQueries - q0);
    if (threadIdx.x <= nq) mine = __ldg(offsets + q0 + threadIdx.x);
    if (threadIdx.x == 0) closing = __ldg(offsets + q0 + nq);
  };
  uint64_t nextMine = 0, nextClosing = 0;
  fetchOffsets(blockIdx.x, nextMine, nextClosing);
  for (uint64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
    const uint64_t q0 = tile * 256;
    const uint32_t nq = (uint32_t)min((uint64_t)256, numQueries - q0);
    __syncthreads();  // previous tile consumed
    if (threadIdx.x <= nq) sOff[threadIdx.x] = nextMine;
    if (threadIdx.x == 0) sOff[nq] = nextClosing;  // (nq == 256: the 257th offset)
    __syncthreads();
    fetchOffsets(tile + gridDim.x, nextMine, nextClosing);
    const uint64_t aligned0 = sOff[0] & ~15ull, byte1 = sOff[nq];
    const bool staged = !AMINO && byte1 >= aligned0 && byte1 - aligned0 <= kSweepVarStageBytes;  // uniform per CTA
    if constexpr (!AMINO) {
      if (staged) {
        const uint64_t g = aligned0 + 32ull * threadIdx.x;
        if (g < byte1) {
          uint32_t w[8];
          if (g + 32 <= totalBytes) {
            const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(letters + g));
            const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(letters + g) + 1);
            w[0] = v0.x, w[1] = v0.y, w[2] = v0.z, w[3] = v0.w, w[4] = v1.x, w[5] = v1.y, w[6] = v1.z, w[7] = v1.w;
          } else {  // never read past the batch's final letter
#pragma unroll
            for (int j = 0; j < 8; j++) w[j] = sweepSafeWord(letters, (g >> 2) + j, totalBytes);
          }
          uint32_t codes[2] = {0, 0}, badGroups = 0;
#pragma unroll
          for (int j = 0; j < 8; j++) {
            uint32_t bad4 = 0;
            codes[j >> 2] = (codes[j >> 2] << 8) | packFourLetters(w[j], bad4);
            badGroups = (badGroups << 1) | (bad4 ? 1u : 0u);
          }
          sCodes[2 * threadIdx.x] = codes[0];
          sCodes[2 * threadIdx.x + 1] = codes[1];
          sBadGroups[threadIdx.x] = (uint8_t)badGroups;
        }
        __syncthreads();
      }
    }
    if (threadIdx.x < nq) {
      const uint64_t q = q0 + threadIdx.x;
      const uint64_t o = sOff[threadIdx.x], len64 = sOff[threadIdx.x + 1] - o;
      uint32_t key = 0, payload = 1u, bad = (len64 < k || len64 - k > kMaxRest) ? 1u : 0u;
      if (!bad) {
        const uint32_t len = (uint32_t)len64;
        if (staged) {
          const uint32_t a = (uint32_t)(o - aligned0), word = a >> 4, sh = 2u * (a & 15u);
          const uint32_t c0 = sCodes[word], c1 = sCodes[word + 1], c2 = sCodes[word + 2];
          const uint64_t window = ((uint64_t)__funnelshift_l(c1, c0, sh) << 32) | __funnelshift_l(c2, c1, sh);  // letters a .. a+31
          const uint64_t Q = window >> (64u - 2u * len);  // 1 <= len <= 31: first letter most significant
          key = (uint32_t)(Q & keyMask);
          payload = (uint32_t)(Q >> (2 * k)) | (1u << (2u * (len - k)));
          const uint32_t g0 = a >> 2, groups = ((a + len - 1u) >> 2) - g0 + 1u;  // <= 9 groups, within 16 bits from g0's byte
          const uint32_t two = ((uint32_t)sBadGroups[g0 >> 3] << 8) | sBadGroups[(g0 >> 3) + 1];
          bad = (two >> (16u - (g0 & 7u) - groups)) & ((1u << groups) - 1u);
        } else {
          sweepPackVarDirect<AMINO>(letters, totalBytes, o, len, k, keyMask, key, payload, bad);
        }
      }
      if (bad) {
        irregularIds[atomicAdd(irregularCount, 1u)] = (uint32_t)q;
        key = 0;
        payload = 1u;
      }
      sweepStorePair(keys, vals, q, key, payload, bad != 0, compactShift);
      hist.add(sortCtrl, key, shiftA);
    }
  }
  hist.flush(sortCtrl);
}

// ---------------------------------------------------------------------------------------------------------------
// sweepStep: one pass.  FIRST: input = sorted (key, payload) pairs, range from the seed table; else input = the
// previous generation's buckets in letter order.  `steps` = LF steps the queries of this pass still have to do
// INCLUDING this pass's (0 only for FIRST with len == k).
// The pass is issue-bound before it is DRAM-bound (profiles/r01_ncu_sweep_*.json), so everything is 32-bit: positions
// (bwtLength < 2^32), record indices (n < 2^31), and the rank works on 32-bit halves of the sector with a selector
// specialised for the four plain letters (same truth table as nucCodeCare: A = b2&b1, C = b2&b0, G = b1&b0,
// T = ~b2&~b1&b0, src/AwFmOccurrence.c:18-35).
// ---------------------------------------------------------------------------------------------------------------
#ifndef AWFM_SWEEP_NEXT_PREFETCH
#define AWFM_SWEEP_NEXT_PREFETCH 0  // 1: next tile's input towards L2, 2: towards L1
#endif
__device__ __forceinline__ void sweepPrefetchNext(const void *p) {
#if AWFM_SWEEP_NEXT_PREFETCH == 2
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
struct SweepSelector {
  uint32_t flipHi;               // T: match zeros in b2 and b1
  uint32_t any0, any1, any2;     // A ignores b0, C ignores b1, G ignores b2
};
__device__ __forceinline__ SweepSelector sweepSelector(uint32_t letter) {
  SweepSelector s;
  s.flipHi = letter == 3u ? 0xFFFFFFFFu : 0u;
  s.any0 = letter == 0u ? 0xFFFFFFFFu : 0u;
  s.any1 = letter == 1u ? 0xFFFFFFFFu : 0u;
  s.any2 = letter == 2u ? 0xFFFFFFFFu : 0u;
  return s;
}
// C[letter] + Occ(letter, p) for letter 0..3: sector read by this thread, superblock row from L1/L2.
// Pos = uint32_t (bwtLength < 2^32) or uint64_t (WIDE passes).
template <typename Pos>
__device__ __forceinline__ Pos sweepSuper(const DevIndex &ix, uint64_t row, uint32_t letter) {
  if constexpr (sizeof(Pos) == 8) return __ldg(ix.superC + row * kSectorSuperStride + letter);
  else return __ldg(reinterpret_cast<const uint32_t *>(ix.superC) + (row * kSectorSuperStride + letter) * 2u);
}
// One sector in registers; `when` == false: nothing is requested (a predicated-off LDG costs no L1 wavefront) and the
// words are undefined — the caller substitutes another sector's.
struct SweepSector {
  uint64_t a, b, c, d;  // words 0-1, 2-3, 4-5 (code bits 0, 1, 2), 6-7 (counts)
};
__device__ __forceinline__ SweepSector sweepLoadSector(const uint4 *sec) {
  SweepSector s;
#ifdef AWFM_NO_LDG256
  const uint4 v0 = __ldg(sec), v1 = __ldg(sec + 1);
  s.a = v0.x | ((uint64_t)v0.y << 32), s.b = v0.z | ((uint64_t)v0.w << 32);
  s.c = v1.x | ((uint64_t)v1.y << 32), s.d = v1.z | ((uint64_t)v1.w << 32);
#else
  // the whole 32-B sector in ONE request (LDG.E.256, sm_100): the passes are bound by the L1/LSU pipe, not by DRAM
  asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(s.a), "=l"(s.b), "=l"(s.c), "=l"(s.d) : "l"(sec));
#endif
  return s;
}
// the sector at `sec` if `when`, else a copy of `other`: the request is PREDICATED (a predicated-off LDG costs no L1
// wavefront and nothing is branched around, so it is in flight together with the request for `other`)
__device__ __forceinline__ SweepSector sweepLoadSectorOr(const uint4 *sec, bool when, const SweepSector &other) {
  SweepSector s;
  asm("{\n\t.reg .pred p;\n\t.reg .b64 t0, t1, t2, t3;\n\t"
      "setp.ne.u32 p, %9, 0;\n\t"
      "mov.b64 t0, 0;\n\tmov.b64 t1, 0;\n\tmov.b64 t2, 0;\n\tmov.b64 t3, 0;\n\t"  // (undefined registers get a stack home)
      "@p ld.global.nc.v4.u64 {t0,t1,t2,t3}, [%8];\n\t"
      "selp.b64 %0, t0, %4, p;\n\t"
      "selp.b64 %1, t1, %5, p;\n\t"
      "selp.b64 %2, t2, %6, p;\n\t"
      "selp.b64 %3, t3, %7, p;\n\t}"
      : "=l"(s.a), "=l"(s.b), "=l"(s.c), "=l"(s.d)
      : "l"(other.a), "l"(other.b), "l"(other.c), "l"(other.d), "l"(sec), "r"((uint32_t)when));
  return s;
}
template <typename Pos>  // `super` = sweepSuper(row of p)
__device__ __forceinline__ Pos sweepRankIn(const SweepSector &v, Pos p, uint32_t letter, const SweepSelector &s, Pos super) {
  const uint32_t b0lo = (uint32_t)v.a, b0hi = (uint32_t)(v.a >> 32), b1lo = (uint32_t)v.b, b1hi = (uint32_t)(v.b >> 32);
  const uint32_t b2lo = (uint32_t)v.c, b2hi = (uint32_t)(v.c >> 32);
  const int local = (int)((uint32_t)p & 63u) + 1;             // positions 0..local-1 of the sector count
  const uint32_t maskLo = lowBits(local), maskHi = lowBits(local - 32);
  const uint32_t lo = ((b1lo ^ s.flipHi) | s.any1) & ((b2lo ^ s.flipHi) | s.any2) & (b0lo | s.any0) & maskLo;
  const uint32_t hi = ((b1hi ^ s.flipHi) | s.any1) & ((b2hi ^ s.flipHi) | s.any2) & (b0hi | s.any0) & maskHi;
  const uint32_t rel = (uint32_t)(v.d >> (letter * 16u)) & 0xFFFFu;
  return super + rel + __popc(lo) + __popc(hi);
}
template <typename Pos>
__device__ __forceinline__ Pos sweepRank(const DevIndex &ix, Pos p, uint32_t letter, const SweepSelector &s, Pos super) {
  return sweepRankIn<Pos>(sweepLoadSector(ix.lines + (uint64_t)(p >> 6) * kSectorU4), p, letter, s, super);
}

// Amino: one thread reads the code bits and the letter's count of one 128-B quarter-line (awfm_device.cuh): uint4 0/1 =
// b0..b3 of the low/high 32 positions, words 8/9 = b4, word 11+c = count of letter c.  Selector = (code, care) of
// kAminoCodeCare, the same the tile kernels use (src/AwFmOccurrence.c:65-134).
struct AminoSweepSelector {
  uint32_t flip[5], any[5];
};
__device__ __forceinline__ AminoSweepSelector aminoSweepSelector(uint32_t codeCare) {
  AminoSweepSelector s;
  const uint32_t code = codeCare & 0xFFu, care = codeCare >> 8;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    s.flip[i] = ((code >> i) & 1u) - 1u;  // code bit 1 -> keep, 0 -> invert
    s.any[i] = ((care >> i) & 1u) - 1u;   // cared -> 0, ignored -> ~0
  }
  return s;
}
__device__ __forceinline__ uint32_t aminoSweepRank(const DevIndex &ix, uint32_t p, uint32_t letter,
                                                   const AminoSweepSelector &s) {
  const uint4 *line = ix.lines + (uint64_t)(p >> 6) * kAminoLineU4;
  uint4 v0, v1;
#ifdef AWFM_AMINO_NO_LDG256
  v0 = __ldg(line), v1 = __ldg(line + 1);
#else
  {  // b0..b3 of all 64 positions in one 256-bit request (see sweepRank)
    uint64_t a, b, c, d;
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(line));
    v0 = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
    v1 = make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)d, (uint32_t)(d >> 32));
  }
#endif
  const uint2 b4 = __ldg(reinterpret_cast<const uint2 *>(line + 2));
  const uint32_t rel = __ldg(reinterpret_cast<const uint32_t *>(line) + kAminoRelWord + letter);
  const uint32_t super = (uint32_t)__ldg(ix.superC + (uint64_t)(p >> kSuperShift) * kAminoSuperStride + letter);
  const int local = (int)(p & 63u) + 1;
  const uint32_t maskLo = lowBits(local), maskHi = lowBits(local - 32);
  const uint32_t lo = ((v0.x ^ s.flip[0]) | s.any[0]) & ((v0.y ^ s.flip[1]) | s.any[1]) & ((v0.z ^ s.flip[2]) | s.any[2]) &
                      ((v0.w ^ s.flip[3]) | s.any[3]) & ((b4.x ^ s.flip[4]) | s.any[4]) & maskLo;
  const uint32_t hi = ((v1.x ^ s.flip[0]) | s.any[0]) & ((v1.y ^ s.flip[1]) | s.any[1]) & ((v1.z ^ s.flip[2]) | s.any[2]) &
                      ((v1.w ^ s.flip[3]) | s.any[3]) & ((b4.y ^ s.flip[4]) | s.any[4]) & maskHi;
  return super + rel + __popc(lo) + __popc(hi);
}

// REC12 (nucleotide, at most 8 letters left of the seed): live records travel as 12 bytes instead of 16 —
// {sp, id} (8 B) in the first cap*8 bytes of a double-ended array, {range width : 16 | remaining letters : 16} (4 B)
// behind them.  A range can only narrow from step to step, so the 16-bit width is checked once: a query whose SEED
// range is wider than 65534 goes to the irregular list (sweepIrregular answers it with the generic per-query search).
// (forcing more resident CTAs per SM through __launch_bounds__ was measured: 5, 6 and 8 CTAs spill and run 20-30 %
// slower than the 64 registers / 4 CTAs the compiler picks on its own, profiles/r02_sweep_probe.jsonl)
// VARLEN (sweepPackVar's payloads, marker bit above the last remaining letter): a record is finished when its payload
// has shrunk to 1; `steps` is then the largest number of steps any record of the batch can still have to do.
// WIDE (nucleotide indexes of 2^32 .. 2^40 positions): 64-bit positions in registers; a record is still 16 bytes —
// {sp bits 0-31, sp bits 32-39 | range width << 8, id, letters} — because a range only narrows from step to step: a
// query whose SEED range is wider than 2^24 - 2 leaves for the irregular list once, in the first pass.
constexpr uint32_t kSweepEmitBits = 4, kSweepEmitBuckets = 1u << kSweepEmitBits;
// EMIT (the LAST pass of a fixed-length nucleotide batch with range output and many survivors — locate workloads): a
// query that has prepended all its letters does not store its count and range at [id] from here — ids are in no order
// by now, so that is two scattered partial-sector stores per query, each a read-modify-write in DRAM (cfg 5, 10 M
// 32-mers all found: 0.70 ms for the last pass against 0.145 for any other) — but is appended once more, to the
// bucket of the SIXTEENTH of the id space its id lies in (id / emitDiv; a slice of n/16 ids holds at most that many
// records, so the slices lie flat in the output generation's first array).  sweepEmit then writes the slices out in
// order: one slice of counts and ranges (20 B per query) stays in L2 until its sectors are complete.
template <bool FIRST, int kSweepItems, bool AMINO = false, bool REC12 = false, bool VARLEN = false, bool WIDE = false,
          bool EMIT = false>
__global__ void __launch_bounds__(kSweepThreads, (!AMINO && !WIDE && kSweepThreads == 256 && kSweepItems <= 4) ? 4 : 0)
    sweepStep(const __grid_constant__ DevIndex ix, const uint32_t *__restrict__ keys, const uint64_t *__restrict__ vals,
              uint64_t numPairs, bool deep, const __grid_constant__ SweepRecs in, const __grid_constant__ SweepRecs out,
              uint32_t steps, uint32_t localBits, uint32_t *__restrict__ counts,
              uint4 *__restrict__ ranges /* or nullptr: every query's final (sp, ep) as the reference leaves it */,
              uint32_t *__restrict__ irregularIds, uint32_t *__restrict__ irregularCount,
              bool rangesOfHitsOnly /* ranges only of queries whose final range is non-empty (locate) */,
              uint32_t emitDiv /* EMIT: query ids per output bucket */) {
  static_assert(!(EMIT && (AMINO || REC12 || VARLEN)), "ordered emit: fixed-length nucleotide batches, 16-byte records");
  static_assert(!(REC12 && AMINO), "12-byte records are a nucleotide format");
  static_assert(!(REC12 && VARLEN), "12-byte records hold 16 bits of letters, no room for the marker bit");
  static_assert(!(WIDE && (AMINO || REC12)), "64-bit positions: nucleotide, 16-byte records");
  using Pos = typename std::conditional<WIDE, uint64_t, uint32_t>::type;
  constexpr uint32_t kSweepTile = kSweepThreads * kSweepItems;
  constexpr uint32_t NB = SweepAlphabet<AMINO>::kCard, LB = SweepAlphabet<AMINO>::kLetterBits;
  // output buckets: one per letter; EMIT: kSweepEmitBuckets slices of the id space, laid out flat (see sweepEmit)
  constexpr uint32_t NBO = EMIT ? kSweepEmitBuckets : NB, LBO = EMIT ? kSweepEmitBits : LB;
  constexpr uint32_t kLetterMask = (1u << LB) - 1u;
  constexpr uint32_t kLineU4 = AMINO ? kAminoLineU4 : kSectorU4;
  __shared__ uint32_t warpCount[kSweepItems][kSweepThreads / 32][NBO];
  __shared__ uint32_t bucketBase[NBO];
  __shared__ uint32_t inPrefixSh[AMINO ? 33 : 1];  // amino: records before bucket b, padded with the total
  __shared__ uint16_t codeCareSh[AMINO ? 32 : 1];  // amino: kAminoCodeCare, lanes index it with different letters
  // first pass only: tile-local counting sort on the key bits the global radix sort left unordered
  __shared__ uint32_t localBins[FIRST ? 1024 : 1];
  __shared__ uint32_t localKeys[FIRST ? kSweepTile : 1];
  __shared__ uint64_t localVals[FIRST ? kSweepTile : 1];
  __shared__ uint32_t localWarpSum[kSweepThreads / 32];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lanesBelow = (1u << lane) - 1u;
  const uint32_t inLast = (uint32_t)in.cap - 1u, outLast = (uint32_t)out.cap - 1u;
  const uint4 *in0 = in.arr[0], *in1 = in.arr[1];
  uint4 *out0 = out.arr[0], *out1 = out.arr[1];

  uint32_t total, before1 = 0, before2 = 0, before3 = 0;  // records in the buckets before bucket 1, 2, 3
  if constexpr (AMINO) {
    if (threadIdx.x < 21) codeCareSh[threadIdx.x] = kAminoCodeCare[threadIdx.x];
    if (threadIdx.x == 0) {
      uint32_t run = 0;
      for (uint32_t b = 0; b < 33; b++) {
        inPrefixSh[b] = run;
        if (!FIRST && b < NB) run += in.count[b];
      }
    }
    __syncthreads();
    total = FIRST ? (uint32_t)numPairs : inPrefixSh[32];
  } else if (FIRST) {
    total = (uint32_t)numPairs;
  } else {
    const uint32_t c0 = in.count[0], c1 = in.count[1], c2 = in.count[2], c3 = in.count[3];
    before1 = c0;
    before2 = c0 + c1;
    before3 = c0 + c1 + c2;
    total = before3 + c3;  // <= number of queries < 2^31
  }

  // Tiles are handed out in order by a ticket counter (word 31 of the output generation's control block), not by
  // static striding: when other kernels hold part of the SMs (NCCL's gather of the previous step's counts at N > 1)
  // some CTAs of this grid start late, and with static striding their tiles would simply run after everyone else's.
  // The ticket of the next tile is drawn while this one is worked on; the barrier that publishes it doubles as the
  // "previous tile's append buffers consumed" barrier.
  // where record i of the input generation lives (16-byte records)
  [[maybe_unused]] auto recordPtr = [&](uint32_t i) -> const uint4 * {
    if constexpr (AMINO) {
      uint32_t b = 0;  // bucket of record i: binary search over the padded prefix counts
#pragma unroll
      for (uint32_t step = 16; step > 0; step >>= 1)
        if (i >= inPrefixSh[b + step]) b += step;
      const uint32_t r = i - inPrefixSh[b];
      return in.arr[b >> 1] + ((b & 1u) ? inLast - r : r);
    } else {
      // buckets 0/2 grow up in arrays 0/1, buckets 1/3 grow down from the end
      const bool ge1 = i >= before1, ge2 = i >= before2, ge3 = i >= before3;
      const uint32_t first = ge3 ? before3 : ge2 ? before2 : ge1 ? before1 : 0u;
      const bool odd = ge1 != ge2 || ge3;  // bucket 1 or 3
      const uint32_t r = i - first;
      return (ge2 ? in1 : in0) + (odd ? inLast - r : r);
    }
  };
  __shared__ uint32_t tileTicket[2];
  uint32_t nextTile = 0;  // thread 0
  if (threadIdx.x == 0) nextTile = atomicAdd(out.count + 31, 1u);
  for (uint32_t round = 0;; round++) {
    if (threadIdx.x == 0) tileTicket[round & 1u] = nextTile;
    __syncthreads();
    const uint32_t tile = tileTicket[round & 1u];
    if ((uint64_t)tile * kSweepTile >= total) break;
    if (threadIdx.x == 0) nextTile = atomicAdd(out.count + 31, 1u);
    const uint32_t base = tile * kSweepTile;
    Pos sp[kSweepItems], ep[kSweepItems];
    uint32_t id[kSweepItems], rest[kSweepItems], bucket[kSweepItems];
    // Every load of a stage is issued for all items before anything waits on it (no branches around the loads:
    // out-of-range items read a clamped, valid address and are disabled afterwards).
    // ---- stage A: this tile's records / (key, payload) pairs ----
    [[maybe_unused]] uint32_t key[kSweepItems];
#pragma unroll
    for (int it = 0; it < kSweepItems; it++) {
      const uint32_t want = base + it * kSweepThreads + threadIdx.x;
      const uint32_t i = min(want, total - 1u);
      if (FIRST) {
        const uint64_t v = __ldg(vals + i);
        key[it] = __ldg(keys + i);
        id[it] = (uint32_t)v;
        rest[it] = (uint32_t)(v >> 32);
        sp[it] = 1;
        ep[it] = 0;
      } else if constexpr (AMINO) {
        uint32_t b = 0;  // bucket of record i: binary search over the padded prefix counts
#pragma unroll
        for (uint32_t step = 16; step > 0; step >>= 1)
          if (i >= inPrefixSh[b + step]) b += step;
        const uint32_t r = i - inPrefixSh[b];
        const uint4 rec = __ldg(in.arr[b >> 1] + ((b & 1u) ? inLast - r : r));
        sp[it] = rec.x;
        ep[it] = rec.x + rec.y;
        id[it] = rec.z;
        rest[it] = rec.w;
      } else {
        // bucket of record i and its slot: buckets 0/2 grow up in arrays 0/1, buckets 1/3 grow down from the end
        const bool ge1 = i >= before1, ge2 = i >= before2, ge3 = i >= before3;
        const uint32_t first = ge3 ? before3 : ge2 ? before2 : ge1 ? before1 : 0u;
        const bool odd = ge1 != ge2 || ge3;  // bucket 1 or 3
        const uint32_t r = i - first;
        const uint32_t slot = odd ? inLast - r : r;
        if constexpr (REC12) {
          const uint4 *base = ge2 ? in1 : in0;
          const uint2 a = __ldg(reinterpret_cast<const uint2 *>(base) + slot);
          const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint2 *>(base) + in.cap) + slot);
          sp[it] = a.x;
          ep[it] = a.x + (w & 0xFFFFu);
          id[it] = a.y;
          rest[it] = w >> 16;
        } else {
          const uint4 rec = __ldg((ge2 ? in1 : in0) + slot);
          if constexpr (WIDE) {
            sp[it] = (uint64_t)rec.x | ((uint64_t)(rec.y & 0xFFu) << 32);
            ep[it] = sp[it] + (rec.y >> 8);
          } else {
            sp[it] = rec.x;
            ep[it] = rec.x + rec.y;
          }
          id[it] = rec.z;
          rest[it] = rec.w;
        }
      }
      if (want >= total) id[it] = kSweepNoId;
    }
    // ---- stage A1 (first pass, optional): finish the ordering inside the tile.  The radix sort ordered the pairs
    //      by the key bits above `localBits`; here the tile's pairs are counting-sorted in shared memory by
    //      (distance of those upper bits from the tile's first pair, clamped to 3) : (low localBits bits), so a
    //      warp's 32 consecutive pairs touch a handful of neighbouring lines instead of 32 scattered ones.  Order is
    //      a locality matter only — the result of a query does not depend on it. ----
    if constexpr (FIRST) {
      // (`localBits` carries two fields: bits 0-7 = how many key bits the tile orders, bits 8-15 = the lowest of them)
      if (const uint32_t lb = localBits & 0xFFu, ls = localBits >> 8; lb) {  // uniform
        uint32_t val32[kSweepItems][2], lk[kSweepItems], pos[kSweepItems];
        const uint32_t firstUpper = __ldg(keys + base) >> (ls + lb);
        for (uint32_t b = threadIdx.x; b < 1024; b += kSweepThreads) localBins[b] = 0;
        __syncthreads();
#pragma unroll
        for (int it = 0; it < kSweepItems; it++) {
          const uint32_t delta = min((key[it] >> (ls + lb)) - firstUpper, 3u);
          lk[it] = base + it * kSweepThreads + threadIdx.x >= total  // padding of the last tile goes to the end
                       ? 1023u : ((delta << lb) | ((key[it] >> ls) & ((1u << lb) - 1u)));
          pos[it] = atomicAdd(&localBins[lk[it]], 1u);
          val32[it][0] = id[it];
          val32[it][1] = rest[it];
        }
        __syncthreads();
        {  // exclusive scan of the 1024 bins: consecutive bins per thread, warp scan, warp totals
          constexpr int kPer = 1024 / kSweepThreads;
          uint32_t bin[kPer], mine = 0;
#pragma unroll
          for (int j = 0; j < kPer; j++) bin[j] = localBins[kPer * threadIdx.x + j], mine += bin[j];
          uint32_t incl = mine;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (unsigned)d) incl += up;
          }
          if (lane == 31) localWarpSum[warp] = incl;
          __syncthreads();
          uint32_t start = incl - mine;
          for (unsigned w = 0; w < warp; w++) start += localWarpSum[w];
#pragma unroll
          for (int j = 0; j < kPer; j++) localBins[kPer * threadIdx.x + j] = start, start += bin[j];
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < kSweepItems; it++) {
          const uint32_t dst = localBins[lk[it]] + pos[it];
          localKeys[dst] = key[it];
          localVals[dst] = ((uint64_t)val32[it][1] << 32) | val32[it][0];
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < kSweepItems; it++) {
          const uint32_t src = it * kSweepThreads + threadIdx.x;
          key[it] = localKeys[src];
          id[it] = (uint32_t)localVals[src];
          rest[it] = (uint32_t)(localVals[src] >> 32);
        }
      }
    }
    // ---- stage A2 (first pass): seed-table entries (any key is inside the table) ----
    if (FIRST) {
      uint64_t s64[kSweepItems], e64[kSweepItems];
#pragma unroll
      for (int it = 0; it < kSweepItems; it++) loadSeedEntry(ix, deep, key[it], s64[it], e64[it]);
#pragma unroll
      for (int it = 0; it < kSweepItems; it++) {
        if (s64[it] > e64[it]) {  // empty seed range: count stays 0, the stored pair is the query's final range
          if (ranges && !rangesOfHitsOnly && id[it] != kSweepNoId)
            ranges[id[it]] = make_uint4((uint32_t)s64[it], (uint32_t)(s64[it] >> 32), (uint32_t)e64[it], (uint32_t)(e64[it] >> 32));
          id[it] = kSweepNoId;
        }
        if ((REC12 || WIDE) && id[it] != kSweepNoId && e64[it] - s64[it] >= (WIDE ? 0xFFFFFFull : 0xFFFFull)) {  // width does not fit its field
          irregularIds[atomicAdd(irregularCount, 1u)] = id[it];
          id[it] = kSweepNoId;
        }
        if (VARLEN && id[it] != kSweepNoId && rest[it] == 1u) {  // len == k: the seed entry is the answer
          counts[id[it]] = (uint32_t)(e64[it] - s64[it] + 1ull);
          if (ranges)
            ranges[id[it]] = make_uint4((uint32_t)s64[it], (uint32_t)(s64[it] >> 32), (uint32_t)e64[it], (uint32_t)(e64[it] >> 32));
          id[it] = kSweepNoId;
        }
        if (id[it] != kSweepNoId) sp[it] = (Pos)s64[it], ep[it] = (Pos)e64[it];
      }
    }
#pragma unroll
    for (int it = 0; it < kSweepItems; it++)
      if (id[it] == kSweepNoId) sp[it] = 1, ep[it] = 0;  // disabled items rank position 0: a valid address
    // ---- stage B: pull the two sectors of every item towards L1 (no registers, no scoreboard held) ----
    // ONE prefetch per item — sp-1 and ep of a sorted query mostly share a 128-B line (or neighbouring ones), and the
    // pass is bound by the L1/LSU pipe (75 % busy, profiles/r02_ncu_sweep_kernels.json), so the second request costs
    // more than it hides: 100 M nucleotide 20-mers 6.82 ms with one, 7.10 with two, 7.10 with none; 50 M amino 8-mers
    // 3.30 with one, 3.39 with two (profiles/r02_sweep_probe.jsonl).
    if (steps > 0) {
#pragma unroll
      for (int it = 0; it < kSweepItems; it++)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(ix.lines + (uint64_t)((sp[it] - 1u) >> 6) * kLineU4));  // (Pos arithmetic)
    }
    // ---- stage C: the LF steps (src/AwFmSearch.c:42-103) ----
#pragma unroll
    for (int it = 0; it < kSweepItems; it++) {
      const uint32_t letter = AMINO ? min(rest[it] & kLetterMask, NB - 1u) : (rest[it] & kLetterMask);
      bool valid = id[it] != kSweepNoId;
      if (steps > 0) {
        Pos nsp, nep;
        if constexpr (AMINO) {
          const AminoSweepSelector sel = aminoSweepSelector(codeCareSh[letter]);
          nsp = aminoSweepRank(ix, sp[it] - 1u, letter, sel);
          nep = aminoSweepRank(ix, ep[it], letter, sel) - 1u;
        } else {
          const SweepSelector sel = sweepSelector(letter);
          // sp-1 and ep nearly always lie in the same 2^16-position superblock: its row is read once
          const Pos pa = sp[it] - 1u, pb = ep[it];
          const uint64_t rowA = (uint64_t)(pa >> kSectorSuperShift), rowB = (uint64_t)(pb >> kSectorSuperShift);
          const Pos superA = sweepSuper<Pos>(ix, rowA, letter);
#ifdef AWFM_SWEEP_TWO_SUPER
          const Pos superB = sweepSuper<Pos>(ix, rowB, letter);
#else
          const Pos superB = rowB == rowA ? superA : sweepSuper<Pos>(ix, rowB, letter);
#endif
#ifndef AWFM_SWEEP_SHARED_SECTOR
          nsp = sweepRank<Pos>(ix, pa, letter, sel, superA);
          nep = sweepRank<Pos>(ix, pb, letter, sel, superB) - 1u;
#else
          // After the first steps sp-1 and ep nearly always lie in the same 64-position sector; here it is requested
          // once (the second request predicated off, nothing branched around).  Measured: NO gain — 0.933 / 0.937 /
          // 0.867 ms against 0.911 / 0.914 / 0.849 for passes 2-4 of 100 M 20-mers — the passes are not bound by the
          // wavefronts of the sector requests; kept for the record, off.
          const bool apart = (pb >> 6) != (pa >> 6);
          const SweepSector secA = sweepLoadSector(ix.lines + (uint64_t)(pa >> 6) * kSectorU4);
          const SweepSector secB = sweepLoadSectorOr(ix.lines + (uint64_t)(pb >> 6) * kSectorU4, apart, secA);
          nsp = sweepRankIn<Pos>(secA, pa, letter, sel, superA);
          nep = sweepRankIn<Pos>(secB, pb, letter, sel, superB) - 1u;
#endif
        }
        sp[it] = nsp;
        ep[it] = nep;
        rest[it] >>= LB;
        if (valid && nep == nsp - 1u) {  // ep == sp - 1 <=> empty: the search stops here (src/AwFmParallelSearch.c:279-311)
          valid = false;
          if (ranges && !rangesOfHitsOnly) {
            if constexpr (WIDE) ranges[id[it]] = make_uint4((uint32_t)nsp, (uint32_t)(nsp >> 32), (uint32_t)nep, (uint32_t)(nep >> 32));
            else ranges[id[it]] = make_uint4(nsp, 0u, nep, nep == 0xFFFFFFFFu ? 0xFFFFFFFFu : 0u);
          }
        }
      }
      bucket[it] = NBO;  // no output
      if (valid) {
        if constexpr (EMIT) {  // (launched for the last pass only) count and range leave through sweepEmit
          bucket[it] = min(id[it] / emitDiv, NBO - 1u);
        } else {
          const bool last = VARLEN ? (steps <= 1 || rest[it] == 1u) : steps <= 1;
          if (last && ranges) ranges[id[it]] = make_uint4((uint32_t)sp[it], (uint32_t)((uint64_t)sp[it] >> 32), (uint32_t)ep[it],
                                                          (uint32_t)((uint64_t)ep[it] >> 32));
          if (last) counts[id[it]] = (uint32_t)(ep[it] - sp[it] + 1u);
          else bucket[it] = letter;  // grouped by the letter just prepended: sp' = C[c] + Occ(c, sp-1) keeps the order
        }
      }
    }
    if (!EMIT && steps <= 1) continue;  // last pass: nothing to append (uniform for the whole grid)
    // ---- stable (inside the tile) append to the output buckets ----
    uint32_t rank[kSweepItems];  // (warpCount / bucketBase of the previous tile were consumed before this tile's ticket barrier)
#pragma unroll
    for (int it = 0; it < kSweepItems; it++) {
      if constexpr (AMINO || EMIT) {
        const uint32_t b = bucket[it];
        unsigned mine = __ballot_sync(0xFFFFFFFFu, b < NBO), forLane = mine;  // lanes in my bucket / in bucket `lane`
#pragma unroll
        for (uint32_t bit = 0; bit < LBO; bit++) {
          const unsigned m = __ballot_sync(0xFFFFFFFFu, (b >> bit) & 1u);
          mine &= ((b >> bit) & 1u) ? m : ~m;
          forLane &= ((lane >> bit) & 1u) ? m : ~m;
        }
        rank[it] = __popc(mine & lanesBelow);
        if (lane < NBO) warpCount[it][warp][lane] = __popc(forLane);
      } else {
        const unsigned live = __ballot_sync(0xFFFFFFFFu, bucket[it] < 4u);
        const unsigned bit0 = __ballot_sync(0xFFFFFFFFu, bucket[it] & 1u), bit1 = __ballot_sync(0xFFFFFFFFu, bucket[it] & 2u);
        const unsigned mine = live & ((bucket[it] & 1u) ? bit0 : ~bit0) & ((bucket[it] & 2u) ? bit1 : ~bit1);
        rank[it] = __popc(mine & lanesBelow);
        if (lane < 4) warpCount[it][warp][lane] = __popc(live & ((lane & 1u) ? bit0 : ~bit0) & ((lane & 2u) ? bit1 : ~bit1));
      }
    }
#if AWFM_SWEEP_NEXT_PREFETCH
    if (threadIdx.x == 0) tileTicket[(round + 1u) & 1u] = nextTile;  // (drawn at the top of this round: long since back)
#endif
    __syncthreads();
#if AWFM_SWEEP_NEXT_PREFETCH
    // The next tile's records (pairs) are pulled towards the SM while this tile waits for its room in the buckets and
    // writes its records out: one DRAM round trip less on the next tile's critical path.
    if constexpr (!REC12) {
      const uint32_t nextBase = tileTicket[(round + 1u) & 1u] * kSweepTile;
      if ((uint64_t)tileTicket[(round + 1u) & 1u] * kSweepTile < total) {
#pragma unroll
        for (int it = 0; it < kSweepItems; it++) {
          const uint32_t i = min(nextBase + it * kSweepThreads + threadIdx.x, total - 1u);
          if (FIRST) {
            sweepPrefetchNext(vals + i);
            if ((threadIdx.x & 1u) == 0) sweepPrefetchNext(keys + i);
          } else {
            sweepPrefetchNext(recordPtr(i));
          }
        }
      }
    }
#endif
    if (threadIdx.x < NBO) {
      uint32_t run = 0;
#pragma unroll
      for (int it = 0; it < kSweepItems; it++)
#pragma unroll
        for (int w = 0; w < kSweepThreads / 32; w++) {
          const uint32_t t = warpCount[it][w][threadIdx.x];
          warpCount[it][w][threadIdx.x] = run;
          run += t;
        }
      bucketBase[threadIdx.x] = run ? atomicAdd(out.count + threadIdx.x, run) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kSweepItems; it++) {
      const uint32_t b = bucket[it];
      if (b < NBO) {
        const uint32_t r = bucketBase[b] + warpCount[it][warp][b] + rank[it];
        // (EMIT: slice b holds ids [b * emitDiv, (b + 1) * emitDiv), so at most emitDiv records: slices lie flat in array 0)
        uint4 *dst = EMIT ? out0 : AMINO ? out.arr[b >> 1] : ((b & 2u) ? out1 : out0);
        const uint32_t slot = EMIT ? b * emitDiv + r : (b & 1u) ? outLast - r : r;
        if constexpr (REC12) {
          reinterpret_cast<uint2 *>(dst)[slot] = make_uint2((uint32_t)sp[it], id[it]);
          reinterpret_cast<uint32_t *>(reinterpret_cast<uint2 *>(dst) + out.cap)[slot] = (uint32_t)(ep[it] - sp[it]) | (rest[it] << 16);
        } else if constexpr (WIDE) {
          dst[slot] = make_uint4((uint32_t)sp[it], (uint32_t)(sp[it] >> 32) | ((uint32_t)(ep[it] - sp[it]) << 8), id[it], rest[it]);
        } else {
          dst[slot] = make_uint4(sp[it], ep[it] - sp[it], id[it], rest[it]);
        }
      }
    }
  }
}

// sweepRefill (fixed-length nucleotide batches with more than 16 letters left of the seed k-mer): the records of
// generation `gen` have prepended the 16 letters they carried; each takes the query's next letters from more[id]
// (written by the pack kernel) into its letters word.  16-byte records, plain or WIDE: id = word 2, letters = word 3.
static __global__ void __launch_bounds__(256) sweepRefill(const __grid_constant__ SweepRecs gen, const uint32_t *__restrict__ more) {
  const uint32_t c0 = gen.count[0], c1 = gen.count[1], c2 = gen.count[2], c3 = gen.count[3];
  const uint32_t before1 = c0, before2 = c0 + c1, before3 = before2 + c2, total = before3 + c3;
  const uint32_t last = (uint32_t)gen.cap - 1u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const bool ge1 = i >= before1, ge2 = i >= before2, ge3 = i >= before3;  // buckets 0/2 grow up, 1/3 down (sweepStep)
    const uint32_t first = ge3 ? before3 : ge2 ? before2 : ge1 ? before1 : 0u;
    const bool odd = ge1 != ge2 || ge3;
    const uint32_t r = i - first;
    uint32_t *rec = reinterpret_cast<uint32_t *>(gen.arr[ge2 ? 1 : 0] + (odd ? last - r : r));
    rec[3] = __ldg(more + rec[2]);
  }
}

// sweepEmit: the survivors of an EMIT last pass, in kSweepEmitBuckets slices of the id space (slice b: gen.count[b]
// records from slot b * emitDiv of array 0), leave for counts[id] and ranges[id].  The grid walks the slices in order (a
// window of gridDim.x * 256 consecutive slots at any moment), so the stores of one moment fall into one slice of the two
// output arrays.
template <bool WIDE>
static __global__ void __launch_bounds__(256) sweepEmit(const __grid_constant__ SweepRecs gen, uint32_t emitDiv,
                                                        uint32_t *__restrict__ counts, uint4 *__restrict__ ranges) {
  __shared__ uint32_t sliceCount[kSweepEmitBuckets];
  if (threadIdx.x < kSweepEmitBuckets) sliceCount[threadIdx.x] = gen.count[threadIdx.x];
  __syncthreads();
  const uint32_t slots = kSweepEmitBuckets * emitDiv;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += gridDim.x * blockDim.x) {
    const uint32_t b = i / emitDiv;
    if (i - b * emitDiv >= sliceCount[b]) continue;
    const uint4 rec = __ldg(gen.arr[0] + i);
    uint64_t sp, width;
    if constexpr (WIDE) sp = (uint64_t)rec.x | ((uint64_t)(rec.y & 0xFFu) << 32), width = rec.y >> 8;
    else sp = rec.x, width = rec.y;
    const uint64_t ep = WIDE ? sp + width : (uint64_t)(uint32_t)(rec.x + rec.y);
    counts[rec.z] = (uint32_t)width + 1u;
    ranges[rec.z] = make_uint4((uint32_t)sp, (uint32_t)(sp >> 32), (uint32_t)ep, (uint32_t)(ep >> 32));
  }
}

// Queries the sweep does not take (ambiguity letters, '$', anything not A/C/G/T/U): the reference's own order of
// business for one query (countKernelV0's body), one thread per listed id.
template <bool AMINO>
__global__ void __launch_bounds__(256)
    sweepIrregular(const __grid_constant__ DevIndex ix, const uint8_t *__restrict__ letters,
                   const uint64_t *__restrict__ offsets /* or nullptr: fixed length */, uint32_t fixedLen,
                   const uint32_t *__restrict__ ids, const uint32_t *__restrict__ numIds,
                   uint32_t *__restrict__ counts, uint4 *__restrict__ ranges, bool rangesOfHitsOnly) {
  const uint32_t n = *numIds;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t q = ids[i];
    const uint64_t off = offsets ? __ldg(offsets + q) : (uint64_t)q * fixedLen;
    const uint64_t len = offsets ? __ldg(offsets + q + 1) - off : (uint64_t)fixedLen;
    const uint8_t *s = letters + off;
    uint64_t sp, ep;
    uint64_t next = openRange<AMINO>(ix, s, len, sp, ep);
    while (next > 0 && sp <= ep) {
      const uint32_t letter = letterIndex<AMINO>(__ldg(s + next - 1));
      if (letter > SweepAlphabet<AMINO>::kCard) {
        sp = 1;
        ep = 0;
        break;
      }
      lfStep<1, AMINO>(ix, sp, ep, letter, 0u, 0xFFFFFFFFu);
      next--;
    }
    counts[q] = (uint32_t)(sp <= ep ? ep - sp + 1 : 0);
    if (ranges && (sp <= ep || !rangesOfHitsOnly))
      ranges[q] = make_uint4((uint32_t)sp, (uint32_t)(sp >> 32), (uint32_t)ep, (uint32_t)(ep >> 32));
  }
}

// The same for the 2-bit packed format (AWFM_QUERY_2BIT): only reached by queries whose seed range is too wide for the
// 12-byte records.  The query's letters are unpacked into letter indices and searched like any other query.
static __global__ void __launch_bounds__(256)
    sweepIrregularBits(const __grid_constant__ DevIndex ix, const uint8_t *__restrict__ packed, uint32_t len,
                       const uint32_t *__restrict__ ids, const uint32_t *__restrict__ numIds,
                       uint32_t *__restrict__ counts, uint4 *__restrict__ ranges, bool rangesOfHitsOnly) {
  const uint32_t n = *numIds, B = (len + 3u) >> 2;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t q = ids[i];
    const uint8_t *src = packed + (uint64_t)q * B;
    uint8_t letters[32];
    for (uint32_t j = 0; j < len && j < 32u; j++) letters[j] = (__ldg(src + (j >> 2)) >> (2u * (j & 3u))) & 3u;
    uint64_t sp, ep;
    uint64_t next = openRange<false, true>(ix, letters, len, sp, ep);
    while (next > 0 && sp <= ep) {
      lfStep<1, false>(ix, sp, ep, letters[next - 1], 0u, 0xFFFFFFFFu);
      next--;
    }
    counts[q] = (uint32_t)(sp <= ep ? ep - sp + 1 : 0);
    if (ranges && (sp <= ep || !rangesOfHitsOnly))
      ranges[q] = make_uint4((uint32_t)sp, (uint32_t)(sp >> 32), (uint32_t)ep, (uint32_t)(ep >> 32));
  }
}

}  // namespace awfm
