// awfm_sweep.cuh — the "sweep" count path for LARGE batches of fixed-length nucleotide queries (sm_100a).
//
// The tile kernels of awfm_kernels.cuh pay one random DRAM line per rank: 1 + 5.64 lines per 20-mer at 3.1 Gbp, and
// the memory system delivers ~42 G random lines/s whatever their size (profiles/r01_granularity_probe.jsonl), so they
// stop at 6.4 G queries/s with the DRAM pipe 65 % busy.  A batch of 100 M queries, however, puts ~8 queries on every
// 128-B line of the index at every step — the misses are only random because the queries are processed in input order.
// Backward search keeps order: if the live queries are ordered by the start of their current range, the positions
// the next LF step ranks (sp-1, ep) are non-decreasing, and after the step the queries that prepended the same letter
// c are still ordered (sp' = C[c] + Occ(c, sp-1) is monotone in sp).  So:
//
//   sweepPack    thread per query: seed-table index of the last k letters (src/AwFmKmerTable.c:21-51) as the sort key,
//                the remaining len-k letters packed 2 bits each (next letter to prepend in the low bits) + query id as
//                the payload.  Queries holding anything but A/C/G/T(/U) go to a side list (sweepIrregular).
//   radix sort   CUB, only the top `sortBits` bits of the key: queries whose seed entries / first ranges share a few
//                KB of the table / of the BWT become neighbours; exact order inside is irrelevant for correctness.
//   sweepStep<FIRST>   reads the seed entry (src/AwFmParallelSearch.c:222-271), does the first LF step
//                (src/AwFmSearch.c:42-103) and appends the still-valid query as a 16-B record {sp, ep-sp, id, letters}
//                to the bucket of the letter it has just prepended;
//   sweepStep<!FIRST>  one pass per further letter over the buckets in letter order A,C,G,T — which is ascending sp
//                order, because ranges of strings starting with A precede those starting with C, ... and inside a
//                bucket the append order is the (ascending) input order — one LF step per record, same append.  A query whose range empties is dropped (its count
//                stays at the zero the output was cleared to, src/AwFmParallelSearch.c:279-311 stops there too); one
//                that has prepended all its letters writes count = ep-sp+1 (u32, :187-190).
//
// Buckets are filled tile by tile through one atomicAdd per bucket per tile (order inside a tile is kept, tiles land
// in roughly launch order), so the order is approximate: the ranks of the tiles in flight at any moment touch a
// window of ~1 % of the BWT, which the 126 MB L2 holds.  The index is then STREAMED once per pass instead of being
// gathered line by line, and all other traffic (keys, records) is sequential.  No spin-waits anywhere.
//
// Exactness: same seed entries, same ranks (sectorRank), same stop rule; only the processing order differs, and the
// result of a query does not depend on it.  Not covered here (the caller falls back to the tile kernels): amino
// indexes, variable-length batches, range output, bwtLength > 2^32, len - k > 16, k > 16.
#pragma once
#include "awfm_kernels.cuh"

namespace awfm {

constexpr int kSweepThreads = 256, kSweepItems = 2, kSweepTile = kSweepThreads * kSweepItems;
constexpr uint32_t kSweepNoId = 0xFFFFFFFFu;
constexpr int kSweepMaxPasses = 18;

// One generation of live records: bucket b (= letter prepended last) lives in arr[b >> 1]; even buckets grow up
// from slot 0, odd ones down from slot cap-1, so four buckets of unknown sizes share 2 x cap slots (total <= cap).
struct SweepRecs {
  uint4 *arr[2];
  uint32_t *count;  // [4] device counters of this generation
  uint64_t cap;
};
__device__ __forceinline__ uint64_t sweepSlot(uint64_t cap, uint32_t bucket, uint32_t r) {
  return (bucket & 1u) ? cap - 1 - (uint64_t)r : (uint64_t)r;
}

// whole-line loads (no .L2::64B hint): the neighbours of this record want the rest of the line
__device__ __forceinline__ NucSector sectorIssueFull(const DevIndex &ix, uint64_t p) {
  const uint4 *s = ix.lines + (p >> 6) * kSectorU4;
  NucSector x;
  x.v0 = __ldg(s);
  x.v1 = __ldg(s + 1);
  return x;
}
// LF step on letters 0..3 with 32-bit positions (bwtLength <= 2^32)
__device__ __forceinline__ void lfStepSweep(const DevIndex &ix, uint32_t &sp, uint32_t &ep, uint32_t letter) {
  const uint64_t pa = (uint64_t)sp - 1, pb = ep;
  const NucSector a = sectorIssueFull(ix, pa), b = sectorIssueFull(ix, pb);
  const uint64_t ca = __ldg(ix.superC + (pa >> kSectorSuperShift) * kSectorSuperStride + letter);
  const uint64_t cb = __ldg(ix.superC + (pb >> kSectorSuperShift) * kSectorSuperStride + letter);
  const uint64_t nsp = ca + sectorCount(a, letter) + sectorPop(a, letter, (uint32_t)pa & 63u);
  const uint64_t nep = cb + sectorCount(b, letter) + sectorPop(b, letter, (uint32_t)pb & 63u) - 1;
  sp = (uint32_t)nsp;  // nsp <= bwtLength - 1 + 1; an empty result has nep == nsp - 1, kept as width 0xFFFFFFFF below
  ep = (uint32_t)nep;
}

// ---------------------------------------------------------------------------------------------------------------
// sweepPackWords: len % 4 == 0 and `letters` 16-B aligned.  One thread per query reads its len/4 words straight from
// global memory (a warp covers 32*len contiguous bytes; the repeats are L1 hits) and translates four letters at a time:
//   code  = ((w >> 1) ^ (w >> 2)) & 3 per byte maps A,C,G,T/U (either case) to 0,1,2,3 (src/AwFmLetter.c:4-22);
//   valid = the byte, lower-cased, equals "acgt"[code] (or 'u' for code 3) — anything else is an irregular query.
// The whole query becomes one number Q, two bits per letter, first letter most significant: the seed-table index of
// the last k letters (src/AwFmKmerTable.c:21-51) is Q mod 4^k, and Q >> 2k has the letter prepended next in its low
// bits.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t packFourLetters(uint32_t w, uint32_t &bad) {
  const uint32_t code = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
  const uint32_t b0 = code & 0x01010101u, b1 = (code >> 1) & 0x01010101u, both = b0 & b1;
  const uint32_t expected = 0x61616161u + b0 * 2u + b1 * 6u + both * 11u;  // 'a', 'c' = +2, 'g' = +6, 't' = +19
  bad |= ((w | 0x20202020u) ^ expected) & ~both;                           // 't' ^ 'u' == 1, allowed where code == 3
  uint32_t t = __byte_perm(code, 0, 0x0123);                               // first letter into the top byte
  t = (t | (t >> 6)) & 0x000F000Fu;
  return (t | (t >> 12)) & 0xFFu;                                          // first letter in bits 7..6
}
__global__ void __launch_bounds__(256)
    sweepPackWords(const uint32_t *__restrict__ words, uint64_t numQueries, uint32_t wordsPerQuery, uint32_t k,
                   uint32_t *__restrict__ keys, uint64_t *__restrict__ vals, uint32_t *__restrict__ irregularIds,
                   uint32_t *__restrict__ irregularCount) {
  const uint64_t keyMask = (k >= 16) ? 0xFFFFFFFFull : ((1ull << (2 * k)) - 1ull);
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < numQueries;
       q += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t *src = words + q * wordsPerQuery;
    uint64_t Q = 0;
    uint32_t bad = 0;
    for (uint32_t i = 0; i < wordsPerQuery; i++) Q = (Q << 8) | packFourLetters(__ldg(src + i), bad);
    uint32_t id = (uint32_t)q;
    if (bad) {
      irregularIds[atomicAdd(irregularCount, 1u)] = id;
      id = kSweepNoId;
    }
    keys[q] = (uint32_t)(Q & keyMask);
    vals[q] = ((Q >> (2 * k)) << 32) | id;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// sweepPack: 256 queries per tile staged through shared memory with coalesced 128-bit loads.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    sweepPack(const uint8_t *__restrict__ letters, uint64_t numQueries, uint32_t len, uint32_t k,
              uint32_t *__restrict__ keys, uint64_t *__restrict__ vals, uint32_t *__restrict__ irregularIds,
              uint32_t *__restrict__ irregularCount) {
  extern __shared__ __align__(16) uint8_t sLetters[];  // 256 * len bytes, rounded up to 16
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
      uint32_t key = 0, packed = 0, bad = 0;
      for (uint32_t i = 0; i < k; i++) {  // leftmost of the last k letters most significant
        const uint32_t l = nucLetterIndex(s[rest + i]);
        bad |= l >> 2;
        key = (key << 2) | (l & 3u);
      }
      for (uint32_t j = 0; j < rest; j++) {  // letter prepended at step j+1 is s[rest-1-j]: low bits first
        const uint32_t l = nucLetterIndex(s[rest - 1 - j]);
        bad |= l >> 2;
        packed |= (l & 3u) << (2 * j);
      }
      uint32_t id = (uint32_t)(q0 + threadIdx.x);
      if (bad) {
        irregularIds[atomicAdd(irregularCount, 1u)] = id;
        id = kSweepNoId;
      }
      keys[q0 + threadIdx.x] = key;
      vals[q0 + threadIdx.x] = ((uint64_t)packed << 32) | id;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// sweepStep: one pass.  FIRST: input = sorted (key, payload) pairs, range from the seed table; else input = the
// previous generation's buckets in letter order.  `steps` = LF steps the queries of this pass still have to do
// INCLUDING this pass's (0 only for FIRST with len == k).
// ---------------------------------------------------------------------------------------------------------------
template <bool FIRST>
__global__ void __launch_bounds__(kSweepThreads)
    sweepStep(const __grid_constant__ DevIndex ix, const uint32_t *__restrict__ keys, const uint64_t *__restrict__ vals,
              uint64_t numPairs, bool deep, const __grid_constant__ SweepRecs in, const __grid_constant__ SweepRecs out,
              uint32_t steps, uint32_t *__restrict__ counts) {
  __shared__ uint32_t warpCount[kSweepItems][kSweepThreads / 32][4];
  __shared__ uint32_t bucketBase[4];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lanesBelow = (1u << lane) - 1u;

  uint64_t total;
  uint32_t before1 = 0, before2 = 0, before3 = 0;  // records in the buckets before bucket 1, 2, 3
  if (FIRST) {
    total = numPairs;
  } else {
    const uint32_t c0 = in.count[0], c1 = in.count[1], c2 = in.count[2], c3 = in.count[3];
    before1 = c0;
    before2 = c0 + c1;
    before3 = c0 + c1 + c2;
    total = (uint64_t)c0 + c1 + c2 + c3;
  }

  for (uint64_t base = (uint64_t)blockIdx.x * kSweepTile; base < total; base += (uint64_t)gridDim.x * kSweepTile) {
    uint32_t sp[kSweepItems], ep[kSweepItems], id[kSweepItems], rest[kSweepItems], bucket[kSweepItems],
        rank[kSweepItems];
#pragma unroll
    for (int it = 0; it < kSweepItems; it++) {
      const uint64_t i = base + (uint64_t)it * kSweepThreads + threadIdx.x;
      bucket[it] = 4;  // no output
      sp[it] = 1;
      ep[it] = 0;
      id[it] = kSweepNoId;
      rest[it] = 0;
      if (i < total) {
        if (FIRST) {
          const uint64_t v = __ldg(vals + i);
          id[it] = (uint32_t)v;
          rest[it] = (uint32_t)(v >> 32);
          if (id[it] != kSweepNoId) {
            uint64_t s64, e64;
            loadSeedEntry(ix, deep, __ldg(keys + i), s64, e64);
            sp[it] = (uint32_t)s64;
            ep[it] = (uint32_t)e64;
            if (s64 > e64) id[it] = kSweepNoId;  // empty seed range: count stays 0
          }
        } else {
          uint32_t b = 0, first = 0;
          if (i >= before1) b = 1, first = before1;
          if (i >= before2) b = 2, first = before2;
          if (i >= before3) b = 3, first = before3;
          const uint4 r = in.arr[b >> 1][sweepSlot(in.cap, b, (uint32_t)i - first)];
          sp[it] = r.x;
          ep[it] = r.x + r.y;
          id[it] = r.z;
          rest[it] = r.w;
        }
      }
    }
#pragma unroll
    for (int it = 0; it < kSweepItems; it++) {
      if (id[it] != kSweepNoId) {
        bool valid = true;
        const uint32_t letter = rest[it] & 3u;
        if (steps > 0) {
          lfStepSweep(ix, sp[it], ep[it], letter);
          rest[it] >>= 2;
          valid = (ep[it] - sp[it]) != 0xFFFFFFFFu;  // ep == sp - 1 <=> empty (bwtLength < 2^32 - 16)
        }
        if (valid) {
          if (steps <= 1) counts[id[it]] = ep[it] - sp[it] + 1u;
          else bucket[it] = letter;  // grouped by the letter just prepended: sp' = C[c] + Occ(c, sp-1) keeps the order
        }
      }
    }
    // ---- stable (inside the tile) append to the four output buckets ----
    __syncthreads();  // warpCount / bucketBase of the previous tile consumed
#pragma unroll
    for (int it = 0; it < kSweepItems; it++) {
      rank[it] = 0;
#pragma unroll
      for (uint32_t b = 0; b < 4; b++) {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, bucket[it] == b);
        if (bucket[it] == b) rank[it] = __popc(m & lanesBelow);
        if (lane == 0) warpCount[it][warp][b] = __popc(m);
      }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      uint32_t run = 0;
#pragma unroll
      for (int it = 0; it < kSweepItems; it++)
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
      if (b < 4) {
        const uint32_t r = bucketBase[b] + warpCount[it][warp][b] + rank[it];
        out.arr[b >> 1][sweepSlot(out.cap, b, r)] = make_uint4(sp[it], ep[it] - sp[it], id[it], rest[it]);
      }
    }
  }
}

// Queries the sweep does not take (ambiguity letters, '$', anything not A/C/G/T/U): the reference's own order of
// business for one query (countKernelV0's body), one thread per listed id.
__global__ void __launch_bounds__(256)
    sweepIrregular(const __grid_constant__ DevIndex ix, const uint8_t *__restrict__ letters, uint32_t len,
                   const uint32_t *__restrict__ ids, const uint32_t *__restrict__ numIds,
                   uint32_t *__restrict__ counts) {
  const uint32_t n = *numIds;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t q = ids[i];
    const uint8_t *s = letters + (uint64_t)q * len;
    uint64_t sp, ep;
    uint64_t next = openRange<false>(ix, s, len, sp, ep);
    while (next > 0 && sp <= ep) {
      const uint32_t letter = nucLetterIndex(__ldg(s + next - 1));
      if (letter > 4u) {
        sp = 1;
        ep = 0;
        break;
      }
      lfStep<1, false>(ix, sp, ep, letter, 0u, 0xFFFFFFFFu);
      next--;
    }
    counts[q] = (uint32_t)(sp <= ep ? ep - sp + 1 : 0);
  }
}

}  // namespace awfm
