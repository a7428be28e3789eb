// awfm_device.cuh — device-side data layout and rank/LF/backtrace primitives (sm_100a).
//
// HBM layout (built once at upload by relayout kernels in awfm_b200.cu from the unchanged reference blocks):
//
//  nucleotide "line"  = 128 B, 128-B aligned, one per 256 BWT positions:
//      8 chunks of 16 B; chunk j = { b0[j], b1[j], b2[j], cnt[j] }   (32-bit words)
//      b_i[j]  = word j of letter bit-vector i  -> bit t of word j is block position 32*j + t
//      cnt[2c], cnt[2c+1] = low / high 32 bits of baseOccurrences[c], c = A,C,G,T
//      baseOccurrences[X] lives in a side array (xBase), the sentinel count is never needed by search.
//    One rank = exactly one 128-B line = 4 sectors; lane j of an 8-lane group issues ONE 128-bit load and
//    owns all three code bits of positions 32j..32j+31 (no bit-gather shuffles).
//    Reference layout being replaced: struct AwFmNucleotideBlock, 160 B, 32-B aligned (src/AwFmIndex.h:61-65).
//
//  amino "line triple" = 384 B, 128-B aligned:
//      [  0,128)  8 chunks { b0[j], b1[j], b2[j], b3[j] }
//      [128,160)  b4[0..7]
//      [160,328)  baseOccurrences[0..20] as u64 (A..Y, Z)
//      [328,384)  padding
//    Reference layout: struct AwFmAminoBlock, 352 B (src/AwFmIndex.h:55-59).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace awfm {

constexpr int kNucLineU4 = 8;     // uint4 per nucleotide line
constexpr int kAminoLineU4 = 24;  // uint4 per amino line triple
constexpr uint32_t kNucSentinel = 5, kAminoSentinel = 21;

struct DevIndex {
  const uint4 *lines;
  const uint64_t *xBase;     // nucleotide only
  const uint4 *seedTable;    // {startLo, startHi, endLo, endHi}
  const uint64_t *sa;        // bit-packed sampled SA viewed as little-endian u64 words (+16 B zero padding)
  uint64_t numBlocks, bwtLength, numSeeds;
  uint64_t prefixSums[24];
  uint32_t saBitWidth, saRatio, saRatioShift /* log2 if power of two else 0xFFFFFFFF */, seedK, amino;
};

// ---- letters (src/AwFmLetter.c:4-22, 55-67) ----
__device__ __forceinline__ uint32_t nucLetterIndex(uint32_t ascii) {
  const uint32_t c = ascii | 0x20u;
  uint32_t r = 4;
  r = (c == 'a') ? 0u : r;
  r = (c == 'c') ? 1u : r;
  r = (c == 'g') ? 2u : r;
  r = (c == 't' || c == 'u') ? 3u : r;
  r = (c == '$') ? 5u : r;
  return r;
}

// table[ascii & 31] of src/AwFmLetter.c:59-61 packed one byte per entry into four u64 (little-endian)
__device__ __forceinline__ uint32_t aminoLetterIndex(uint32_t ascii) {
  if (ascii == '$') return kAminoSentinel;
  const uint32_t i = ascii & 31u;
  const uint64_t t0 = 0x0504030201140014ull;  // entries 0..7  : 20,0,20,1,2,3,4,5
  const uint64_t t1 = 0x140b0a0908140706ull;  // entries 8..15 : 6,7,20,8,9,10,11,20
  const uint64_t t2 = 0x121114100f0e0d0cull;  // entries 16..23: 12,13,14,15,16,20,17,18
  const uint64_t t3 = 0x1414141414141314ull;  // entries 24..31: 20,19,20,20,20,20,20,20
  const uint64_t lo = (i & 8u) ? t1 : t0, hi = (i & 8u) ? t3 : t2;
  const uint64_t t = (i & 16u) ? hi : lo;
  return (uint32_t)(t >> ((i & 7u) * 8u)) & 0xFFu;
}

template <bool AMINO>
__device__ __forceinline__ uint32_t letterIndex(uint32_t ascii) {
  return AMINO ? aminoLetterIndex(ascii) : nucLetterIndex(ascii);
}

// ---- selectors: (code, care) per letter; a position matches when (stored ^ code) & care == 0.
// Derived from the boolean forms of src/AwFmOccurrence.c:18-35 (nucleotide) and :65-134 (amino). ----
__device__ __forceinline__ uint32_t nucCodeCare(uint32_t letter) {
  // nibble pairs, low nibble = code, high nibble = care: A(6,6) C(5,5) G(3,3) T(1,7) X(2,7)
  return (0x7271335566ull >> (letter * 8u)) & 0xFFu;
}
static __constant__ uint16_t kAminoCodeCare[21] = {
    // (care << 8) | code
    0x1C0C, 0x0F17, 0x1303, 0x1606, 0x0F1E, 0x151A, 0x0F1B, 0x1619, 0x1A15, 0x131C, 0x0F1D,
    0x0F08, 0x1909, 0x0F04, 0x1C13, 0x1A0A, 0x1505, 0x1916, 0x0F01, 0x0F02, 0x0F1F};
// code -> letter index (src/AwFmLetter.c:49-53, :89-96)
__device__ __forceinline__ uint32_t nucCodeToLetter(uint32_t code) { return (0x00152435u >> (code * 4u)) & 0xFu; }
static __constant__ uint8_t kAminoCodeToLetter[32] = {21, 18, 19, 2,  13, 16, 3,  20, 11, 12, 15, 20, 0, 20, 20, 20,
                                               20, 20, 20, 14, 20, 8,  17, 1,  20, 7,  5,  6,  9, 10, 4,  20};

// mask of bits 0..rel inclusive of a 32-bit word whose first bit is block position 32*chunk, for an inclusive
// block-local query position `local` (AwFmMaskedVectorPopcount semantics, src/AwFmSimdConfig.c:89-114)
__device__ __forceinline__ uint32_t inclusiveMask(uint32_t local, uint32_t chunk) {
  const int rel = (int)local - (int)(chunk * 32u);
  return rel >= 31 ? 0xFFFFFFFFu : (rel < 0 ? 0u : ((2u << rel) - 1u));
}

__device__ __forceinline__ uint4 ldLine(const uint4 *p) { return __ldg(p); }

template <int LPQ>
__device__ __forceinline__ unsigned groupMaskOf() {
  if (LPQ == 32) return 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u;
  return ((1u << LPQ) - 1u) << (lane / LPQ * LPQ);
}

template <int LPQ>
__device__ __forceinline__ uint64_t groupSum(uint64_t v, unsigned mask) {
#pragma unroll
  for (int d = LPQ / 2; d > 0; d >>= 1) v += __shfl_xor_sync(mask, v, d, LPQ);
  return v;
}

// Selector masks for one LF/backtrace step, expanded to 32-bit lanes: x_i = b_i ^ flip_i, y_i = x_i | dontcare_i
struct Selector {
  uint32_t flip[5], dontcare[5];
};
template <bool AMINO>
__device__ __forceinline__ Selector makeSelector(uint32_t letter) {
  Selector s;
  uint32_t code, care;
  if (AMINO) {
    const uint32_t cc = kAminoCodeCare[letter];
    code = cc & 0xFFu;
    care = cc >> 8;
  } else {
    const uint32_t cc = nucCodeCare(letter);
    code = cc & 0xFu;
    care = cc >> 4;
  }
#pragma unroll
  for (int i = 0; i < (AMINO ? 5 : 3); i++) {
    s.flip[i] = ((code >> i) & 1u) - 1u;      // code bit 1 -> 0 (keep), 0 -> ~0 (invert)
    s.dontcare[i] = ((care >> i) & 1u) - 1u;  // cared -> 0, ignored -> ~0
  }
  return s;
}

// ---- rank of `letter` at inclusive position `pos`, cooperative over an LPQ-lane group ----
// Every lane returns baseOccurrences[letter] + popcount(select(letter) & bits 0..pos%256).
template <int LPQ>
struct NucLoad {
  uint4 v[8 / LPQ];
};
template <int LPQ>
__device__ __forceinline__ NucLoad<LPQ> nucIssue(const DevIndex &ix, uint64_t block, unsigned sub) {
  NucLoad<LPQ> l;
  const uint4 *line = ix.lines + block * kNucLineU4;
#pragma unroll
  for (int i = 0; i < 8 / LPQ; i++) l.v[i] = ldLine(line + sub + LPQ * i);
  return l;
}
template <int LPQ>
__device__ __forceinline__ uint64_t nucPartial(const NucLoad<LPQ> &l, const Selector &s, uint32_t letter,
                                               uint32_t local, unsigned sub) {
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8 / LPQ; i++) {
    const uint32_t chunk = sub + LPQ * i;
    const uint4 v = l.v[i];
    const uint32_t sel = ((v.x ^ s.flip[0]) | s.dontcare[0]) & ((v.y ^ s.flip[1]) | s.dontcare[1]) &
                         ((v.z ^ s.flip[2]) | s.dontcare[2]);
    acc += __popc(sel & inclusiveMask(local, chunk));
    acc += (chunk == 2u * letter) ? (uint64_t)v.w : 0ull;
    acc += (chunk == 2u * letter + 1u) ? ((uint64_t)v.w << 32) : 0ull;
  }
  return acc;
}

template <int LPQ>
struct AminoLoad {
  uint4 v[8 / LPQ];
  uint32_t b4[8 / LPQ];
  uint64_t base;
};
template <int LPQ>
__device__ __forceinline__ AminoLoad<LPQ> aminoIssue(const DevIndex &ix, uint64_t block, uint32_t letter,
                                                     unsigned sub) {
  AminoLoad<LPQ> l;
  const uint4 *line = ix.lines + block * kAminoLineU4;
#pragma unroll
  for (int i = 0; i < 8 / LPQ; i++) {
    l.v[i] = ldLine(line + sub + LPQ * i);
    l.b4[i] = __ldg(reinterpret_cast<const uint32_t *>(line + 8) + sub + LPQ * i);
  }
  l.base = __ldg(reinterpret_cast<const uint64_t *>(line + 10) + letter);
  return l;
}
template <int LPQ>
__device__ __forceinline__ uint64_t aminoPartial(const AminoLoad<LPQ> &l, const Selector &s, uint32_t local,
                                                 unsigned sub) {
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8 / LPQ; i++) {
    const uint32_t chunk = sub + LPQ * i;
    const uint4 v = l.v[i];
    const uint32_t sel = ((v.x ^ s.flip[0]) | s.dontcare[0]) & ((v.y ^ s.flip[1]) | s.dontcare[1]) &
                         ((v.z ^ s.flip[2]) | s.dontcare[2]) & ((v.w ^ s.flip[3]) | s.dontcare[3]) &
                         ((l.b4[i] ^ s.flip[4]) | s.dontcare[4]);
    acc += __popc(sel & inclusiveMask(local, chunk));
  }
  return acc;
}

// One LF-mapping step (src/AwFmSearch.c:42-159): sp' = C[c] + Occ(c, sp-1), ep' = C[c] + Occ(c, ep) - 1.
// Both block lines are requested before either is consumed (two independent misses in flight per group).
template <int LPQ, bool AMINO>
__device__ __forceinline__ void lfStep(const DevIndex &ix, uint64_t &sp, uint64_t &ep, uint32_t letter,
                                       unsigned sub, unsigned mask) {
  const uint64_t pa = sp - 1, pb = ep;
  const uint64_t ba = pa >> 8, bb = pb >> 8;
  const Selector s = makeSelector<AMINO>(letter);
  uint64_t ra, rb;
  if (AMINO) {
    const AminoLoad<LPQ> la = aminoIssue<LPQ>(ix, ba, letter, sub);
    const AminoLoad<LPQ> lb = aminoIssue<LPQ>(ix, bb, letter, sub);
    ra = groupSum<LPQ>(aminoPartial<LPQ>(la, s, (uint32_t)pa & 255u, sub), mask) + la.base;
    rb = groupSum<LPQ>(aminoPartial<LPQ>(lb, s, (uint32_t)pb & 255u, sub), mask) + lb.base;
  } else {
    const NucLoad<LPQ> la = nucIssue<LPQ>(ix, ba, sub);
    const NucLoad<LPQ> lb = nucIssue<LPQ>(ix, bb, sub);
    uint64_t xa = 0, xb = 0;
    if (letter == 4u) {  // ambiguity letter: base count from the side array
      xa = __ldg(ix.xBase + ba);
      xb = __ldg(ix.xBase + bb);
    }
    ra = groupSum<LPQ>(nucPartial<LPQ>(la, s, letter, (uint32_t)pa & 255u, sub), mask) + xa;
    rb = groupSum<LPQ>(nucPartial<LPQ>(lb, s, letter, (uint32_t)pb & 255u, sub), mask) + xb;
  }
  const uint64_t c = ix.prefixSums[letter];
  sp = c + ra;
  ep = c + rb - 1;
}

// One backtrace step (src/AwFmSearch.c:369-427): c = BWT[p]; sentinel -> 0; else C[c] + Occ(c, p) - 1.
// The letter and the rank come from the SAME line load.
template <int LPQ, bool AMINO>
__device__ __forceinline__ uint64_t backtraceStep(const DevIndex &ix, uint64_t p, unsigned sub, unsigned mask) {
  const uint64_t block = p >> 8;
  const uint32_t local = (uint32_t)p & 255u, ownerChunk = local >> 5, bit = local & 31u;
  const unsigned groupBase = (threadIdx.x & 31u) / LPQ * LPQ;
  if (AMINO) {
    const uint4 *line = ix.lines + block * kAminoLineU4;
    AminoLoad<LPQ> l;
    uint32_t code = 0;
#pragma unroll
    for (int i = 0; i < 8 / LPQ; i++) {
      l.v[i] = ldLine(line + sub + LPQ * i);
      l.b4[i] = __ldg(reinterpret_cast<const uint32_t *>(line + 8) + sub + LPQ * i);
    }
#pragma unroll
    for (int i = 0; i < 8 / LPQ; i++) {
      const uint4 v = l.v[i];
      const uint32_t c = ((v.x >> bit) & 1u) | (((v.y >> bit) & 1u) << 1) | (((v.z >> bit) & 1u) << 2) |
                         (((v.w >> bit) & 1u) << 3) | (((l.b4[i] >> bit) & 1u) << 4);
      code = (sub + LPQ * i == ownerChunk) ? c : code;
    }
    code = __shfl_sync(mask, code, groupBase + (ownerChunk % LPQ));
    const uint32_t letter = kAminoCodeToLetter[code];
    if (letter == kAminoSentinel) return 0;
    const uint64_t base = __ldg(reinterpret_cast<const uint64_t *>(line + 10) + letter);
    const Selector s = makeSelector<true>(letter);
    const uint64_t r = groupSum<LPQ>(aminoPartial<LPQ>(l, s, local, sub), mask) + base;
    return ix.prefixSums[letter] + r - 1;
  } else {
    const NucLoad<LPQ> l = nucIssue<LPQ>(ix, block, sub);
    uint32_t code = 0;
#pragma unroll
    for (int i = 0; i < 8 / LPQ; i++) {
      const uint4 v = l.v[i];
      const uint32_t c = ((v.x >> bit) & 1u) | (((v.y >> bit) & 1u) << 1) | (((v.z >> bit) & 1u) << 2);
      code = (sub + LPQ * i == ownerChunk) ? c : code;
    }
    code = __shfl_sync(mask, code, groupBase + (ownerChunk % LPQ));
    const uint32_t letter = nucCodeToLetter(code);
    if (letter == kNucSentinel) return 0;
    const Selector s = makeSelector<false>(letter);
    uint64_t r = groupSum<LPQ>(nucPartial<LPQ>(l, s, letter, local, sub), mask);
    if (letter == 4u) r += __ldg(ix.xBase + block);
    return ix.prefixSums[letter] + r - 1;
  }
}

// sampled-SA field j: w-bit little-endian field at bit j*w (src/AwFmSuffixArray.c:22-39, 114-142)
__device__ __forceinline__ uint64_t saValue(const DevIndex &ix, uint64_t j) {
  const uint32_t w = ix.saBitWidth;
  const uint64_t bit = j * w;  // < 2^64 for any index that fits in memory
  const uint64_t word = bit >> 6;
  const uint32_t sh = (uint32_t)bit & 63u;
  const uint64_t lo = __ldg(ix.sa + word);
  uint64_t v = lo >> sh;
  if (sh + w > 64u) v |= __ldg(ix.sa + word + 1) << (64u - sh);
  return w == 64u ? v : (v & ((1ull << w) - 1ull));
}

__device__ __forceinline__ bool isSampled(const DevIndex &ix, uint64_t p) {  // src/AwFmIndexStruct.c:88-91
  if (ix.saRatioShift != 0xFFFFFFFFu) return (p & ((1ull << ix.saRatioShift) - 1ull)) == 0;
  return (p % (uint64_t)ix.saRatio) == 0;
}
__device__ __forceinline__ uint64_t sampleIndexOf(const DevIndex &ix, uint64_t p) {
  if (ix.saRatioShift != 0xFFFFFFFFu) return p >> ix.saRatioShift;
  return p / (uint64_t)ix.saRatio;
}

}  // namespace awfm
