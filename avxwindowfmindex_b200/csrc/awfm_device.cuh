// awfm_device.cuh — device-side data layout and rank/LF/backtrace primitives (sm_100a).
//
// HBM layout (built once at upload by relayout kernels in awfm_kernels.cuh from the unchanged reference blocks):
//
//  nucleotide "sector" = 32 B, 32-B aligned, one per 64 BWT positions (0.5 B per position; a 128-B line = 256 positions):
//      words 0-1 / 2-3 / 4-5 : bits 0 / 1 / 2 of the letter codes as 64-bit vectors (bit t <-> position 64*s + t)
//      word 6 = count(A) | count(C) << 16,  word 7 = count(G) | count(T) << 16: occurrences in BWT[S * 2^16, 64*s)
//               where S = s >> 10 is the enclosing 2^16-position superblock (so every count fits 16 bits);
//      superC[S][c] = C[c] + occurrences of c before the superblock (64 B per 2^16 positions: 3 MB at 3.1 Gbp,
//               L2-resident; the address does not depend on the sector, so the read overlaps the sector's miss);
//      xRel16[s] = the ambiguity letter's 16-bit count (side array, touched only when a query or the BWT holds X).
//    One rank = ONE 32-B sector read by ONE lane (two 128-bit loads of the same sector = one request), the letter
//    selector is three 64-bit LOP3s, the count one shift/mask: no shuffles inside a rank.  A 2-lane group does an
//    LF step (lane 0 ranks sp-1, lane 1 ranks ep, one exchange); a single thread owns a located hit's whole walk.
//    History of this layout on B200 (profiles/): 128-B lines per 256 positions with 64-bit counts (3.9 G queries/s,
//    instruction-bound) -> 64-B half-lines per 128 positions with 32-bit counts read by 2-4 lanes (5.6 G) -> sectors
//    (6.4 G; locate 2.6 -> 4.6 G hits/s because a walk needs one lane, so 2048 walks are in flight per SM).
//    Reference layout: struct AwFmNucleotideBlock, 160 B per 256 positions, 32-B aligned (src/AwFmIndex.h:61-65).
//
//  amino "quarter-line" = 128 B, 128-B aligned, one per 64 BWT positions (32 words):
//      w[0..3]   b0..b3 of positions 64*q +  0..31     (bit t <-> position t)
//      w[4..7]   b0..b3 of positions 64*q + 32..63
//      w[8], w[9] b4 of the low / high 32 positions;  w[10] = 0
//      w[11 + c] occurrences of letter c (A..Y = 0..19, Z = 20) in BWT[0, 64*q), relative to the enclosing
//                2^31-position superblock (32 bits); absolute 64-bit superblock counts + C[c] sit in superC.
//    One rank = ONE 128-B line (the unit the memory system fetches per missing request anyway, see
//    profiles/r01_granularity_probe.json): code bits, the letter's count and nothing else.  The first amino layout
//    of this path kept the reference's 256-position granularity (384-B line triples: 128 B of b0..b3, then b4,
//    then 21 64-bit counts) and paid ~3 line misses per rank.
//    Reference layout: struct AwFmAminoBlock, 352 B per 256 positions (src/AwFmIndex.h:55-59).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace awfm {

constexpr int kAminoLineU4 = 8;   // uint4 per amino quarter-line (64 positions)
constexpr int kAminoRelWord = 11;  // first relative-count word of a quarter-line
constexpr int kAminoSuperStride = 24;  // u64 per row of the amino superC
constexpr int kSuperShift = 31;        // amino: superblock of the 32-bit relative counts = 2^31 positions
constexpr int kSectorU4 = 2;           // uint4 per nucleotide sector (64 positions)
constexpr int kSectorSuperShift = 16;  // nucleotide: superblock of the 16-bit sector counts = 2^16 positions
constexpr int kSectorSuperStride = 8;  // u64 per row of the nucleotide superC
constexpr uint32_t kNucSentinel = 5, kAminoSentinel = 21;

struct DevIndex {
  const uint4 *lines;        // nucleotide: sectors; amino: quarter-lines
  const uint16_t *xRel16;    // nucleotide only: ambiguity-letter count per sector
  const uint64_t *superC;    // [superblock][8 | 24] = C[c] + count of c before the superblock (c = 0..4 | 0..20);
                             // superblock = 2^16 positions (nucleotide) / 2^31 (amino)
  const uint4 *seedTable;    // {startLo, startHi, endLo, endHi}
  const void *deepSeedTable; // derived at upload time (awfm_gpu_ctx_extend_seed_table), or nullptr: the range the
                             // reference holds after the last deepSeedK letters; uint2 {start, end} when
                             // bwtLength <= 2^32, else uint4 like seedTable
  uint32_t deepSeedK, deepSeedWide;
  const uint64_t *sa;        // bit-packed sampled SA viewed as little-endian u64 words (+16 B zero padding)
  const uint64_t *sequenceEnds;  // cumulative end offset of every FASTA record (incl. its separator), or nullptr
  uint64_t numSequences;
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
  return (uint32_t)(0x7271335566ull >> (letter * 8u)) & 0xFFu;
}
static __constant__ uint16_t kAminoCodeCare[21] = {
    // (care << 8) | code
    0x1C0C, 0x0F17, 0x1303, 0x1606, 0x0F1E, 0x151A, 0x0F1B, 0x1619, 0x1A15, 0x131C, 0x0F1D,
    0x0F08, 0x1909, 0x0F04, 0x1C13, 0x1A0A, 0x1505, 0x1916, 0x0F01, 0x0F02, 0x0F1F};
// code -> letter index (src/AwFmLetter.c:49-53, :89-96)
__device__ __forceinline__ uint32_t nucCodeToLetter(uint32_t code) { return (0x00152435u >> (code * 4u)) & 0xFu; }
static __constant__ uint8_t kAminoCodeToLetter[32] = {21, 18, 19, 2,  13, 16, 3,  20, 11, 12, 15, 20, 0, 20, 20, 20,
                                                      20, 20, 20, 14, 20, 8,  17, 1,  20, 7,  5,  6,  9, 10, 4,  20};

// mask of the low `n` bits with n clamped to [0, 32] (one BMSK after a max with zero)
__device__ __forceinline__ uint32_t lowBits(int n) {
  uint32_t m;
  const uint32_t width = (uint32_t)max(n, 0);
  asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(m) : "r"(0u), "r"(width));
  return m;
}
// mask of the bits of a 32-bit word (first bit = block position 32*chunk) at block positions <= local, INCLUSIVE
// (AwFmMaskedVectorPopcount semantics, src/AwFmSimdConfig.c:89-114)
__device__ __forceinline__ uint32_t inclusiveMask(uint32_t local, uint32_t chunk) {
  return lowBits((int)local + 1 - (int)(chunk * 32u));
}

template <int LPQ>
__device__ __forceinline__ unsigned groupMaskOf() {
  if (LPQ == 32) return 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u;
  return ((1u << LPQ) - 1u) << (lane / LPQ * LPQ);
}
template <int LPQ>
__device__ __forceinline__ unsigned groupBaseLane() {
  return (threadIdx.x & 31u) / LPQ * LPQ;
}

template <int LPQ>
__device__ __forceinline__ uint64_t groupSum(uint64_t v, unsigned mask) {
#pragma unroll
  for (int d = LPQ / 2; d > 0; d >>= 1) v += __shfl_xor_sync(mask, v, d, LPQ);
  return v;
}
template <int LPQ>
__device__ __forceinline__ uint32_t groupSum32(uint32_t v, unsigned mask) {
#pragma unroll
  for (int d = LPQ / 2; d > 0; d >>= 1) v += __shfl_xor_sync(mask, v, d, LPQ);
  return v;
}

// Selector masks for one LF/backtrace step, expanded to 32-bit lanes: y_i = (b_i ^ flip_i) | dontcare_i (one LOP3)
struct Selector {
  uint32_t flip[5], dontcare[5];
};
template <bool AMINO>
__device__ __forceinline__ Selector makeSelector(uint32_t letter) {
  Selector s;
  uint32_t code, care;
  if constexpr (AMINO) {
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

// 128-bit read-only load that asks L2 to fetch only the 64-B half of the line it misses on (PTX prefetch-size hint
// .L2::64B).  By default a missing sector brings its whole 128-B line from DRAM (profiles/r01_granularity_probe*:
// 125 B per random read, 63 B with the hint); on this path the other half is rarely wanted.
__device__ __forceinline__ uint4 ldgHalfLine(const uint4 *p) {
  uint4 v;
#ifdef AWFM_NO_L2_64B_HINT
  v = __ldg(p);
#else
  asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#endif
  return v;
}

// ================================================================================================ nucleotide
struct NucSector {
  uint4 v0, v1;  // v0 = {b0.lo, b0.hi, b1.lo, b1.hi}, v1 = {b2.lo, b2.hi, A|C<<16, G|T<<16}
};
__device__ __forceinline__ NucSector sectorIssue(const DevIndex &ix, uint64_t p) {
  const uint4 *s = ix.lines + (p >> 6) * kSectorU4;
  NucSector x;
  x.v0 = ldgHalfLine(s);
  x.v1 = ldgHalfLine(s + 1);
  return x;
}
// popcount(select(letter) & positions 0..local inclusive), local in [0, 64)
__device__ __forceinline__ uint32_t sectorPop(const NucSector &x, uint32_t letter, uint32_t local) {
  const uint32_t cc = nucCodeCare(letter), code = cc & 0xFu, care = cc >> 4;
  const uint64_t b0 = (uint64_t)x.v0.x | ((uint64_t)x.v0.y << 32), b1 = (uint64_t)x.v0.z | ((uint64_t)x.v0.w << 32),
                 b2 = (uint64_t)x.v1.x | ((uint64_t)x.v1.y << 32);
  const uint64_t f0 = (uint64_t)((code >> 0) & 1u) - 1ull, f1 = (uint64_t)((code >> 1) & 1u) - 1ull,
                 f2 = (uint64_t)((code >> 2) & 1u) - 1ull;
  const uint64_t d0 = (uint64_t)((care >> 0) & 1u) - 1ull, d1 = (uint64_t)((care >> 1) & 1u) - 1ull,
                 d2 = (uint64_t)((care >> 2) & 1u) - 1ull;
  const uint64_t sel = ((b0 ^ f0) | d0) & ((b1 ^ f1) | d1) & ((b2 ^ f2) | d2);
  const uint64_t mask = (2ull << local) - 1ull;  // local == 63 -> all ones
  return (uint32_t)__popcll(sel & mask);
}
// 16-bit count of letter 0..3 at the sector start, relative to its 2^16-position superblock
__device__ __forceinline__ uint32_t sectorCount(const NucSector &x, uint32_t letter) {
  return (((letter & 2u) ? x.v1.w : x.v1.z) >> ((letter & 1u) * 16u)) & 0xFFFFu;
}
// Occ(letter, p) + C[letter] for letter 0..4
__device__ __forceinline__ uint64_t sectorRank(const DevIndex &ix, const NucSector &x, uint64_t super, uint32_t letter,
                                               uint64_t p) {
  const uint32_t rel = (letter == 4u) ? (uint32_t)__ldg(ix.xRel16 + (p >> 6)) : sectorCount(x, letter);
  return super + rel + sectorPop(x, letter, (uint32_t)p & 63u);
}
// one LF step owned by one thread (src/AwFmSearch.c:42-103)
__device__ __forceinline__ void lfStepSector(const DevIndex &ix, uint64_t &sp, uint64_t &ep, uint32_t letter) {
  const uint64_t pa = sp - 1, pb = ep;
  const NucSector a = sectorIssue(ix, pa), b = sectorIssue(ix, pb);
  const uint64_t ca = __ldg(ix.superC + (pa >> kSectorSuperShift) * kSectorSuperStride + letter);
  const uint64_t cb = __ldg(ix.superC + (pb >> kSectorSuperShift) * kSectorSuperStride + letter);
  sp = sectorRank(ix, a, ca, letter, pa);
  ep = sectorRank(ix, b, cb, letter, pb) - 1;
}
// one LF step by a 2-lane group: lane 0 ranks sp-1, lane 1 ranks ep, one exchange
__device__ __forceinline__ void lfStepSectorPair(const DevIndex &ix, uint64_t &sp, uint64_t &ep, uint32_t letter,
                                                 unsigned sub, unsigned mask) {
  const uint64_t p = sub == 0 ? sp - 1 : ep;
  const NucSector x = sectorIssue(ix, p);
  const uint64_t c = __ldg(ix.superC + (p >> kSectorSuperShift) * kSectorSuperStride + letter);
  const uint64_t mine = sectorRank(ix, x, c, letter, p);
  const uint64_t other = __shfl_xor_sync(mask, mine, 1);
  sp = sub == 0 ? mine : other;
  ep = (sub == 0 ? other : mine) - 1;
}
// one backtrace step owned by one thread (src/AwFmSearch.c:369-397)
__device__ __forceinline__ uint64_t backtraceStepSector(const DevIndex &ix, uint64_t p) {
  const uint64_t *row = ix.superC + (p >> kSectorSuperShift) * kSectorSuperStride;
  const NucSector x = sectorIssue(ix, p);
  asm volatile("prefetch.global.L1 [%0];" ::"l"(row));      // the row entry depends on the letter being read:
  asm volatile("prefetch.global.L1 [%0];" ::"l"(row + 4));  // have both of its sectors on the way meanwhile
  const uint32_t local = (uint32_t)p & 63u, bit = local & 31u;
  const bool hi = local >= 32u;
  const uint32_t w0 = hi ? x.v0.y : x.v0.x, w1 = hi ? x.v0.w : x.v0.z, w2 = hi ? x.v1.y : x.v1.x;
  const uint32_t code = ((w0 >> bit) & 1u) | (((w1 >> bit) & 1u) << 1) | (((w2 >> bit) & 1u) << 2);
  const uint32_t letter = nucCodeToLetter(code);
  if (letter == kNucSentinel) return 0;
  return sectorRank(ix, x, __ldg(row + letter), letter, p) - 1;
}

// ================================================================================================ amino
// LPQ lanes (1 or 2) cooperate on one quarter-line; lane `sub` holds the 32-position chunks sub, sub+LPQ.
template <int LPQ>
struct AminoLoad {
  uint4 v[2 / LPQ];
  uint32_t b4[2 / LPQ];
};
template <int LPQ>
__device__ __forceinline__ AminoLoad<LPQ> aminoIssue(const uint4 *line, unsigned sub) {
  AminoLoad<LPQ> l;
#pragma unroll
  for (int i = 0; i < 2 / LPQ; i++) {
    l.v[i] = __ldg(line + sub + LPQ * i);
    l.b4[i] = __ldg(reinterpret_cast<const uint32_t *>(line + 2) + sub + LPQ * i);
  }
  return l;
}
__device__ __forceinline__ uint32_t aminoRel(const uint4 *line, uint32_t letter) {
  return __ldg(reinterpret_cast<const uint32_t *>(line) + kAminoRelWord + letter);
}
template <int LPQ>
__device__ __forceinline__ uint32_t aminoPop(const AminoLoad<LPQ> &l, const Selector &s, uint32_t local,
                                             unsigned sub) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 2 / LPQ; i++) {
    const uint4 v = l.v[i];
    const uint32_t sel = ((v.x ^ s.flip[0]) | s.dontcare[0]) & ((v.y ^ s.flip[1]) | s.dontcare[1]) &
                         ((v.z ^ s.flip[2]) | s.dontcare[2]) & ((v.w ^ s.flip[3]) | s.dontcare[3]) &
                         ((l.b4[i] ^ s.flip[4]) | s.dontcare[4]);
    acc += __popc(sel & inclusiveMask(local, sub + LPQ * i));
  }
  return acc;
}

// ---- 4-lane amino groups: the whole quarter-line in ONE pair of instructions (lane j holds uint4 j and j+4; each
// instruction covers two adjacent sectors of 8 lines per warp, the request shape that sustains the full random-line
// rate in profiles/r01_gather_shapes_ncu.txt).  Lanes 0/1 own the low/high 32 positions, lane 2 also holds b4 and
// count 0, lanes 3,0,1,2,3 (second load) the other counts.
struct AminoLine4 {
  uint4 a, b;
};
__device__ __forceinline__ AminoLine4 aminoIssue4(const uint4 *line, unsigned sub) {
  AminoLine4 l;
  l.a = __ldg(line + sub);
  l.b = __ldg(line + sub + 4);
  return l;
}
// b4 word of this lane's 32 positions (meaningful in lanes 0 and 1 of the group)
__device__ __forceinline__ uint32_t aminoB4Of(const AminoLine4 &l, unsigned sub, unsigned mask) {
  const unsigned src = groupBaseLane<4>() + 2;
  const uint32_t lo = __shfl_sync(mask, l.a.x, src), hi = __shfl_sync(mask, l.a.y, src);
  return sub == 0 ? lo : hi;
}
__device__ __forceinline__ uint32_t aminoPop4(const AminoLine4 &l, uint32_t b4, const Selector &s, uint32_t local,
                                              unsigned sub) {
  const uint4 v = l.a;
  const uint32_t sel = ((v.x ^ s.flip[0]) | s.dontcare[0]) & ((v.y ^ s.flip[1]) | s.dontcare[1]) &
                       ((v.z ^ s.flip[2]) | s.dontcare[2]) & ((v.w ^ s.flip[3]) | s.dontcare[3]) &
                       ((b4 ^ s.flip[4]) | s.dontcare[4]);
  return sub < 2 ? __popc(sel & inclusiveMask(local, sub)) : 0u;
}
// relative count word of `letter`: word 11 + letter of the line = uint4 (w >> 2), component (w & 3)
__device__ __forceinline__ uint32_t aminoRel4(const AminoLine4 &l, uint32_t letter, unsigned mask) {
  const uint32_t w = kAminoRelWord + letter, u = w >> 2, comp = w & 3u;
  const uint4 v = (u >> 2) ? l.b : l.a;
  const uint32_t mine = comp == 0 ? v.x : comp == 1 ? v.y : comp == 2 ? v.z : v.w;
  return __shfl_sync(mask, mine, groupBaseLane<4>() + (u & 3u));
}

// ================================================================================================ LF step
// One LF-mapping step (src/AwFmSearch.c:42-159): sp' = C[c] + Occ(c, sp-1), ep' = C[c] + Occ(c, ep) - 1.
// Both blocks are requested before either is consumed (two independent misses in flight per group); the two
// partial popcounts travel through the group reduction packed in one 32-bit register.
// Nucleotide: one thread (both ranks) or a 2-lane group (one sector per lane); amino groups are 1 or 2 lanes
// (quarter-line = 2 chunks of 32 positions) or 4 lanes (whole line in registers, see AminoLine4).
template <int LPQ, bool AMINO>
__device__ __forceinline__ void lfStep(const DevIndex &ix, uint64_t &sp, uint64_t &ep, uint32_t letter,
                                       unsigned sub, unsigned mask) {
  [[maybe_unused]] const uint64_t pa = sp - 1, pb = ep;
  [[maybe_unused]] Selector s;
  if constexpr (AMINO) s = makeSelector<true>(letter);
  if constexpr (AMINO && LPQ == 4) {
    const uint4 *lineA = ix.lines + (pa >> 6) * kAminoLineU4, *lineB = ix.lines + (pb >> 6) * kAminoLineU4;
    const AminoLine4 la = aminoIssue4(lineA, sub);
    const AminoLine4 lb = aminoIssue4(lineB, sub);
    const uint64_t ca = __ldg(ix.superC + (pa >> kSuperShift) * kAminoSuperStride + letter);
    const uint64_t cb = __ldg(ix.superC + (pb >> kSuperShift) * kAminoSuperStride + letter);
    const uint32_t ra = aminoRel4(la, letter, mask), rb = aminoRel4(lb, letter, mask);
    const uint32_t b4a = aminoB4Of(la, sub, mask), b4b = aminoB4Of(lb, sub, mask);
    const uint32_t packed = groupSum32<4>(aminoPop4(la, b4a, s, (uint32_t)pa & 63u, sub) |
                                              (aminoPop4(lb, b4b, s, (uint32_t)pb & 63u, sub) << 16),
                                          mask);
    sp = ca + ra + (packed & 0xFFFFu);
    ep = cb + rb + (packed >> 16) - 1;
  } else if constexpr (AMINO) {
    static_assert(!AMINO || LPQ <= 2, "2-chunk amino groups are 1 or 2 lanes");
    const uint4 *lineA = ix.lines + (pa >> 6) * kAminoLineU4, *lineB = ix.lines + (pb >> 6) * kAminoLineU4;
    const AminoLoad<LPQ> la = aminoIssue<LPQ>(lineA, sub);
    const AminoLoad<LPQ> lb = aminoIssue<LPQ>(lineB, sub);
    const uint32_t ra = aminoRel(lineA, letter), rb = aminoRel(lineB, letter);
    const uint64_t ca = __ldg(ix.superC + (pa >> kSuperShift) * kAminoSuperStride + letter);
    const uint64_t cb = __ldg(ix.superC + (pb >> kSuperShift) * kAminoSuperStride + letter);
    const uint32_t packed = groupSum32<LPQ>(aminoPop<LPQ>(la, s, (uint32_t)pa & 63u, sub) |
                                                (aminoPop<LPQ>(lb, s, (uint32_t)pb & 63u, sub) << 16),
                                            mask);
    sp = ca + ra + (packed & 0xFFFFu);
    ep = cb + rb + (packed >> 16) - 1;
  } else {
    static_assert(AMINO || LPQ <= 2, "nucleotide LF steps are done by one thread or by a 2-lane group");
    if constexpr (LPQ == 1) lfStepSector(ix, sp, ep, letter);
    else lfStepSectorPair(ix, sp, ep, letter, sub, mask);
  }
}

// One backtrace step (src/AwFmSearch.c:369-427): c = BWT[p]; sentinel -> 0; else C[c] + Occ(c, p) - 1.
// The letter and the rank come from the SAME block load.
template <int LPQ, bool AMINO>
__device__ __forceinline__ uint64_t backtraceStep(const DevIndex &ix, uint64_t p, unsigned sub, unsigned mask) {
  if constexpr (AMINO && LPQ == 4) {
    const uint32_t local = (uint32_t)p & 63u, ownerChunk = local >> 5, bit = local & 31u;
    const uint4 *line = ix.lines + (p >> 6) * kAminoLineU4;
    const AminoLine4 l = aminoIssue4(line, sub);
    const uint32_t b4 = aminoB4Of(l, sub, mask);
    const uint32_t c = ((l.a.x >> bit) & 1u) | (((l.a.y >> bit) & 1u) << 1) | (((l.a.z >> bit) & 1u) << 2) |
                       (((l.a.w >> bit) & 1u) << 3) | (((b4 >> bit) & 1u) << 4);
    const uint32_t code = __shfl_sync(mask, c, groupBaseLane<4>() + ownerChunk);
    const uint32_t letter = kAminoCodeToLetter[code];
    if (letter == kAminoSentinel) return 0;
    const uint32_t rel = aminoRel4(l, letter, mask);
    const Selector s = makeSelector<true>(letter);
    const uint32_t pop = groupSum32<4>(aminoPop4(l, b4, s, local, sub), mask);
    return __ldg(ix.superC + (p >> kSuperShift) * kAminoSuperStride + letter) + rel + pop - 1;
  } else if constexpr (AMINO) {
    const uint32_t local = (uint32_t)p & 63u, ownerChunk = local >> 5, bit = local & 31u;
    const uint4 *line = ix.lines + (p >> 6) * kAminoLineU4;
    const AminoLoad<LPQ> l = aminoIssue<LPQ>(line, sub);
    // the count word depends on the letter just being read: pull the line's other two sectors towards L1 now so
    // that dependent load does not pay a second trip to L2
    asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const uint8_t *>(line) + 64));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const uint8_t *>(line) + 96));
    uint32_t code = 0;
#pragma unroll
    for (int i = 0; i < 2 / LPQ; i++) {
      const uint4 v = l.v[i];
      const uint32_t c = ((v.x >> bit) & 1u) | (((v.y >> bit) & 1u) << 1) | (((v.z >> bit) & 1u) << 2) |
                         (((v.w >> bit) & 1u) << 3) | (((l.b4[i] >> bit) & 1u) << 4);
      code = (sub + LPQ * i == ownerChunk) ? c : code;
    }
    if (LPQ > 1) code = __shfl_sync(mask, code, groupBaseLane<LPQ>() + (ownerChunk % LPQ));
    const uint32_t letter = kAminoCodeToLetter[code];
    if (letter == kAminoSentinel) return 0;
    const uint32_t rel = aminoRel(line, letter);
    const Selector s = makeSelector<true>(letter);
    const uint32_t pop = groupSum32<LPQ>(aminoPop<LPQ>(l, s, local, sub), mask);
    return __ldg(ix.superC + (p >> kSuperShift) * kAminoSuperStride + letter) + rel + pop - 1;
  } else {
    static_assert(AMINO || LPQ == 1, "a nucleotide walk is owned by one thread");
    return backtraceStepSector(ix, p);
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
