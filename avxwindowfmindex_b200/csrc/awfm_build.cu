// awfm_build.cu — device-side construction of the index arrays consumed by the search path (SURVEY.md §8 row f4,
// "next"): suffix array by radix sort of packed prefixes, BWT bit-vector blocks + base occurrences, prefix sums,
// k-mer seed table and the bit-packed sampled suffix array, all in the reference's own formats
// (src/AwFmCreate.c:31-450, src/AwFmSuffixArray.c:58-112) so the result can be written as an unchanged `.awfmi`.
//
// Scope: texts with bwtLength < 2^32.  Suffixes are first ordered by one radix sort of their first 22 (nucleotide) /
// 13 (amino) symbols per leading-symbol bucket; the groups that still tie are finished ON THE DEVICE by prefix
// doubling (Manber-Myers / Larsson-Sadakane): with rank_h = the position of a suffix's group in the h-order, the tied
// suffixes are sorted by (rank_h[i], rank_h[i + h]) — one radix sort of the tied elements only — which gives the
// 2h-order, so a repeat of length L costs log2(L / 22) rounds instead of a comparison of L symbols per pair (tandem
// arrays, satellite DNA, duplicated segments; the reference's libdivsufsort is O(n log n) as well).  Amino texts must
// be single-case over the 20 letters plus ambiguity codes, as the reference's own sanitizer assumes
// (src/AwFmLetter.c:69-79).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include <string>
#include <vector>

#include "../../include/awfm_gpu.h"
#include "awfm_device.cuh"

using namespace awfm;

extern int awfm_set_error(int code, const char *what, const char *detail);

#define CUB_TRY(call)                                                                                              \
  do {                                                                                                             \
    cudaError_t e_ = (call);                                                                                       \
    if (e_ != cudaSuccess) {                                                                                       \
      cudaGetLastError();                                                                                          \
      return awfm_set_error(e_ == cudaErrorMemoryAllocation ? AWFM_GPU_ERR_ALLOC : AWFM_GPU_ERR_CUDA, #call,       \
                            cudaGetErrorString(e_));                                                               \
    }                                                                                                              \
  } while (0)

namespace {

struct Buf {
  void *p = nullptr;
  ~Buf() { cudaFree(p); }
  cudaError_t alloc(size_t bytes) {
    cudaFree(p);
    p = nullptr;
    return cudaMalloc(&p, bytes ? bytes : 16);
  }
  template <typename T>
  T *as() { return (T *)p; }
};

// ---- synthetic letters, identical to avxwindowfmindex_b200/synth.py ----
__global__ void synthLetters(uint8_t *out, uint64_t count, uint64_t seed, uint64_t start, int amino) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  uint64_t z = seed + (start + i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  if (amino) {
    const char *a = "ACDEFGHIKLMNPQRSTVWY";
    out[i] = a[((z >> 32) * 20ull) >> 32];
  } else {
    const char *a = "ACGT";
    out[i] = a[z >> 62];
  }
}

// sym[i] = letter index + 1 (sort rank; the sentinel is 0), zero padding after the sentinel
template <bool AMINO>
__global__ void textToSymbols(const uint8_t *text, uint64_t n, uint64_t padded, uint8_t *sym) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= padded) return;
  uint8_t s = 0;
  if (i < n) {
    uint32_t l = letterIndex<AMINO>(text[i]);
    if (l > (AMINO ? 20u : 4u)) l = AMINO ? 20u : 4u;  // a literal '$' inside the text is sanitised like any other byte
    s = (uint8_t)(l + 1);
  }
  sym[i] = s;
}

__global__ void symbolHistogram(const uint8_t *sym, uint64_t n, unsigned long long *hist /*32*/) {
  __shared__ unsigned int local[32];
  if (threadIdx.x < 32) local[threadIdx.x] = 0;
  __syncthreads();
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    atomicAdd(&local[sym[i] & 31], 1u);
  __syncthreads();
  if (threadIdx.x < 32 && local[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)local[threadIdx.x]);
}

struct FirstSymbolIs {
  const uint8_t *sym;
  uint8_t s;
  __device__ bool operator()(uint32_t i) const { return sym[i] == s; }
};

// key = symbols [pos+1, pos+1+D) packed most-significant first
__global__ void packKeys(const uint8_t *sym, const uint32_t *pos, uint64_t count, int depth, int bits, uint64_t *keys) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  const uint8_t *s = sym + pos[j] + 1;
  uint64_t k = 0;
  for (int d = 0; d < depth; d++) k = (k << bits) | s[d];
  keys[j] = k;
}

// per SA position of a bucket: bit 0 = first of its group of equal keys, bit 1 = the group has more than one member
__global__ void groupFlags(const uint64_t *keys, uint64_t count, uint8_t *flags) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  const uint64_t k = keys[j];
  const bool head = j == 0 || keys[j - 1] != k;
  const bool tied = !head || (j + 1 < count && keys[j + 1] == k);
  flags[j] = (head ? 1u : 0u) | (tied ? 2u : 0u);
}
struct IsTied {
  __host__ __device__ unsigned long long operator()(uint8_t f) const { return (f >> 1) & 1u; }
};
struct IsTiedFlag {
  __host__ __device__ bool operator()(uint8_t f) const { return (f & 2u) != 0; }
};
struct HeadPosition {  // SA position j if j starts a group, else 0: the running maximum is the rank of j's group
  const uint8_t *flags;
  __host__ __device__ uint32_t operator()(uint32_t j) const { return (flags[j] & 1u) ? j : 0u; }
};
struct MaxU32 {
  __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};
__global__ void scatterRanks(const uint32_t *sa, const uint32_t *rank, uint64_t count, uint32_t *isa) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < count) isa[sa[j]] = rank[j];
}
// one doubling round over the tied elements (t = index into the list of tied SA positions, ascending)
__global__ void doublingKeys(const uint32_t *tieIdx, uint64_t m, const uint32_t *sa, const uint32_t *isa, uint64_t h,
                             uint64_t *keys, uint32_t *pos) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const uint32_t p = sa[tieIdx[t]];
  pos[t] = p;
  // tied at depth h => the first h symbols hold no sentinel => p + h <= n: a valid suffix
  keys[t] = ((uint64_t)isa[p] << 32) | isa[(uint64_t)p + h];
}
struct DoublingHead {  // SA position of element t if it starts a group of equal (rank_h, rank_h at +h) pairs, else 0
  const uint64_t *keys;
  const uint32_t *tieIdx;
  __host__ __device__ uint32_t operator()(uint32_t t) const {
    return (t == 0 || keys[t] != keys[t - 1]) ? tieIdx[t] : 0u;
  }
};
__global__ void doublingScatter(const uint32_t *tieIdx, uint64_t m, const uint64_t *keys, const uint32_t *pos,
                                const uint32_t *rank, uint32_t *sa, uint32_t *isa, uint8_t *still) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  sa[tieIdx[t]] = pos[t];
  isa[pos[t]] = rank[t];
  const bool head = t == 0 || keys[t] != keys[t - 1];
  still[t] = !head || (t + 1 < m && keys[t + 1] == keys[t]);
}

// One CTA (256 threads) per BWT block: letter bit-vectors in the reference layout + per-block letter counts.
// src/AwFmCreate.c:281-405.  counts is letter-major: counts[letter * numBlocks + block].
template <bool AMINO>
__global__ void buildBlocks(const uint8_t *sym, const uint32_t *sa, uint64_t bwtLength, uint64_t numBlocks,
                            uint8_t *rawBlocks, unsigned long long *counts) {
  constexpr int NVEC = AMINO ? 5 : 3, NLET = AMINO ? 22 : 6, BLOCK = AMINO ? 352 : 160;
  __shared__ unsigned int hist[NLET];
  const uint64_t blk = blockIdx.x;
  const uint64_t p = blk * 256 + threadIdx.x;
  if (threadIdx.x < NLET) hist[threadIdx.x] = 0;
  __syncthreads();
  uint32_t code = 0, letter = 0xFFu;
  if (p < bwtLength) {
    const uint32_t s = sa[p];
    letter = (s == 0) ? (AMINO ? 21u : 5u) : (uint32_t)sym[s - 1] - 1u;
    if (AMINO) {
      code = letter == 21u ? 0u : (kAminoCodeCare[letter] & 0xFFu);  // sentinel 0b00000 (src/AwFmLetter.c:81-87)
    } else {
      code = letter == 5u ? 4u : (nucCodeCare(letter) & 0xFu);  // sentinel 0b100 (src/AwFmLetter.c:44-47)
    }
    atomicAdd(&hist[letter], 1u);
  }
  uint32_t *vec = reinterpret_cast<uint32_t *>(rawBlocks + blk * BLOCK);
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
#pragma unroll
  for (int b = 0; b < NVEC; b++) {
    const unsigned word = __ballot_sync(0xFFFFFFFFu, (code >> b) & 1u);
    if (lane == 0) vec[8 * b + warp] = word;
  }
  __syncthreads();
  if (threadIdx.x < NLET) counts[(uint64_t)threadIdx.x * numBlocks + blk] = hist[threadIdx.x];
}

template <bool AMINO>
__global__ void writeBaseOccurrences(const unsigned long long *base /*letter-major exclusive sums*/, uint64_t numBlocks,
                                     uint8_t *rawBlocks) {
  constexpr int NLET = AMINO ? 22 : 6, NSLOT = AMINO ? 24 : 8, BLOCK = AMINO ? 352 : 160, OFF = AMINO ? 160 : 96;
  const uint64_t blk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= numBlocks) return;
  unsigned long long *dst = reinterpret_cast<unsigned long long *>(rawBlocks + blk * BLOCK + OFF);
  for (int l = 0; l < NSLOT; l++) dst[l] = l < NLET ? base[(uint64_t)l * numBlocks + blk] : 0ull;
}

// seed table by counting (equivalent to the reference's recursive backward steps, src/AwFmCreate.c:407-450):
//   table[K] = [ #suffixes < K , #suffixes < K + #suffixes prefixed by K - 1 ]
// hist[K]  = suffixes whose first k symbols are the regular letters of K
// delta[v] = suffixes that hit the sentinel / an ambiguity letter within their first k symbols and sort below every K >= v
template <bool AMINO>
__global__ void seedHistogram(const uint8_t *sym, uint64_t n, int k, uint64_t tableLen, unsigned int *hist,
                              unsigned int *delta) {
  constexpr uint64_t CARD = AMINO ? 20 : 4;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  uint64_t p = 0;
  for (int j = 0; j < k; j++) {
    const uint32_t s = sym[i + j];  // padded with zeros beyond the sentinel
    if (s == 0 || s > CARD) {
      uint64_t scale = 1;
      for (int t = j; t < k; t++) scale *= CARD;
      uint64_t v = p * scale;  // prefix padded with the smallest letter
      if (s != 0) v += scale;  // ambiguity letter sorts above every regular letter
      if (v < tableLen) atomicAdd(&delta[v], 1u);
      return;
    }
    p = p * CARD + (s - 1);
  }
  atomicAdd(&hist[p], 1u);
}

__global__ void seedCombine(const unsigned long long *cumHist, const unsigned long long *cumDelta, const unsigned int *hist,
                            uint64_t tableLen, uint64_t *table) {
  const uint64_t kmer = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kmer >= tableLen) return;
  const uint64_t lb = cumHist[kmer] + cumDelta[kmer];
  table[2 * kmer] = lb;
  table[2 * kmer + 1] = lb + hist[kmer] - 1;
}

struct U32ToU64 {
  __host__ __device__ unsigned long long operator()(unsigned int v) const { return v; }
};

// bit-packed sampled SA (src/AwFmSuffixArray.c:58-112): sample j = SA[j*ratio] in bits [j*w, (j+1)*w); one thread per
// output 64-bit word.  The 8 padding bytes after the packed stream hold, as in the reference's in-place packing of the
// full 64-bit suffix array, the untouched bytes of that array at the same offsets.
__global__ void packSampledSa(const uint32_t *sa, uint64_t bwtLength, uint32_t ratio, uint32_t w, uint64_t numSamples,
                              uint64_t packedBytes /* without padding */, uint64_t numWords, uint64_t *out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= numWords) return;
  const uint64_t bit0 = t * 64, bit1 = bit0 + 64;
  uint64_t word = 0;
  for (uint64_t j = bit0 / w; j < numSamples && j * w < bit1; j++) {
    const uint64_t v = sa[j * ratio];
    const uint64_t start = j * w;
    if (start >= bit0) word |= v << (start - bit0);
    else word |= v >> (bit0 - start);
  }
  // leftover bytes of the original u64 suffix array in [packedBytes, packedBytes + 8)
  for (int b = 0; b < 8; b++) {
    const uint64_t byte = t * 8 + b;
    if (byte >= packedBytes && byte < packedBytes + 8) {
      const uint64_t idx = byte / 8;
      const uint64_t v = idx < bwtLength ? (uint64_t)sa[idx] : 0ull;
      word |= ((v >> (8 * (byte % 8))) & 0xFFull) << (8 * b);
    }
  }
  out[t] = word;
}

inline unsigned gridOf(uint64_t n, unsigned threads = 256) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

struct awfm_built_index {
  int device = 0;
  uint8_t alphabet = 0, seedK = 0, saRatio = 0, saBitWidth = 0;
  uint64_t n = 0, bwtLength = 0, numBlocks = 0, numSeeds = 0, saBytes = 0;
  uint64_t prefixSums[24] = {0};
  Buf blocks, seedTable, sa, prefix;
  uint64_t tieSuffixes = 0;  // suffixes in groups that tied after the first radix sort
  uint32_t tieRounds = 0;    // prefix-doubling rounds that finished them
  double buildMs = 0;
};

extern "C" int awfm_gpu_synth_letters(int device, uint8_t *dOut, uint64_t count, uint64_t seed, uint64_t start, int amino) {
  CUB_TRY(cudaSetDevice(device));
  if (count) synthLetters<<<gridOf(count), 256>>>(dOut, count, seed, start, amino);
  CUB_TRY(cudaGetLastError());
  return AWFM_GPU_OK;
}

extern "C" void awfm_gpu_built_destroy(awfm_built_index *b) {
  if (!b) return;
  cudaSetDevice(b->device);
  delete b;
}

template <bool AMINO>
static int buildImpl(awfm_built_index *B, const uint8_t *dText) {
  constexpr int BITS = AMINO ? 5 : 3, DEPTH = AMINO ? 12 : 21, NSYM = AMINO ? 22 : 6, NLET = AMINO ? 22 : 6;
  constexpr uint64_t CARD = AMINO ? 20 : 4;
  const uint64_t n = B->n, N = B->bwtLength;
  cudaEvent_t e0, e1;
  CUB_TRY(cudaEventCreate(&e0));
  CUB_TRY(cudaEventCreate(&e1));
  CUB_TRY(cudaEventRecord(e0));

  // 1. symbols (rank order), padded so every suffix has DEPTH+1 readable symbols
  const uint64_t padded = N + 64;
  Buf sym;
  CUB_TRY(sym.alloc(padded));
  textToSymbols<AMINO><<<gridOf(padded), 256>>>(dText, n, padded, sym.as<uint8_t>());
  Buf dHist;
  CUB_TRY(dHist.alloc(32 * 8));
  CUB_TRY(cudaMemset(dHist.p, 0, 32 * 8));
  symbolHistogram<<<1184, 256>>>(sym.as<uint8_t>(), N, dHist.as<unsigned long long>());
  unsigned long long hist[32];
  CUB_TRY(cudaMemcpy(hist, dHist.p, sizeof hist, cudaMemcpyDeviceToHost));

  // 2. suffix array: SA[0] is the sentinel suffix; then one radix sort per leading symbol
  Buf sa;
  CUB_TRY(sa.alloc(N * 4));
  {
    const uint32_t sentinelPos = (uint32_t)n;
    CUB_TRY(cudaMemcpy(sa.p, &sentinelPos, 4, cudaMemcpyHostToDevice));
  }
  uint64_t maxBucket = 0;
  for (int s = 1; s < NSYM; s++) maxBucket = std::max<uint64_t>(maxBucket, hist[s]);
  Buf posA, posB, keyA, keyB, flags, numSel, temp;
  CUB_TRY(posA.alloc(maxBucket * 4));
  CUB_TRY(posB.alloc(maxBucket * 4));
  CUB_TRY(keyA.alloc(maxBucket * 8));
  CUB_TRY(keyB.alloc(maxBucket * 8));
  CUB_TRY(flags.alloc(N));  // per SA position: bit 0 = head of its group in the (DEPTH+1)-order, bit 1 = group size > 1
  CUB_TRY(numSel.alloc(8));
  size_t tempBytes = 0;
  auto ensureTemp = [&](size_t bytes) -> cudaError_t {
    if (bytes <= tempBytes) return cudaSuccess;
    tempBytes = bytes;
    return temp.alloc(bytes);
  };
  {
    size_t a = 0, b = 0;
    cub::CountingInputIterator<uint32_t> counting(0);
    CUB_TRY(cub::DeviceSelect::If(nullptr, a, counting, posA.as<uint32_t>(), numSel.as<uint64_t>(), (int64_t)N,
                                  FirstSymbolIs{sym.as<uint8_t>(), 1}));
    CUB_TRY(cub::DeviceRadixSort::SortPairs(nullptr, b, keyA.as<uint64_t>(), keyB.as<uint64_t>(), posA.as<uint32_t>(),
                                            posB.as<uint32_t>(), (int64_t)maxBucket, 0, DEPTH * BITS));
    CUB_TRY(ensureTemp(std::max(a, b)));
  }
  {
    const uint8_t sentinelFlags = 1;  // SA[0]: a group of its own
    CUB_TRY(cudaMemcpy(flags.p, &sentinelFlags, 1, cudaMemcpyHostToDevice));
  }
  uint64_t bucketStart = 1;
  for (int s = 1; s < NSYM; s++) {
    const uint64_t cnt = hist[s];
    if (cnt == 0) continue;
    cub::CountingInputIterator<uint32_t> counting(0);
    size_t tb = tempBytes;
    CUB_TRY(cub::DeviceSelect::If(temp.p, tb, counting, posA.as<uint32_t>(), numSel.as<uint64_t>(), (int64_t)N,
                                  FirstSymbolIs{sym.as<uint8_t>(), (uint8_t)s}));
    packKeys<<<gridOf(cnt), 256>>>(sym.as<uint8_t>(), posA.as<uint32_t>(), cnt, DEPTH, BITS, keyA.as<uint64_t>());
    tb = tempBytes;
    CUB_TRY(cub::DeviceRadixSort::SortPairs(temp.p, tb, keyA.as<uint64_t>(), keyB.as<uint64_t>(), posA.as<uint32_t>(),
                                            posB.as<uint32_t>(), (int64_t)cnt, 0, DEPTH * BITS));
    groupFlags<<<gridOf(cnt), 256>>>(keyB.as<uint64_t>(), cnt, flags.as<uint8_t>() + bucketStart);
    CUB_TRY(cudaMemcpyAsync(sa.as<uint32_t>() + bucketStart, posB.p, cnt * 4, cudaMemcpyDeviceToDevice));
    bucketStart += cnt;
  }
  CUB_TRY(cudaDeviceSynchronize());
  posA.alloc(0), posB.alloc(0), keyA.alloc(0), keyB.alloc(0);

  // 2b. groups that still tie after DEPTH+1 symbols: prefix doubling over the tied elements only
  {
    cub::TransformInputIterator<unsigned long long, IsTied, const uint8_t *> tiedCount(flags.as<uint8_t>(), IsTied());
    size_t tb = 0;
    CUB_TRY(cub::DeviceReduce::Sum(nullptr, tb, tiedCount, numSel.as<unsigned long long>(), (int64_t)N));
    CUB_TRY(ensureTemp(tb));
    tb = tempBytes;
    CUB_TRY(cub::DeviceReduce::Sum(temp.p, tb, tiedCount, numSel.as<unsigned long long>(), (int64_t)N));
    uint64_t m = 0;
    CUB_TRY(cudaMemcpy(&m, numSel.p, 8, cudaMemcpyDeviceToHost));
    B->tieSuffixes = m;
    if (m > 0) {
      cub::CountingInputIterator<uint32_t> counting(0);
      // rank of every suffix in the (DEPTH+1)-order = SA position of the head of its group
      Buf isa, rank;
      CUB_TRY(isa.alloc(N * 4));
      CUB_TRY(rank.alloc(N * 4));
      {
        cub::TransformInputIterator<uint32_t, HeadPosition, cub::CountingInputIterator<uint32_t>> heads(
            counting, HeadPosition{flags.as<uint8_t>()});
        tb = 0;
        CUB_TRY(cub::DeviceScan::InclusiveScan(nullptr, tb, heads, rank.as<uint32_t>(), MaxU32(), (int64_t)N));
        CUB_TRY(ensureTemp(tb));
        tb = tempBytes;
        CUB_TRY(cub::DeviceScan::InclusiveScan(temp.p, tb, heads, rank.as<uint32_t>(), MaxU32(), (int64_t)N));
        scatterRanks<<<gridOf(N), 256>>>(sa.as<uint32_t>(), rank.as<uint32_t>(), N, isa.as<uint32_t>());
        CUB_TRY(cudaGetLastError());
      }
      Buf tieIdx[2], keys[2], pos[2], still;
      CUB_TRY(tieIdx[0].alloc(m * 4));
      CUB_TRY(tieIdx[1].alloc(m * 4));
      {
        cub::TransformInputIterator<bool, IsTiedFlag, const uint8_t *> tiedFlag(flags.as<uint8_t>(), IsTiedFlag());
        tb = 0;
        CUB_TRY(cub::DeviceSelect::Flagged(nullptr, tb, counting, tiedFlag, tieIdx[0].as<uint32_t>(), numSel.as<uint64_t>(),
                                           (int64_t)N));
        CUB_TRY(ensureTemp(tb));
        tb = tempBytes;
        CUB_TRY(cub::DeviceSelect::Flagged(temp.p, tb, counting, tiedFlag, tieIdx[0].as<uint32_t>(), numSel.as<uint64_t>(),
                                           (int64_t)N));
      }
      CUB_TRY(cudaDeviceSynchronize());
      rank.alloc(0);
      flags.alloc(0);
      for (int i = 0; i < 2; i++) {
        CUB_TRY(keys[i].alloc(m * 8));
        CUB_TRY(pos[i].alloc(m * 4));
      }
      CUB_TRY(rank.alloc(m * 4));
      CUB_TRY(still.alloc(m));
      {
        size_t a = 0, b = 0, c = 0;
        CUB_TRY(cub::DeviceRadixSort::SortPairs(nullptr, a, keys[0].as<uint64_t>(), keys[1].as<uint64_t>(), pos[0].as<uint32_t>(),
                                                pos[1].as<uint32_t>(), (int64_t)m, 0, 64));
        cub::TransformInputIterator<uint32_t, DoublingHead, cub::CountingInputIterator<uint32_t>> heads(
            counting, DoublingHead{keys[1].as<uint64_t>(), tieIdx[0].as<uint32_t>()});
        CUB_TRY(cub::DeviceScan::InclusiveScan(nullptr, b, heads, rank.as<uint32_t>(), MaxU32(), (int64_t)m));
        CUB_TRY(cub::DeviceSelect::Flagged(nullptr, c, tieIdx[0].as<uint32_t>(), still.as<uint8_t>(), tieIdx[1].as<uint32_t>(),
                                           numSel.as<uint64_t>(), (int64_t)m));
        CUB_TRY(ensureTemp(std::max(a, std::max(b, c))));
      }
      int cur = 0;
      uint32_t rounds = 0;
      for (uint64_t h = DEPTH + 1; m > 0; h *= 2, rounds++) {
        if (h > N) return awfm_set_error(AWFM_GPU_ERR_CUDA, "suffix sort: ties left beyond the text length", nullptr);
        uint32_t *idx = tieIdx[cur].as<uint32_t>();
        doublingKeys<<<gridOf(m), 256>>>(idx, m, sa.as<uint32_t>(), isa.as<uint32_t>(), h, keys[0].as<uint64_t>(),
                                         pos[0].as<uint32_t>());
        tb = tempBytes;
        CUB_TRY(cub::DeviceRadixSort::SortPairs(temp.p, tb, keys[0].as<uint64_t>(), keys[1].as<uint64_t>(), pos[0].as<uint32_t>(),
                                                pos[1].as<uint32_t>(), (int64_t)m, 0, 64));
        cub::TransformInputIterator<uint32_t, DoublingHead, cub::CountingInputIterator<uint32_t>> heads(
            counting, DoublingHead{keys[1].as<uint64_t>(), idx});
        tb = tempBytes;
        CUB_TRY(cub::DeviceScan::InclusiveScan(temp.p, tb, heads, rank.as<uint32_t>(), MaxU32(), (int64_t)m));
        doublingScatter<<<gridOf(m), 256>>>(idx, m, keys[1].as<uint64_t>(), pos[1].as<uint32_t>(), rank.as<uint32_t>(),
                                            sa.as<uint32_t>(), isa.as<uint32_t>(), still.as<uint8_t>());
        tb = tempBytes;
        CUB_TRY(cub::DeviceSelect::Flagged(temp.p, tb, idx, still.as<uint8_t>(), tieIdx[cur ^ 1].as<uint32_t>(),
                                           numSel.as<uint64_t>(), (int64_t)m));
        CUB_TRY(cudaMemcpy(&m, numSel.p, 8, cudaMemcpyDeviceToHost));
        cur ^= 1;
      }
      B->tieRounds = rounds;
    }
  }
  CUB_TRY(cudaDeviceSynchronize());
  flags.alloc(0);

  // 3. BWT blocks + base occurrences
  const uint64_t numBlocks = B->numBlocks;
  const uint64_t blockBytes = AMINO ? 352 : 160;
  CUB_TRY(B->blocks.alloc(numBlocks * blockBytes));
  CUB_TRY(cudaMemset(B->blocks.p, 0, numBlocks * blockBytes));
  Buf counts, bases;
  CUB_TRY(counts.alloc((uint64_t)NLET * numBlocks * 8));
  CUB_TRY(bases.alloc((uint64_t)NLET * numBlocks * 8));
  buildBlocks<AMINO><<<(unsigned)numBlocks, 256>>>(sym.as<uint8_t>(), sa.as<uint32_t>(), N, numBlocks,
                                                    B->blocks.as<uint8_t>(), counts.as<unsigned long long>());
  {
    size_t tb = 0;
    CUB_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.as<unsigned long long>(), bases.as<unsigned long long>(),
                                          (int64_t)numBlocks));
    if (tb > tempBytes) {
      CUB_TRY(temp.alloc(tb));
      tempBytes = tb;
    }
    for (int l = 0; l < NLET; l++) {
      size_t t2 = tempBytes;
      CUB_TRY(cub::DeviceScan::ExclusiveSum(temp.p, t2, counts.as<unsigned long long>() + (uint64_t)l * numBlocks,
                                            bases.as<unsigned long long>() + (uint64_t)l * numBlocks, (int64_t)numBlocks));
    }
  }
  writeBaseOccurrences<AMINO><<<gridOf(numBlocks), 256>>>(bases.as<unsigned long long>(), numBlocks, B->blocks.as<uint8_t>());
  counts.alloc(0), bases.alloc(0);

  // 4. prefix sums (src/AwFmCreate.c:338-343): [1, 1+#A, ..., bwtLength]
  B->prefixSums[0] = 1;
  for (uint64_t c = 0; c <= CARD; c++) B->prefixSums[c + 1] = B->prefixSums[c] + hist[c + 1];
  CUB_TRY(B->prefix.alloc(24 * 8));
  CUB_TRY(cudaMemcpy(B->prefix.p, B->prefixSums, 24 * 8, cudaMemcpyHostToDevice));

  // 5. seed table
  const uint64_t tableLen = B->numSeeds;
  {
    Buf kh, kd, ch, cd;
    CUB_TRY(kh.alloc(tableLen * 4));
    CUB_TRY(kd.alloc(tableLen * 4));
    CUB_TRY(ch.alloc(tableLen * 8));
    CUB_TRY(cd.alloc(tableLen * 8));
    CUB_TRY(cudaMemset(kh.p, 0, tableLen * 4));
    CUB_TRY(cudaMemset(kd.p, 0, tableLen * 4));
    seedHistogram<AMINO><<<gridOf(N), 256>>>(sym.as<uint8_t>(), n, B->seedK, tableLen, kh.as<unsigned int>(),
                                             kd.as<unsigned int>());
    cub::TransformInputIterator<unsigned long long, U32ToU64, const unsigned int *> hIt(kh.as<unsigned int>(), U32ToU64());
    cub::TransformInputIterator<unsigned long long, U32ToU64, const unsigned int *> dIt(kd.as<unsigned int>(), U32ToU64());
    size_t tb = 0;
    CUB_TRY(cub::DeviceScan::InclusiveSum(nullptr, tb, dIt, cd.as<unsigned long long>(), (int64_t)tableLen));
    if (tb > tempBytes) {
      CUB_TRY(temp.alloc(tb));
      tempBytes = tb;
    }
    size_t t2 = tempBytes;
    CUB_TRY(cub::DeviceScan::ExclusiveSum(temp.p, t2, hIt, ch.as<unsigned long long>(), (int64_t)tableLen));
    t2 = tempBytes;
    CUB_TRY(cub::DeviceScan::InclusiveSum(temp.p, t2, dIt, cd.as<unsigned long long>(), (int64_t)tableLen));
    CUB_TRY(B->seedTable.alloc(tableLen * 16));
    seedCombine<<<gridOf(tableLen), 256>>>(ch.as<unsigned long long>(), cd.as<unsigned long long>(), kh.as<unsigned int>(),
                                           tableLen, B->seedTable.as<uint64_t>());
  }

  // 6. sampled, bit-packed suffix array
  {
    const uint32_t w = B->saBitWidth, r = B->saRatio;
    const uint64_t numSamples = (N + r - 1) / r;
    const uint64_t packedBytes = (numSamples * w + 7) / 8;
    B->saBytes = packedBytes + 8;  // src/AwFmSuffixArray.c:41-53
    const uint64_t numWords = (B->saBytes + 7) / 8;
    CUB_TRY(B->sa.alloc(numWords * 8));
    packSampledSa<<<gridOf(numWords), 256>>>(sa.as<uint32_t>(), N, r, w, numSamples, packedBytes, numWords,
                                             B->sa.as<uint64_t>());
  }
  CUB_TRY(cudaGetLastError());
  CUB_TRY(cudaEventRecord(e1));
  CUB_TRY(cudaEventSynchronize(e1));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  B->buildMs = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return AWFM_GPU_OK;
}

extern "C" int awfm_gpu_build_index(awfm_built_index **out, int device, const uint8_t *dText, uint64_t n, uint8_t alphabet,
                                    uint8_t seedK, uint8_t saRatio) {
  if (!out || !dText) return awfm_set_error(AWFM_GPU_ERR_ARG, "null argument", nullptr);
  if (alphabet < 1 || alphabet > 3 || saRatio == 0 || seedK == 0) return awfm_set_error(AWFM_GPU_ERR_ARG, "bad configuration", nullptr);
  if (n < 1 || n + 1 >= (1ull << 32)) return awfm_set_error(AWFM_GPU_ERR_ARG, "device builder supports 1 <= length < 2^32 - 1", nullptr);
  CUB_TRY(cudaSetDevice(device));
  awfm_built_index *B = new awfm_built_index();
  B->device = device;
  B->alphabet = alphabet, B->seedK = seedK, B->saRatio = saRatio;
  B->n = n, B->bwtLength = n + 1;
  B->numBlocks = 1 + (B->bwtLength - 1) / 256;
  B->saBitWidth = (uint8_t)std::max(1, 64 - __builtin_clzll(B->bwtLength - 1));  // src/AwFmSuffixArray.c:12-18
  B->numSeeds = 1;
  for (int i = 0; i < seedK; i++) B->numSeeds *= (alphabet == 1 ? 20 : 4);
  const int rc = alphabet == 1 ? buildImpl<true>(B, dText) : buildImpl<false>(B, dText);
  if (rc != AWFM_GPU_OK) {
    delete B;
    return rc;
  }
  *out = B;
  return AWFM_GPU_OK;
}

extern "C" int awfm_gpu_build_index_host(awfm_built_index **out, int device, const uint8_t *text, uint64_t n, uint8_t alphabet,
                                         uint8_t seedK, uint8_t saRatio) {
  if (!text) return awfm_set_error(AWFM_GPU_ERR_ARG, "null argument", nullptr);
  CUB_TRY(cudaSetDevice(device));
  Buf dText;
  CUB_TRY(dText.alloc(n));
  CUB_TRY(cudaMemcpy(dText.p, text, n, cudaMemcpyHostToDevice));
  return awfm_gpu_build_index(out, device, dText.as<uint8_t>(), n, alphabet, seedK, saRatio);
}

// device-pointer view for awfm_gpu_ctx_create_from_device; valid while `b` lives
extern "C" int awfm_gpu_built_view(awfm_built_index *b, awfm_index_view *v, uint64_t *tieSuffixes, double *buildMs) {
  if (!b || !v) return awfm_set_error(AWFM_GPU_ERR_ARG, "null argument", nullptr);
  memset(v, 0, sizeof *v);
  v->blocks = b->blocks.p;
  v->numBlocks = b->numBlocks;
  v->prefixSums = (const uint64_t *)b->prefix.p;
  v->seedTable = b->seedTable.p;
  v->saBytes = (const uint8_t *)b->sa.p;
  v->saByteLength = b->saBytes;
  v->bwtLength = b->bwtLength;
  v->saBitWidth = b->saBitWidth;
  v->saRatio = b->saRatio;
  v->seedK = b->seedK;
  v->alphabet = b->alphabet;
  if (tieSuffixes) *tieSuffixes = b->tieSuffixes;
  if (buildMs) *buildMs = b->buildMs;
  return AWFM_GPU_OK;
}

extern "C" uint32_t awfm_gpu_built_tie_rounds(const awfm_built_index *b) { return b ? b->tieRounds : 0; }

// copies the arrays into caller-provided HOST buffers sized from the view (any pointer may be NULL to skip)
extern "C" int awfm_gpu_built_download(awfm_built_index *b, void *blocks, uint64_t *prefixSums, void *seedTable, uint8_t *saBytes) {
  if (!b) return awfm_set_error(AWFM_GPU_ERR_ARG, "null argument", nullptr);
  CUB_TRY(cudaSetDevice(b->device));
  const uint64_t blockBytes = b->alphabet == 1 ? 352 : 160;
  if (blocks) CUB_TRY(cudaMemcpy(blocks, b->blocks.p, b->numBlocks * blockBytes, cudaMemcpyDeviceToHost));
  if (prefixSums) memcpy(prefixSums, b->prefixSums, (b->alphabet == 1 ? 22 : 6) * 8);
  if (seedTable) CUB_TRY(cudaMemcpy(seedTable, b->seedTable.p, b->numSeeds * 16, cudaMemcpyDeviceToHost));
  if (saBytes) CUB_TRY(cudaMemcpy(saBytes, b->sa.p, b->saBytes, cudaMemcpyDeviceToHost));
  return AWFM_GPU_OK;
}
