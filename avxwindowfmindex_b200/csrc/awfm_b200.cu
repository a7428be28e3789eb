// awfm_b200.cu — C-ABI (include/awfm_gpu.h) over the sm_100a kernels: index residency, launch dispatch,
// packed-batch host/device entry points, and the pipelined search-list engine used by the drop-in shim.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Xcompiler -fPIC,-fopenmp -shared
// There is NO CPU fallback anywhere in this file: every entry point fails when CUDA is unavailable.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <omp.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <cub/device/device_radix_sort.cuh>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "awfm_internal.cuh"

using namespace awfm;

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string gLastError;
int awfm_fail(int code, const char *what, const char *detail) {
  gLastError = what;
  if (detail) {
    gLastError += ": ";
    gLastError += detail;
  }
  return code;
}
int awfm_set_error(int code, const char *what, const char *detail) { return awfm_fail(code, what, detail); }

extern "C" const char *awfm_gpu_last_error(void) { return gLastError.c_str(); }
extern "C" int awfm_gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// ------------------------------------------------------------------------------------------------ lanes
int GrowBuf::ensure(size_t bytes) {
  if (cap >= bytes && p) return AWFM_GPU_OK;
  release();
  const size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = host ? cudaHostAlloc(&p, want, cudaHostAllocPortable) : cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    p = nullptr;
    return awfm_fail(AWFM_GPU_ERR_ALLOC, host ? "pinned host staging allocation failed" : "device scratch allocation failed",
                     cudaGetErrorString(e));
  }
  cap = want;
  return AWFM_GPU_OK;
}
void GrowBuf::release() {
  if (p) {
    if (host) cudaFreeHost(p);
    else cudaFree(p);
  }
  p = nullptr;
  cap = 0;
}

int awfm_lane_prepare(awfm_gpu_ctx *c, Lane &L) {
  (void)c;
  if (L.ready) return AWFM_GPU_OK;
  if (!L.sc.dWorkCounter) CU(cudaMalloc(&L.sc.dWorkCounter, 64));
  if (!L.unpackDone) CU(cudaEventCreateWithFlags(&L.unpackDone, cudaEventDisableTiming));
  for (auto &s : L.slots) {
    if (!s.stream) CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (!s.done) CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    if (!s.offsetsDone) CU(cudaEventCreateWithFlags(&s.offsetsDone, cudaEventDisableTiming));
    if (!s.sc.dWorkCounter) CU(cudaMalloc(&s.sc.dWorkCounter, 64));
  }
  L.ready = true;
  return AWFM_GPU_OK;
}

LaneHold::LaneHold(awfm_gpu_ctx *ctx) : c(ctx) {
  rc = awfm_set_device(c);
  if (rc) return;
  // host-facing calls use lanes 1..kLanes-1 (lane 0 belongs to the asynchronous device-buffer entry points)
  for (int i = 1; i < awfm_gpu_ctx::kLanes && !lane; i++)
    if (c->lanes[i].mu.try_lock()) lane = &c->lanes[i];
  if (!lane) {
    lane = &c->lanes[1];
    lane->mu.lock();
  }
  rc = awfm_lane_prepare(c, *lane);
  c->lastLane.store((int)(lane - c->lanes));
}
LaneHold::~LaneHold() {
  if (lane) lane->mu.unlock();
}
AllLanesHold::AllLanesHold(awfm_gpu_ctx *ctx) : c(ctx) {
  c->mu.lock();
  for (auto &l : c->lanes) l.mu.lock();
}
AllLanesHold::~AllLanesHold() {
  for (auto &l : c->lanes) l.mu.unlock();
  c->mu.unlock();
}

int awfm_set_device(const awfm_gpu_ctx *c) {
  CU(cudaSetDevice(c->device));
  return AWFM_GPU_OK;
}

static uint64_t numSeedsOf(uint8_t alphabet, uint8_t k) {
  uint64_t n = 1;
  for (int i = 0; i < k; i++) n *= (alphabet == 1 ? 20u : 4u);
  return n;
}

static int ctxCreateCommon(awfm_gpu_ctx **out, int device, const awfm_index_view *v, bool fromDevice) {
  if (!out || !v || !v->blocks || !v->prefixSums || !v->seedTable) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  if (v->alphabet < 1 || v->alphabet > 3) return awfm_fail(AWFM_GPU_ERR_ARG, "alphabetType must be 1 (amino), 2 (DNA) or 3 (RNA)");
  if (v->bwtLength < 2 || v->numBlocks != 1 + (v->bwtLength - 1) / 256) return awfm_fail(AWFM_GPU_ERR_ARG, "numBlocks does not match bwtLength");
  if (v->saBytes && (v->saRatio == 0 || v->saBitWidth == 0 || v->saBitWidth > 64)) return awfm_fail(AWFM_GPU_ERR_ARG, "bad SA ratio / bit width");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return awfm_fail(AWFM_GPU_ERR_NO_DEVICE, "no such CUDA device");
  CU(cudaSetDevice(device));
  awfm_gpu_ctx *c = new awfm_gpu_ctx();
  c->device = device;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  c->numSMs = prop.multiProcessorCount;
  const bool amino = v->alphabet == 1;
  const uint32_t rawBlockBytes = amino ? 352u : 160u;
  const uint32_t numPrefix = amino ? 22u : 6u;
  const cudaMemcpyKind kind = fromDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  auto bail = [&](int code) {
    awfm_gpu_ctx_destroy(c);
    return code;
  };
#define CUB_(call)                                                     \
  do {                                                                 \
    cudaError_t e_ = (call);                                           \
    if (e_ != cudaSuccess) {                                           \
      cudaGetLastError();                                              \
      return bail(awfm_fail(e_ == cudaErrorMemoryAllocation ? AWFM_GPU_ERR_ALLOC : AWFM_GPU_ERR_CUDA, #call, cudaGetErrorString(e_))); \
    }                                                                  \
  } while (0)

  // prefix sums first: the nucleotide superblock table folds C[c] in
  uint64_t prefix[24] = {0};
  if (fromDevice) CUB_(cudaMemcpy(prefix, v->prefixSums, numPrefix * 8, cudaMemcpyDeviceToHost));
  else memcpy(prefix, v->prefixSums, numPrefix * 8);

  // blocks -> sectors / quarter-lines.  The raw copy is staged in slabs so peak extra memory stays small.
  const uint64_t lineBytes = amino ? 512u : 128u;  // per 256 positions: 4 amino quarter-lines / 4 nucleotide sectors
  CUB_(cudaMalloc(&c->dLines, v->numBlocks * lineBytes));
  c->deviceBytes += v->numBlocks * lineBytes;
  uint64_t *dSuperCounts = nullptr, *dPrefix = nullptr;
  if (amino) {
    // counts at the first block of every 2^31-position superblock; superC folds the prefix sums C[c] in
    const uint64_t numSuper = ((v->bwtLength - 1) >> kSuperShift) + 1;
    std::vector<uint64_t> superCounts(numSuper * kAminoSuperStride, 0), superC(numSuper * kAminoSuperStride, 0);
    for (uint64_t s = 0; s < numSuper; s++) {
      const uint8_t *src = (const uint8_t *)v->blocks + ((s << kSuperShift) >> 8) * rawBlockBytes + 160;
      if (fromDevice) CUB_(cudaMemcpy(&superCounts[s * kAminoSuperStride], src, 21 * 8, cudaMemcpyDeviceToHost));
      else memcpy(&superCounts[s * kAminoSuperStride], src, 21 * 8);
      for (int l = 0; l < 21; l++) superC[s * kAminoSuperStride + l] = prefix[l] + superCounts[s * kAminoSuperStride + l];
    }
    CUB_(cudaMalloc(&dSuperCounts, numSuper * kAminoSuperStride * 8));
    CUB_(cudaMemcpy(dSuperCounts, superCounts.data(), numSuper * kAminoSuperStride * 8, cudaMemcpyHostToDevice));
    CUB_(cudaMalloc(&c->dSuperC, numSuper * kAminoSuperStride * 8));
    CUB_(cudaMemcpy(c->dSuperC, superC.data(), numSuper * kAminoSuperStride * 8, cudaMemcpyHostToDevice));
  } else {
    // one row per 2^16 positions, filled on the device from the slab that holds the superblock's first block
    const uint64_t numSuper = ((v->numBlocks * 256 - 1) >> kSectorSuperShift) + 1;
    CUB_(cudaMalloc(&c->dXRel16, v->numBlocks * 4 * 2));
    CUB_(cudaMalloc(&c->dSuperC, numSuper * kSectorSuperStride * 8));
    CUB_(cudaMalloc(&dSuperCounts, numSuper * kSectorSuperStride * 8));
    CUB_(cudaMalloc(&dPrefix, 8 * 8));
    CUB_(cudaMemcpy(dPrefix, prefix, 6 * 8, cudaMemcpyHostToDevice));
    c->deviceBytes += v->numBlocks * 8 + numSuper * kSectorSuperStride * 8;
  }
  {
    const uint64_t slabBlocks = std::min<uint64_t>(v->numBlocks, 1u << 20);  // <= 352 MB staging; multiple of 256
    uint8_t *dRaw = nullptr;
    CUB_(cudaMalloc(&dRaw, slabBlocks * rawBlockBytes));
    for (uint64_t b0 = 0; b0 < v->numBlocks; b0 += slabBlocks) {
      const uint64_t nb = std::min(slabBlocks, v->numBlocks - b0);
      cudaError_t e = cudaMemcpy(dRaw, (const uint8_t *)v->blocks + b0 * rawBlockBytes, nb * rawBlockBytes, kind);
      if (e == cudaSuccess) {
        const unsigned grid = (unsigned)((nb + 255) / 256);
        if (amino) {
          relayoutAmino<<<grid, 256>>>(dRaw, nb, b0, dSuperCounts, (uint4 *)c->dLines);
        } else {
          const unsigned superGrid = (unsigned)(((nb + 255) / 256 + 255) / 256);
          sectorSuperRows<<<superGrid, 256>>>(dRaw, nb, b0, dPrefix, dSuperCounts, (uint64_t *)c->dSuperC);
          relayoutNucleotideSectors<<<grid, 256>>>(dRaw, nb, b0, dSuperCounts, (uint4 *)c->dLines,
                                                   (uint16_t *)c->dXRel16);
        }
        e = cudaDeviceSynchronize();
      }
      if (e != cudaSuccess) {
        cudaFree(dRaw);
        cudaFree(dSuperCounts);
        cudaFree(dPrefix);
        CUB_(e);
      }
    }
    cudaFree(dRaw);
    cudaFree(dSuperCounts);
    cudaFree(dPrefix);
  }
  // seed table
  const uint64_t numSeeds = numSeedsOf(v->alphabet, v->seedK);
  CUB_(cudaMalloc(&c->dSeed, numSeeds * 16));
  CUB_(cudaMemcpy(c->dSeed, v->seedTable, numSeeds * 16, kind));
  c->deviceBytes += numSeeds * 16;
  // sampled SA (+ zero padding so the two-word read never leaves the allocation)
  if (v->saBytes) {
    const uint64_t padded = ((v->saByteLength + 15) & ~15ull) + 16;
    CUB_(cudaMalloc(&c->dSa, padded));
    CUB_(cudaMemset(c->dSa, 0, padded));
    CUB_(cudaMemcpy(c->dSa, v->saBytes, v->saByteLength, kind));
    c->deviceBytes += padded;
    c->hasSa = true;
  }
  DevIndex &ix = c->ix;
  ix.lines = (const uint4 *)c->dLines;
  ix.xRel16 = (const uint16_t *)c->dXRel16;
  ix.superC = (const uint64_t *)c->dSuperC;
  ix.seedTable = (const uint4 *)c->dSeed;
  ix.sa = (const uint64_t *)c->dSa;
  ix.numBlocks = v->numBlocks;
  ix.bwtLength = v->bwtLength;
  ix.numSeeds = numSeeds;
  memcpy(ix.prefixSums, prefix, sizeof ix.prefixSums);
  ix.saBitWidth = v->saBitWidth;
  ix.saRatio = v->saRatio ? v->saRatio : 1;
  ix.saRatioShift = 0xFFFFFFFFu;
  if ((ix.saRatio & (ix.saRatio - 1)) == 0) {
    ix.saRatioShift = 0;
    while ((1u << ix.saRatioShift) < ix.saRatio) ix.saRatioShift++;
  }
  ix.seedK = v->seedK;
  ix.amino = amino;
  // defaults measured on B200 (profiles/r01_locate_sweep*.jsonl): amino reads whole quarter-lines with 4-lane groups;
  // nucleotide LF steps use 2-lane groups (one sector per lane), a nucleotide walk is always one thread
  c->countLpq = amino ? 4 : 2;
  c->locateLpq = amino ? 4 : 1;
  ix.deepSeedTable = nullptr;
  ix.deepSeedK = ix.deepSeedWide = 0;
  if (int r = awfm_lane_prepare(c, c->lanes[0])) return bail(r);  // the other lanes are prepared on first use
#undef CUB_
  *out = c;
  return AWFM_GPU_OK;
}

extern "C" int awfm_gpu_ctx_create(awfm_gpu_ctx **ctx, int device, const awfm_index_view *view) {
  return ctxCreateCommon(ctx, device, view, false);
}
extern "C" int awfm_gpu_ctx_create_from_device(awfm_gpu_ctx **ctx, int device, const awfm_index_view *view) {
  return ctxCreateCommon(ctx, device, view, true);
}

// ------------------------------------------------------------------------------------------------ .awfmi loader
// SURVEY.md §8 row f3.  The reference reads an index with fread() into freshly malloc'ed host arrays
// (awFmReadIndexFromFile, src/AwFmFile.c:195-449) and the drop-in then uploads those.  Here the version-8 file
// (layout: src/AwFmFile.c:48-187, offsets :524-558) is mapped read-only and its sections are copied from the page
// cache straight into device staging, re-laid out slab by slab; no host copy of the index is ever built.
struct AwfmiSections {
  uint64_t fileBytes, bwtLength, numBlocks, blockBytes, blocksOff, prefixOff, numPrefix, seedOff, numSeeds, saOff,
      saBytes, headerOff, headerBytes, metaOff, numSequences;
  uint32_t version, featureFlags;
  uint8_t saRatio, seedK, alphabet, storesSequence, saBitWidth;
};

static int parseAwfmi(const uint8_t *f, uint64_t fileBytes, AwfmiSections *o) {
  memset(o, 0, sizeof *o);
  o->fileBytes = fileBytes;
  if (fileBytes < 30 || memcmp(f, "AwFmIndex\n", 10) != 0) return awfm_fail(AWFM_GPU_ERR_ARG, "not an .awfmi file (bad magic)");
  memcpy(&o->version, f + 10, 4);
  memcpy(&o->featureFlags, f + 14, 4);
  if (o->version != 8) return awfm_fail(AWFM_GPU_ERR_ARG, "unsupported .awfmi version (this loader reads version 8)");
  o->saRatio = f[18], o->seedK = f[19], o->alphabet = f[20], o->storesSequence = f[21];
  memcpy(&o->bwtLength, f + 22, 8);
  if (o->alphabet < 1 || o->alphabet > 3 || o->saRatio == 0 || o->bwtLength < 2) return awfm_fail(AWFM_GPU_ERR_ARG, "corrupt .awfmi header");
  const bool amino = o->alphabet == 1;
  o->numBlocks = 1 + (o->bwtLength - 1) / 256;
  o->blockBytes = amino ? 352 : 160;
  o->numPrefix = amino ? 22 : 6;
  if (o->seedK > (amino ? 14 : 31)) return awfm_fail(AWFM_GPU_ERR_ARG, "corrupt .awfmi header (seed length)");
  o->numSeeds = numSeedsOf(o->alphabet, o->seedK);
  o->blocksOff = 30;
  o->prefixOff = o->blocksOff + o->numBlocks * o->blockBytes;
  o->seedOff = o->prefixOff + o->numPrefix * 8;
  o->saOff = o->seedOff + o->numSeeds * 16 + (o->storesSequence ? o->bwtLength - 1 : 0);
  o->saBitWidth = (uint8_t)(64 - __builtin_clzll(o->bwtLength - 1 ? o->bwtLength - 1 : 1));  // src/AwFmSuffixArray.c:12-18
  const uint64_t samples = (o->bwtLength + o->saRatio - 1) / o->saRatio;                     // :144-147
  o->saBytes = (samples * o->saBitWidth + 7) / 8 + 8;                                        // :41-53
  uint64_t end = o->saOff + o->saBytes;
  if (end > fileBytes) return awfm_fail(AWFM_GPU_ERR_ARG, "truncated .awfmi file");
  if (o->featureFlags & 1u) {  // FastaVector section: header length, record count, header chars, record table
    if (end + 16 > fileBytes) return awfm_fail(AWFM_GPU_ERR_ARG, "truncated .awfmi file (FastaVector section)");
    memcpy(&o->headerBytes, f + end, 8);
    memcpy(&o->numSequences, f + end + 8, 8);
    o->headerOff = end + 16;
    o->metaOff = o->headerOff + o->headerBytes;
    if (o->metaOff < o->headerOff || o->numSequences > (fileBytes - o->metaOff) / 16 || o->metaOff > fileBytes)
      return awfm_fail(AWFM_GPU_ERR_ARG, "truncated .awfmi file (FastaVector section)");
  }
  return AWFM_GPU_OK;
}

extern "C" int awfm_gpu_ctx_create_from_file(awfm_gpu_ctx **ctx, int device, const char *path, int wantSuffixArray,
                                             awfm_file_info *info) {
  if (!ctx || !path) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return awfm_fail(AWFM_GPU_ERR_ARG, "cannot open index file", path);
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 30) {
    close(fd);
    return awfm_fail(AWFM_GPU_ERR_ARG, "not an .awfmi file (too short)", path);
  }
  const uint64_t bytes = (uint64_t)st.st_size;
  void *map = mmap(nullptr, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return awfm_fail(AWFM_GPU_ERR_ALLOC, "mmap of the index file failed", path);
  madvise(map, bytes, MADV_SEQUENTIAL);
  const uint8_t *f = (const uint8_t *)map;
  AwfmiSections sec;
  int rc = parseAwfmi(f, bytes, &sec);  // format errors are reported before any CUDA call
  if (rc == AWFM_GPU_OK) {
    uint64_t prefix[24] = {0};
    memcpy(prefix, f + sec.prefixOff, sec.numPrefix * 8);  // unaligned in the file
    awfm_index_view v;
    memset(&v, 0, sizeof v);
    v.blocks = f + sec.blocksOff;
    v.numBlocks = sec.numBlocks;
    v.prefixSums = prefix;
    v.seedTable = f + sec.seedOff;
    v.saBytes = wantSuffixArray ? f + sec.saOff : nullptr;
    v.saByteLength = sec.saBytes;
    v.bwtLength = sec.bwtLength;
    v.saBitWidth = sec.saBitWidth;
    v.saRatio = sec.saRatio;
    v.seedK = sec.seedK;
    v.alphabet = sec.alphabet;
    rc = ctxCreateCommon(ctx, device, &v, false);
    if (rc == AWFM_GPU_OK && sec.numSequences) {
      std::vector<uint64_t> meta(sec.numSequences * 2);
      memcpy(meta.data(), f + sec.metaOff, sec.numSequences * 16);
      rc = awfm_gpu_ctx_set_sequences(*ctx, meta.data(), sec.numSequences);
      if (rc != AWFM_GPU_OK) {
        awfm_gpu_ctx_destroy(*ctx);
        *ctx = nullptr;
      }
    }
  }
  if (info) {
    info->bwtLength = sec.bwtLength;
    info->numSequences = sec.numSequences;
    info->suffixArrayByteLength = sec.saBytes;
    info->versionNumber = sec.version;
    info->featureFlags = sec.featureFlags;
    info->suffixArrayCompressionRatio = sec.saRatio;
    info->kmerLengthInSeedTable = sec.seedK;
    info->alphabetType = sec.alphabet;
    info->storeOriginalSequence = sec.storesSequence;
  }
  munmap(map, bytes);
  return rc;
}

static void freeScratch(LocateScratch &sc) {
  cudaFree(sc.scanTemp);
  cudaFree(sc.dWorkCounter);
  sc = LocateScratch();
}

static void freeSweep(SweepScratch &w) {
  cudaFree(w.arena);
  cudaFree(w.ctrl);
  cudaFree(w.sortCtrl);
  cudaFree(w.sortTemp);
  if (w.done) cudaEventDestroy(w.done);
  for (int i = 0; i < w.numStages; i++) cudaEventDestroy(w.stage[i]);
  w = SweepScratch();
}

static void freeSlot(PipeSlot &s) {
  freeScratch(s.sc);
  if (s.hHit) cudaFreeHost(s.hHit);
  if (s.hPos) cudaFreeHost(s.hPos);
  cudaFree(s.dHit);
  cudaFree(s.dPos);
  if (s.offsetsDone) cudaEventDestroy(s.offsetsDone);
  if (s.hLetters) cudaFreeHost(s.hLetters);
  if (s.hOffsets) cudaFreeHost(s.hOffsets);
  if (s.hCounts) cudaFreeHost(s.hCounts);
  free(s.hOld);
  cudaFree(s.dLetters);
  cudaFree(s.dOffsets);
  cudaFree(s.dCounts);
  cudaFree(s.dRanges);
  if (s.stream) cudaStreamDestroy(s.stream);
  if (s.done) cudaEventDestroy(s.done);
  s = PipeSlot();
}

static void freeLane(Lane &L) {
  for (auto &s : L.slots) freeSlot(s);
  for (auto &e : L.kernelEvents) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  L.kernelEvents.clear();
  freeScratch(L.sc);
  freeSweep(L.sweep);
  if (L.hBigPos) cudaFreeHost(L.hBigPos);
  cudaFree(L.dBigPos);
  L.hBigPos = L.dBigPos = nullptr, L.bigPosCap = 0;
  if (L.unpackDone) cudaEventDestroy(L.unpackDone);
  L.unpackDone = nullptr;
  for (GrowBuf *b : {&L.dLetters, &L.dOffsets, &L.dCounts, &L.dRanges, &L.dHits, &L.dPositions, &L.dUnpacked}) b->release();
  L.ready = false;
}

extern "C" void awfm_gpu_ctx_destroy(awfm_gpu_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  {
    AllLanesHold all(c);  // waits for calls still in flight on any lane
    for (auto &l : c->lanes) freeLane(l);
  }
  cudaFree(c->dLines);
  cudaFree(c->dXRel16);
  cudaFree(c->dSuperC);
  cudaFree(c->dSeed);
  cudaFree(c->dSa);
  cudaFree(c->dSequenceEnds);
  cudaFree(c->dDeepSeed);
  cudaFree(c->dDenseSa);
  cudaGetLastError();
  delete c;
}

extern "C" uint64_t awfm_gpu_ctx_device_bytes(const awfm_gpu_ctx *c) { return c ? c->deviceBytes.load() : 0; }

extern "C" int awfm_gpu_ctx_set_tuning(awfm_gpu_ctx *c, const char *key, int64_t value) {
  if (!c || !key) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  const std::string k(key);
  auto lpqOk = [](int64_t v) { return v == 1 || v == 2 || v == 4 || v == 8; };
  if (k == "count_lpq" && lpqOk(value)) c->countLpq = (int)value;
  else if (k == "locate_lpq" && lpqOk(value)) c->locateLpq = (int)value;
  else if (k == "count_variant" && (value == 0 || value == 1)) c->countVariant = (int)value;
  else if (k == "locate_variant" && (value == 0 || value == 1)) c->locateVariant = (int)value;
  else if (k == "blocks_per_sm" && value >= 0 && value <= 32) c->blocksPerSm = (int)value;
  else if (k == "sweep_min_queries") c->sweepMinQueries = value;
  else if (k == "sweep_max_batch" && value >= 256 && value <= (1ll << 30)) c->sweepMaxBatch = value;
  else if (k == "sweep_sort_bits" && value >= 0 && value <= 32) c->sweepSortBits = (int)value;
  else if (k == "sweep_profile" && (value == 0 || value == 1)) c->sweepProfile = (int)value;
  else if (k == "sweep_own_sort" && (value == 0 || value == 1)) c->sweepOwnSort = (int)value;
  else if (k == "sweep_record12" && (value == 0 || value == 1)) c->sweepRecord12 = (int)value;
  else if (k == "sweep_ordered_emit" && (value == 0 || value == 1)) c->sweepOrderedEmit = (int)value;
  else if (k == "sweep_compact_pairs" && (value == 0 || value == 1)) c->sweepCompactPairs = (int)value;
  else if (k == "sweep_variable" && (value == 0 || value == 1)) c->sweepVariable = (int)value;
  else if (k == "sweep_wide" && (value == 0 || value == 1)) c->sweepWide = (int)value;
  else if (k == "sweep_local_bits" && value >= -1 && value <= 8) c->sweepLocalBits = (int)value;
  else if (k == "sweep_items" && (value == 1 || value == 2 || value == 4 || value == 8)) c->sweepItems = (int)value;
  else if (k == "sweep_first_items" && (value == 1 || value == 2 || value == 4 || value == 8)) c->sweepFirstItems = (int)value;
  else if (k == "chunk_queries" && (value == 0 || (value >= 64 && value <= (1ll << 30)))) c->chunkQueries = value;
  else if (k == "locate_chunk_queries" && value >= 64 && value <= (1ll << 30)) c->locateChunkQueries = value;
  else if (k == "locate_inline_hits" && value >= 0 && value <= (1ll << 32)) c->locateInlineHits = value;
  else if (k == "locate_window_hits" && value >= 1 && value <= (1ll << 32)) c->locateWindowHits = value;
  else if (k == "use_deep_seed_table" && (value == 0 || value == 1)) {  // A/B switch for a table already derived
    const bool on = value && c->dDeepSeed;
    c->ix.deepSeedTable = on ? c->dDeepSeed : nullptr;
    c->ix.deepSeedK = on ? c->deepSeedKBuilt : 0;
  }
  else return awfm_fail(AWFM_GPU_ERR_ARG, "unknown tuning key or bad value", key);
  return AWFM_GPU_OK;
}

// ---- kernel event bookkeeping (device time of OUR kernels on the launching stream) ----
void awfm_begin_call(Lane &L) {
  L.eventsUsed = 0;
  L.stats = awfm_gpu_stats{};
}
static EventPair *nextEvents(Lane &L) {
  if (L.eventsUsed == L.kernelEvents.size()) {
    EventPair p;
    if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return nullptr;
    L.kernelEvents.push_back(p);
  }
  return &L.kernelEvents[L.eventsUsed++];
}

// Stats of the most recent call on the context (the lane it ran on).  Not meaningful while another call is in flight.
extern "C" int awfm_gpu_ctx_get_stats(const awfm_gpu_ctx *cc, awfm_gpu_stats *out) {
  awfm_gpu_ctx *c = const_cast<awfm_gpu_ctx *>(cc);
  if (!c || !out) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  if (awfm_set_device(c)) return AWFM_GPU_ERR_CUDA;
  Lane &L = c->lanes[c->lastLane.load()];
  double ms = 0;
  for (size_t i = 0; i < L.eventsUsed; i++) {
    CU(cudaEventSynchronize(L.kernelEvents[i].b));
    float f = 0;
    CU(cudaEventElapsedTime(&f, L.kernelEvents[i].a, L.kernelEvents[i].b));
    ms += f;
  }
  L.stats.kernelMs = ms;
  *out = L.stats;
  return AWFM_GPU_OK;
}

// ------------------------------------------------------------------------------------------------ launches
template <typename K>
static int gridFor(awfm_gpu_ctx *c, K kernel, int threads, int *grid, size_t dynamicSmem = 0) {
  int perSm = c->blocksPerSm;
  if (perSm == 0 || dynamicSmem) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, threads, dynamicSmem));
  if (perSm < 1) perSm = 1;
  *grid = c->numSMs * perSm;  // whole multiples of the SM count: persistent grid-stride CTAs
  return AWFM_GPU_OK;
}

constexpr int kTile = 256, kTileLetterBytes = 8192;

template <int LPQ, bool AMINO>
static int launchCount(awfm_gpu_ctx *c, const QueryBatch &qb, uint32_t *dCounts, uint4 *dRanges, cudaStream_t st) {
  int grid = 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(qb.letters) & 15u) == 0;
  if (c->countVariant == 1 && aligned) {
    auto k = countKernelV1<LPQ, AMINO, kTile, kTileLetterBytes>;
    if (int r = gridFor(c, k, 256, &grid)) return r;
    const uint64_t tiles = (qb.numQueries + kTile - 1) / kTile;
    grid = (int)std::min<uint64_t>((uint64_t)grid, tiles);
    k<<<grid, 256, 0, st>>>(c->ix, qb, dCounts, dRanges);
  } else {
    auto k = countKernelV0<LPQ, AMINO>;
    if (int r = gridFor(c, k, 256, &grid)) return r;
    const uint64_t need = (qb.numQueries * LPQ + 255) / 256;
    grid = (int)std::min<uint64_t>((uint64_t)grid, need);
    k<<<grid, 256, 0, st>>>(c->ix, qb, dCounts, dRanges);
  }
  CU(cudaGetLastError());
  return AWFM_GPU_OK;
}

// the walk itself: dPos holds BWT positions on entry, text positions on exit
template <int LPQ, bool AMINO>
static int launchWalk(awfm_gpu_ctx *c, LocateScratch &sc, uint64_t numHits, uint64_t *dPos, cudaStream_t st) {
  int grid = 0;
  if (c->locateVariant == 1) {  // group per hit with refill from a chunk dispenser
    auto k = locateKernelRefill<LPQ, AMINO>;
    if (int r = gridFor(c, k, 256, &grid)) return r;
    const uint64_t need = (numHits * LPQ + 255) / 256;
    grid = (int)std::min<uint64_t>((uint64_t)grid, need);
    CU(cudaMemsetAsync(sc.dWorkCounter, 0, sizeof(unsigned long long), st));
    k<<<grid, 256, 0, st>>>(c->ix, numHits, dPos, sc.dWorkCounter);
    CU(cudaGetLastError());
    return AWFM_GPU_OK;
  }
  auto k = locateKernel<LPQ, AMINO>;
  if (int r = gridFor(c, k, 256, &grid)) return r;
  const uint64_t need = (numHits * LPQ + 255) / 256;
  grid = (int)std::min<uint64_t>((uint64_t)grid, need);
  k<<<grid, 256, 0, st>>>(c->ix, numHits, dPos);
  CU(cudaGetLastError());
  return AWFM_GPU_OK;
}

template <int LPQ, bool AMINO>
static int launchLocate(awfm_gpu_ctx *c, LocateScratch &sc, const uint4 *dRanges, const uint64_t *dHitOffsets,
                        uint64_t n, uint64_t hb, uint64_t he, uint64_t *dPos, cudaStream_t st) {
  {  // BWT start position of every hit of the window, written into the output buffer itself
    const uint64_t warps = (n + 31) / 32;
    const int g = (int)std::min<uint64_t>((warps + 7) / 8, (uint64_t)c->numSMs * 8);
    expandHits<<<std::max(g, 1), 256, 0, st>>>(dRanges, dHitOffsets, n, hb, he, dPos);
    CU(cudaGetLastError());
  }
  return launchWalk<LPQ, AMINO>(c, sc, he - hb, dPos, st);
}

// lanes per query / per hit: amino quarter-lines are read by 1, 2 or 4 lanes; a nucleotide LF step by 1 or 2 lanes
// (one sector each), a nucleotide walk by one thread.  Larger requests are clamped.
#define DISPATCH_COUNT(fn, lpq, amino, ...)                                \
  ((amino) ? ((lpq) == 1   ? fn<1, true>(__VA_ARGS__)                      \
              : (lpq) == 2 ? fn<2, true>(__VA_ARGS__)                      \
                           : fn<4, true>(__VA_ARGS__))                     \
           : ((lpq) == 1   ? fn<1, false>(__VA_ARGS__)                     \
                           : fn<2, false>(__VA_ARGS__)))
#define DISPATCH_LOCATE(fn, lpq, amino, ...)                               \
  ((amino) ? ((lpq) == 1   ? fn<1, true>(__VA_ARGS__)                      \
              : (lpq) == 2 ? fn<2, true>(__VA_ARGS__)                      \
                           : fn<4, true>(__VA_ARGS__))                     \
           : fn<1, false>(__VA_ARGS__))

static int locateDeviceRaw(awfm_gpu_ctx *c, uint64_t numHits, uint64_t *dPos, cudaStream_t st) {
  if (numHits == 0) return AWFM_GPU_OK;
  return DISPATCH_LOCATE(launchWalk, c->locateLpq, c->ix.amino != 0, c, c->lanes[0].sc, numHits, dPos, st);
}

// ---- sweep count path (awfm_sweep.cuh): large fixed-length nucleotide batches, counts only ----
static uint32_t sweepSeedK(const awfm_gpu_ctx *c, uint32_t len) {
  return (c->ix.deepSeedK && len >= c->ix.deepSeedK) ? c->ix.deepSeedK : c->ix.seedK;
}
static uint32_t sweepKeyBits(const awfm_gpu_ctx *c, uint32_t k) {  // bits of a seed-table index of depth k
  if (!c->ix.amino) return 2 * k;
  uint64_t entries = 1;
  for (uint32_t i = 0; i < k; i++) entries *= 20;
  uint32_t bits = 0;
  while ((1ull << bits) < entries) bits++;
  return bits;
}
static bool sweepEligible(const awfm_gpu_ctx *c, const uint8_t *dLetters, const uint64_t *dOffsets, uint32_t len,
                          uint64_t n, const awfm_range *dRanges) {
  if (c->sweepMinQueries < 0 || c->countVariant != 1) return false;
  if ((reinterpret_cast<uintptr_t>(dLetters) & 15u) != 0) return false;
  if (n >= 0x70000000ull) return false;  // 32-bit record indices
  // 32-bit positions; nucleotide indexes beyond that take the passes with 64-bit positions (40 bits in a record)
  if (c->ix.bwtLength >= 0xFFFFFFF0ull && (c->ix.amino || c->ix.bwtLength >= (1ull << 40))) return false;
  if (dOffsets) {  // variable lengths (sweepPackVar): the index's own seed table only; lengths are checked per query
    if (!c->sweepVariable || c->ix.deepSeedK || c->ix.seedK == 0 || c->ix.seedK > (c->ix.amino ? 7u : 16u)) return false;
  } else {
    const uint32_t k = sweepSeedK(c, len);
    if (c->ix.amino) {  // 5 bits per remaining letter in a 32-bit payload, 20^k seed entries in a 32-bit key
      if (k == 0 || k > 7 || len < k || len - k > 6 || len > 64) return false;
    } else if (k == 0 || k > 16 || len < k || len - k > kSweepMaxRestNuc || len > 32) {
      return false;  // (17..24 letters left of the seed: sweepRefill; the pack kernels hold a query in 64 bits)
    }
  }
  if (c->sweepMinQueries > 0) return n >= (uint64_t)c->sweepMinQueries;
  // automatic: pays off once the batch puts about one query on every second 128-B line of the index (measured
  // break-even at 3.1 Gbp: 6 M queries for 16- and 20-mers alike, profiles/r02_sweep_probe.jsonl; round 1's slower
  // sweep: 12 M).
  // (with range output every query also pays a scattered 16-B store: the break-even stays where round 1 measured it;
  // with a derived deep seed table the tile kernel has fewer steps left to pay for: twice the batch)
  // (range output that leaves through the ordered emit — fixed-length nucleotide batches of up to 2^24 queries — costs
  // less than the scattered stores did: locate's front end for 10 M 16-mers at 3.1 Gbp 1.34 ms through the sweep, 1.41
  // through the tile kernel; threshold at three quarters of the old one)
  const bool deepActive = !dOffsets && c->ix.deepSeedK && len >= c->ix.deepSeedK;
  const bool emits = dRanges && !c->ix.amino && !dOffsets && c->sweepOrderedEmit && n <= (1ull << 24);
  const uint64_t perLine = emits ? (c->ix.bwtLength >> 9) + (c->ix.bwtLength >> 10)
                                 : c->ix.bwtLength >> (c->ix.amino ? 6 : dRanges ? 8 : 9);
  return n >= std::max<uint64_t>(1ull << 22, perLine) << (deepActive ? 1 : 0);
}

static int ensureSweep(awfm_gpu_ctx *c, Lane &L, uint64_t n, int arrays) {
  SweepScratch &w = L.sweep;
  if (!w.done) CU(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
  while (w.numStages < kSweepMaxPasses + 4) {
    CU(cudaEventCreate(&w.stage[w.numStages]));
    w.numStages++;
  }
  if (!w.ctrl) CU(cudaMalloc(&w.ctrl, (kSweepMaxPasses * kSweepCtrlStride + 4) * sizeof(uint32_t)));
  if (!w.sortCtrl) CU(cudaMalloc(&w.sortCtrl, sizeof(SortCtrl) + 2 * (sizeof(uint32_t) << (2 * kSortMaxDigitBits))));
  if (w.cap >= n && w.arrays == arrays) return AWFM_GPU_OK;
  CU(cudaDeviceSynchronize());
  cudaFree(w.arena);
  w.arena = nullptr;
  c->deviceBytes -= w.bytes;
  w.bytes = 0;
  w.cap = 0;
  const uint64_t cap = (n + (n >> 4) + 1024 + 15) & ~15ull;  // multiple of 16 records: every carved buffer 64-B aligned
  const uint64_t bytes = cap * (2 * 4 + 2 * 8 + 2 * (uint64_t)arrays * 16 + 4 + 4);
  if (cudaMalloc(&w.arena, bytes) != cudaSuccess) {
    cudaGetLastError();
    w.arena = nullptr;
    return awfm_fail(AWFM_GPU_ERR_ALLOC, "sweep scratch does not fit in device memory");
  }
  uint8_t *p = static_cast<uint8_t *>(w.arena);
  for (int g = 0; g < 2; g++)
    for (int a = 0; a < arrays; a++) w.recs[g][a] = reinterpret_cast<uint4 *>(p), p += cap * 16;
  for (int i = 0; i < 2; i++) w.vals[i] = reinterpret_cast<uint64_t *>(p), p += cap * 8;
  for (int i = 0; i < 2; i++) w.keys[i] = reinterpret_cast<uint32_t *>(p), p += cap * 4;
  w.irregularIds = reinterpret_cast<uint32_t *>(p), p += cap * 4;
  w.more = reinterpret_cast<uint32_t *>(p);
  w.cap = cap;
  w.arrays = arrays;
  w.bytes = bytes;
  c->deviceBytes += bytes;
  return AWFM_GPU_OK;
}

// dOffsets != nullptr: a variable-length batch (n+1 letter offsets into dLetters); `len` is then the longest query of
// the batch if the caller knows it, else 0.
template <bool AMINO>
static int sweepCountBatch(awfm_gpu_ctx *c, Lane &L, const uint8_t *dLetters, const uint64_t *dOffsets, uint32_t format,
                           uint32_t len, uint64_t n, uint32_t *dCounts, uint4 *dRanges, bool hitsOnly, cudaStream_t st) {
  SweepScratch &w = L.sweep;
  const bool variable = dOffsets != nullptr;
  constexpr uint32_t kVarMaxRest = AMINO ? kSweepVarMaxRestAmino : kSweepVarMaxRestNuc;
  const uint32_t k = variable ? c->ix.seedK : sweepSeedK(c, len);
  // variable: as many passes as the longest query the records can hold needs (a pass over no records costs a launch)
  const uint32_t steps = !variable ? len - k : (len > k ? std::min(len - k, kVarMaxRest) : len ? 1u : kVarMaxRest);
  const bool deep = !variable && c->ix.deepSeedK && len >= c->ix.deepSeedK;
  int stage = 0;
  auto mark = [&]() {
    if (c->sweepProfile && stage < w.numStages) cudaEventRecord(w.stage[stage++], st);
  };
  CU(cudaStreamWaitEvent(st, w.done, 0));
  mark();
  CU(cudaMemsetAsync(dCounts, 0, n * sizeof(uint32_t), st));
  CU(cudaMemsetAsync(w.ctrl, 0, (kSweepMaxPasses * kSweepCtrlStride + 4) * sizeof(uint32_t), st));
  uint32_t *irregularCount = w.ctrl + kSweepMaxPasses * kSweepCtrlStride;
  // ---- how the pairs get ordered: the top bits of the key globally, up to 8 more inside each tile of the first pass ----
  const int endBit = (int)sweepKeyBits(c, k);
  // Low key bits left to the first pass's tile-local sort.  Automatic: as many as keep a group of equal upper bits
  // within about two 1024-pair tiles (beyond that the tiles of a group interleave too many runs for the warps to
  // coalesce), then trimmed to what saves a whole 8-bit radix pass.
  int wantLocal = c->sweepLocalBits;
  if (wantLocal < 0) {
    wantLocal = 0;
    while (wantLocal < 8 && wantLocal < endBit && (double)n * (double)(2u << wantLocal) <= 2048.0 * (double)(1ull << endBit))
      wantLocal++;
    const int radixPasses = (endBit - wantLocal + 7) / 8;
    wantLocal = std::max(0, endBit - 8 * radixPasses);
  }
  uint32_t localBits = (uint32_t)std::min({wantLocal, endBit, 8});
  int beginBit = std::max(0, endBit - std::min(c->sweepSortBits, endBit - (int)localBits));
  // Seed tables deeper than 2^24 entries (k = 13..16, the index's own or a derived one): the two bucket passes order
  // the top 16 key bits, the tiles of the first pass the 8 below them, and the lowest endBit - 24 bits stay unordered —
  // neighbouring seed entries are neighbouring ranges of the BWT, so that costs no locality worth a third pass (or
  // CUB's four: 2.5 ms instead of 1.3 per 100 M pairs).
  uint32_t localShift = 0;
  if (c->sweepLocalBits < 0 && c->sweepSortBits >= 32 && c->sweepOwnSort && endBit > 3 * kSortMaxDigitBits) {
    localBits = kSortMaxDigitBits;
    localShift = (uint32_t)endBit - 3 * kSortMaxDigitBits;
    beginBit = endBit - 2 * kSortMaxDigitBits;
  }
  const int sortedBits = endBit - beginBit;
  // our own two bucket passes (awfm_sort.cuh) order up to 16 bits; deeper seed tables go through CUB
  const bool ownSort = c->sweepOwnSort && sortedBits >= 1 && sortedBits <= 2 * kSortMaxDigitBits;
  const uint32_t dA = ownSort ? (uint32_t)(sortedBits <= kSortMaxDigitBits ? sortedBits : sortedBits - sortedBits / 2) : 0u;
  const uint32_t dB = ownSort ? (uint32_t)sortedBits - dA : 0u;
  const uint32_t shiftA = (uint32_t)endBit - dA, shiftB = (uint32_t)beginBit;
  SortCtrl *sortCtrl = ownSort ? reinterpret_cast<SortCtrl *>(w.sortCtrl) : nullptr;
  uint32_t *countB = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(w.sortCtrl) + sizeof(SortCtrl));
  uint32_t *cursorB = countB + (1u << (2 * kSortMaxDigitBits));
  if (ownSort) CU(cudaMemsetAsync(w.sortCtrl, 0, sizeof(SortCtrl) + (dB ? sizeof(uint32_t) << (dA + dB) : 0), st));
  // Compact pairs between the pack kernel and the second bucket pass (awfm_sort.cuh): one 8-byte word per pair when key,
  // payload and query id fit it once the first digit has left the key.
  SortCompact compact{};
  compact.keyBits = (uint32_t)endBit, compact.lowBits = (uint32_t)endBit - dA;
  compact.restBits = variable ? SweepAlphabet<AMINO>::kLetterBits * kVarMaxRest + 1u : SweepAlphabet<AMINO>::kLetterBits * steps;
  compact.idBits = 1;
  while (compact.idBits < 32 && n > (1ull << compact.idBits) - 1ull) compact.idBits++;  // all ones = irregular query
  // more than 16 letters left of the seed (nucleotide, fixed length): letters 17.. wait in w.more for sweepRefill
  uint32_t *more = (!AMINO && !variable && steps > 16) ? w.more : nullptr;
  const bool compactPairs = ownSort && dB > 0 && !more && c->sweepCompactPairs && compact.keyBits + compact.restBits <= 63 &&
                            compact.lowBits + compact.restBits + compact.idBits <= 64 && n <= (1ull << compact.idBits) - 1ull;
  const uint32_t compactShift = compactPairs ? compact.keyBits : 0u;
  {
    const uint64_t tiles = (n + 255) / 256;
    const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)c->numSMs * 8);
    if (variable) {
      sweepPackVar<AMINO><<<grid, 256, 0, st>>>(dLetters, dOffsets, n, k, w.keys[0], w.vals[0], w.irregularIds, irregularCount, sortCtrl, shiftA, compactShift);
    } else if (format == AWFM_QUERY_2BIT) {  // nucleotide only (sweepEligible): the packed bytes straight into (key, payload)
      const size_t smem = ((size_t)256 * ((len + 3) / 4) + 15) & ~(size_t)15;
      sweepPackBits<<<grid, 256, smem, st>>>(dLetters, n, len, k, w.keys[0], w.vals[0], sortCtrl, shiftA, compactShift, more);
    } else if (AMINO && len % 4 == 0 && len <= 12) {
      const uint32_t *words = reinterpret_cast<const uint32_t *>(dLetters);
      switch (len / 4) {
        case 1: sweepPackWordsAmino<1><<<grid, 256, 0, st>>>(words, n, k, w.keys[0], w.vals[0], w.irregularIds, irregularCount, sortCtrl, shiftA, compactShift); break;
        case 2: sweepPackWordsAmino<2><<<grid, 256, 0, st>>>(words, n, k, w.keys[0], w.vals[0], w.irregularIds, irregularCount, sortCtrl, shiftA, compactShift); break;
        default: sweepPackWordsAmino<3><<<grid, 256, 0, st>>>(words, n, k, w.keys[0], w.vals[0], w.irregularIds, irregularCount, sortCtrl, shiftA, compactShift); break;
      }
    } else if (!AMINO && len % 4 == 0) {
      const uint32_t *words = reinterpret_cast<const uint32_t *>(dLetters);
      switch (len / 4) {
#define AWFM_PACK_CASE(W)                                                                                         \
  case W:                                                                                                         \
    sweepPackWords<W><<<grid, 256, 0, st>>>(words, n, k, w.keys[0], w.vals[0], w.irregularIds, irregularCount, sortCtrl, shiftA, compactShift, more); \
    break;
        AWFM_PACK_CASE(1) AWFM_PACK_CASE(2) AWFM_PACK_CASE(3) AWFM_PACK_CASE(4) AWFM_PACK_CASE(5) AWFM_PACK_CASE(6)
        AWFM_PACK_CASE(7) AWFM_PACK_CASE(8)
#undef AWFM_PACK_CASE
        default: return awfm_fail(AWFM_GPU_ERR_ARG, "sweep: query length out of range");
      }
    } else {
      const size_t smem = ((size_t)256 * len + 15) & ~(size_t)15;
      sweepPack<AMINO><<<grid, 256, smem, st>>>(dLetters, n, len, k, w.keys[0], w.vals[0], w.irregularIds, irregularCount, sortCtrl, shiftA, compactShift, more);
    }
    CU(cudaGetLastError());
  }
  mark();
  int cur = 0;
  uint32_t sortLaunches = 0;
  if (ownSort) {
    int grid = 0;
    if (kSortSmemBytes + 6 * 1024 > 48 * 1024) {  // more than the default dynamic shared memory: opt in (once per device is enough)
      CU(cudaFuncSetAttribute(sortPass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmemBytes));
      CU(cudaFuncSetAttribute(sortPass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmemBytes));
    }
    if (kSortCompactSmemBytes + 6 * 1024 > 48 * 1024) {
      CU(cudaFuncSetAttribute(sortPassCompact<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortCompactSmemBytes));
      CU(cudaFuncSetAttribute(sortPassCompact<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortCompactSmemBytes));
    }
    const uint32_t tilePairs = compactPairs ? (uint32_t)kSortCompactTile : (uint32_t)kSortTile;
    if (compactPairs) {
      if (int r = gridFor(c, sortPassCompact<false>, kSortThreads, &grid, kSortCompactSmemBytes)) return r;
    } else if (int r = gridFor(c, sortPass<false>, kSortThreads, &grid, kSortSmemBytes)) {
      return r;
    }
    grid = (int)std::min<uint64_t>((uint64_t)grid, (n + tilePairs - 1) / tilePairs);
    sortBases<0><<<1, kSortThreads, 0, st>>>(sortCtrl, countB, cursorB, dA, dB, tilePairs);
    if (compactPairs) {  // (dB > 0) pack words in vals[0] -> pass A words in vals[1] -> (key, payload | id) pairs in keys[0] / vals[0]
      int countGrid = 0;
      if (int r = gridFor(c, sortDigitCounts<true, kSortCompactItems>, kSortThreads, &countGrid)) return r;
      countGrid = (int)std::min<uint64_t>((uint64_t)countGrid, (n + tilePairs - 1) / tilePairs + (1u << dA));
      sortPassCompact<false><<<grid, kSortThreads, kSortCompactSmemBytes, st>>>(w.vals[0], w.vals[1], nullptr, nullptr, (uint32_t)n,
                                                                               sortCtrl, cursorB, dA, dB, shiftA, compact);
      sortDigitCounts<true, kSortCompactItems><<<countGrid, kSortThreads, 0, st>>>(w.vals[1], sortCtrl, countB, dA, dB, shiftB);
      sortBases<1><<<1u << dA, kSortThreads, 0, st>>>(sortCtrl, countB, cursorB, dA, dB, tilePairs);
      sortPassCompact<true><<<grid, kSortThreads, kSortCompactSmemBytes, st>>>(w.vals[1], nullptr, w.keys[0], w.vals[0], (uint32_t)n,
                                                                              sortCtrl, cursorB, dA, dB, shiftB, compact);
      cur = 0;
      sortLaunches = 5;
    } else {
      sortPass<false><<<grid, kSortThreads, kSortSmemBytes, st>>>(w.keys[0], w.vals[0], w.keys[1], w.vals[1], (uint32_t)n, sortCtrl,
                                                                 cursorB, dA, dB, shiftA);
      cur = 1;
      sortLaunches = 2;
      if (dB) {
        int countGrid = 0;
        if (int r = gridFor(c, sortDigitCounts<false, kSortItems>, kSortThreads, &countGrid)) return r;
        countGrid = (int)std::min<uint64_t>((uint64_t)countGrid, (n + tilePairs - 1) / tilePairs + (1u << dA));
        sortDigitCounts<false, kSortItems><<<countGrid, kSortThreads, 0, st>>>(w.keys[1], sortCtrl, countB, dA, dB, shiftB);
        sortBases<1><<<1u << dA, kSortThreads, 0, st>>>(sortCtrl, countB, cursorB, dA, dB, tilePairs);
        sortPass<true><<<grid, kSortThreads, kSortSmemBytes, st>>>(w.keys[1], w.vals[1], w.keys[0], w.vals[0], (uint32_t)n, sortCtrl,
                                                                  cursorB, dA, dB, shiftB);
        cur = 0;
        sortLaunches = 5;
      }
    }
    CU(cudaGetLastError());
  } else if (endBit > beginBit) {
    cub::DoubleBuffer<uint32_t> dk(w.keys[0], w.keys[1]);
    cub::DoubleBuffer<uint64_t> dv(w.vals[0], w.vals[1]);
    size_t need = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)n, beginBit, endBit, st));
    if (need > w.sortTempBytes) {
      CU(cudaStreamSynchronize(st));
      cudaFree(w.sortTemp);
      w.sortTemp = nullptr;
      w.sortTempBytes = 0;
      CU(cudaMalloc(&w.sortTemp, need));
      w.sortTempBytes = need;
    }
    size_t have = w.sortTempBytes;
    CU(cub::DeviceRadixSort::SortPairs(w.sortTemp, have, dk, dv, (int)n, beginBit, endBit, st));
    cur = dk.selector;
    sortLaunches = 3;
  }
  mark();
  auto gen = [&](int g, int pass) {
    SweepRecs r;
    for (int a = 0; a < kSweepMaxArrays; a++) r.arr[a] = w.recs[g][a];
    r.count = w.ctrl + kSweepCtrlStride * pass;
    r.cap = w.cap;
    return r;
  };
  // 12-byte records (nucleotide, at most 8 letters left of the seed table's k-mer): see sweepStep
  const bool wide = !AMINO && (c->ix.bwtLength >= 0xFFFFFFF0ull || c->sweepWide);
  const bool rec12 = !AMINO && !variable && !wide && steps <= 8 && c->sweepRecord12;
  uint32_t refills = 0;
  // Ordered emit (sweepStep<EMIT>, sweepEmit): range output of a fixed-length nucleotide batch small enough for a
  // sixteenth of its counts and ranges (20 B per query) to stay in L2 while it is written.
  const bool emit = !AMINO && !variable && !rec12 && dRanges && steps >= 1 && c->sweepOrderedEmit && n <= (1ull << 24);
  const uint32_t emitDiv = (uint32_t)((n + kSweepEmitBuckets - 1) / kSweepEmitBuckets);  // (16 * emitDiv <= n + 15 <= w.cap)
  auto launchPass = [&](auto first, auto items, auto small, auto var, auto big, auto ordered, uint32_t pass) -> int {
    constexpr bool FIRST = decltype(first)::value;
    constexpr int ITEMS = decltype(items)::value;
    constexpr bool VARLEN = decltype(var)::value;
    constexpr bool WIDE = decltype(big)::value && !AMINO;
    constexpr bool REC12 = decltype(small)::value && !AMINO && !VARLEN && !WIDE;
    constexpr bool EMIT = decltype(ordered)::value && !AMINO && !VARLEN && !REC12;
    auto kf = sweepStep<FIRST, ITEMS, AMINO, REC12, VARLEN, WIDE, EMIT>;
    int grid = 0;
    if (int r = gridFor(c, kf, kSweepThreads, &grid)) return r;
    const uint64_t tile = (uint64_t)kSweepThreads * ITEMS;
    grid = (int)std::min<uint64_t>((uint64_t)grid, (n + tile - 1) / tile);
    if (FIRST)
      kf<<<grid, kSweepThreads, 0, st>>>(c->ix, w.keys[cur], w.vals[cur], n, deep, gen(1, kSweepMaxPasses - 1), gen(0, 0),
                                         steps, localBits | (localShift << 8), dCounts, dRanges, w.irregularIds, irregularCount,
                                         hitsOnly, emitDiv);
    else  // pass p does LF step p+1 of the queries still alive
      kf<<<grid, kSweepThreads, 0, st>>>(c->ix, nullptr, nullptr, 0, deep, gen((pass - 1) & 1, pass - 1),
                                         gen(pass & 1, pass), steps - pass, 0u, dCounts, dRanges, w.irregularIds, irregularCount,
                                         hitsOnly, emitDiv);
    CU(cudaGetLastError());
    return AWFM_GPU_OK;
  };
  auto launchPassRec = [&](auto first, auto items, uint32_t pass) -> int {
    if (emit && pass + 1 == steps)  // the last pass hands its survivors to sweepEmit
      return launchPass(first, items, std::false_type(), std::false_type(), std::false_type(), std::true_type(), pass);
    return rec12 ? launchPass(first, items, std::true_type(), std::false_type(), std::false_type(), std::false_type(), pass)
                 : launchPass(first, items, std::false_type(), std::false_type(), std::false_type(), std::false_type(), pass);
  };
  auto launchPassItems = [&](auto first, uint32_t pass) -> int {
    // variable lengths / 64-bit positions: one instantiation per pass kind, 4 records per thread
    constexpr auto four = std::integral_constant<int, 4>();
    if (wide) {
      if (variable) return launchPass(first, four, std::false_type(), std::true_type(), std::true_type(), std::false_type(), pass);
      if (emit && pass + 1 == steps)
        return launchPass(first, four, std::false_type(), std::false_type(), std::true_type(), std::true_type(), pass);
      return launchPass(first, four, std::false_type(), std::false_type(), std::true_type(), std::false_type(), pass);
    }
    if (variable) return launchPass(first, four, std::false_type(), std::true_type(), std::false_type(), std::false_type(), pass);
    switch (decltype(first)::value ? c->sweepFirstItems : c->sweepItems) {
      case 1: return launchPassRec(first, std::integral_constant<int, 1>(), pass);
      case 2: return launchPassRec(first, std::integral_constant<int, 2>(), pass);
#if AWFM_SWEEP_THREADS <= 256  // 8 records per thread of a 512-thread CTA would need more than 48 KB of static shared memory
      case 8: return launchPassRec(first, std::integral_constant<int, 8>(), pass);
#endif
      default: return launchPassRec(first, std::integral_constant<int, 4>(), pass);
    }
  };
  auto emitSurvivors = [&](uint32_t pass) -> int {  // after the last pass (pass + 1 == steps)
    if (!emit || pass + 1 != steps) return AWFM_GPU_OK;
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)c->numSMs * 8);
    if (wide) sweepEmit<true><<<grid, 256, 0, st>>>(gen(pass & 1, pass), emitDiv, dCounts, dRanges);
    else sweepEmit<false><<<grid, 256, 0, st>>>(gen(pass & 1, pass), emitDiv, dCounts, dRanges);
    CU(cudaGetLastError());
    refills++;
    return AWFM_GPU_OK;
  };
  if (int r = launchPassItems(std::true_type(), 0)) return r;
  if (int r = emitSurvivors(0)) return r;
  mark();
  for (uint32_t pass = 1; pass < steps; pass++) {
    if (int r = launchPassItems(std::false_type(), pass)) return r;
    if (int r = emitSurvivors(pass)) return r;
    if (more && pass == 15 && steps > 16) {  // generation 15 has prepended its 16 letters: the next ones from w.more
      sweepRefill<<<(int)std::min<uint64_t>((n + 255) / 256, (uint64_t)c->numSMs * 8), 256, 0, st>>>(gen(pass & 1, pass), more);
      CU(cudaGetLastError());
      refills++;
    }
    mark();
  }
  if (format == AWFM_QUERY_ASCII) {
    sweepIrregular<AMINO><<<c->numSMs * 2, 256, 0, st>>>(c->ix, dLetters, dOffsets, len, w.irregularIds, irregularCount, dCounts,
                                                         dRanges, hitsOnly);
    CU(cudaGetLastError());
  } else if (rec12 || wide) {  // the 2-bit format has no irregular letters, but a seed range may be too wide for the record
    sweepIrregularBits<<<c->numSMs * 2, 256, 0, st>>>(c->ix, dLetters, len, w.irregularIds, irregularCount, dCounts, dRanges,
                                                      hitsOnly);
    CU(cudaGetLastError());
  }
  mark();
  CU(cudaEventRecord(w.done, st));
  w.stagesRecorded = stage;
  w.lastSteps = steps, w.lastBuckets = AMINO ? 20 : 4, w.lastQueries = n;
  L.stats.launches += 2 + ((format == AWFM_QUERY_ASCII || rec12 || wide) ? 1 : 0) + (steps > 1 ? steps - 1 : 0) + sortLaunches + refills;
  return AWFM_GPU_OK;
}

// Batches larger than "sweep_max_batch" queries go through the scratch in slices (96 B of scratch per query of a
// slice for nucleotide indexes, 352 B for amino ones: two generations of 2 | 10 record arrays + the sort buffers).
static int sweepCount(awfm_gpu_ctx *c, Lane &L, const uint8_t *dLetters, const uint64_t *dOffsets, uint32_t format,
                      uint32_t len, uint64_t n, uint32_t *dCounts, uint4 *dRanges, bool hitsOnly, cudaStream_t st) {
  const bool amino = c->ix.amino != 0;
  const uint64_t maxBatch = amino ? std::min<int64_t>(c->sweepMaxBatch, 1ll << 26) : c->sweepMaxBatch;
  const uint64_t slice = std::min<uint64_t>(n, maxBatch & ~255ull);  // slices start 16-B aligned
  const uint64_t queryBytes = dOffsets ? 0 : awfm_query_bytes(format, len);  // offsets are absolute: same letter base
  if (int r = ensureSweep(c, L, slice, amino ? 10 : 2)) return r;
  for (uint64_t first = 0; first < n; first += slice) {
    const uint64_t m = std::min(slice, n - first);
    uint4 *ranges = dRanges ? dRanges + first : nullptr;
    const uint64_t *offsets = dOffsets ? dOffsets + first : nullptr;
    const int r = amino ? sweepCountBatch<true>(c, L, dLetters + first * queryBytes, offsets, format, len, m, dCounts + first, ranges, hitsOnly, st)
                        : sweepCountBatch<false>(c, L, dLetters + first * queryBytes, offsets, format, len, m, dCounts + first, ranges, hitsOnly, st);
    if (r) return r;
  }
  return AWFM_GPU_OK;
}

// One batch of queries on the device: counts and (optionally) every query's final range.  2-/5-bit batches that do not
// take the sweep path are first expanded to ASCII letters in the lane's scratch (unpackQueries).
int awfm_count_device_impl(awfm_gpu_ctx *c, Lane &L, const PackedBatch &batch, uint32_t *dCounts, awfm_range *dRanges,
                           cudaStream_t st) {
  const uint64_t n = batch.numQueries;
  if (n == 0) return AWFM_GPU_OK;
  if (batch.format != AWFM_QUERY_ASCII && batch.format != AWFM_QUERY_2BIT && batch.format != AWFM_QUERY_5BIT)
    return awfm_fail(AWFM_GPU_ERR_ARG, "unknown query format");
  if (batch.format != AWFM_QUERY_ASCII) {
    if (batch.offsets || batch.length == 0) return awfm_fail(AWFM_GPU_ERR_ARG, "2-/5-bit query batches are fixed-length");
    if ((batch.format == AWFM_QUERY_5BIT) != (c->ix.amino != 0))
      return awfm_fail(AWFM_GPU_ERR_ARG, "query format does not match the index alphabet (2-bit: nucleotide, 5-bit: amino)");
  }
  EventPair *ev = nextEvents(L);
  if (ev) CU(cudaEventRecord(ev->a, st));
  const uint8_t *dLetters = batch.data;
  uint32_t format = batch.format;
  const uint32_t fixedLen = batch.length;
  // (measured on cfg 3, 10 M 16-mers with hits for half of the queries: 1.46 ms through the sweep, 1.37 through the tile
  // kernel — range output, of all queries or of the hits only, keeps the higher threshold)
  const awfm_range *thresholdRanges = dRanges;
  const bool directBits = format == AWFM_QUERY_2BIT && fixedLen <= 32 &&
                          sweepEligible(c, dLetters, nullptr, fixedLen, n, thresholdRanges);
  const bool unpack = format != AWFM_QUERY_ASCII && !directBits;
  if (unpack) {
    CU(cudaStreamWaitEvent(st, L.unpackDone, 0));
    if (L.dUnpacked.cap < n * (uint64_t)fixedLen + 16) CU(cudaDeviceSynchronize());  // about to be reallocated
    if (int r = L.dUnpacked.ensure(n * (uint64_t)fixedLen + 16)) return r;
    const uint64_t words = (n * (uint64_t)fixedLen + 3) / 4;
    const int grid = (int)std::min<uint64_t>((words + 255) / 256, (uint64_t)c->numSMs * 16);
    if (format == AWFM_QUERY_2BIT) unpackQueries<2><<<grid, 256, 0, st>>>(dLetters, n, fixedLen, (uint8_t *)L.dUnpacked.p);
    else unpackQueries<5><<<grid, 256, 0, st>>>(dLetters, n, fixedLen, (uint8_t *)L.dUnpacked.p);
    CU(cudaGetLastError());
    L.stats.launches += 1;
    dLetters = (const uint8_t *)L.dUnpacked.p;
    format = AWFM_QUERY_ASCII;
  }
  QueryBatch qb{dLetters, batch.offsets, n, fixedLen, batch.rangesOfHitsOnly ? 1u : 0u};
  int r;
  const bool sweep = directBits || sweepEligible(c, dLetters, batch.offsets, fixedLen, n, thresholdRanges);
  // (a variable-length batch hands the sweep its longest query's length if the caller stated one, else 0)
  r = sweep ? sweepCount(c, L, dLetters, batch.offsets, format, batch.offsets ? batch.maxLength : fixedLen, n, dCounts,
                         (uint4 *)dRanges, batch.rangesOfHitsOnly, st)
            : AWFM_GPU_OK;
  if (!sweep || r == AWFM_GPU_ERR_ALLOC) {  // no room for the sweep's scratch: the tile kernel needs none
    L.sweep.stagesRecorded = 0;
    if (format != AWFM_QUERY_ASCII) return awfm_fail(AWFM_GPU_ERR_ALLOC, "sweep scratch does not fit in device memory");
    r = DISPATCH_COUNT(launchCount, c->countLpq, c->ix.amino != 0, c, qb, dCounts, (uint4 *)dRanges, st);
    L.stats.launches += 1;
  }
  if (unpack) CU(cudaEventRecord(L.unpackDone, st));
  if (ev) CU(cudaEventRecord(ev->b, st));
  L.stats.queries += n;
  return r;
}

extern "C" int awfm_gpu_ctx_sweep_stage_ms(awfm_gpu_ctx *c, double *ms, int capacity) {
  if (!c || !ms) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  if (int r = awfm_set_device(c)) return r;
  SweepScratch &w = c->lanes[c->lastLane.load()].sweep;
  int n = 0;
  for (int i = 0; i + 1 < w.stagesRecorded && n < capacity; i++, n++) {
    if (cudaEventSynchronize(w.stage[i + 1]) != cudaSuccess) break;
    float f = 0;
    if (cudaEventElapsedTime(&f, w.stage[i], w.stage[i + 1]) != cudaSuccess) break;
    ms[n] = f;
  }
  cudaGetLastError();
  return n;
}

// Live records after every pass of the most recent sweep count call (what the compulsory-traffic model of the bench is
// computed from): live[0] = queries of the batch, live[p] = records appended by pass p (still searching after LF step
// p), p = 1 .. passes-1 (the last pass appends nothing).  *irregular = queries answered by sweepIrregular.
extern "C" int awfm_gpu_ctx_sweep_live(awfm_gpu_ctx *c, uint64_t *live, int capacity, uint64_t *irregular) {
  if (!c || !live || capacity < 1) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  if (int r = awfm_set_device(c)) return r;
  SweepScratch &w = c->lanes[c->lastLane.load()].sweep;
  if (!w.ctrl || w.stagesRecorded == 0 && w.lastQueries == 0) return 0;
  CU(cudaDeviceSynchronize());
  std::vector<uint32_t> ctrl(kSweepMaxPasses * kSweepCtrlStride + 4);
  CU(cudaMemcpy(ctrl.data(), w.ctrl, ctrl.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  int n = 0;
  live[n++] = w.lastQueries;
  for (uint32_t p = 0; p + 1 < w.lastSteps && n < capacity; p++) {
    uint64_t sum = 0;
    for (uint32_t b = 0; b < w.lastBuckets; b++) sum += ctrl[p * kSweepCtrlStride + b];
    live[n++] = sum;
  }
  if (irregular) *irregular = ctrl[kSweepMaxPasses * kSweepCtrlStride];
  return n;
}

extern "C" int awfm_gpu_count_device(awfm_gpu_ctx *c, const uint8_t *dLetters, const uint64_t *dOffsets,
                                     uint32_t fixedLen, uint64_t n, uint32_t *dCounts, awfm_range *dRanges,
                                     void *stream) {
  return awfm_gpu_count_device_format(c, dLetters, AWFM_QUERY_ASCII, dOffsets, fixedLen, n, dCounts, dRanges, stream);
}

extern "C" int awfm_gpu_count_device_format(awfm_gpu_ctx *c, const void *dQueries, uint32_t format,
                                            const uint64_t *dOffsets, uint32_t fixedLen, uint64_t n, uint32_t *dCounts,
                                            awfm_range *dRanges, void *stream) {
  if (!c || !dCounts || (n && !dQueries)) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  if (int r = awfm_set_device(c)) return r;
  Lane &L = c->lanes[0];
  c->lastLane.store(0);
  awfm_begin_call(L);
  PackedBatch b;
  b.data = (const uint8_t *)dQueries, b.offsets = dOffsets, b.format = format, b.length = fixedLen, b.numQueries = n;
  return awfm_count_device_impl(c, L, b, dCounts, dRanges, (cudaStream_t)stream);
}

// hitOffsets[q] = base + sum of the (u32-truncated) hit-list lengths of the queries before q; hitOffsets[n] = base + total
int awfm_scan_impl(awfm_gpu_ctx *c, Lane &L, LocateScratch &sc, const void *lengthSource, bool fromCounts, uint64_t n,
                   uint64_t *dHitOffsets, uint64_t base, cudaStream_t st) {
  (void)c;
  const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
  const size_t need = (tiles + 1) * sizeof(uint64_t);
  if (need > sc.scanTempBytes) {
    CU(cudaStreamSynchronize(st));
    cudaFree(sc.scanTemp);
    sc.scanTemp = nullptr;
    sc.scanTempBytes = 0;
    CU(cudaMalloc(&sc.scanTemp, need + need / 4));
    sc.scanTempBytes = need + need / 4;
  }
  uint64_t *tileSums = (uint64_t *)sc.scanTemp;
  if (tiles) {
    if (fromCounts) scanTileSums<true><<<(unsigned)tiles, 256, 0, st>>>(lengthSource, n, tileSums);
    else scanTileSums<false><<<(unsigned)tiles, 256, 0, st>>>(lengthSource, n, tileSums);
  }
  scanTileBases<<<1, 256, 0, st>>>(tileSums, tiles, base);
  if (tiles) {
    if (fromCounts) scanTileOffsets<true><<<(unsigned)tiles, 256, 0, st>>>(lengthSource, n, tileSums, dHitOffsets);
    else scanTileOffsets<false><<<(unsigned)tiles, 256, 0, st>>>(lengthSource, n, tileSums, dHitOffsets);
  } else {
    CU(cudaMemcpyAsync(dHitOffsets, tileSums, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
  }
  CU(cudaGetLastError());
  L.stats.launches += tiles ? 3 : 1;
  return AWFM_GPU_OK;
}

// The front end of a device-resident locate in one call: search, ranges of the queries WITH hits only, hit offsets
// scanned from the u32 counts (a quarter of the bytes of the ranges, and no scattered 16-B store per empty query).
extern "C" int awfm_gpu_locate_prepare_device(awfm_gpu_ctx *c, const void *dQueries, uint32_t format,
                                              const uint64_t *dOffsets, uint32_t fixedLen, uint64_t n, uint32_t *dCounts,
                                              awfm_range *dRanges, uint64_t *dHitOffsets, void *stream) {
  if (!c || !dCounts || !dRanges || !dHitOffsets || (n && !dQueries)) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  if (int r = awfm_set_device(c)) return r;
  Lane &L = c->lanes[0];
  c->lastLane.store(0);
  awfm_begin_call(L);
  PackedBatch b;
  b.data = (const uint8_t *)dQueries, b.offsets = dOffsets, b.format = format, b.length = fixedLen, b.numQueries = n;
  b.rangesOfHitsOnly = true;
  if (int r = awfm_count_device_impl(c, L, b, dCounts, dRanges, (cudaStream_t)stream)) return r;
  return awfm_scan_impl(c, L, L.sc, dCounts, true, n, dHitOffsets, 0, (cudaStream_t)stream);
}

extern "C" int awfm_gpu_scan_ranges_device(awfm_gpu_ctx *c, const awfm_range *dRanges, uint64_t n,
                                           uint64_t *dHitOffsets, void *stream) {
  if (!c || !dHitOffsets || (n && !dRanges)) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  if (int r = awfm_set_device(c)) return r;
  return awfm_scan_impl(c, c->lanes[0], c->lanes[0].sc, dRanges, false, n, dHitOffsets, 0, (cudaStream_t)stream);
}

// Backtrace walk of the flat hit indices [hb, he) (in the numbering of dHitOffsets) into dPos[h - hb].
int awfm_locate_device_impl(awfm_gpu_ctx *c, Lane &L, LocateScratch &sc, const awfm_range *dRanges,
                            const uint64_t *dHitOffsets, uint64_t n, uint64_t hb, uint64_t he, uint64_t *dPos,
                            cudaStream_t st) {
  if (!c->hasSa) return awfm_fail(AWFM_GPU_ERR_NO_SA, "context was created without a sampled suffix array");
  if (he <= hb) return AWFM_GPU_OK;
  EventPair *ev = nextEvents(L);
  if (ev) CU(cudaEventRecord(ev->a, st));
  int r = DISPATCH_LOCATE(launchLocate, c->locateLpq, c->ix.amino != 0, c, sc, (const uint4 *)dRanges, dHitOffsets, n,
                          hb, he, dPos, st);
  if (ev) CU(cudaEventRecord(ev->b, st));
  L.stats.launches += 2;
  L.stats.hits += he - hb;
  return r;
}

extern "C" int awfm_gpu_locate_device(awfm_gpu_ctx *c, const awfm_range *dRanges, const uint64_t *dHitOffsets,
                                      uint64_t n, uint64_t hb, uint64_t he, uint64_t *dPos, void *stream) {
  if (!c || !dRanges || !dHitOffsets || (he > hb && !dPos)) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  if (int r = awfm_set_device(c)) return r;
  return awfm_locate_device_impl(c, c->lanes[0], c->lanes[0].sc, dRanges, dHitOffsets, n, hb, he, dPos, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ derived structures
static uint32_t ratioShiftOf(uint32_t ratio) {
  if (ratio & (ratio - 1)) return 0xFFFFFFFFu;
  uint32_t s = 0;
  while ((1u << s) < ratio) s++;
  return s;
}

template <bool AMINO>
static int extendSeedLevels(awfm_gpu_ctx *c, uint32_t depth, bool wide) {
  const uint64_t card = AMINO ? 20 : 4;
  const uint64_t entryBytes = wide ? 16 : 8;
  const void *src = c->dSeed;
  bool srcWide = true;
  uint64_t numSrc = c->ix.numSeeds;
  void *prev = nullptr;  // intermediate level owned here
  for (uint32_t level = c->ix.seedK; level < depth; level++) {
    void *dst = nullptr;
    cudaError_t e = cudaMalloc(&dst, numSrc * card * entryBytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      cudaFree(prev);
      return awfm_fail(AWFM_GPU_ERR_ALLOC, "derived seed table does not fit in device memory", cudaGetErrorString(e));
    }
    const int grid = c->numSMs * 8;
    if (srcWide && wide) extendSeedTable<AMINO, true, true><<<grid, 256>>>(c->ix, src, numSrc, dst);
    else if (srcWide) extendSeedTable<AMINO, true, false><<<grid, 256>>>(c->ix, src, numSrc, dst);
    else if (wide) extendSeedTable<AMINO, false, true><<<grid, 256>>>(c->ix, src, numSrc, dst);
    else extendSeedTable<AMINO, false, false><<<grid, 256>>>(c->ix, src, numSrc, dst);
    e = cudaDeviceSynchronize();
    cudaFree(prev);
    if (e != cudaSuccess) {
      cudaGetLastError();
      cudaFree(dst);
      return awfm_fail(AWFM_GPU_ERR_CUDA, "extendSeedTable", cudaGetErrorString(e));
    }
    prev = dst;
    src = dst;
    srcWide = wide;
    numSrc *= card;
  }
  c->dDeepSeed = prev;
  c->deepSeedBytes = numSrc * entryBytes;
  return AWFM_GPU_OK;
}

extern "C" int awfm_gpu_ctx_extend_seed_table(awfm_gpu_ctx *c, uint32_t depth, double *buildMs) {
  if (!c) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  AllLanesHold all(c);
  if (int r = awfm_set_device(c)) return r;
  CU(cudaDeviceSynchronize());
  const double t0 = omp_get_wtime();
  if (buildMs) *buildMs = 0;
  // drop what is there
  cudaFree(c->dDeepSeed);
  c->dDeepSeed = nullptr;
  c->deviceBytes -= c->deepSeedBytes;
  c->deepSeedBytes = 0;
  c->ix.deepSeedTable = nullptr;
  c->ix.deepSeedK = 0;
  c->deepSeedKBuilt = 0;
  if (depth <= c->ix.seedK) return AWFM_GPU_OK;
  const bool amino = c->ix.amino != 0;
  if (depth > (amino ? 9u : 20u)) return awfm_fail(AWFM_GPU_ERR_ARG, "seed table depth too large (max 20 nucleotide, 9 amino)");
  const bool wide = c->ix.bwtLength > (1ull << 32);
  int rc = amino ? extendSeedLevels<true>(c, depth, wide) : extendSeedLevels<false>(c, depth, wide);
  if (rc) return rc;
  c->deviceBytes += c->deepSeedBytes;
  c->ix.deepSeedTable = c->dDeepSeed;
  c->ix.deepSeedK = depth;
  c->deepSeedKBuilt = depth;
  c->ix.deepSeedWide = wide;
  if (buildMs) *buildMs = 1e3 * (omp_get_wtime() - t0);
  return AWFM_GPU_OK;
}

static int locateDeviceRaw(awfm_gpu_ctx *c, uint64_t numHits, uint64_t *dPos, cudaStream_t st);

extern "C" int awfm_gpu_ctx_densify_suffix_array(awfm_gpu_ctx *c, uint32_t newRatio, double *buildMs) {
  if (!c) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  AllLanesHold all(c);
  if (int r = awfm_set_device(c)) return r;
  if (!c->hasSa) return awfm_fail(AWFM_GPU_ERR_NO_SA, "context was created without a sampled suffix array");
  CU(cudaDeviceSynchronize());
  const double t0 = omp_get_wtime();
  if (buildMs) *buildMs = 0;
  if (c->dDenseSa) {  // back to the index's own samples
    cudaFree(c->dDenseSa);
    c->dDenseSa = nullptr;
    c->deviceBytes -= c->denseSaBytes;
    c->denseSaBytes = 0;
    c->ix.sa = (const uint64_t *)c->origSa;
    c->ix.saBitWidth = c->origSaBitWidth;
    c->ix.saRatio = c->origSaRatio;
    c->ix.saRatioShift = c->origSaRatioShift;
  }
  if (newRatio == 0 || newRatio >= c->ix.saRatio) return AWFM_GPU_OK;
  // samples at BWT positions j*newRatio, stored as aligned 32- or 64-bit fields (a packed SA of that width)
  const uint64_t n = c->ix.bwtLength;
  const uint64_t samples = (n + newRatio - 1) / newRatio;
  const uint32_t width = c->ix.saBitWidth <= 32 ? 32 : 64;
  const uint64_t bytes = ((samples * (width / 8) + 15) & ~15ull) + 16;
  void *dense = nullptr;
  uint64_t *work = nullptr;
  const uint64_t slab = std::min<uint64_t>(samples, 1ull << 28);  // 2 GiB of walk state at a time
  if (cudaMalloc(&dense, bytes) != cudaSuccess || cudaMalloc(&work, slab * 8) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(dense);
    cudaFree(work);
    return awfm_fail(AWFM_GPU_ERR_ALLOC, "dense suffix array does not fit in device memory");
  }
  cudaMemset(dense, 0, bytes);
  cudaStream_t st = c->lanes[0].slots[0].stream;
  int rc = AWFM_GPU_OK;
  for (uint64_t first = 0; first < samples && rc == AWFM_GPU_OK; first += slab) {
    const uint64_t count = std::min(slab, samples - first);
    saIota<<<c->numSMs * 8, 256, 0, st>>>(work, first, count, newRatio);
    rc = locateDeviceRaw(c, count, work, st);
    if (rc == AWFM_GPU_OK) {
      if (width == 32) saNarrow<uint32_t><<<c->numSMs * 8, 256, 0, st>>>(work, count, (uint32_t *)dense + first);
      else saNarrow<uint64_t><<<c->numSMs * 8, 256, 0, st>>>(work, count, (uint64_t *)dense + first);
      if (cudaStreamSynchronize(st) != cudaSuccess) rc = awfm_fail(AWFM_GPU_ERR_CUDA, "densify suffix array", cudaGetErrorString(cudaGetLastError()));
    }
  }
  cudaFree(work);
  if (rc) {
    cudaFree(dense);
    return rc;
  }
  c->origSa = c->dSa;
  c->origSaBitWidth = c->ix.saBitWidth;
  c->origSaRatio = c->ix.saRatio;
  c->origSaRatioShift = c->ix.saRatioShift;
  c->dDenseSa = dense;
  c->denseSaBytes = bytes;
  c->deviceBytes += bytes;
  c->ix.sa = (const uint64_t *)dense;
  c->ix.saBitWidth = width;
  c->ix.saRatio = newRatio;
  c->ix.saRatioShift = ratioShiftOf(newRatio);
  if (buildMs) *buildMs = 1e3 * (omp_get_wtime() - t0);
  return AWFM_GPU_OK;
}

// ------------------------------------------------------------------------------------------------ contig mapping
extern "C" int awfm_gpu_ctx_set_sequences(awfm_gpu_ctx *c, const void *metadata, uint64_t numSequences) {
  if (!c || (numSequences && !metadata)) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  AllLanesHold all(c);
  if (int r = awfm_set_device(c)) return r;
  CU(cudaDeviceSynchronize());
  cudaFree(c->dSequenceEnds);
  c->dSequenceEnds = nullptr;
  c->ix.sequenceEnds = nullptr;
  c->ix.numSequences = 0;
  if (numSequences == 0) return AWFM_GPU_OK;
  // struct FastaVectorMetadata { size_t headerEndPosition; size_t sequenceEndPosition; } -> dense array of ends
  std::vector<uint64_t> ends(numSequences);
  const uint64_t *m = (const uint64_t *)metadata;
  for (uint64_t i = 0; i < numSequences; i++) {
    ends[i] = m[2 * i + 1];
    if (i && ends[i] < ends[i - 1]) return awfm_fail(AWFM_GPU_ERR_ARG, "sequenceEndPosition must be non-decreasing");
  }
  CU(cudaMalloc(&c->dSequenceEnds, numSequences * 8));
  CU(cudaMemcpy(c->dSequenceEnds, ends.data(), numSequences * 8, cudaMemcpyHostToDevice));
  c->deviceBytes += numSequences * 8;
  c->ix.sequenceEnds = c->dSequenceEnds;
  c->ix.numSequences = numSequences;
  return AWFM_GPU_OK;
}

int awfm_map_device_impl(awfm_gpu_ctx *c, Lane &L, const uint64_t *dPos, uint64_t n, uint64_t *dSeq, uint64_t *dLocal,
                         cudaStream_t st) {
  if (!c->ix.sequenceEnds) return awfm_fail(AWFM_GPU_ERR_ARG, "context has no sequence table (awfm_gpu_ctx_set_sequences)");
  if (n == 0) return AWFM_GPU_OK;
  const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)c->numSMs * 8);
  EventPair *ev = nextEvents(L);
  if (ev) CU(cudaEventRecord(ev->a, st));
  mapPositionsKernel<<<grid, 256, 0, st>>>(c->ix.sequenceEnds, c->ix.numSequences, dPos, n, dSeq, dLocal);
  CU(cudaGetLastError());
  if (ev) CU(cudaEventRecord(ev->b, st));
  L.stats.launches += 1;
  return AWFM_GPU_OK;
}

extern "C" int awfm_gpu_map_positions_device(awfm_gpu_ctx *c, const uint64_t *dPositions, uint64_t n,
                                             uint64_t *dSequenceIndex, uint64_t *dLocalPosition, void *stream) {
  if (!c || (n && (!dPositions || !dSequenceIndex || !dLocalPosition))) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  if (int r = awfm_set_device(c)) return r;
  return awfm_map_device_impl(c, c->lanes[0], dPositions, n, dSequenceIndex, dLocalPosition, (cudaStream_t)stream);
}

extern "C" int awfm_gpu_map_positions_host(awfm_gpu_ctx *c, const uint64_t *positions, uint64_t n,
                                           uint64_t *sequenceIndex, uint64_t *localPosition, uint64_t *numIllegal) {
  if (!c || (n && (!positions || !sequenceIndex || !localPosition))) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  LaneHold hold(c);
  if (hold.rc) return hold.rc;
  Lane &L = *hold;
  awfm_begin_call(L);
  if (numIllegal) *numIllegal = 0;
  if (!c->ix.sequenceEnds) return awfm_fail(AWFM_GPU_ERR_ARG, "context has no sequence table (awfm_gpu_ctx_set_sequences)");
  if (n == 0) return AWFM_GPU_OK;
  cudaStream_t st = L.slots[0].stream;
  const uint64_t batch = std::min<uint64_t>(n, 1ull << 27);  // bounded staging: 3 x 1 GiB
  if (int r = L.dPositions.ensure(batch * 24)) return r;
  uint64_t *d = (uint64_t *)L.dPositions.p;
  int rc = AWFM_GPU_OK;
  uint64_t illegal = 0;
  for (uint64_t b0 = 0; b0 < n && rc == AWFM_GPU_OK; b0 += batch) {
    const uint64_t nb = std::min(batch, n - b0);
    cudaError_t e = cudaMemcpyAsync(d, positions + b0, nb * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = awfm_map_device_impl(c, L, d, nb, d + batch, d + 2 * batch, st);
    if (e == cudaSuccess && rc == AWFM_GPU_OK) e = cudaMemcpyAsync(sequenceIndex + b0, d + batch, nb * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == AWFM_GPU_OK) e = cudaMemcpyAsync(localPosition + b0, d + 2 * batch, nb * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == AWFM_GPU_OK) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = awfm_fail(AWFM_GPU_ERR_CUDA, "map positions", cudaGetErrorString(e));
    if (rc == AWFM_GPU_OK && numIllegal)
      for (uint64_t i = b0; i < b0 + nb; i++) illegal += sequenceIndex[i] == ~0ull;
  }
  if (numIllegal) *numIllegal = illegal;
  L.stats.h2dBytes = n * 8;
  L.stats.d2hBytes = n * 16;
  return rc;
}

// ------------------------------------------------------------------------------------------------ host-buffer calls
// Single-device, synchronous, one batch at a time; device buffers persist in the lane between calls.  (The pipelined,
// multi-GPU engine for large host batches is awfm_gpu_group_* in awfm_multi.cu.)
static uint64_t totalLetters(const uint64_t *offsets, uint32_t fixedLen, uint64_t n) {
  return offsets ? offsets[n] : n * (uint64_t)fixedLen;
}

extern "C" int awfm_gpu_count_host(awfm_gpu_ctx *c, const uint8_t *letters, const uint64_t *offsets,
                                   uint32_t fixedLen, uint64_t n, uint32_t *counts, awfm_range *ranges) {
  if (!c || !counts || (n && !letters)) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  LaneHold hold(c);
  if (hold.rc) return hold.rc;
  Lane &L = *hold;
  awfm_begin_call(L);
  if (n == 0) return AWFM_GPU_OK;
  const uint64_t nLetters = totalLetters(offsets, fixedLen, n);
  if (int r = L.dLetters.ensure(nLetters + 16)) return r;
  if (int r = L.dCounts.ensure(n * 4)) return r;
  if (offsets)
    if (int r = L.dOffsets.ensure((n + 1) * 8)) return r;
  if (ranges)
    if (int r = L.dRanges.ensure(n * 16)) return r;
  cudaStream_t st = L.slots[0].stream;
  CU(cudaMemcpyAsync(L.dLetters.p, letters, nLetters, cudaMemcpyHostToDevice, st));
  if (offsets) CU(cudaMemcpyAsync(L.dOffsets.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st));
  PackedBatch b;
  b.data = (const uint8_t *)L.dLetters.p, b.offsets = offsets ? (const uint64_t *)L.dOffsets.p : nullptr;
  b.length = fixedLen, b.numQueries = n;
  if (int r = awfm_count_device_impl(c, L, b, (uint32_t *)L.dCounts.p, ranges ? (awfm_range *)L.dRanges.p : nullptr, st))
    return r;
  CU(cudaMemcpyAsync(counts, L.dCounts.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (ranges) CU(cudaMemcpyAsync(ranges, L.dRanges.p, n * 16, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  L.stats.h2dBytes = nLetters + (offsets ? (n + 1) * 8 : 0);
  L.stats.d2hBytes = n * 4 + (ranges ? n * 16 : 0);
  return AWFM_GPU_OK;
}

extern "C" int awfm_gpu_locate_host(awfm_gpu_ctx *c, const uint8_t *letters, const uint64_t *offsets,
                                    uint32_t fixedLen, uint64_t n, uint64_t *hitOffsets, uint64_t *positions,
                                    uint64_t positionsCapacity, awfm_range *ranges) {
  if (!c || !hitOffsets || (n && !letters)) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  LaneHold hold(c);
  if (hold.rc) return hold.rc;
  Lane &L = *hold;
  if (!c->hasSa && positions) return awfm_fail(AWFM_GPU_ERR_NO_SA, "context was created without a sampled suffix array");
  awfm_begin_call(L);
  hitOffsets[0] = 0;
  if (n == 0) return AWFM_GPU_OK;
  const uint64_t nLetters = totalLetters(offsets, fixedLen, n);
  if (int r = L.dLetters.ensure(nLetters + 16)) return r;
  if (int r = L.dCounts.ensure(n * 4)) return r;
  if (int r = L.dRanges.ensure(n * 16)) return r;
  if (int r = L.dHits.ensure((n + 1) * 8)) return r;
  if (offsets)
    if (int r = L.dOffsets.ensure((n + 1) * 8)) return r;
  cudaStream_t st = L.slots[0].stream;
  CU(cudaMemcpyAsync(L.dLetters.p, letters, nLetters, cudaMemcpyHostToDevice, st));
  if (offsets) CU(cudaMemcpyAsync(L.dOffsets.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st));
  PackedBatch b;
  b.data = (const uint8_t *)L.dLetters.p, b.offsets = offsets ? (const uint64_t *)L.dOffsets.p : nullptr;
  b.length = fixedLen, b.numQueries = n;
  b.rangesOfHitsOnly = ranges == nullptr;  // the walk only needs the ranges of queries with hits
  if (int r = awfm_count_device_impl(c, L, b, (uint32_t *)L.dCounts.p, (awfm_range *)L.dRanges.p, st)) return r;
  if (int r = awfm_scan_impl(c, L, L.sc, b.rangesOfHitsOnly ? L.dCounts.p : L.dRanges.p, b.rangesOfHitsOnly, n,
                             (uint64_t *)L.dHits.p, 0, st))
    return r;
  CU(cudaMemcpyAsync(hitOffsets, L.dHits.p, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (ranges) CU(cudaMemcpyAsync(ranges, L.dRanges.p, n * 16, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const uint64_t total = hitOffsets[n];
  L.stats.h2dBytes = nLetters + (offsets ? (n + 1) * 8 : 0);
  L.stats.d2hBytes = (n + 1) * 8 + (ranges ? n * 16 : 0);
  if (!positions || total == 0) return AWFM_GPU_OK;
  if (positionsCapacity < total) return awfm_fail(AWFM_GPU_ERR_ARG, "positions buffer smaller than hitOffsets[numQueries]");
  // bounded device staging: at most 1 Gi hits (8 GB) per launch
  const uint64_t batch = std::min<uint64_t>(total, 1ull << 30);
  if (int r = L.dPositions.ensure(batch * 8)) return r;
  for (uint64_t hb = 0; hb < total; hb += batch) {
    const uint64_t he = std::min(total, hb + batch);
    if (int r = awfm_locate_device_impl(c, L, L.sc, (const awfm_range *)L.dRanges.p, (const uint64_t *)L.dHits.p, n, hb, he,
                                        (uint64_t *)L.dPositions.p, st))
      return r;
    CU(cudaMemcpyAsync(positions + hb, L.dPositions.p, (he - hb) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  L.stats.d2hBytes += total * 8;
  return AWFM_GPU_OK;
}

// ------------------------------------------------------------------------------------------------ search-list engine
// The reference API hands over an array of 32-B structs with pointers to arbitrary host strings
// (src/AwFmIndex.h:111-123).  Chunks of the list are packed into pinned staging by `numThreads` host threads,
// shipped and searched on a per-slot stream, and their counts scattered back while the next chunks are in flight.

static int ensureSlot(PipeSlot &s, uint64_t queries, uint64_t letterBytes, bool wantRanges) {
  if (s.queryCap < queries) {
    if (s.hOffsets) cudaFreeHost(s.hOffsets);
    if (s.hCounts) cudaFreeHost(s.hCounts);
    cudaFree(s.dOffsets);
    cudaFree(s.dCounts);
    cudaFree(s.dRanges);
    free(s.hOld);
    s.hOffsets = nullptr, s.hCounts = nullptr, s.dOffsets = nullptr, s.dCounts = nullptr, s.dRanges = nullptr, s.hOld = nullptr;
    s.queryCap = 0;
    if (!(s.hOld = static_cast<uint32_t *>(malloc(queries * 4)))) return awfm_fail(AWFM_GPU_ERR_ALLOC, "host memory for a pipeline slot");
    CU(cudaHostAlloc(&s.hOffsets, (queries + 1) * 8, cudaHostAllocPortable));
    CU(cudaHostAlloc(&s.hCounts, queries * 4, cudaHostAllocPortable));
    CU(cudaMalloc(&s.dOffsets, (queries + 1) * 8));
    CU(cudaMalloc(&s.dCounts, queries * 4));
    s.queryCap = queries;
  }
  if (wantRanges && !s.dRanges) CU(cudaMalloc(&s.dRanges, s.queryCap * 16));
  if (s.lettersCap < letterBytes) {
    if (s.hLetters) cudaFreeHost(s.hLetters);
    s.hLetters = nullptr;
    s.lettersCap = 0;
    const uint64_t cap = letterBytes + letterBytes / 4 + 64;
    CU(cudaHostAlloc(&s.hLetters, cap, cudaHostAllocPortable));
    s.lettersCap = cap;
  }
  return AWFM_GPU_OK;
}

static int ensureDeviceLetters(PipeSlot &s, uint64_t letterBytes) {
  if (s.dLettersCap < letterBytes + 16) {
    cudaFree(s.dLetters);
    s.dLetters = nullptr;
    s.dLettersCap = 0;
    const uint64_t cap = letterBytes + letterBytes / 4 + 64;
    CU(cudaMalloc(&s.dLetters, cap));
    s.dLettersCap = cap;
  }
  return AWFM_GPU_OK;
}

// Small fixed-length copy without a libc call: two overlapping 16-B (or 8-/4-B) moves cover any length <= 32.
static inline void copyLetters(uint8_t *dst, const uint8_t *src, uint64_t len) {
  if (len >= 16 && len <= 32) {
    uint64_t a0, a1, b0, b1;
    memcpy(&a0, src, 8), memcpy(&a1, src + 8, 8);
    memcpy(&b0, src + len - 16, 8), memcpy(&b1, src + len - 8, 8);
    memcpy(dst, &a0, 8), memcpy(dst + 8, &a1, 8);
    memcpy(dst + len - 16, &b0, 8), memcpy(dst + len - 8, &b1, 8);
  } else if (len >= 8 && len < 16) {
    uint64_t a, b;
    memcpy(&a, src, 8), memcpy(&b, src + len - 8, 8);
    memcpy(dst, &a, 8), memcpy(dst + len - 8, &b, 8);
  } else {
    memcpy(dst, src, len);
  }
}

struct Packed {
  uint32_t uniformLen = 0;        // != 0: every query of the chunk has this length (offsets not needed)
  uint64_t letterBytes = 0;
  const uint8_t *source = nullptr;  // where the H2D copy reads from: the slot's staging or the caller's own buffer
};

bool awfm_is_pinned_host(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// ---- packing one chunk of the list by the engine's OpenMP team ----
// Shared by the whole team; written by the master thread only, and only between barriers.
// Fast path (one pass over the 32-B entries): all lengths equal; if the strings are also laid out back to back in
// page-locked memory the copy is skipped and the DMA reads the caller's buffer directly.  General path: two passes
// (lengths -> offsets, then copy).
struct TeamPack {
  const awfm_kmer_search_data *d0 = nullptr;
  uint64_t n = 0, len0 = 0;
  bool optimistic = false, direct = false, fallback = false, wantRanges = false;
  uint8_t *staging = nullptr;
  uint32_t *old = nullptr;  // count engine: every entry's current `count` is noted while its line is being read anyway
  const uint8_t *base = nullptr;
  int uniformAll = 1, contiguousAll = 1;
  std::vector<uint64_t> partSum;
  int rc = AWFM_GPU_OK;
};

// master thread, before the barrier that precedes teamPack()
static void teamPackPrepare(TeamPack &tp, PipeSlot &s, const awfm_kmer_search_data *d0, uint64_t n, bool sourcePinned,
                            bool wantRanges) {
  tp.d0 = d0, tp.n = n, tp.wantRanges = wantRanges, tp.rc = AWFM_GPU_OK;
  tp.len0 = d0[0].kmerLength;
  tp.optimistic = tp.len0 >= 1 && tp.len0 <= 65536;
  tp.base = (const uint8_t *)d0[0].kmerString;
  // cheap probe before committing to the copy-free variant
  tp.direct = tp.optimistic && sourcePinned &&
              (const uint8_t *)d0[n - 1].kmerString == tp.base + (n - 1) * tp.len0 &&
              (const uint8_t *)d0[n / 2].kmerString == tp.base + (n / 2) * tp.len0;
  // every DEVICE buffer of the slot is sized here, by the thread that has the chunk's device current; the team may
  // later only grow the (portable, page-locked) host staging
  tp.rc = ensureSlot(s, n, (!tp.optimistic || tp.direct) ? 16 : n * tp.len0 + 16, wantRanges);
  tp.staging = s.hLetters;
  tp.old = wantRanges ? nullptr : s.hOld;
  tp.uniformAll = tp.contiguousAll = 1;
  tp.fallback = !tp.optimistic;
}

// every thread of the team (t of T); contains barriers, so all threads must call it under the same conditions
static void teamPack(TeamPack &tp, PipeSlot &s, int t, int T) {
  const awfm_kmer_search_data *d0 = tp.d0;
  const uint64_t n = tp.n, len0 = tp.len0;
  const bool worker = t >= 0;  // t < 0: a thread that only keeps the team's barriers (the count engine's driver)
  const uint64_t a = worker ? n * t / T : 0, b = worker ? n * (t + 1) / T : 0;
  uint32_t *old = tp.rc == AWFM_GPU_OK ? tp.old : nullptr;
  if (n && tp.rc == AWFM_GPU_OK && tp.optimistic) {
    bool uni = true, con = true;
    if (tp.direct) {
      const uint8_t *base = tp.base;
      if (old) {
        for (uint64_t i = a; i < b; i++) {
          uni &= d0[i].kmerLength == len0;
          con &= (const uint8_t *)d0[i].kmerString == base + i * len0;
          old[i] = d0[i].count;
        }
      } else {
        for (uint64_t i = a; i < b; i++) {
          uni &= d0[i].kmerLength == len0;
          con &= (const uint8_t *)d0[i].kmerString == base + i * len0;
        }
      }
    } else {
      uint8_t *staging = tp.staging;
      for (uint64_t i = a; i < b && uni; i++) {
        uni = d0[i].kmerLength == len0;
        if (uni) copyLetters(staging + i * len0, (const uint8_t *)d0[i].kmerString, len0);
        if (old) old[i] = d0[i].count;
      }
    }
    if (!uni) {
#pragma omp atomic write
      tp.uniformAll = 0;
    }
    if (!con) {
#pragma omp atomic write
      tp.contiguousAll = 0;
    }
  }
#pragma omp barrier
  if (n && tp.rc == AWFM_GPU_OK) {
    // a direct (copy-free) attempt that found a gap, or mixed lengths: redo this chunk on the general path
    if (tp.optimistic && tp.uniformAll && tp.direct && !tp.contiguousAll) {
#pragma omp barrier
#pragma omp master
      {
        tp.direct = false;
        tp.rc = ensureSlot(s, n, n * len0 + 16, tp.wantRanges);
        tp.staging = s.hLetters;
      }
#pragma omp barrier
      if (tp.rc == AWFM_GPU_OK) {
        uint8_t *staging = tp.staging;
        for (uint64_t i = a; i < b; i++) copyLetters(staging + i * len0, (const uint8_t *)d0[i].kmerString, len0);
      }
#pragma omp barrier
    } else if (!tp.optimistic || !tp.uniformAll) {
      uint64_t sum = 0;
      for (uint64_t i = a; i < b; i++) {
        sum += d0[i].kmerLength;
        if (old) old[i] = d0[i].count;
      }
      if (worker) tp.partSum[t + 1] = sum;
#pragma omp barrier
#pragma omp master
      {
        tp.partSum[0] = 0;
        for (int k = 0; k < T; k++) tp.partSum[k + 1] += tp.partSum[k];
        tp.rc = ensureSlot(s, n, tp.partSum[T] + 16, tp.wantRanges);
        tp.staging = s.hLetters;
        tp.fallback = true;
      }
#pragma omp barrier
      if (worker && tp.rc == AWFM_GPU_OK) {
        uint64_t o = tp.partSum[t];
        uint64_t *offs = s.hOffsets;
        uint8_t *staging = tp.staging;
        for (uint64_t i = a; i < b; i++) {
          offs[i] = o;
          memcpy(staging + o, d0[i].kmerString, d0[i].kmerLength);
          o += d0[i].kmerLength;
        }
      }
#pragma omp barrier
    }
  }
}

// master thread, after teamPack() and a barrier: what to ship
static Packed teamPackResult(TeamPack &tp, PipeSlot &s, int T) {
  Packed pk;
  if (tp.fallback) {
    s.hOffsets[tp.n] = tp.partSum[T];
    pk.uniformLen = 0, pk.letterBytes = tp.partSum[T], pk.source = s.hLetters;
  } else {
    pk.uniformLen = (uint32_t)tp.len0, pk.letterBytes = tp.n * tp.len0, pk.source = tp.direct ? tp.base : s.hLetters;
  }
  return pk;
}

static int submitCount(awfm_gpu_ctx *c, Lane &L, PipeSlot &s, const Packed &pk, bool wantRanges) {
  const bool fixed = pk.uniformLen != 0;
  if (int r = ensureDeviceLetters(s, pk.letterBytes)) return r;
  CU(cudaMemcpyAsync(s.dLetters, pk.source, pk.letterBytes, cudaMemcpyHostToDevice, s.stream));
  if (!fixed) CU(cudaMemcpyAsync(s.dOffsets, s.hOffsets, (s.n + 1) * 8, cudaMemcpyHostToDevice, s.stream));
  PackedBatch b;
  b.data = s.dLetters, b.offsets = fixed ? nullptr : s.dOffsets, b.length = pk.uniformLen, b.numQueries = s.n;
  b.rangesOfHitsOnly = wantRanges;  // the locate engine scans the counts and reads the ranges of queries with hits only
  if (int r = awfm_count_device_impl(c, L, b, s.dCounts, wantRanges ? (awfm_range *)s.dRanges : nullptr, s.stream))
    return r;
  L.stats.h2dBytes += pk.letterBytes + (fixed ? 0 : (s.n + 1) * 8);
  return AWFM_GPU_OK;
}

static int teamSize(uint32_t numThreads, uint64_t n, uint64_t chunk) {
  return (int)std::max<uint64_t>(1, std::min<uint64_t>(std::max<uint32_t>(1, numThreads), std::min(n, chunk) / 64));
}

#include "awfm_list_engine.inc"

// ------------------------------------------------------------------------------------------------ gather probe
template <int BYTES>
static int runGather(int lanes, const uint4 *d, uint64_t numRecords, uint64_t numReads, uint64_t *sink, int grid) {
  switch (lanes) {
    case 1: gatherProbe<BYTES, 1><<<grid, 256>>>(d, numRecords, numReads, sink); break;
    case 2: gatherProbe<BYTES, 2><<<grid, 256>>>(d, numRecords, numReads, sink); break;
    case 4: gatherProbe<BYTES, 4><<<grid, 256>>>(d, numRecords, numReads, sink); break;
    case 8: gatherProbe<BYTES, 8><<<grid, 256>>>(d, numRecords, numReads, sink); break;
    default: return awfm_fail(AWFM_GPU_ERR_ARG, "lanesPerRead must be 1, 2, 4 or 8");
  }
  CU(cudaGetLastError());
  return AWFM_GPU_OK;
}

// L2 -> DRAM fetch granularity hint of the current device (cudaLimitMaxL2FetchGranularity: 32, 64 or 128 bytes).
extern "C" int awfm_gpu_set_l2_fetch_granularity(int device, int bytes, int *actual) {
  CU(cudaSetDevice(device));
  if (bytes > 0) CU(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
  size_t v = 0;
  CU(cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity));
  if (actual) *actual = (int)v;
  return AWFM_GPU_OK;
}

extern "C" int awfm_gpu_gather_bandwidth(int device, uint64_t arrayBytes, uint32_t bytesPerRead, uint64_t numReads,
                                         int lanesPerRead, double *gbps) {
  if (!gbps || arrayBytes < 4096) return awfm_fail(AWFM_GPU_ERR_ARG, "bad argument");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  GrowBuf data, sink;
  struct Release {
    GrowBuf &a, &b;
    ~Release() { a.release(), b.release(); }
  } release{data, sink};
  if (int r = data.ensure(arrayBytes)) return r;
  if (int r = sink.ensure(64)) return r;
  CU(cudaMemset(data.p, 1, arrayBytes));
  const uint64_t numRecords = arrayBytes / bytesPerRead;
  const int grid = prop.multiProcessorCount * 8;
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a));
  CU(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CU(cudaEventRecord(a));
    int r;
    switch (bytesPerRead) {
      case 16: r = runGather<16>(lanesPerRead, (const uint4 *)data.p, numRecords, numReads, (uint64_t *)sink.p, grid); break;
      case 32: r = runGather<32>(lanesPerRead, (const uint4 *)data.p, numRecords, numReads, (uint64_t *)sink.p, grid); break;
      case 64: r = runGather<64>(lanesPerRead, (const uint4 *)data.p, numRecords, numReads, (uint64_t *)sink.p, grid); break;
      case 128: r = runGather<128>(lanesPerRead, (const uint4 *)data.p, numRecords, numReads, (uint64_t *)sink.p, grid); break;
      case 256: r = runGather<256>(lanesPerRead, (const uint4 *)data.p, numRecords, numReads, (uint64_t *)sink.p, grid); break;
      default: r = awfm_fail(AWFM_GPU_ERR_ARG, "bytesPerRead must be 16, 32, 64, 128 or 256");
    }
    if (r) return r;
    CU(cudaEventRecord(b));
    CU(cudaEventSynchronize(b));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *gbps = (double)numReads * bytesPerRead / (best * 1e-3) / 1e9;
  return AWFM_GPU_OK;
}
