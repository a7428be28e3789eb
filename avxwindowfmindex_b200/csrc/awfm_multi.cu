// awfm_multi.cu — device groups (include/awfm_gpu.h, awfm_gpu_group_*): one call fanned out over the GPUs of the box
// from a single process, and the pipelined packed-batch engine (SURVEY.md §8 rows e and f1).
//
// Partitioning follows SURVEY.md §8e: the index is replicated in every GPU's HBM, query i goes to GPU floor(i*G/N) in
// contiguous shards, every GPU has its own streams and DMA queue, and results are written by each GPU's copy engine
// straight into the caller's host arrays — no device-to-device exchange exists on this path.  One host thread per GPU
// drives that GPU's pipeline; with one GPU the calling thread does it itself.
//
// Per GPU, a shard is worked through in chunks on four slots (stream + buffers): while chunk i is searched, chunk i+1
// is on its way in and the counts of chunk i-1 on their way out.  Every H2D goes through the device's one copy-in
// stream, in chunk order (copies queued on several streams are interleaved by the copy engine: with six slots and a
// copy per slot stream the call took 16.3-18.8 ms instead of 13.7).  The searches of successive chunks serialise on the
// device (the sweep path's scratch is one per lane), the copies overlap them.  Chunks are 3 * 2^23 queries: the sweep
// streams the index once per LF step whatever the batch, 0.7 ms per chunk at 3.1 Gbp, so the link (11 M 2-bit 20-mers
// per ms) and the search (2.3 ms per 25 M chunk) run at the same pace; measured 13.7 ms per 100 M against 14.2 with
// 2^24-query chunks, the same with 3, 4 or 6 slots.
//
// There is no CPU fallback: every entry point fails when CUDA is unavailable.
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <thread>

#include "awfm_internal.cuh"

using namespace awfm;

namespace {

struct PendingCopy {  // a pageable destination: the D2H went to page-locked staging, the host finishes it
  void *dst;
  const void *src;
  size_t bytes;
};

struct PackSlot {  // one in-flight chunk (count) or walk window (locate) of a device's pipeline
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr, arrived = nullptr;  // arrived: the chunk's H2D (on the device's copy-in stream) is complete
  GrowBuf dQueries, dOffsets, dCounts, dPos, dSeq, dLoc;
  GrowBuf hIn, hOffsets, hOut;
  LocateScratch sc;
  std::vector<PendingCopy> pending;
  bool busy = false;
  PackSlot() { hIn.host = hOffsets.host = hOut.host = true; }
};

struct GroupDevice {
  awfm_gpu_ctx *ctx = nullptr;
  bool owned = false, ready = false;
#ifndef AWFM_GROUP_SLOTS
#define AWFM_GROUP_SLOTS 4
#endif
  // one more slot than stages (H2D, search, D2H): a chunk's H2D can be queued while the counts of the chunk three
  // before it are still on their way out
  static constexpr int kSlots = AWFM_GROUP_SLOTS;
  PackSlot slots[kSlots];
  cudaStream_t copyIn = nullptr;  // every H2D of the pipeline, in chunk order: chunk i is complete before chunk i+1 starts
  GrowBuf dRanges, dHit, dCounts;  // locate: the shard's ranges, counts and hit offsets stay on the device between the two phases
  uint64_t *hTotal = nullptr;  // page-locked
  cudaEvent_t rebased = nullptr;  // locate phase B: the shard's hit offsets carry their global base
  uint64_t total = 0, base = 0;
  int rc = AWFM_GPU_OK;
  std::string err;
  awfm_gpu_stats stats{};
};

}  // namespace

struct awfm_gpu_group {
  std::vector<std::unique_ptr<GroupDevice>> dev;
  int64_t chunkQueries = 3ll << 23, minShard = 1ll << 16, windowHits = 1ll << 21;
  std::mutex mu;  // one packed-batch call at a time per group
};

namespace {

int prepareDevice(GroupDevice &D) {
  if (D.ready) return AWFM_GPU_OK;
  CU(cudaSetDevice(D.ctx->device));
  for (auto &s : D.slots) {
    if (!s.stream) CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (!s.done) CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    if (!s.arrived) CU(cudaEventCreateWithFlags(&s.arrived, cudaEventDisableTiming));
    if (!s.sc.dWorkCounter) CU(cudaMalloc(&s.sc.dWorkCounter, 64));
  }
  if (!D.copyIn) CU(cudaStreamCreateWithFlags(&D.copyIn, cudaStreamNonBlocking));
  if (!D.hTotal) CU(cudaHostAlloc(&D.hTotal, 64, cudaHostAllocPortable));
  if (!D.rebased) CU(cudaEventCreateWithFlags(&D.rebased, cudaEventDisableTiming));
  D.ready = true;
  return AWFM_GPU_OK;
}

void releaseDevice(GroupDevice &D) {
  if (!D.ctx) return;
  cudaSetDevice(D.ctx->device);
  for (auto &s : D.slots) {
    if (s.stream) {
      cudaStreamSynchronize(s.stream);
      cudaStreamDestroy(s.stream);
    }
    if (s.done) cudaEventDestroy(s.done);
    if (s.arrived) cudaEventDestroy(s.arrived);
    s.arrived = nullptr;
    for (GrowBuf *b : {&s.dQueries, &s.dOffsets, &s.dCounts, &s.dPos, &s.dSeq, &s.dLoc, &s.hIn, &s.hOffsets, &s.hOut}) b->release();
    cudaFree(s.sc.scanTemp);
    cudaFree(s.sc.dWorkCounter);
    s.sc = LocateScratch();
    s.stream = nullptr, s.done = nullptr;
  }
  if (D.copyIn) {
    cudaStreamSynchronize(D.copyIn);
    cudaStreamDestroy(D.copyIn);
    D.copyIn = nullptr;
  }
  D.dRanges.release();
  D.dHit.release();
  D.dCounts.release();
  if (D.hTotal) cudaFreeHost(D.hTotal);
  D.hTotal = nullptr;
  if (D.rebased) cudaEventDestroy(D.rebased);
  D.rebased = nullptr;
  cudaGetLastError();
  if (D.owned) awfm_gpu_ctx_destroy(D.ctx);
  D.ctx = nullptr;
}

// what the caller handed in, shared read-only by the device threads
struct Job {
  const uint8_t *queries = nullptr;
  const uint64_t *offsets = nullptr;
  uint32_t format = AWFM_QUERY_ASCII, fixedLen = 0;
  uint64_t n = 0, queryBytes = 0;
  bool inPinned = false, offsetsPinned = false;
  // count
  uint32_t *counts = nullptr;
  bool countsPinned = false;
  // locate
  uint64_t *hitOffsets = nullptr, *positions = nullptr, *sequenceIndex = nullptr, *localPosition = nullptr;
  bool hitPinned = false, posPinned = false, seqPinned = false, locPinned = false, walk = false;
};

int finishSlot(PackSlot &s) {  // the slot's D2H copies are complete; pageable destinations are filled from staging
  if (!s.busy) return AWFM_GPU_OK;
  CU(cudaEventSynchronize(s.done));
  for (const PendingCopy &p : s.pending) memcpy(p.dst, p.src, p.bytes);
  s.pending.clear();
  s.busy = false;
  return AWFM_GPU_OK;
}

// device -> caller's host array; page-locked destinations are written in place, pageable ones through `stage`
int copyOut(PackSlot &s, GrowBuf &stage, size_t stageOffset, void *dst, bool dstPinned, const void *dSrc, size_t bytes) {
  if (bytes == 0) return AWFM_GPU_OK;
  if (dstPinned) {
    CU(cudaMemcpyAsync(dst, dSrc, bytes, cudaMemcpyDeviceToHost, s.stream));
  } else {
    uint8_t *h = (uint8_t *)stage.p + stageOffset;
    CU(cudaMemcpyAsync(h, dSrc, bytes, cudaMemcpyDeviceToHost, s.stream));
    s.pending.push_back(PendingCopy{dst, h, bytes});
  }
  return AWFM_GPU_OK;
}

// H2D of queries [q0, q0+m) into the slot, and the batch descriptor the kernels take.  The copy is queued on the
// device's ONE copy-in stream (copies queued on several streams are interleaved by the copy engine, so every chunk
// would arrive late); the slot's stream — search and D2H — waits for it.  The caller has made sure the slot's previous
// chunk is finished (its buffers are rewritten here).
int shipChunk(const Job &job, cudaStream_t copyIn, PackSlot &s, uint64_t q0, uint64_t m, PackedBatch *batch, uint64_t *h2dBytes) {
  const bool variable = job.offsets != nullptr;
  const uint64_t b0 = variable ? job.offsets[q0] : q0 * job.queryBytes;
  const uint64_t nbytes = variable ? job.offsets[q0 + m] - b0 : m * job.queryBytes;
  // variable length: the kernels index the letters by the caller's own offsets, so the chunk is placed where
  // (device base - b0) is 16-B aligned and the batch's letter pointer is that (virtual) base
  const uint64_t pad = variable ? (b0 & 15u) : 0;
  if (int r = s.dQueries.ensure(nbytes + pad + 32)) return r;
  const uint8_t *src = job.queries + b0;
  if (!job.inPinned && nbytes) {
    if (int r = s.hIn.ensure(nbytes)) return r;
    memcpy(s.hIn.p, src, nbytes);
    src = (const uint8_t *)s.hIn.p;
  }
  if (nbytes) CU(cudaMemcpyAsync((uint8_t *)s.dQueries.p + pad, src, nbytes, cudaMemcpyHostToDevice, copyIn));
  *h2dBytes += nbytes;
  batch->format = job.format;
  batch->length = job.fixedLen;
  batch->numQueries = m;
  batch->offsets = nullptr;
  batch->data = (const uint8_t *)s.dQueries.p;
  if (variable) {
    if (int r = s.dOffsets.ensure((m + 1) * 8)) return r;
    const uint64_t *osrc = job.offsets + q0;
    if (!job.offsetsPinned) {
      if (int r = s.hOffsets.ensure((m + 1) * 8)) return r;
      memcpy(s.hOffsets.p, osrc, (m + 1) * 8);
      osrc = (const uint64_t *)s.hOffsets.p;
    }
    CU(cudaMemcpyAsync(s.dOffsets.p, osrc, (m + 1) * 8, cudaMemcpyHostToDevice, copyIn));
    *h2dBytes += (m + 1) * 8;
    batch->offsets = (const uint64_t *)s.dOffsets.p;
    batch->data = (const uint8_t *)s.dQueries.p + pad - b0;  // only ever dereferenced at + offsets[q] >= b0
  }
  CU(cudaEventRecord(s.arrived, copyIn));
  CU(cudaStreamWaitEvent(s.stream, s.arrived, 0));
  return AWFM_GPU_OK;
}

void addStats(awfm_gpu_stats &into, const awfm_gpu_stats &s) {
  into.launches += s.launches;
  into.queries += s.queries;
  into.hits += s.hits;
  into.h2dBytes += s.h2dBytes;
  into.d2hBytes += s.d2hBytes;
  into.kernelMs += s.kernelMs;
}

// Chunk boundaries of a shard.  The first and the last chunk are half-sized: the pipeline's fill (H2D of the first chunk,
// nothing to overlap it with) and drain (D2H of the last chunk's results) shrink, while the chunks in between stay large
// enough for the sweep path's per-call cost (the index is streamed once per LF step whatever the batch size).
std::vector<uint64_t> chunkStarts(uint64_t qa, uint64_t qb, uint64_t chunk) {
  std::vector<uint64_t> starts{qa};
  const uint64_t half = std::max<uint64_t>(256, (chunk / 2 + 255) & ~255ull);
  if (chunk < 1024) {  // too small to be worth a ramp (tests): plain chunks
    for (uint64_t pos = qa + chunk; pos < qb; pos += chunk) starts.push_back(pos);
  } else if (qb - qa <= chunk + half) {
    if (qb - qa > chunk) starts.push_back(qa + (((qb - qa) / 2 + 255) & ~255ull));
  } else {
    uint64_t pos = qa + half;
    while (qb - pos > chunk + half) {
      starts.push_back(pos);
      pos += chunk;
    }
    starts.push_back(pos);
    if (qb - pos > chunk) starts.push_back(qb - half - ((qb - half - pos) & 255ull));
  }
  starts.push_back(qb);
  return starts;
}

// ---- count: one device's shard [qa, qb) ----
int countShard(awfm_gpu_group *g, GroupDevice &D, const Job &job, uint64_t qa, uint64_t qb) {
  awfm_gpu_ctx *c = D.ctx;
  LaneHold hold(c);  // makes the device current
  if (hold.rc) return hold.rc;
  Lane &L = *hold;
  awfm_begin_call(L);
  if (int r = prepareDevice(D)) return r;
  const std::vector<uint64_t> starts = chunkStarts(qa, qb, (uint64_t)g->chunkQueries);
  uint64_t h2d = 0, d2h = 0;
  int rc = AWFM_GPU_OK;
  const bool verbose = getenv("AWFM_GPU_VERBOSE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  auto ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  for (size_t k = 0; k + 1 < starts.size() && rc == AWFM_GPU_OK; k++) {
    PackSlot &s = D.slots[k % GroupDevice::kSlots];
    const uint64_t q0 = starts[k], m = starts[k + 1] - q0;
    const double tw = ms();
    if ((rc = finishSlot(s))) break;
    if (verbose)
      fprintf(stderr, "[awfm_gpu] packed count dev %d chunk %zu (%llu queries): waited for its slot %.3f -> %.3f ms\n",
              c->device, k, (unsigned long long)m, tw, ms());
    PackedBatch b;
    if ((rc = shipChunk(job, D.copyIn, s, q0, m, &b, &h2d))) break;
    if ((rc = s.dCounts.ensure(m * 4))) break;
    if ((rc = awfm_count_device_impl(c, L, b, (uint32_t *)s.dCounts.p, nullptr, s.stream))) break;
    if (!job.countsPinned && (rc = s.hOut.ensure(m * 4))) break;
    if ((rc = copyOut(s, s.hOut, 0, job.counts + q0, job.countsPinned, s.dCounts.p, m * 4))) break;
    cudaError_t e = cudaEventRecord(s.done, s.stream);
    if (e != cudaSuccess) {
      rc = awfm_fail(AWFM_GPU_ERR_CUDA, "packed count pipeline", cudaGetErrorString(e));
      break;
    }
    s.busy = true;
    d2h += m * 4;
    if (verbose) fprintf(stderr, "[awfm_gpu] packed count dev %d chunk %zu enqueued at %.3f ms\n", c->device, k, ms());
  }
  for (auto &s : D.slots) {  // drain (also on errors: never leave a DMA into the caller's memory in flight)
    const int r = finishSlot(s);
    if (rc == AWFM_GPU_OK) rc = r;
    if (r) cudaStreamSynchronize(s.stream), s.busy = false, s.pending.clear();
    if (verbose) fprintf(stderr, "[awfm_gpu] packed count dev %d: a slot drained at %.3f ms\n", c->device, ms());
  }
  if (rc != AWFM_GPU_OK) cudaStreamSynchronize(D.copyIn);  // (a copy queued for a chunk that was never searched)
  L.stats.h2dBytes = h2d;
  L.stats.d2hBytes = d2h;
  addStats(D.stats, L.stats);
  return rc;
}

// ---- locate, phase A: ranges of the whole shard, their scan, the shard's hit total ----
int locateShardRanges(awfm_gpu_group *g, GroupDevice &D, const Job &job, uint64_t qa, uint64_t qb) {
  awfm_gpu_ctx *c = D.ctx;
  LaneHold hold(c);
  if (hold.rc) return hold.rc;
  Lane &L = *hold;
  awfm_begin_call(L);
  if (int r = prepareDevice(D)) return r;
  const uint64_t shard = qb - qa;
  D.total = 0;
  if (shard == 0) return AWFM_GPU_OK;
  if (int r = D.dRanges.ensure(shard * 16)) return r;
  if (int r = D.dHit.ensure((shard + 1) * 8)) return r;
  if (int r = D.dCounts.ensure(shard * 4)) return r;
  const std::vector<uint64_t> starts = chunkStarts(qa, qb, (uint64_t)g->chunkQueries);
  uint64_t h2d = 0;
  for (size_t k = 0; k + 1 < starts.size(); k++) {
    PackSlot &s = D.slots[k % GroupDevice::kSlots];
    const uint64_t q0 = starts[k], m = starts[k + 1] - q0;
    if (k >= GroupDevice::kSlots) CU(cudaStreamSynchronize(s.stream));  // the slot's input buffer is about to be rewritten
    PackedBatch b;
    if (int r = shipChunk(job, D.copyIn, s, q0, m, &b, &h2d)) return r;
    b.rangesOfHitsOnly = true;  // the hit offsets are scanned from the counts; the walk reads the ranges of hits only
    if (int r = awfm_count_device_impl(c, L, b, (uint32_t *)D.dCounts.p + (q0 - qa), (awfm_range *)D.dRanges.p + (q0 - qa), s.stream))
      return r;
  }
  for (auto &s : D.slots) CU(cudaStreamSynchronize(s.stream));
  PackSlot &s0 = D.slots[0];
  if (int r = awfm_scan_impl(c, L, s0.sc, D.dCounts.p, true, shard, (uint64_t *)D.dHit.p, 0, s0.stream)) return r;
  CU(cudaMemcpyAsync(D.hTotal, (uint64_t *)D.dHit.p + shard, 8, cudaMemcpyDeviceToHost, s0.stream));
  CU(cudaStreamSynchronize(s0.stream));
  D.total = *D.hTotal;
  L.stats.h2dBytes = h2d;
  L.stats.d2hBytes = 8;
  addStats(D.stats, L.stats);
  return AWFM_GPU_OK;
}

// ---- locate, phase B: global hit offsets out, then the walk in windows of flat hit indices ----
int locateShardWalk(awfm_gpu_group *g, GroupDevice &D, const Job &job, uint64_t qa, uint64_t qb, bool lastShard) {
  awfm_gpu_ctx *c = D.ctx;
  LaneHold hold(c);
  if (hold.rc) return hold.rc;
  Lane &L = *hold;
  awfm_begin_call(L);
  const uint64_t shard = qb - qa;
  if (shard == 0) return AWFM_GPU_OK;
  uint64_t d2h = 0;
  int rc = AWFM_GPU_OK;
  PackSlot &s0 = D.slots[0];
  uint64_t *dHit = (uint64_t *)D.dHit.p;
  if (D.base) {
    addHitBase<<<std::min<unsigned>((unsigned)((shard + 256) / 256), (unsigned)c->numSMs * 8u), 256, 0, s0.stream>>>(dHit, shard + 1, D.base);
    CU(cudaGetLastError());
    L.stats.launches += 1;
  }
  {  // hitOffsets[qa .. qb) (+ the final entry from the last shard)
    // the walk windows below run on other streams: they must see the rebased offsets, but need not wait for the copy
    CU(cudaEventRecord(D.rebased, s0.stream));
    for (int i = 1; i < GroupDevice::kSlots; i++) CU(cudaStreamWaitEvent(D.slots[i].stream, D.rebased, 0));
    const uint64_t entries = shard + (lastShard ? 1 : 0);
    if (!job.hitPinned && (rc = s0.hOut.ensure(entries * 8))) return rc;
    if ((rc = copyOut(s0, s0.hOut, 0, job.hitOffsets + qa, job.hitPinned, dHit, entries * 8))) return rc;
    CU(cudaEventRecord(s0.done, s0.stream));
    s0.busy = true;
    d2h += entries * 8;
  }
  if (job.walk && D.total) {
    const uint64_t window = (uint64_t)g->windowHits;
    const bool mapped = job.sequenceIndex != nullptr;
    uint64_t k = 1;  // slot 0 is still busy with the hit offsets: start on slot 1
    for (uint64_t hb = D.base; hb < D.base + D.total && rc == AWFM_GPU_OK; hb += window, k++) {
      PackSlot &s = D.slots[k % GroupDevice::kSlots];
      const uint64_t he = std::min(D.base + D.total, hb + window), cnt = he - hb;
      if ((rc = finishSlot(s))) break;
      if ((rc = s.dPos.ensure(cnt * 8))) break;
      if ((rc = awfm_locate_device_impl(c, L, s.sc, (const awfm_range *)D.dRanges.p, dHit, shard, hb, he, (uint64_t *)s.dPos.p, s.stream))) break;
      const int outs = mapped ? 3 : 1;
      const bool staged = !job.posPinned || (mapped && (!job.seqPinned || !job.locPinned));
      if (staged && (rc = s.hOut.ensure(cnt * 8 * outs))) break;
      if ((rc = copyOut(s, s.hOut, 0, job.positions + hb, job.posPinned, s.dPos.p, cnt * 8))) break;
      if (mapped) {
        if ((rc = s.dSeq.ensure(cnt * 8)) || (rc = s.dLoc.ensure(cnt * 8))) break;
        if ((rc = awfm_map_device_impl(c, L, (const uint64_t *)s.dPos.p, cnt, (uint64_t *)s.dSeq.p, (uint64_t *)s.dLoc.p, s.stream))) break;
        if ((rc = copyOut(s, s.hOut, cnt * 8, job.sequenceIndex + hb, job.seqPinned, s.dSeq.p, cnt * 8))) break;
        if ((rc = copyOut(s, s.hOut, cnt * 16, job.localPosition + hb, job.locPinned, s.dLoc.p, cnt * 8))) break;
      }
      cudaError_t e = cudaEventRecord(s.done, s.stream);
      if (e != cudaSuccess) {
        rc = awfm_fail(AWFM_GPU_ERR_CUDA, "packed locate pipeline", cudaGetErrorString(e));
        break;
      }
      s.busy = true;
      d2h += cnt * 8 * outs;
    }
  }
  for (auto &s : D.slots) {
    const int r = finishSlot(s);
    if (rc == AWFM_GPU_OK) rc = r;
    if (r) cudaStreamSynchronize(s.stream), s.busy = false, s.pending.clear();
  }
  L.stats.d2hBytes = d2h;
  addStats(D.stats, L.stats);
  return rc;
}

// Runs fn(deviceIndex) for every device with a non-empty shard: inline for one, one host thread per device otherwise.
template <typename F>
int forEachDevice(awfm_gpu_group *g, int used, F fn) {
  if (used <= 1) {
    GroupDevice &D = *g->dev[0];
    D.rc = fn(0);
    if (D.rc) D.err = awfm_gpu_last_error();
    return D.rc;
  }
  std::vector<std::thread> threads;
  for (int d = 0; d < used; d++)
    threads.emplace_back([g, d, &fn]() {
      GroupDevice &D = *g->dev[d];
      D.rc = fn(d);
      if (D.rc) D.err = awfm_gpu_last_error();  // the message is thread-local: carry it to the caller
    });
  for (auto &t : threads) t.join();
  for (int d = 0; d < used; d++)
    if (g->dev[d]->rc) return awfm_fail(g->dev[d]->rc, g->dev[d]->err.c_str());
  return AWFM_GPU_OK;
}

struct Shards {
  uint64_t per = 0;
  int used = 1;
  uint64_t begin(int d, uint64_t n) const { return std::min(n, (uint64_t)d * per); }
  uint64_t end(int d, uint64_t n) const { return std::min(n, (uint64_t)(d + 1) * per); }
};
// contiguous shards, multiples of 256 queries (every shard and chunk of a fixed-length batch then starts 16-B aligned),
// none smaller than "packed_min_shard" (a tiny batch is not worth a second GPU's launch latency)
Shards makeShards(const awfm_gpu_group *g, uint64_t n) {
  Shards sh;
  const uint64_t G = g->dev.size();
  uint64_t per = (n + G - 1) / G;
  per = std::max<uint64_t>(per, (uint64_t)g->minShard);
  per = (per + 255) & ~255ull;
  sh.per = per;
  sh.used = (int)std::max<uint64_t>(1, std::min<uint64_t>(G, (n + per - 1) / per));
  return sh;
}

int checkJob(const awfm_gpu_group *g, const void *queries, uint32_t format, const uint64_t *offsets, uint32_t fixedLen,
             uint64_t n, Job *job) {
  if (!g || g->dev.empty()) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  if (n && !queries) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  if (format != AWFM_QUERY_ASCII && format != AWFM_QUERY_2BIT && format != AWFM_QUERY_5BIT)
    return awfm_fail(AWFM_GPU_ERR_ARG, "unknown query format");
  if (format != AWFM_QUERY_ASCII && offsets) return awfm_fail(AWFM_GPU_ERR_ARG, "2-/5-bit query batches are fixed-length");
  if (!offsets && fixedLen == 0 && n) return awfm_fail(AWFM_GPU_ERR_ARG, "fixedLen must be > 0 when offsets is NULL");
  const bool amino = g->dev[0]->ctx->ix.amino != 0;
  if (format != AWFM_QUERY_ASCII && (format == AWFM_QUERY_5BIT) != amino)
    return awfm_fail(AWFM_GPU_ERR_ARG, "query format does not match the index alphabet (2-bit: nucleotide, 5-bit: amino)");
  job->queries = (const uint8_t *)queries;
  job->offsets = offsets;
  job->format = format;
  job->fixedLen = fixedLen;
  job->n = n;
  job->queryBytes = awfm_query_bytes(format, fixedLen);
  job->inPinned = n && awfm_is_pinned_host(queries);
  job->offsetsPinned = offsets && awfm_is_pinned_host(offsets);
  return AWFM_GPU_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ group lifecycle
extern "C" int awfm_gpu_group_create_from_contexts(awfm_gpu_group **out, awfm_gpu_ctx *const *contexts, int count) {
  if (!out || !contexts || count < 1) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  for (int i = 0; i < count; i++) {
    if (!contexts[i]) return awfm_fail(AWFM_GPU_ERR_ARG, "null context");
    for (int j = 0; j < i; j++)
      if (contexts[j] == contexts[i]) return awfm_fail(AWFM_GPU_ERR_ARG, "a context may appear in a group only once");
    if (contexts[i]->ix.bwtLength != contexts[0]->ix.bwtLength || contexts[i]->ix.amino != contexts[0]->ix.amino ||
        contexts[i]->ix.seedK != contexts[0]->ix.seedK)
      return awfm_fail(AWFM_GPU_ERR_ARG, "the contexts of a group must hold the same index");
  }
  awfm_gpu_group *g = new awfm_gpu_group();
  for (int i = 0; i < count; i++) {
    g->dev.emplace_back(new GroupDevice());
    g->dev.back()->ctx = contexts[i];
  }
  *out = g;
  return AWFM_GPU_OK;
}

extern "C" int awfm_gpu_group_create(awfm_gpu_group **out, const int *devices, int numDevices, const awfm_index_view *view) {
  if (!out || !view) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  std::vector<int> list;
  if (devices && numDevices > 0) list.assign(devices, devices + numDevices);
  else {
    const int n = awfm_gpu_device_count();
    if (n <= 0) return awfm_fail(AWFM_GPU_ERR_NO_DEVICE, "no CUDA device");
    for (int i = 0; i < n; i++) list.push_back(i);
  }
  // the replicas are uploaded concurrently: every GPU has its own PCIe link
  std::vector<awfm_gpu_ctx *> ctxs(list.size(), nullptr);
  std::vector<int> rcs(list.size(), AWFM_GPU_OK);
  std::vector<std::string> errs(list.size());
  if (list.size() == 1) {
    rcs[0] = awfm_gpu_ctx_create(&ctxs[0], list[0], view);
    if (rcs[0]) errs[0] = awfm_gpu_last_error();
  } else {
    std::vector<std::thread> threads;
    for (size_t i = 0; i < list.size(); i++)
      threads.emplace_back([&, i]() {
        rcs[i] = awfm_gpu_ctx_create(&ctxs[i], list[i], view);
        if (rcs[i]) errs[i] = awfm_gpu_last_error();
      });
    for (auto &t : threads) t.join();
  }
  for (size_t i = 0; i < list.size(); i++)
    if (rcs[i]) {
      for (auto *c : ctxs) awfm_gpu_ctx_destroy(c);
      return awfm_fail(rcs[i], errs[i].c_str());
    }
  awfm_gpu_group *g = new awfm_gpu_group();
  for (auto *c : ctxs) {
    g->dev.emplace_back(new GroupDevice());
    g->dev.back()->ctx = c;
    g->dev.back()->owned = true;
  }
  *out = g;
  return AWFM_GPU_OK;
}

extern "C" void awfm_gpu_group_destroy(awfm_gpu_group *g) {
  if (!g) return;
  {
    std::lock_guard<std::mutex> lock(g->mu);
    for (auto &d : g->dev) releaseDevice(*d);
  }
  delete g;
}

extern "C" int awfm_gpu_group_size(const awfm_gpu_group *g) { return g ? (int)g->dev.size() : 0; }
extern "C" awfm_gpu_ctx *awfm_gpu_group_context(awfm_gpu_group *g, int i) {
  return (g && i >= 0 && i < (int)g->dev.size()) ? g->dev[i]->ctx : nullptr;
}

extern "C" int awfm_gpu_group_set_sequences(awfm_gpu_group *g, const void *metadata, uint64_t numSequences) {
  if (!g) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  for (auto &d : g->dev)
    if (int r = awfm_gpu_ctx_set_sequences(d->ctx, metadata, numSequences)) return r;
  return AWFM_GPU_OK;
}

extern "C" int awfm_gpu_group_set_tuning(awfm_gpu_group *g, const char *key, int64_t value) {
  if (!g || !key) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  const std::string k(key);
  if (k == "packed_chunk_queries") {
    if (value < 256 || value > (1ll << 30)) return awfm_fail(AWFM_GPU_ERR_ARG, "bad value", key);
    g->chunkQueries = (value + 255) & ~255ll;
  } else if (k == "packed_min_shard") {
    if (value < 1 || value > (1ll << 40)) return awfm_fail(AWFM_GPU_ERR_ARG, "bad value", key);
    g->minShard = value;
  } else if (k == "packed_window_hits") {
    if (value < 1 || value > (1ll << 32)) return awfm_fail(AWFM_GPU_ERR_ARG, "bad value", key);
    g->windowHits = value;
  } else {
    for (auto &d : g->dev)
      if (int r = awfm_gpu_ctx_set_tuning(d->ctx, key, value)) return r;
  }
  return AWFM_GPU_OK;
}

extern "C" int awfm_gpu_group_get_stats(awfm_gpu_group *g, awfm_gpu_stats *out) {
  if (!g || !out) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  *out = awfm_gpu_stats{};
  for (auto &d : g->dev) addStats(*out, d->stats);
  return AWFM_GPU_OK;
}

// ------------------------------------------------------------------------------------------------ packed batches
extern "C" int awfm_gpu_group_count(awfm_gpu_group *g, const void *queries, uint32_t format, const uint64_t *offsets,
                                    uint32_t fixedLen, uint64_t n, uint32_t *counts) {
  Job job;
  if (int r = checkJob(g, queries, format, offsets, fixedLen, n, &job)) return r;
  if (n && !counts) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> lock(g->mu);
  for (auto &d : g->dev) d->stats = awfm_gpu_stats{}, d->rc = AWFM_GPU_OK;
  if (n == 0) return AWFM_GPU_OK;
  job.counts = counts;
  job.countsPinned = awfm_is_pinned_host(counts);
  const Shards sh = makeShards(g, n);
  return forEachDevice(g, sh.used, [&](int d) { return countShard(g, *g->dev[d], job, sh.begin(d, n), sh.end(d, n)); });
}

extern "C" int awfm_gpu_group_locate(awfm_gpu_group *g, const void *queries, uint32_t format, const uint64_t *offsets,
                                     uint32_t fixedLen, uint64_t n, uint64_t *hitOffsets, uint64_t *positions,
                                     uint64_t positionsCapacity, uint64_t *sequenceIndex, uint64_t *localPosition,
                                     uint64_t *totalHits) {
  Job job;
  if (int r = checkJob(g, queries, format, offsets, fixedLen, n, &job)) return r;
  if (!hitOffsets) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  if ((sequenceIndex == nullptr) != (localPosition == nullptr))
    return awfm_fail(AWFM_GPU_ERR_ARG, "sequenceIndex and localPosition go together");
  std::lock_guard<std::mutex> lock(g->mu);
  for (auto &d : g->dev) {
    d->stats = awfm_gpu_stats{}, d->rc = AWFM_GPU_OK, d->total = d->base = 0;
    if (!d->ctx->hasSa && positions) return awfm_fail(AWFM_GPU_ERR_NO_SA, "context was created without a sampled suffix array");
    if (sequenceIndex && !d->ctx->ix.sequenceEnds)
      return awfm_fail(AWFM_GPU_ERR_ARG, "context has no sequence table (awfm_gpu_group_set_sequences)");
  }
  if (totalHits) *totalHits = 0;
  hitOffsets[0] = 0;
  if (n == 0) return AWFM_GPU_OK;
  job.hitOffsets = hitOffsets, job.positions = positions, job.sequenceIndex = sequenceIndex, job.localPosition = localPosition;
  job.hitPinned = awfm_is_pinned_host(hitOffsets);
  const Shards sh = makeShards(g, n);
  if (int r = forEachDevice(g, sh.used, [&](int d) { return locateShardRanges(g, *g->dev[d], job, sh.begin(d, n), sh.end(d, n)); }))
    return r;
  uint64_t total = 0;
  for (int d = 0; d < sh.used; d++) {
    g->dev[d]->base = total;
    total += g->dev[d]->total;
  }
  if (totalHits) *totalHits = total;
  job.walk = positions != nullptr && positionsCapacity >= total && total > 0;
  if (job.walk) {
    job.posPinned = awfm_is_pinned_host(positions);
    job.seqPinned = sequenceIndex && awfm_is_pinned_host(sequenceIndex);
    job.locPinned = localPosition && awfm_is_pinned_host(localPosition);
  }
  return forEachDevice(g, sh.used, [&](int d) {
    return locateShardWalk(g, *g->dev[d], job, sh.begin(d, n), sh.end(d, n), d == sh.used - 1);
  });
}

// ------------------------------------------------------------------------------------------------ the list layout
static int groupList(awfm_gpu_group *g, awfm_kmer_search_data *data, uint64_t n, uint32_t numThreads, bool locate) {
  if (!g || g->dev.empty()) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  std::vector<awfm_gpu_ctx *> ctxs;
  // a list too short to give every device a chunk uses fewer devices
  const int64_t countChunk = g->dev[0]->ctx->chunkQueries > 0 ? g->dev[0]->ctx->chunkQueries : (1ll << 16);  // 0 = automatic, at least 2^16
  const uint64_t chunk = (uint64_t)(locate ? g->dev[0]->ctx->locateChunkQueries : countChunk);
  const uint64_t chunks = std::max<uint64_t>(1, (n + chunk - 1) / chunk);
  for (size_t d = 0; d < g->dev.size() && d < chunks; d++) ctxs.push_back(g->dev[d]->ctx);
  const int rc = awfm_search_list_run(ctxs.data(), (int)ctxs.size(), data, n, numThreads, locate);
  for (auto &d : g->dev) d->stats = awfm_gpu_stats{};
  for (size_t d = 0; d < ctxs.size(); d++) awfm_gpu_ctx_get_stats(ctxs[d], &g->dev[d]->stats);
  return rc;
}
extern "C" int awfm_gpu_group_search_list_count(awfm_gpu_group *g, awfm_kmer_search_data *data, uint64_t n,
                                                uint32_t numThreads) {
  return groupList(g, data, n, numThreads, false);
}
extern "C" int awfm_gpu_group_search_list_locate(awfm_gpu_group *g, awfm_kmer_search_data *data, uint64_t n,
                                                 uint32_t numThreads) {
  return groupList(g, data, n, numThreads, true);
}

// ------------------------------------------------------------------------------------------------ host / peer memory
extern "C" int awfm_gpu_host_alloc(void **p, uint64_t bytes) {
  if (!p) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  CU(cudaHostAlloc(p, bytes ? bytes : 16, cudaHostAllocPortable));
  return AWFM_GPU_OK;
}
extern "C" void awfm_gpu_host_free(void *p) {
  if (p) cudaFreeHost(p);
  cudaGetLastError();
}
extern "C" int awfm_gpu_host_register(void *p, uint64_t bytes) {
  if (!p) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  CU(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return AWFM_GPU_OK;
}
extern "C" int awfm_gpu_host_unregister(void *p) {
  if (!p) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  CU(cudaHostUnregister(p));
  return AWFM_GPU_OK;
}

extern "C" int awfm_gpu_device_malloc(int device, void **dPtr, uint64_t bytes) {
  if (!dPtr) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  CU(cudaSetDevice(device));
  CU(cudaMalloc(dPtr, bytes ? bytes : 16));
  return AWFM_GPU_OK;
}
extern "C" int awfm_gpu_device_free(int device, void *dPtr) {
  CU(cudaSetDevice(device));
  CU(cudaFree(dPtr));
  return AWFM_GPU_OK;
}
extern "C" int awfm_gpu_ipc_export(int device, const void *dPtr, void *handle64) {
  if (!dPtr || !handle64) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  CU(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, const_cast<void *>(dPtr)));
  memcpy(handle64, &h, 64);
  return AWFM_GPU_OK;
}
extern "C" int awfm_gpu_ipc_open(int device, const void *handle64, void **dPtr) {
  if (!dPtr || !handle64) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  CU(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CU(cudaIpcOpenMemHandle(dPtr, h, cudaIpcMemLazyEnablePeerAccess));
  return AWFM_GPU_OK;
}
extern "C" int awfm_gpu_ipc_close(int device, void *dPtr) {
  CU(cudaSetDevice(device));
  CU(cudaIpcCloseMemHandle(dPtr));
  return AWFM_GPU_OK;
}
extern "C" int awfm_gpu_peer_copy_async(int device, void *dst, const void *src, uint64_t bytes, void *stream) {
  if (bytes && (!dst || !src)) return awfm_fail(AWFM_GPU_ERR_ARG, "null argument");
  CU(cudaSetDevice(device));
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return AWFM_GPU_OK;
}
