// granularity_probe.cu — measurement tool (not product): how many independent pseudo-random records per second
// can one B200 read from a >L2 array, as a function of record size, lanes per record, records in flight per lane
// group and the PTX load flavour (cache operators / L2 prefetch-size hints)?  Answers whether the 128 B that
// ncu sees fetched from DRAM per missing 64-B half-line request (profiles/r01_ncu_count_halfline.json) can be
// avoided by a load qualifier.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o granularity_probe
// granularity_probe.cu ; run under gpurun, one JSON line per configuration.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));    \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

enum Flavor { NC = 0, NC_L2_64, NC_L2_128, NC_L2_256, CG, CV, NC_NOALLOC, NC_EVICT_FIRST_HINT, LU, NUM_FLAVORS };
static const char *kFlavorNames[] = {"ld.global.nc", "ld.global.nc.L2::64B", "ld.global.nc.L2::128B",
                                     "ld.global.nc.L2::256B", "ld.global.cg", "ld.global.cv",
                                     "ld.global.nc.L1::no_allocate", "ld.global.nc.L2::cache_hint(evict_first)",
                                     "ld.global.lu"};

template <int F>
__device__ __forceinline__ uint4 load16(const uint4 *p, uint64_t policy) {
  uint4 v;
  if constexpr (F == NC)
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if constexpr (F == NC_L2_64)
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if constexpr (F == NC_L2_128)
    asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if constexpr (F == NC_L2_256)
    asm volatile("ld.global.nc.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if constexpr (F == CG)
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if constexpr (F == CV)
    asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if constexpr (F == NC_NOALLOC)
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if constexpr (F == NC_EVICT_FIRST_HINT)
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(policy));
  else
    asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// BYTES per record, LANES consecutive lanes share a record (16 B per lane per load), UNROLL records in flight.
template <int BYTES, int LANES, int UNROLL, int F>
__global__ void __launch_bounds__(256) probe(const uint4 *__restrict__ data, uint64_t numRecords, uint64_t numReads,
                                             uint64_t *__restrict__ sink) {
  constexpr int U4 = BYTES / 16;
  constexpr int PER_LANE = (U4 + LANES - 1) / LANES;
  uint64_t policy = 0;
  if constexpr (F == NC_EVICT_FIRST_HINT) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  const unsigned sub = threadIdx.x % LANES;
  const uint64_t numGroups = (uint64_t)gridDim.x * blockDim.x / LANES;
  uint64_t acc = 0;
  for (uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES; i0 < numReads; i0 += numGroups * UNROLL) {
    uint4 v[UNROLL][PER_LANE];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const uint64_t i = i0 + (uint64_t)u * numGroups;
      const uint64_t rec = mix(i) % numRecords;
#pragma unroll
      for (int k = 0; k < PER_LANE; k++) {
        const int c = sub + LANES * k;
        v[u][k] = (c < U4 && i < numReads) ? load16<F>(data + rec * U4 + c, policy) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
#pragma unroll
      for (int k = 0; k < PER_LANE; k++) acc += v[u][k].x ^ v[u][k].y ^ v[u][k].z ^ v[u][k].w;
  }
  if (acc == 0x1234567ull) sink[0] = acc;
}

template <int BYTES, int LANES, int UNROLL, int F>
static void run(const uint4 *data, uint64_t arrayBytes, uint64_t numReads, uint64_t *sink, int ctasPerSm) {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<BYTES, LANES, UNROLL, F>, 256, 0));
  if (ctasPerSm > 0 && ctasPerSm < occ) occ = ctasPerSm;
  const int grid = sms * occ;
  const uint64_t numRecords = arrayBytes / BYTES;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(a));
    probe<BYTES, LANES, UNROLL, F><<<grid, 256>>>(data, numRecords, numReads, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  printf("{\"record_bytes\": %d, \"lanes\": %d, \"unroll\": %d, \"flavor\": \"%s\", \"ctas_per_sm\": %d, \"array_gb\": %.2f, "
         "\"g_records_per_s\": %.2f, \"gb_per_s\": %.1f}\n",
         BYTES, LANES, UNROLL, kFlavorNames[F], occ, arrayBytes / 1e9, numReads / (best * 1e-3) / 1e9,
         (double)numReads * BYTES / (best * 1e-3) / 1e9);
  fflush(stdout);
}

template <int F>
static void sweepFlavor(const uint4 *d, uint64_t bytes, uint64_t n, uint64_t *sink) {
  run<64, 4, 4, F>(d, bytes, n, sink, 0);
  run<32, 2, 4, F>(d, bytes, n, sink, 0);
  run<128, 8, 4, F>(d, bytes, n, sink, 0);
}

int main(int argc, char **argv) {
  const uint64_t bytes = (argc > 1 ? strtoull(argv[1], 0, 10) : 4096ull) << 20;
  const uint64_t n = (argc > 2 ? strtoull(argv[2], 0, 10) : 512ull) << 20;
  const int mode = argc > 3 ? atoi(argv[3]) : 0;  // 1 = single configuration for ncu (64 B, flavour argv[4])
  uint4 *d;
  uint64_t *sink;
  CK(cudaMalloc(&d, bytes));
  CK(cudaMemset(d, 1, bytes));
  CK(cudaMalloc(&sink, 8));
  if (mode == 1) {
    const int f = argc > 4 ? atoi(argv[4]) : 0;
    switch (f) {
      case NC: run<64, 4, 4, NC>(d, bytes, n, sink, 0); break;
      case NC_L2_64: run<64, 4, 4, NC_L2_64>(d, bytes, n, sink, 0); break;
      case CG: run<64, 4, 4, CG>(d, bytes, n, sink, 0); break;
      case CV: run<64, 4, 4, CV>(d, bytes, n, sink, 0); break;
      case NC_NOALLOC: run<64, 4, 4, NC_NOALLOC>(d, bytes, n, sink, 0); break;
      case NC_EVICT_FIRST_HINT: run<64, 4, 4, NC_EVICT_FIRST_HINT>(d, bytes, n, sink, 0); break;
      case 100: run<32, 2, 4, NC>(d, bytes, n, sink, 0); break;
      case 101: run<128, 8, 4, NC>(d, bytes, n, sink, 0); break;
      case 102: run<16, 1, 4, NC>(d, bytes, n, sink, 0); break;
      case 103: run<64, 2, 4, NC>(d, bytes, n, sink, 0); break;
      case 104: run<64, 1, 4, NC>(d, bytes, n, sink, 0); break;
      case 105: run<32, 1, 8, NC>(d, bytes, n, sink, 0); break;
      case 106: run<128, 4, 4, NC>(d, bytes, n, sink, 0); break;
      case 107: run<128, 2, 4, NC>(d, bytes, n, sink, 0); break;
      case 108: run<128, 1, 2, NC>(d, bytes, n, sink, 0); break;
      default: run<64, 4, 4, LU>(d, bytes, n, sink, 0); break;
    }
    return 0;
  }
  for (int g = 0; g < 4; g++) {  // device-wide L2 fetch-granularity limit: default, 32, 64, 128
    size_t actual = 0;
    if (g > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g == 1 ? 32 : g == 2 ? 64 : 128);
    CK(cudaDeviceGetLimit(&actual, cudaLimitMaxL2FetchGranularity));
    printf("{\"max_l2_fetch_granularity\": %zu}\n", actual);
    run<64, 4, 4, NC>(d, bytes, n, sink, 0);
    run<32, 2, 4, NC>(d, bytes, n, sink, 0);
  }
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 64);
  sweepFlavor<NC>(d, bytes, n, sink);
  sweepFlavor<NC_L2_64>(d, bytes, n, sink);
  sweepFlavor<NC_L2_128>(d, bytes, n, sink);
  sweepFlavor<NC_L2_256>(d, bytes, n, sink);
  sweepFlavor<CG>(d, bytes, n, sink);
  sweepFlavor<CV>(d, bytes, n, sink);
  sweepFlavor<NC_NOALLOC>(d, bytes, n, sink);
  sweepFlavor<NC_EVICT_FIRST_HINT>(d, bytes, n, sink);
  sweepFlavor<LU>(d, bytes, n, sink);
  // records in flight per group and lanes per record, default flavour
  run<64, 4, 1, NC>(d, bytes, n, sink, 0);
  run<64, 4, 2, NC>(d, bytes, n, sink, 0);
  run<64, 4, 8, NC>(d, bytes, n, sink, 0);
  run<64, 2, 4, NC>(d, bytes, n, sink, 0);
  run<64, 1, 4, NC>(d, bytes, n, sink, 0);
  run<64, 1, 8, NC>(d, bytes, n, sink, 0);
  run<32, 1, 8, NC>(d, bytes, n, sink, 0);
  run<16, 1, 8, NC>(d, bytes, n, sink, 0);
  run<16, 1, 16, NC>(d, bytes, n, sink, 0);
  run<128, 4, 4, NC>(d, bytes, n, sink, 0);
  run<256, 8, 4, NC>(d, bytes, n, sink, 0);
  return 0;
}
