/* Host-side floor of the reference's list layout (no GPU): how fast can T threads walk N 32-B AwFmKmerSearchData
 * entries (src/AwFmIndex.h:111-117) once to read {kmerString, kmerLength} and once to write `count`?  Variants:
 *   whole   two passes over the whole list (entries leave the caches in between)
 *   chunked pass 1 of chunk i+LAG and pass 2 of chunk i interleaved, as the drop-in engine does
 *   fused   both in one touch (what an engine with zero GPU latency could do) - the lower bound
 * gcc -O2 -fopenmp tools/host_list_floor.c -o /tmp/host_list_floor && /tmp/host_list_floor 100000000 16 */
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct Entry {
  char *kmerString;
  uint64_t kmerLength;
  uint64_t *positionList;
  uint32_t count, capacity;
};

static double now(void) { return omp_get_wtime(); }

int main(int argc, char **argv) {
  const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
  const int T = argc > 2 ? atoi(argv[2]) : omp_get_max_threads();
  const uint64_t len = 20;
  struct Entry *e = aligned_alloc(64, n * sizeof *e);
  uint32_t *counts = aligned_alloc(64, n * 4);
  char *letters = (char *)0x100000;
#pragma omp parallel for num_threads(T) schedule(static)
  for (uint64_t i = 0; i < n; i++) {
    e[i].kmerString = letters + i * len, e[i].kmerLength = len, e[i].positionList = 0, e[i].count = 0, e[i].capacity = 4;
    counts[i] = (uint32_t)i;
  }
  for (int rep = 0; rep < 3; rep++) {
    int ok = 1;
    double t0 = now();
#pragma omp parallel num_threads(T) reduction(&& : ok)
    {
      const int t = omp_get_thread_num();
      const uint64_t a = n * t / T, b = n * (t + 1) / T;
      int u = 1;
      for (uint64_t i = a; i < b; i++) u &= (e[i].kmerLength == len) & (e[i].kmerString == letters + i * len);
      ok = u;
#pragma omp barrier
      for (uint64_t i = a; i < b; i++) e[i].count = counts[i];
    }
    double t1 = now();
    printf("whole    : %.1f ms (ok=%d)\n", 1e3 * (t1 - t0), ok);
    for (uint64_t chunk = 1 << 14; chunk <= 1 << 20; chunk <<= 2) {
      const uint64_t nc = (n + chunk - 1) / chunk, LAG = 3;
      t0 = now();
#pragma omp parallel num_threads(T) reduction(&& : ok)
      {
        const int t = omp_get_thread_num();
        int u = 1;
        for (uint64_t c = 0; c < nc + LAG; c++) {
          if (c >= LAG) {
            const uint64_t f = (c - LAG) * chunk, m = (f + chunk <= n ? chunk : n - f);
            for (uint64_t i = f + m * t / T; i < f + m * (t + 1) / T; i++) e[i].count = counts[i];
          }
          if (c < nc) {
            const uint64_t f = c * chunk, m = (f + chunk <= n ? chunk : n - f);
            for (uint64_t i = f + m * t / T; i < f + m * (t + 1) / T; i++)
              u &= (e[i].kmerLength == len) & (e[i].kmerString == letters + i * len);
          }
#pragma omp barrier
        }
        ok = u;
      }
      t1 = now();
      printf("chunked %7llu: %.1f ms (ok=%d)\n", (unsigned long long)chunk, 1e3 * (t1 - t0), ok);
    }
    t0 = now();
#pragma omp parallel for num_threads(T) schedule(static) reduction(&& : ok)
    for (uint64_t i = 0; i < n; i++) {
      ok = ok && (e[i].kmerLength == len) & (e[i].kmerString == letters + i * len);
      e[i].count = counts[i];
    }
    t1 = now();
    printf("fused    : %.1f ms (ok=%d)\n", 1e3 * (t1 - t0), ok);
  }
  return 0;
}
