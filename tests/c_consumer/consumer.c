/* A plain C user of the reference's public API (src/AwFmIndex.h), written the way README.md's example is: build an
 * index with awFmCreateIndex, fill an AwFmKmerSearchList, call awFmParallelSearchCount and awFmParallelSearchLocate,
 * print what came back.  It contains nothing of ours.  tests/test_c_consumer.py builds it three ways:
 *   consumer_ref      linked against the reference only                      -> the expected output
 *   consumer_ref      run with LD_PRELOAD=libawfm_b200.so                    -> INTEGRATION.md section 2
 *   consumer_linked   linked with libawfm_b200.so BEFORE the reference       -> INTEGRATION.md section 1
 * and requires byte-identical stdout.  TEST INFRASTRUCTURE (compiled against /root/reference/src/AwFmIndex.h where that
 * tree exists; the binaries travel to the GPU box under oracle/_ref/). */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "AwFmIndex.h"

static uint64_t state = 0x2545F4914F6CDD1Dull;
static uint64_t next(void) {
  state ^= state << 13;
  state ^= state >> 7;
  state ^= state << 17;
  return state;
}

int main(int argc, char **argv) {
  const int amino = argc > 1 && strcmp(argv[1], "amino") == 0;
  const char *path = argc > 2 ? argv[2] : "/tmp/awfm_c_consumer.awfmi";
  const size_t textLength = amino ? 60000 : 200000, numKmers = 20000;
  const char *letters = amino ? "ACDEFGHIKLMNPQRSTVWY" : "ACGT";
  const size_t card = strlen(letters);
  uint8_t *text = malloc(textLength);
  for (size_t i = 0; i < textLength; i++) text[i] = (uint8_t)letters[next() % card];
  for (size_t i = 0; i < textLength; i += 997) text[i] = amino ? 'X' : 'N'; /* ambiguity letters in the text */

  struct AwFmIndexConfiguration config = {.suffixArrayCompressionRatio = 5, .kmerLengthInSeedTable = amino ? 3 : 7,
                                          .alphabetType = amino ? AwFmAlphabetAmino : AwFmAlphabetDna,
                                          .keepSuffixArrayInMemory = true, .storeOriginalSequence = false};
  struct AwFmIndex *index = NULL;
  remove(path);
  enum AwFmReturnCode rc = awFmCreateIndex(&index, &config, text, textLength, path);
  if (rc < 0) {
    fprintf(stderr, "awFmCreateIndex failed: %d\n", rc);
    return 2;
  }

  /* queries: substrings of the text (hits), random strings (mostly misses), some shorter than the seed length,
   * some with an ambiguity letter, lower case */
  struct AwFmKmerSearchList *list = awFmCreateKmerSearchList(numKmers);
  char *pool = malloc(numKmers * 40);
  for (size_t i = 0; i < numKmers; i++) {
    const size_t len = 1 + next() % 30;
    char *k = pool + i * 40;
    if (next() % 3) {
      const size_t start = next() % (textLength - len);
      memcpy(k, text + start, len);
    } else {
      for (size_t j = 0; j < len; j++) k[j] = letters[next() % card];
    }
    if (next() % 16 == 0) k[next() % len] = amino ? 'x' : 'n';
    if (!amino && next() % 8 == 0)
      for (size_t j = 0; j < len; j++) k[j] |= 0x20;
    list->kmerSearchData[i].kmerString = k;
    list->kmerSearchData[i].kmerLength = len;
  }
  list->count = numKmers;

  awFmParallelSearchCount(index, list, 4);
  uint64_t countSum = 0, countHash = 1469598103934665603ull;
  for (size_t i = 0; i < numKmers; i++) {
    countSum += list->kmerSearchData[i].count;
    countHash = (countHash ^ list->kmerSearchData[i].count) * 1099511628211ull;
  }
  printf("count: kmers=%zu total=%llu hash=%016llx\n", numKmers, (unsigned long long)countSum,
         (unsigned long long)countHash);

  rc = awFmParallelSearchLocate(index, list, 4);
  uint64_t hits = 0, posHash = 1469598103934665603ull;
  for (size_t i = 0; i < numKmers; i++) {
    const struct AwFmKmerSearchData *d = &list->kmerSearchData[i];
    hits += d->count;
    if (d->capacity < d->count) {
      printf("capacity violated at %zu\n", i);
      return 3;
    }
    for (uint32_t j = 0; j < d->count; j++) posHash = (posHash ^ d->positionList[j]) * 1099511628211ull;
  }
  printf("locate: rc=%d hits=%llu hash=%016llx\n", rc, (unsigned long long)hits, (unsigned long long)posHash);
  for (size_t i = 0; i < 5; i++) {
    const struct AwFmKmerSearchData *d = &list->kmerSearchData[i];
    printf("kmer %zu: len=%llu count=%u capacity=%u first=%llu\n", i, (unsigned long long)d->kmerLength, d->count,
           d->capacity, d->count ? (unsigned long long)d->positionList[0] : 0ull);
  }
  /* which implementation answered: the additive awFmGpu* symbols exist only in the drop-in (stderr, not compared) */
  int (*numDevices)(const struct AwFmIndex *) = (int (*)(const struct AwFmIndex *))dlsym(RTLD_DEFAULT, "awFmGpuNumDevices");
  void (*release)(const struct AwFmIndex *) = (void (*)(const struct AwFmIndex *))dlsym(RTLD_DEFAULT, "awFmGpuReleaseIndex");
  fprintf(stderr, "engine: %s", numDevices ? "b200 drop-in" : "reference");
  if (numDevices) fprintf(stderr, " devices=%d", numDevices(index));
  fprintf(stderr, "\n");
  if (release) release(index);
  awFmDeallocKmerSearchList(list);
  awFmDeallocIndex(index);
  free(pool);
  free(text);
  return 0;
}
