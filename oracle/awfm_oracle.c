/*
 * awfm_oracle.c — TEST INFRASTRUCTURE (see awfm_oracle.h).  Scalar restatement of the reference's batched
 * exact-match k-mer search, written from the behaviour documented in SURVEY.md App. A/B; each function cites
 * the reference lines it follows.  No SIMD, no prefetch, no chunk-of-8 interleave: those do not change results.
 */
#include "awfm_oracle.h"
#include <string.h>

enum { NUC_BLOCK_BYTES = 160, AMINO_BLOCK_BYTES = 352, NUC_BASE_OFFSET = 96, AMINO_BASE_OFFSET = 160 };
enum { NUC_SENTINEL = 5, AMINO_SENTINEL = 21 };

static int is_amino(const awfm_oracle_index *ix) { return ix->alphabet == 1; }

/* ---- letters: src/AwFmLetter.c:4-22 (nucleotide), :55-67 (amino) ---- */
static const uint8_t kAminoAsciiToIndex[32] = {20, 0,  20, 1,  2,  3,  4,  5,  6,  7,  20, 8,  9,  10, 11, 20,
                                               12, 13, 14, 15, 16, 20, 17, 18, 20, 19, 20, 20, 20, 20, 20, 20};

uint8_t awfm_oracle_letter_index(uint8_t alphabet, uint8_t ascii) {
  if (alphabet == 1) {
    if (ascii == '$') return AMINO_SENTINEL;
    return kAminoAsciiToIndex[ascii & 0x1F];
  }
  switch (ascii | 0x20) {
    case 'a': return 0;
    case 'c': return 1;
    case 'g': return 2;
    case 't':
    case 'u': return 3;
    case '$': return NUC_SENTINEL; /* '$' | 0x20 == '$' */
    default: return 4;
  }
}

/* src/AwFmLetter.c:98-125 — used only for seed-table eligibility (src/AwFmKmerTable.c:4-19).  tolower() there only
 * changes 'A'..'Z' in the C locale. */
int awfm_oracle_letter_is_ambiguous(uint8_t alphabet, uint8_t ascii) {
  uint8_t lower = (ascii >= 'A' && ascii <= 'Z') ? (uint8_t)(ascii + 32) : ascii;
  if (alphabet == 1) return lower == 'z' || lower == 'x' || lower == 'b';
  return !(lower == 'a' || lower == 'c' || lower == 'g' || lower == 't' || lower == 'u');
}

/* ---- rank selectors ----
 * The reference selects a letter with a boolean combination of the bit-vectors that tests only SOME of the code
 * bits (src/AwFmOccurrence.c:18-35 nucleotide, :65-134 amino).  Tabulated here as (code, care-mask): a position
 * matches when (storedCode ^ code) & care == 0.  On a well-formed index this equals exact code equality. */
static const uint8_t kNucCode[5] = {6, 5, 3, 1, 2}; /* src/AwFmLetter.c:44-47 */
static const uint8_t kNucCare[5] = {6, 5, 3, 7, 7};
static const uint8_t kAminoCode[21] = {0x0C, 0x17, 0x03, 0x06, 0x1E, 0x1A, 0x1B, 0x19, 0x15, 0x1C, 0x1D,
                                       0x08, 0x09, 0x04, 0x13, 0x0A, 0x05, 0x16, 0x01, 0x02, 0x1F}; /* :81-87 */
static const uint8_t kAminoCare[21] = {0x1C, 0x0F, 0x13, 0x16, 0x0F, 0x15, 0x0F, 0x16, 0x1A, 0x13, 0x0F,
                                       0x0F, 0x19, 0x0F, 0x1C, 0x1A, 0x15, 0x19, 0x0F, 0x0F, 0x0F};
/* code -> letter index: src/AwFmLetter.c:49-53 and :89-96 */
static const uint8_t kNucCodeToIndex[8] = {5, 3, 4, 2, 5, 1, 0, 0 /* 0b111 unused; the reference reads past its table */};
static const uint8_t kAminoCodeToIndex[32] = {21, 18, 19, 2,  13, 16, 3,  20, 11, 12, 15, 20, 0, 20, 20, 20,
                                              20, 20, 20, 14, 20, 8,  17, 1,  20, 7,  5,  6,  9, 10, 4,  20};

static uint32_t popcount8(uint8_t v) { return (uint32_t)__builtin_popcount(v); }

/* popcount of positions 0..localPosition (INCLUSIVE, src/AwFmSimdConfig.c:89-114) of the block holding `letter` */
uint32_t awfm_oracle_block_popcount(const uint8_t *block, uint8_t alphabet, uint8_t letter, uint8_t localPosition) {
  const int amino = alphabet == 1;
  const int numVectors = amino ? 5 : 3;
  const uint8_t code = amino ? kAminoCode[letter] : kNucCode[letter];
  const uint8_t care = amino ? kAminoCare[letter] : kNucCare[letter];
  uint32_t total = 0;
  const uint32_t lastByte = localPosition / 8u;
  for (uint32_t byte = 0; byte <= lastByte; byte++) {
    uint8_t match = 0xFF;
    for (int v = 0; v < numVectors; v++) {
      if (!((care >> v) & 1)) continue;
      const uint8_t bits = block[32 * v + byte];
      match &= ((code >> v) & 1) ? bits : (uint8_t)~bits;
    }
    if (byte == lastByte) match &= (uint8_t)(0xFFu >> (7u - (localPosition % 8u)));
    total += popcount8(match);
  }
  return total;
}

static const uint8_t *block_ptr(const awfm_oracle_index *ix, uint64_t position) {
  return ix->blocks + (position / 256u) * (is_amino(ix) ? AMINO_BLOCK_BYTES : NUC_BLOCK_BYTES);
}

static uint64_t base_occurrence(const awfm_oracle_index *ix, const uint8_t *block, uint8_t letter) {
  uint64_t v;
  memcpy(&v, block + (is_amino(ix) ? AMINO_BASE_OFFSET : NUC_BASE_OFFSET) + 8u * letter, 8);
  return v;
}

/* Occ(letter, position) = baseOccurrences[letter] + masked popcount (src/AwFmSearch.c:57-64) */
uint64_t awfm_oracle_occ(const awfm_oracle_index *ix, uint8_t letter, uint64_t position) {
  const uint8_t *block = block_ptr(ix, position);
  return base_occurrence(ix, block, letter) +
         awfm_oracle_block_popcount(block, ix->alphabet, letter, (uint8_t)(position % 256u));
}

/* src/AwFmOccurrence.c:170-217 */
uint8_t awfm_oracle_letter_at(const awfm_oracle_index *ix, uint64_t position) {
  const uint8_t *block = block_ptr(ix, position);
  const uint32_t local = (uint32_t)(position % 256u), byte = local / 8u, bit = local % 8u;
  const int numVectors = is_amino(ix) ? 5 : 3;
  uint8_t code = 0;
  for (int v = 0; v < numVectors; v++) code |= (uint8_t)(((block[32 * v + byte] >> bit) & 1u) << v);
  return is_amino(ix) ? kAminoCodeToIndex[code] : kNucCodeToIndex[code];
}

/* One LF-mapping step, src/AwFmSearch.c:42-103 (nucleotide) / :105-159 (amino):
 *   sp' = C[c] + Occ(c, sp-1),  ep' = C[c] + Occ(c, ep) - 1 */
void awfm_oracle_step(const awfm_oracle_index *ix, uint64_t *sp, uint64_t *ep, uint8_t letter) {
  const uint64_t c = ix->prefixSums[letter];
  const uint64_t newSp = c + awfm_oracle_occ(ix, letter, *sp - 1);
  const uint64_t newEp = c + awfm_oracle_occ(ix, letter, *ep) - 1;
  *sp = newSp;
  *ep = newEp;
}

/* src/AwFmSearch.c:369-427: LF(p) = C[c] + Occ(c, p) - 1 with c = BWT[p]; the sentinel maps to position 0 */
uint64_t awfm_oracle_backtrace_step(const awfm_oracle_index *ix, uint64_t position) {
  const uint8_t letter = awfm_oracle_letter_at(ix, position);
  if (letter == (is_amino(ix) ? AMINO_SENTINEL : NUC_SENTINEL)) return 0;
  return ix->prefixSums[letter] + awfm_oracle_occ(ix, letter, position) - 1;
}

/* src/AwFmSuffixArray.c:22-39,114-142: sample j is the w-bit little-endian field at bit j*w of the byte stream */
uint64_t awfm_oracle_sa_value(const awfm_oracle_index *ix, uint64_t sampleIndex) {
  const uint32_t w = ix->saBitWidth;
  uint64_t value = 0;
  /* walk bit by bit over bytes: independent of the reference's 8-byte-load-and-patch strategy, same field */
  const uint64_t firstBit = (sampleIndex / 8u) * w * 8u + (sampleIndex % 8u) * w;
  uint32_t got = 0;
  uint64_t byte = firstBit / 8u;
  uint32_t bitInByte = (uint32_t)(firstBit % 8u);
  while (got < w) {
    const uint32_t take = (8u - bitInByte) < (w - got) ? (8u - bitInByte) : (w - got);
    const uint64_t chunk = ((uint64_t)ix->saBytes[byte] >> bitInByte) & ((1ull << take) - 1ull);
    value |= chunk << got;
    got += take;
    bitInByte = 0;
    byte++;
  }
  return value;
}

static int range_valid(uint64_t sp, uint64_t ep) { return sp <= ep; } /* src/AwFmIndexStruct.c:99-102 */

static void account_step(uint64_t sp, uint64_t ep, awfm_oracle_work *work) {
  if (!work) return;
  work->lfSteps++;
  work->lfBlockReads += ((sp - 1) / 256u == ep / 256u) ? 1u : 2u;
}

/* src/AwFmParallelSearch.c:222-313 for one query:
 *   seedable (len >= k and no ambiguous letter among the last k, src/AwFmKmerTable.c:4-19)
 *       -> table[sum idx(q[len-k+j]) * |A|^(k-1-j)]                      (src/AwFmKmerTable.c:21-51)
 *   else -> non-seeded search over the last min(len,k) letters            (src/AwFmSearch.c:485-520)
 *   then letters len-(k+1), len-(k+2), ... 0 while the range stays valid  (src/AwFmParallelSearch.c:279-311)
 * Both branches stop at the first invalid range and keep it, so after the start they are one loop. */
void awfm_oracle_search(const awfm_oracle_index *ix, const uint8_t *kmer, uint64_t len, uint64_t *spOut,
                        uint64_t *epOut, awfm_oracle_work *work) {
  const uint8_t alphabet = ix->alphabet;
  const uint64_t k = ix->seedK;
  const uint64_t cardinality = is_amino(ix) ? 20 : 4;
  uint64_t sp, ep, next; /* next = number of leading letters still to be consumed */
  if (work) {
    work->queries++;
    work->queryLetters += len;
  }
  if (len == 0) { /* the reference reads kmer[-1] here (undefined); this build defines the result as empty */
    *spOut = 1;
    *epOut = 0;
    return;
  }
  int seedable = len >= k;
  if (seedable) {
    for (uint64_t i = len - k; i < len; i++) {
      if (awfm_oracle_letter_is_ambiguous(alphabet, kmer[i])) seedable = 0;
      /* amino letters that pass the predicate yet map to index 20 (j, o, u, ...) would index past the table in
       * the reference (undefined); this build routes them through the non-seeded path instead. */
      if (awfm_oracle_letter_index(alphabet, kmer[i]) >= cardinality) seedable = 0;
    }
  }
  if (seedable) {
    uint64_t tableIndex = 0;
    for (uint64_t i = len - k; i < len; i++) tableIndex = tableIndex * cardinality + awfm_oracle_letter_index(alphabet, kmer[i]);
    sp = ix->seedTable[2 * tableIndex];
    ep = ix->seedTable[2 * tableIndex + 1];
    next = len - k;
    if (work) work->seeded++;
  } else {
    const uint8_t last = awfm_oracle_letter_index(alphabet, kmer[len - 1]);
    if (last > cardinality) { /* '$' in a query: prefixSums[last+1] is out of bounds in the reference (undefined) */
      *spOut = 1;
      *epOut = 0;
      return;
    }
    sp = ix->prefixSums[last];
    ep = ix->prefixSums[last + 1] - 1;
    next = len - 1;
  }
  while (next > 0 && range_valid(sp, ep)) {
    const uint8_t letter = awfm_oracle_letter_index(alphabet, kmer[next - 1]);
    if (letter > cardinality) { /* '$' (undefined in the reference): defined here as "no match" */
      sp = 1;
      ep = 0;
      break;
    }
    account_step(sp, ep, work);
    awfm_oracle_step(ix, &sp, &ep, letter);
    next--;
  }
  *spOut = sp;
  *epOut = ep;
}

static const uint8_t *query_ptr(const uint8_t *letters, const uint64_t *offsets, uint32_t fixedLen, uint64_t i,
                                uint64_t *len) {
  if (offsets) {
    *len = offsets[i + 1] - offsets[i];
    return letters + offsets[i];
  }
  *len = fixedLen;
  return letters + i * (uint64_t)fixedLen;
}

static void merge_work(awfm_oracle_work *into, const awfm_oracle_work *from) {
  into->queries += from->queries;
  into->seeded += from->seeded;
  into->lfSteps += from->lfSteps;
  into->lfBlockReads += from->lfBlockReads;
  into->queryLetters += from->queryLetters;
  into->hits += from->hits;
  into->backtraceSteps += from->backtraceSteps;
}

static void finish_work(const awfm_oracle_index *ix, awfm_oracle_work *w) {
  /* per rank call 3*32+8 = 104 B (nucleotide) / 5*32+8 = 168 B (amino) */
  const uint64_t occBytes = is_amino(ix) ? 168 : 104;
  w->countBytes = w->queryLetters + 16 * w->seeded + occBytes * w->lfBlockReads + 4 * w->queries;
  w->locateBytes = occBytes * w->backtraceSteps + (((uint64_t)ix->saBitWidth + 7) / 8 + 8) * w->hits;
}

/* src/AwFmParallelSearch.c:159-220; count = (uint32) range length (:187-190) */
void awfm_oracle_count(const awfm_oracle_index *ix, const uint8_t *letters, const uint64_t *offsets,
                       uint32_t fixedLen, uint64_t n, uint32_t *counts, uint64_t *ranges, awfm_oracle_work *work,
                       int numThreads) {
  awfm_oracle_work total;
  memset(&total, 0, sizeof total);
  (void)numThreads;
#pragma omp parallel num_threads(numThreads > 0 ? numThreads : 1)
  {
    awfm_oracle_work local;
    memset(&local, 0, sizeof local);
#pragma omp for schedule(static)
    for (uint64_t i = 0; i < n; i++) {
      uint64_t len, sp, ep;
      const uint8_t *q = query_ptr(letters, offsets, fixedLen, i, &len);
      awfm_oracle_search(ix, q, len, &sp, &ep, &local);
      counts[i] = (uint32_t)(range_valid(sp, ep) ? ep - sp + 1 : 0);
      if (ranges) {
        ranges[2 * i] = sp;
        ranges[2 * i + 1] = ep;
      }
    }
#pragma omp critical
    merge_work(&total, &local);
  }
  if (work) {
    finish_work(ix, &total);
    *work = total;
  }
}

/* src/AwFmParallelSearch.c:315-365 + src/AwFmSuffixArray.c:179-203, CSR output in SA order */
uint64_t awfm_oracle_locate(const awfm_oracle_index *ix, const uint8_t *letters, const uint64_t *offsets,
                            uint32_t fixedLen, uint64_t n, uint64_t *hitOffsets, uint64_t *positions,
                            awfm_oracle_work *work, int numThreads) {
  awfm_oracle_work total;
  memset(&total, 0, sizeof total);
  uint64_t *starts = positions ? (uint64_t *)__builtin_malloc(n * sizeof(uint64_t) + 8) : 0;
  uint64_t running = 0;
  for (uint64_t i = 0; i < n; i++) { /* pass 1: ranges -> offsets (serial; tests are small) */
    uint64_t len, sp, ep;
    const uint8_t *q = query_ptr(letters, offsets, fixedLen, i, &len);
    awfm_oracle_search(ix, q, len, &sp, &ep, &total);
    hitOffsets[i] = running;
    if (starts) starts[i] = sp;
    /* the reference truncates the range length to uint32 in setPositionListCount (:328, :367-368) */
    running += (uint32_t)(range_valid(sp, ep) ? ep - sp + 1 : 0);
  }
  hitOffsets[n] = running;
  total.hits = running;
  if (positions) {
    (void)numThreads;
#pragma omp parallel num_threads(numThreads > 0 ? numThreads : 1)
    {
      uint64_t steps = 0;
#pragma omp for schedule(dynamic, 64)
      for (uint64_t i = 0; i < n; i++) {
        for (uint64_t h = hitOffsets[i]; h < hitOffsets[i + 1]; h++) {
          uint64_t p = starts[i] + (h - hitOffsets[i]), offset = 0;
          while (p % ix->saRatio != 0) { /* src/AwFmIndexStruct.c:88-91 */
            p = awfm_oracle_backtrace_step(ix, p);
            offset++;
          }
          steps += offset;
          positions[h] = (awfm_oracle_sa_value(ix, p / ix->saRatio) + offset) % ix->bwtLength;
        }
      }
#pragma omp atomic
      total.backtraceSteps += steps;
    }
    __builtin_free(starts);
  }
  if (work) {
    finish_work(ix, &total);
    *work = total;
  }
  return running;
}

/* lib/FastaVector/src/FastaVector.c:338-381: first s with g < E[s]; local = g - E[s-1] */
int awfm_oracle_contig_of(const uint64_t *ends, uint64_t numSequences, uint64_t g, uint64_t *sequenceIndex,
                          uint64_t *localPosition) {
  if (numSequences == 0 || g > ends[numSequences - 1]) return -1;
  uint64_t lo = 0, hi = numSequences; /* invariant: all s < lo have E[s] <= g */
  while (lo < hi) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if (g < ends[mid]) hi = mid;
    else lo = mid + 1;
  }
  *sequenceIndex = lo;
  *localPosition = lo == 0 ? g : g - ends[lo - 1];
  return 0;
}
