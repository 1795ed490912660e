/* tests/support/corpusgen.c — deterministic synthetic corpora.
 *
 * Shared by the tests and bench.py; the SAME bytes feed the CPU oracle and the
 * GPU. Every document has its own generator state derived from (seed, global
 * document index) with splitmix64, then xoshiro256**, so a corpus is identical
 * no matter how many threads or shards generate it.
 *
 *   kind 0 "cjk":   doc = L code points, L uniform in [min_len, max_len], each an
 *                   ideograph U+4E00 + r, r ~ Zipf(s) over `alphabet` symbols (3 bytes each)
 *   kind 1 "ascii": doc = W words, W uniform in [min_len, max_len], joined by one
 *                   space; word = vocabulary[r], r ~ Zipf(s) over `alphabet` words;
 *                   vocabulary word i = 3..9 lowercase letters derived from (seed, i)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t* x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t next_u64(rng_t* r) {
  uint64_t* s = r->s;
  const uint64_t result = rotl(s[1] * 5, 7) * 9;
  const uint64_t t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
  return result;
}
static inline void seed_rng(rng_t* r, uint64_t seed, uint64_t stream) {
  uint64_t x = seed ^ (stream * 0xD6E8FEB86659FD93ULL + 0x2545F4914F6CDD1DULL);
  for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&x);
}
static inline double next_unit(rng_t* r) { return (double)(next_u64(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t next_range(rng_t* r, uint32_t lo, uint32_t hi) { /* inclusive */
  return lo + (uint32_t)(next_u64(r) % (uint64_t)(hi - lo + 1));
}

typedef struct {
  int kind; uint64_t seed; uint32_t alphabet; double zipf_s; uint32_t min_len; uint32_t max_len;
  double* cdf;        /* [alphabet] */
  char* vocab;        /* ascii: alphabet x 10 bytes (len byte + up to 9 letters) */
} gen_t;

static uint32_t zipf_draw(const gen_t* g, rng_t* r) {
  const double u = next_unit(r) * g->cdf[g->alphabet - 1];
  uint32_t lo = 0, hi = g->alphabet - 1;
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (g->cdf[mid] <= u) lo = mid + 1; else hi = mid; }
  return lo;
}

void* corpus_gen_create(int kind, uint64_t seed, uint32_t alphabet, double zipf_s, uint32_t min_len, uint32_t max_len) {
  gen_t* g = (gen_t*)calloc(1, sizeof(gen_t));
  g->kind = kind; g->seed = seed; g->alphabet = alphabet; g->zipf_s = zipf_s; g->min_len = min_len; g->max_len = max_len;
  g->cdf = (double*)malloc(sizeof(double) * alphabet);
  double acc = 0.0;
  for (uint32_t i = 0; i < alphabet; ++i) { acc += 1.0 / pow((double)(i + 1), zipf_s); g->cdf[i] = acc; }
  if (kind == 1) {
    g->vocab = (char*)malloc((size_t)alphabet * 10);
    for (uint32_t i = 0; i < alphabet; ++i) {
      rng_t r; seed_rng(&r, seed ^ 0xA5C11ULL, 0x100000000ULL + i);
      const uint32_t len = next_range(&r, 3, 9);
      g->vocab[(size_t)i * 10] = (char)len;
      for (uint32_t j = 0; j < len; ++j) g->vocab[(size_t)i * 10 + 1 + j] = (char)('a' + next_u64(&r) % 26);
    }
  }
  return g;
}
void corpus_gen_destroy(void* h) { gen_t* g = (gen_t*)h; if (!g) return; free(g->cdf); free(g->vocab); free(g); }

/* Writes one document (or only measures it when out == NULL). Returns its byte length. */
static uint64_t gen_doc(const gen_t* g, uint64_t doc, uint8_t* out) {
  rng_t r; seed_rng(&r, g->seed, doc);
  const uint32_t n = next_range(&r, g->min_len, g->max_len);
  uint64_t pos = 0;
  if (g->kind == 0) {
    if (out == NULL) return (uint64_t)n * 3;
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t cp = 0x4E00u + zipf_draw(g, &r);
      out[pos++] = (uint8_t)(0xE0 | (cp >> 12)); out[pos++] = (uint8_t)(0x80 | ((cp >> 6) & 0x3F)); out[pos++] = (uint8_t)(0x80 | (cp & 0x3F));
    }
    return pos;
  }
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t w = zipf_draw(g, &r);
    const char* word = g->vocab + (size_t)w * 10;
    const uint32_t len = (uint32_t)word[0];
    if (i > 0) { if (out) out[pos] = ' '; ++pos; }
    if (out) memcpy(out + pos, word + 1, len);
    pos += len;
  }
  return pos;
}

typedef struct { const gen_t* g; uint64_t first; uint64_t begin; uint64_t end; uint64_t* offsets; uint8_t* text; int fill; } job_t;
static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (uint64_t d = j->begin; d < j->end; ++d) {
    if (j->fill) gen_doc(j->g, j->first + d, j->text + j->offsets[d]);
    else j->offsets[d + 1] = gen_doc(j->g, j->first + d, NULL);
  }
  return NULL;
}
static void run(const gen_t* g, uint64_t first, uint64_t n, uint64_t* offsets, uint8_t* text, int fill, int threads) {
  if (threads < 1) threads = 1;
  if ((uint64_t)threads > n) threads = n ? (int)n : 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * threads);
  for (int t = 0; t < threads; ++t) {
    jobs[t] = (job_t){g, first, n * t / threads, n * (t + 1) / threads, offsets, text, fill};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

/* offsets[n_docs+1]; returns total bytes. Documents are global indices first_doc .. first_doc+n_docs-1. */
uint64_t corpus_gen_sizes(void* h, uint64_t first_doc, uint64_t n_docs, uint64_t* offsets, int threads) {
  const gen_t* g = (const gen_t*)h;
  offsets[0] = 0;
  run(g, first_doc, n_docs, offsets, NULL, 0, threads);
  for (uint64_t d = 0; d < n_docs; ++d) offsets[d + 1] += offsets[d];
  return offsets[n_docs];
}
void corpus_gen_fill(void* h, uint64_t first_doc, uint64_t n_docs, const uint64_t* offsets, uint8_t* text, int threads) {
  run((const gen_t*)h, first_doc, n_docs, (uint64_t*)offsets, text, 1, threads);
}

/* Arbitrary documents by global index (used to cut query terms out of random documents without
 * materialising the corpus): offsets[n+1] relative to 0; call with text == NULL to size. */
uint64_t corpus_gen_docs(void* h, const uint64_t* doc_indices, uint64_t n, uint64_t* offsets, uint8_t* text) {
  const gen_t* g = (const gen_t*)h;
  uint64_t pos = 0;
  for (uint64_t i = 0; i < n; ++i) {
    offsets[i] = pos;
    pos += gen_doc(g, doc_indices[i], text ? text + pos : NULL);
  }
  offsets[n] = pos;
  return pos;
}
