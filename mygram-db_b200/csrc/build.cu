// build.cu — index construction on the device.
//
// Replaces Index::AddDocumentBatch (reference src/index/index.cpp:76-119) for a
// whole shard: GenerateHybridNgrams (utils/string_utils.cpp:452-509) per
// document, per-document sort+unique (index.cpp:88-91), term -> ascending doc
// ids (index.cpp:104-118, posting_list.cpp:291-324).
//
//   K1 tokenize_flat<COUNT>  4 KB tiles of the text arena, one 16-byte chunk per thread, document boundaries from
//                            a shared-memory table: code points per document (= CountCodePoints,
//                            string_utils.cpp:655-669), valid bytes per document, n-grams per tile
//      scan                  tile -> first n-gram slot
//   K2 tokenize_flat<EMIT>   same decode; writes (packed key << 22 | byte offset in document, doc) in text order
//   K3 radix sort            stable by key => docs ascending inside a key, occurrences in text order (primitives.cu)
//   K4 csr                   segmented unique (drops duplicate (key, doc) = the per-document unique) + compaction
//                            into CSR; first / second occurrence positions of every posting
//   K5 bitmaps               doc bitmaps for lists with density >= dense_threshold
//   K6 journal merge         resident corpus + mutation journal -> merged corpus (then K1..K5 again)
#include <algorithm>
#include <chrono>
#include <cstdlib>

#include "mgx_internal.cuh"

namespace mgx {

namespace {

__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[5], int j) {
  return (w[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
}

__device__ __forceinline__ uint64_t umin_u64(uint64_t a, uint64_t b) { return a < b ? a : b; }

__device__ __forceinline__ uint4 ld_stream_16(const uint8_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ------------------------------------------------------------------ flat tokenizer
// The arena is one byte stream (documents back to back), so tokenisation does not have to walk it document by
// document: a CTA takes a 4 KB tile, every thread one 16-byte chunk, and the document of each byte comes from the
// tile's document offsets staged in shared memory (documents average ~200 bytes, so a warp per document leaves
// most lanes idle). Strict UTF-8 decoding is a LOCAL predicate per byte (see parse_utf8: every non-continuation
// byte is reached by the reference's skip-one-byte scan, string_utils.cpp:206-216, and a character may not run
// past the end of ITS document). Per tile:
//   A  decode: code point + document of every character start of the chunk;
//   B  block-wide compaction of the characters (code point, byte position, document) into shared memory, plus the
//      first (width-1) characters after the tile that belong to the tile's last document (windows may need them);
//   C  one thread per character: GenerateHybridNgrams' window rule (string_utils.cpp:452-509) -> emit flag;
//      COUNT sums the flags per tile, EMIT writes (packed key << pos_bits | byte offset in the document, document)
//      at the tile's scanned base + the flag's rank, i.e. in text order.
// COUNT also accumulates, per document, its code points (= CountCodePoints, BM25's dl) and the bytes covered by
// valid characters (a document is valid UTF-8 iff they equal its length). A tile whose documents do not fit the
// staged table (more than kFlatDocCap starts in 4 KB) is processed in several rounds over ranges of documents.
constexpr int kFlatThreads = 256;
constexpr uint32_t kFlatTile = kFlatThreads * 16;   // 4096 bytes
constexpr uint32_t kFlatDocCap = 1024;
constexpr int kFlatPerThread = 4;                    // characters a thread takes per strip of the window phase
constexpr uint32_t kFlatHalo = kMaxNgramSize - 1;   // characters after the tile a window can reach

struct FlatSmem {
  uint32_t cp[kFlatTile + kFlatHalo + 1];
  uint16_t pos[kFlatTile + kFlatHalo + 1];   // byte position relative to the tile start (halo: >= tile length)
  uint16_t doc[kFlatTile + kFlatHalo + 1];   // index into docrel
  int32_t docrel[kFlatDocCap + 2];           // start of document (round_first + i) relative to the tile start
  uint32_t doc_cps[kFlatDocCap + 1];
  uint32_t doc_valid[kFlatDocCap + 1];
  uint32_t warp_cnt[kFlatThreads / 32];
  uint32_t n_halo;
  uint32_t prev_cp;  // decoded character right before the tile in its first document (signatures), kNoCp if none
};
constexpr uint32_t kNoCp = 0xFFFFFFFFu;

__global__ void flat_tile_first_doc_kernel(const uint64_t* __restrict__ text_off, uint64_t n_docs, uint64_t n_tiles,
                                           uint32_t* __restrict__ out) {
  const uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_tiles) {
    return;
  }
  const uint64_t byte = t * kFlatTile;
  uint64_t lo = 0;  // invariant: text_off[lo] <= byte
  uint64_t hi = n_docs;
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (text_off[mid] <= byte) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  out[t] = static_cast<uint32_t>(lo);
}

__device__ __forceinline__ uint32_t flat_block_scan(uint32_t v, uint32_t* warp_cnt, uint32_t* total) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, s);
    if (lane >= static_cast<unsigned>(s)) {
      inc += o;
    }
  }
  __syncthreads();  // warp_cnt may still be read from the previous use
  if (lane == 31) {
    warp_cnt[warp] = inc;
  }
  __syncthreads();
  uint32_t prefix = 0;
  uint32_t tot = 0;
#pragma unroll
  for (int w = 0; w < kFlatThreads / 32; ++w) {
    const uint32_t c = warp_cnt[w];
    if (static_cast<unsigned>(w) < warp) {
      prefix += c;
    }
    tot += c;
  }
  *total = tot;
  return prefix + inc - v;
}

// MODE 0 = COUNT, 1 = EMIT (after COUNT: exact slots), 2 = FUSED: one pass that both gathers the per-document
// totals and writes the n-grams, into slots whose per-tile bases come from an UPPER bound (the tile's character-start
// bytes, lead_count_kernel); the unused slots at the end of every tile get the placeholder key 0, which no n-gram
// has, sorts first and is skipped by the CSR kernels.
constexpr int kTokCount = 0;
constexpr int kTokEmit = 1;
constexpr int kTokFused = 2;
template <int MODE>
__global__ void __launch_bounds__(kFlatThreads, 4)
tokenize_flat_kernel(const uint8_t* __restrict__ text, const uint64_t* __restrict__ text_off, uint64_t n_docs,
                     uint64_t text_bytes, uint64_t n_tiles, const uint32_t* __restrict__ tile_first_doc, int ngram,
                     int kanji, int cross, int width, int pos_bits, uint32_t* __restrict__ doc_len,
                     uint32_t* __restrict__ doc_valid_bytes, uint32_t* __restrict__ tile_cnt,
                     const uint64_t* __restrict__ tile_off, uint64_t* __restrict__ keys_out,
                     uint32_t* __restrict__ docs_out, int wide_words, uint64_t wide_stride, int sig_cfg) {
  // wide_words > 0 (width > 3): word w of the n-gram in slot s goes to keys_out[w * wide_stride + s]
  // sig_cfg != 0 (index build, pos_bits > 0): the pos_bits payload bits are split into offset (sig_cfg & 0xFF bits),
  // next signature ((sig_cfg >> 8) & 0xFF bits) and prev signature ((sig_cfg >> 16) & 0xFF bits), SigLayout
  const int sig_pos = sig_cfg & 0xFF;
  const int sig_next = (sig_cfg >> 8) & 0xFF;
  const int sig_prev = (sig_cfg >> 16) & 0xFF;
  const uint32_t halo_want = static_cast<uint32_t>(width - 1) + (sig_cfg != 0 ? 1u : 0u);
  constexpr bool EMIT = MODE != kTokCount;    // writes n-grams
  constexpr bool STATS = MODE != kTokEmit;    // gathers the per-document totals
  __shared__ FlatSmem sm;
  const unsigned lane = threadIdx.x & 31u;
  const uint64_t pos_max = pos_bits > 0 ? ((1ULL << (sig_cfg != 0 ? sig_pos : pos_bits)) - 1) : 0;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t tile_b = tile * kFlatTile;
    const uint32_t tile_len = static_cast<uint32_t>(umin_u64(kFlatTile, text_bytes - tile_b));
    const uint32_t first_doc = tile_first_doc[tile];
    const uint32_t last_doc = tile + 1 < n_tiles ? tile_first_doc[tile + 1] : static_cast<uint32_t>(n_docs - 1);
    uint64_t emitted_tile = 0;  // n-grams of the rounds done so far (same value in every thread)
    // rounds over ranges of at most kFlatDocCap documents (one round unless the tile holds very many tiny documents)
    for (uint32_t r_first = first_doc; r_first <= last_doc; r_first += kFlatDocCap) {
      const uint32_t r_docs = min(kFlatDocCap, last_doc - r_first + 1u);  // documents of this round
      __syncthreads();  // the previous round / tile is done with shared memory
      for (uint32_t i = threadIdx.x; i <= r_docs; i += kFlatThreads) {
        const uint64_t off = text_off[static_cast<uint64_t>(r_first) + i];  // i == r_docs: end of the last document
        int64_t rel = static_cast<int64_t>(off) - static_cast<int64_t>(tile_b);
        rel = rel < -0x40000000LL ? -0x40000000LL : (rel > 0x40000000LL ? 0x40000000LL : rel);
        sm.docrel[i] = static_cast<int32_t>(rel);
        if (i < r_docs) {
          sm.doc_cps[i] = 0;
          sm.doc_valid[i] = 0;
        }
      }
      if (threadIdx.x == 0) {
        sm.n_halo = 0;
        sm.prev_cp = kNoCp;
      }
      __syncthreads();
      // byte range of the tile covered by this round's documents
      const int32_t r_begin = max(sm.docrel[0], 0);
      const int32_t r_end = min(sm.docrel[r_docs], static_cast<int32_t>(tile_len));

      // ---- A: decode the chunk
      const int32_t c0 = static_cast<int32_t>(threadIdx.x) * 16;
      uint32_t flags = 0;
      uint32_t cps[16];   // code point (21 bits) | (document - dj0) << 21: a 16-byte chunk spans at most 17 documents
      uint32_t dj0 = 0;
      if (c0 < r_end && c0 + 16 > r_begin) {
        const uint4 v = ld_stream_16(text + tile_b + c0);
        const uint32_t w4 = *reinterpret_cast<const uint32_t*>(text + tile_b + c0 + 16);  // the arena is padded
        const uint32_t w[5] = {v.x, v.y, v.z, v.w, w4};
        // document of the first byte this thread looks at: largest j with docrel[j] <= p
        const int32_t p0 = max(c0, r_begin);
        uint32_t lo = 0;
        uint32_t hi = r_docs;  // docrel[r_docs] = r_end' > p0
        while (hi - lo > 1) {
          const uint32_t mid = (lo + hi) >> 1;
          if (sm.docrel[mid] <= p0) {
            lo = mid;
          } else {
            hi = mid;
          }
        }
        uint32_t dj = lo;
        dj0 = lo;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int32_t p = c0 + j;
          cps[j] = 0;
          const uint32_t b0 = byte_of(w, j);
          // a continuation byte (10xxxxxx) starts no character, whatever document it lies in: two of three bytes of
          // CJK text leave here, before the document lookup and the decoder
          if ((b0 & 0xC0u) != 0x80u && p >= r_begin && p < r_end) {
            while (p >= sm.docrel[dj + 1]) {  // next document (empty documents are stepped over)
              ++dj;
            }
            uint32_t cp = 0;
            const int len = parse_utf8(b0, byte_of(w, j + 1), byte_of(w, j + 2), byte_of(w, j + 3),
                                       static_cast<uint64_t>(sm.docrel[dj + 1] - p), &cp);
            if (len > 0) {
              flags |= 1u << j;
              cps[j] = cp | ((dj - dj0) << 21);
              if (STATS) {
                atomicAdd(&sm.doc_cps[dj], 1u);
                atomicAdd(&sm.doc_valid[dj], static_cast<uint32_t>(len));
              }
            }
          }
        }
      }
      // ---- B: compaction in text order
      uint32_t n_chars = 0;
      uint32_t at = flat_block_scan(static_cast<uint32_t>(__popc(flags)), sm.warp_cnt, &n_chars);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (flags & (1u << j)) {
          sm.cp[at] = cps[j] & 0x1FFFFFu;
          sm.pos[at] = static_cast<uint16_t>(c0 + j);
          sm.doc[at] = static_cast<uint16_t>(dj0 + (cps[j] >> 21));
          ++at;
        }
      }
      // the character right before the tile, when the tile starts inside a document: the nearest non-continuation
      // byte behind the tile start, if it decodes (what precedes the tile's first character when the text is
      // contiguous there -- all a signature has to be right about)
      if (threadIdx.x == 32 && EMIT && sig_prev > 0 && r_first == first_doc && sm.docrel[0] < 0) {
        const int32_t doc_b = sm.docrel[0];
        for (int32_t p = -1; p >= doc_b && p >= -4; --p) {
          const uint8_t* q = text + tile_b + p;  // tile_b + p >= start of the document >= 0
          if ((q[0] & 0xC0u) != 0x80u) {
            uint32_t cp = 0;
            if (parse_utf8(q[0], q[1], q[2], q[3], static_cast<uint64_t>(sm.docrel[1] - p), &cp) > 0) {
              sm.prev_cp = cp;
            }
            break;
          }
        }
      }
      // halo: the first (width - 1) characters after the tile that still belong to the round's last document (one
      // more when signatures are written: the character after the last window)
      if (threadIdx.x < 32 && halo_want > 0 && sm.docrel[r_docs] > static_cast<int32_t>(tile_len) &&
          r_end == static_cast<int32_t>(tile_len)) {
        const int32_t doc_end = sm.docrel[r_docs];
        uint32_t found = 0;
        for (int32_t base = static_cast<int32_t>(tile_len); base < doc_end && found < halo_want;
             base += 32) {
          const int32_t p = base + static_cast<int32_t>(lane);
          uint32_t cp = 0;
          int len = 0;
          if (p < doc_end) {
            const uint8_t* q = text + tile_b + p;  // within the arena + padding
            len = parse_utf8(q[0], q[1], q[2], q[3], static_cast<uint64_t>(doc_end - p), &cp);
          }
          const unsigned m = __ballot_sync(0xffffffffu, len > 0);
          const uint32_t rank = __popc(m & ((1u << lane) - 1u));
          if (len > 0 && found + rank < halo_want) {
            sm.cp[n_chars + found + rank] = cp;
            sm.pos[n_chars + found + rank] = static_cast<uint16_t>(min(p, 0xFFFF));
            sm.doc[n_chars + found + rank] = static_cast<uint16_t>(r_docs - 1);
          }
          found += __popc(m);
        }
        if (lane == 0) {
          sm.n_halo = min(found, halo_want);
        }
      }
      __syncthreads();
      const uint32_t n_all = n_chars + sm.n_halo;

      // ---- C: windows (GenerateHybridNgrams, string_utils.cpp:452-509). A strip is kFlatPerThread * 256 characters,
      // every thread takes kFlatPerThread CONSECUTIVE ones (text order = thread order, then the order inside a
      // thread), so a tile of CJK text needs two block-wide scans instead of six.
      uint64_t strip_base = EMIT ? tile_off[tile] + emitted_tile : 0;  // slot of the strip's first n-gram
      for (uint32_t k0 = 0; k0 < n_chars; k0 += kFlatThreads * kFlatPerThread) {
        uint32_t ok_mask = 0;
        uint64_t key[kFlatPerThread];
#pragma unroll
        for (int c4 = 0; c4 < kFlatPerThread; ++c4) {
          const uint32_t k = k0 + threadIdx.x * kFlatPerThread + c4;
          key[c4] = 0;
          if (k < n_chars) {
            const uint32_t c = sm.cp[k];
            const uint32_t dk = sm.doc[k];
            const bool cjk = is_cjk_ideograph(c);
            const int size = cjk ? kanji : ngram;  // :484-485: chosen by the START code point
            // :487 the window must fit the document: its last character exists and is in the same document
            if (k + static_cast<uint32_t>(size) <= n_all && sm.doc[k + size - 1] == dk) {
              bool ok = true;
              uint64_t kk = static_cast<uint64_t>(c) + 1;
              if (wide_words > 0) {
                for (int j = 1; j < size; ++j) {
                  if (!cross && is_cjk_ideograph(sm.cp[k + j]) != cjk) {  // :491-503
                    ok = false;
                  }
                }
              }
              for (int j = 1; j < width && wide_words == 0; ++j) {
                uint64_t field = 0;
                if (j < size) {
                  const uint32_t cj = sm.cp[k + j];
                  if (!cross && is_cjk_ideograph(cj) != cjk) {  // :491-503 legacy boundary rejection
                    ok = false;
                  }
                  field = static_cast<uint64_t>(cj) + 1;
                }
                kk = (kk << 21) | field;
              }
              key[c4] = kk;
              ok_mask |= ok ? (1u << c4) : 0u;
            }
          }
        }
        uint32_t n_emit = 0;
        uint32_t rank = flat_block_scan(static_cast<uint32_t>(__popc(ok_mask)), sm.warp_cnt, &n_emit);
        if (EMIT) {
#pragma unroll
          for (int c4 = 0; c4 < kFlatPerThread; ++c4) {
            if ((ok_mask >> c4) & 1u) {
              const uint32_t k = k0 + threadIdx.x * kFlatPerThread + c4;
              const uint32_t dk = sm.doc[k];
              const uint64_t slot = strip_base + rank;
              ++rank;
              if (wide_words > 0) {
                const int size = is_cjk_ideograph(sm.cp[k]) ? kanji : ngram;
                for (int w = 0; w < wide_words; ++w) {
                  uint64_t word = 0;
                  for (int f = 0; f < 3; ++f) {
                    const int j = 3 * w + f;
                    word = (word << 21) | (j < size ? static_cast<uint64_t>(sm.cp[k + j]) + 1 : 0ULL);
                  }
                  keys_out[static_cast<uint64_t>(w) * wide_stride + slot] = word;
                }
              } else {
                const uint64_t in_doc = static_cast<uint64_t>(static_cast<int64_t>(sm.pos[k]) - sm.docrel[dk]);
                uint64_t low = umin_u64(in_doc, pos_max);
                if (sig_cfg != 0) {
                  const uint32_t size = static_cast<uint32_t>(is_cjk_ideograph(sm.cp[k]) ? kanji : ngram);
                  uint32_t nx = kNoCp;  // next decoded character of the same document
                  if (k + size < n_all && sm.doc[k + size] == dk) {
                    nx = sm.cp[k + size];
                  }
                  uint32_t pv = kNoCp;  // previous one
                  if (k > 0) {
                    pv = sm.doc[k - 1] == dk ? sm.cp[k - 1] : kNoCp;
                  } else if (dk == 0) {
                    pv = sm.prev_cp;
                  }
                  const uint64_t s_next = (nx != kNoCp && sig_next > 0) ? sig_of_hash(sig_hash7(nx), sig_next) : 0u;
                  const uint64_t s_prev = (pv != kNoCp && sig_prev > 0) ? sig_of_hash(sig_hash7(pv), sig_prev) : 0u;
                  low |= (s_next << sig_pos) | (s_prev << (sig_pos + sig_next));
                }
                keys_out[slot] = pos_bits > 0 ? ((key[c4] << pos_bits) | low) : key[c4];
              }
              docs_out[slot] = r_first + dk;
            }
          }
        }
        strip_base += n_emit;
        emitted_tile += n_emit;
      }
      // ---- per-document totals of this round
      if (STATS) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < r_docs; i += kFlatThreads) {
          const uint32_t c = sm.doc_cps[i];
          const uint32_t vb = sm.doc_valid[i];
          if (c != 0) {
            atomicAdd(&doc_len[static_cast<uint64_t>(r_first) + i], c);
          }
          if (vb != 0) {
            atomicAdd(&doc_valid_bytes[static_cast<uint64_t>(r_first) + i], vb);
          }
        }
      }
    }
    if (MODE == kTokCount && threadIdx.x == 0) {
      tile_cnt[tile] = static_cast<uint32_t>(emitted_tile);
    }
    if (MODE == kTokFused) {
      // placeholders for the slots of the tile's upper bound that no n-gram took
      for (uint64_t i = tile_off[tile] + emitted_tile + threadIdx.x; i < tile_off[tile + 1]; i += kFlatThreads) {
        keys_out[i] = 0;
        for (int w = 1; w < wide_words; ++w) {
          keys_out[static_cast<uint64_t>(w) * wide_stride + i] = 0;
        }
        docs_out[i] = 0;
      }
    }
  }
}

// Upper bound of a tile's n-grams: its character-start (non-continuation) bytes. One 16-byte chunk per thread.
__global__ void __launch_bounds__(kFlatThreads) lead_count_kernel(const uint8_t* __restrict__ text, uint64_t text_bytes,
                                                                  uint64_t n_tiles, uint32_t* __restrict__ tile_cnt) {
  __shared__ uint32_t warp_cnt[kFlatThreads / 32];
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t byte0 = tile * kFlatTile + static_cast<uint64_t>(threadIdx.x) * 16;
    uint32_t c = 0;
    if (byte0 < text_bytes) {
      const uint4 v = ld_stream_16(text + byte0);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      const uint32_t n = static_cast<uint32_t>(umin_u64(16, text_bytes - byte0));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // bytes that are NOT 10xxxxxx; bytes beyond the arena are masked off
        uint32_t lead = ~__vcmpeq4(w[k] & 0xC0C0C0C0u, 0x80808080u) & 0x01010101u;
        const int valid = static_cast<int>(n) - 4 * k;
        if (valid < 4) {
          lead &= valid <= 0 ? 0u : ((1u << (8 * valid)) - 1u);
        }
        c += __popc(lead);
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      c += __shfl_xor_sync(0xffffffffu, c, s);
    }
    __syncthreads();
    if ((threadIdx.x & 31u) == 0) {
      warp_cnt[threadIdx.x >> 5] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
      for (int w2 = 0; w2 < kFlatThreads / 32; ++w2) {
        t += warp_cnt[w2];
      }
      tile_cnt[tile] = t;
    }
  }
}

// After COUNT: corpus counters from the per-document totals. counters: [0] non-empty docs, [1] docs with invalid
// bytes, [2] code points.
__global__ void __launch_bounds__(256) flat_doc_stats_kernel(const uint64_t* __restrict__ text_off,
                                                             const uint32_t* __restrict__ doc_len,
                                                             const uint32_t* __restrict__ doc_valid_bytes,
                                                             uint64_t n_docs, unsigned long long* __restrict__ counters) {
  unsigned long long nonempty = 0, invalid = 0, cps = 0;
  for (uint64_t d = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; d < n_docs;
       d += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t bytes = text_off[d + 1] - text_off[d];
    nonempty += bytes != 0 ? 1 : 0;
    invalid += doc_valid_bytes[d] != bytes ? 1 : 0;
    cps += doc_len[d];
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    nonempty += __shfl_xor_sync(0xffffffffu, nonempty, s);
    invalid += __shfl_xor_sync(0xffffffffu, invalid, s);
    cps += __shfl_xor_sync(0xffffffffu, cps, s);
  }
  if ((threadIdx.x & 31u) == 0) {
    if (nonempty != 0) atomicAdd(&counters[0], nonempty);
    if (invalid != 0) atomicAdd(&counters[1], invalid);
    if (cps != 0) atomicAdd(&counters[2], cps);
  }
}

// ------------------------------------------------------------------ CSR
constexpr int kCsrThreads = 256;
constexpr int kCsrItems = 8;
constexpr int kCsrTile = kCsrThreads * kCsrItems;

struct Heads {
  uint32_t pairs;
  uint32_t terms;
};

// pb = position bits carried in the low end of every key (0 if none): n-gram identity is key >> pb.
//
// Both CSR kernels walk the sorted pairs WARP-STRIPED: a CTA takes kCsrTile pairs, warp w of it the 256 consecutive
// pairs [w * 256, (w + 1) * 256) as kCsrItems rows of 32 (row k, lane l = pair w * 256 + k * 32 + l). Every load is a
// full-width coalesced row, the neighbours a head flag needs come from the next lane's register (a shuffle) instead
// of a second and third load of the same array, and the rank of a head inside a row is a ballot and a popcount.
struct CsrRows {
  uint64_t key[kCsrItems + 1];  // row kCsrItems: lanes 0, 1 hold the two pairs after the warp's range
  uint32_t doc[kCsrItems + 1];
  uint64_t prev_key;            // lane 0: the pair before the warp's range (kInvalidKey if none)
  uint32_t prev_doc;
};

__device__ __forceinline__ void csr_load_rows(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ docs,
                                              uint64_t wbase, uint64_t n, unsigned lane, bool tail, CsrRows* r) {
#pragma unroll
  for (int k = 0; k < kCsrItems; ++k) {
    const uint64_t i = wbase + static_cast<uint64_t>(k) * 32 + lane;
    const bool in = i < n;
    r->key[k] = in ? keys[i] : kInvalidKey;
    r->doc[k] = in ? docs[i] : 0u;
  }
  r->key[kCsrItems] = kInvalidKey;
  r->doc[kCsrItems] = 0;
  if (tail && lane < 2) {
    const uint64_t i = wbase + static_cast<uint64_t>(kCsrItems) * 32 + lane;
    if (i < n) {
      r->key[kCsrItems] = keys[i];
      r->doc[kCsrItems] = docs[i];
    }
  }
  r->prev_key = kInvalidKey;
  r->prev_doc = 0;
  if (lane == 0 && wbase > 0 && wbase - 1 < n) {
    r->prev_key = keys[wbase - 1];
    r->prev_doc = docs[wbase - 1];
  }
}

// head flags of row k: bit 0 = first pair of an (n-gram, document) run, bit 1 = first pair of an n-gram
__device__ __forceinline__ uint32_t csr_row_heads(const CsrRows& r, int k, uint64_t wbase, unsigned lane, int pb) {
  const uint64_t key = r.key[k];
  // the pair before this one: the previous lane, the last lane of the previous row, or the pair before the range
  uint64_t pk = __shfl_up_sync(0xffffffffu, key, 1);
  uint32_t pd = __shfl_up_sync(0xffffffffu, r.doc[k], 1);
  const uint64_t wrap_k = __shfl_sync(0xffffffffu, k > 0 ? r.key[k > 0 ? k - 1 : 0] : r.prev_key, k > 0 ? 31 : 0);
  const uint32_t wrap_d = __shfl_sync(0xffffffffu, k > 0 ? r.doc[k > 0 ? k - 1 : 0] : r.prev_doc, k > 0 ? 31 : 0);
  if (lane == 0) {
    pk = wrap_k;
    pd = wrap_d;
  }
  if (key == kInvalidKey || (key >> pb) == 0) {
    return 0;  // beyond the end / placeholder slots (fused tokenizer: key 0; no n-gram has it)
  }
  const bool first = wbase == 0 && k == 0 && lane == 0;
  const bool term_head = first || pk == kInvalidKey || (pk >> pb) != (key >> pb);
  const bool pair_head = term_head || pd != r.doc[k];
  return (pair_head ? 1u : 0u) | (term_head ? 2u : 0u);
}

__global__ void __launch_bounds__(kCsrThreads) csr_count_kernel(const uint64_t* __restrict__ keys,
                                                                const uint32_t* __restrict__ docs, uint64_t n, int pb,
                                                                uint64_t* __restrict__ block_pairs,
                                                                uint64_t* __restrict__ block_terms) {
  __shared__ uint32_t sp[kCsrThreads / 32];
  __shared__ uint32_t st[kCsrThreads / 32];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const uint64_t wbase = static_cast<uint64_t>(blockIdx.x) * kCsrTile + static_cast<uint64_t>(warp) * (32 * kCsrItems);
  CsrRows r;
  csr_load_rows(keys, docs, wbase, n, lane, false, &r);
  uint32_t pairs = 0;
  uint32_t terms = 0;
#pragma unroll
  for (int k = 0; k < kCsrItems; ++k) {
    const uint32_t h = csr_row_heads(r, k, wbase, lane, pb);
    pairs += __popc(__ballot_sync(0xffffffffu, (h & 1u) != 0));
    terms += __popc(__ballot_sync(0xffffffffu, (h & 2u) != 0));
  }
  if (lane == 0) {
    sp[warp] = pairs;
    st[warp] = terms;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t p = 0;
    uint64_t t = 0;
    for (int w = 0; w < kCsrThreads / 32; ++w) {
      p += sp[w];
      t += st[w];
    }
    block_pairs[blockIdx.x] = p;
    block_terms[blockIdx.x] = t;
  }
}

// single CTA: exclusive scan of both block-count arrays in place; totals -> totals[0..1]
__global__ void __launch_bounds__(1024) csr_scan_kernel(uint64_t* __restrict__ block_pairs,
                                                        uint64_t* __restrict__ block_terms, uint64_t n_blocks,
                                                        uint64_t* __restrict__ totals) {
  __shared__ uint64_t warp_tot[2][32];
  __shared__ uint64_t carry[2];
  if (threadIdx.x < 2) {
    carry[threadIdx.x] = 0;
  }
  __syncthreads();
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  for (uint64_t base = 0; base < n_blocks; base += 1024) {
    const uint64_t i = base + threadIdx.x;
    uint64_t v[2] = {i < n_blocks ? block_pairs[i] : 0, i < n_blocks ? block_terms[i] : 0};
    uint64_t inc[2] = {v[0], v[1]};
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const uint64_t o0 = __shfl_up_sync(0xffffffffu, inc[0], s);
      const uint64_t o1 = __shfl_up_sync(0xffffffffu, inc[1], s);
      if (lane >= static_cast<unsigned>(s)) {
        inc[0] += o0;
        inc[1] += o1;
      }
    }
    if (lane == 31) {
      warp_tot[0][warp] = inc[0];
      warp_tot[1][warp] = inc[1];
    }
    __syncthreads();
    if (warp < 2) {
      uint64_t w = warp_tot[warp][lane];
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const uint64_t o = __shfl_up_sync(0xffffffffu, w, s);
        if (lane >= static_cast<unsigned>(s)) {
          w += o;
        }
      }
      warp_tot[warp][lane] = w;
    }
    __syncthreads();
    const uint64_t c0 = carry[0];
    const uint64_t c1 = carry[1];
    if (i < n_blocks) {
      block_pairs[i] = c0 + (warp == 0 ? 0 : warp_tot[0][warp - 1]) + inc[0] - v[0];
      block_terms[i] = c1 + (warp == 0 ? 0 : warp_tot[1][warp - 1]) + inc[1] - v[1];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      carry[0] = c0 + warp_tot[0][31];
      carry[1] = c1 + warp_tot[1][31];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    totals[0] = carry[0];
    totals[1] = carry[1];
  }
}

__global__ void __launch_bounds__(kCsrThreads) csr_write_kernel(const uint64_t* __restrict__ keys,
                                                                const uint32_t* __restrict__ docs, uint64_t n, int pb,
                                                                const uint64_t* __restrict__ block_pairs,
                                                                const uint64_t* __restrict__ block_terms,
                                                                uint64_t* __restrict__ term_keys,
                                                                uint64_t* __restrict__ term_off,
                                                                uint32_t* __restrict__ postings,
                                                                uint32_t* __restrict__ post_pos,
                                                                uint32_t* __restrict__ post_pos2, int sig_cfg) {
  // sig_cfg: split of the pb payload bits (tokenize_flat_kernel); 0 = all of them are the offset
  __shared__ uint32_t sp[kCsrThreads / 32];
  __shared__ uint32_t st[kCsrThreads / 32];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint64_t wbase = static_cast<uint64_t>(blockIdx.x) * kCsrTile + static_cast<uint64_t>(warp) * (32 * kCsrItems);
  CsrRows r;
  csr_load_rows(keys, docs, wbase, n, lane, pb > 0, &r);
  uint32_t heads[kCsrItems];
  uint32_t pairs = 0;
  uint32_t terms = 0;
#pragma unroll
  for (int k = 0; k < kCsrItems; ++k) {
    heads[k] = csr_row_heads(r, k, wbase, lane, pb);
    pairs += __popc(__ballot_sync(0xffffffffu, (heads[k] & 1u) != 0));
    terms += __popc(__ballot_sync(0xffffffffu, (heads[k] & 2u) != 0));
  }
  if (lane == 0) {
    sp[warp] = pairs;
    st[warp] = terms;
  }
  __syncthreads();
  uint64_t pp = block_pairs[blockIdx.x];  // output slot of the warp's next pair head / term head
  uint64_t tp = block_terms[blockIdx.x];
  for (unsigned w = 0; w < warp; ++w) {
    pp += sp[w];
    tp += st[w];
  }
  const int off_bits = sig_cfg != 0 ? (sig_cfg & 0xFF) : pb;
  const int next_bits = (sig_cfg >> 8) & 0xFF;
  const uint64_t pos_mask = pb > 0 ? (1ULL << off_bits) - 1 : 0;
  const uint64_t pay_mask = pb > 0 ? (1ULL << pb) - 1 : 0;
  const uint32_t next_mask = (1u << next_bits) - 1u;
#pragma unroll
  for (int k = 0; k < kCsrItems; ++k) {
    const unsigned pair_mask = __ballot_sync(0xffffffffu, (heads[k] & 1u) != 0);
    const unsigned term_mask = __ballot_sync(0xffffffffu, (heads[k] & 2u) != 0);
    const uint64_t key = r.key[k];
    const uint32_t doc = r.doc[k];
    // the two pairs after this one (only the position payload looks at them)
    uint64_t k1 = kInvalidKey, k2 = kInvalidKey;
    uint32_t d1 = 0, d2 = 0;
    if (pb > 0) {
      k1 = __shfl_down_sync(0xffffffffu, key, 1);
      d1 = __shfl_down_sync(0xffffffffu, doc, 1);
      k2 = __shfl_down_sync(0xffffffffu, key, 2);
      d2 = __shfl_down_sync(0xffffffffu, doc, 2);
      const uint64_t nk = __shfl_sync(0xffffffffu, r.key[k + 1], (lane + 2) & 31);  // lanes 30, 31 read lanes 0, 1
      const uint32_t nd = __shfl_sync(0xffffffffu, r.doc[k + 1], (lane + 2) & 31);
      const uint64_t nk0 = __shfl_sync(0xffffffffu, r.key[k + 1], 0);
      const uint32_t nd0 = __shfl_sync(0xffffffffu, r.doc[k + 1], 0);
      if (lane == 31) {
        k1 = nk0;
        d1 = nd0;
      }
      if (lane >= 30) {
        k2 = nk;
        d2 = nd;
      }
    }
    if ((heads[k] & 1u) != 0) {
      const uint64_t at = pp + __popc(pair_mask & lt_mask);
      postings[at] = doc;
      if (pb > 0) {
        // The sort is stable and the tokenizer emits a document's n-grams in text order, so the head of a
        // (n-gram, document) run is the first occurrence; a second entry of the run means "more than once".
        const bool multi = k1 != kInvalidKey && (k1 >> pb) == (key >> pb) && d1 == doc;
        const bool third = multi && k2 != kInvalidKey && (k2 >> pb) == (key >> pb) && d2 == doc;
        const uint64_t pos = key & pos_mask;
        const uint64_t pos2 = multi ? (k1 & pos_mask) : kPosUnknown;
        // neighbour signatures of the two occurrences: next in bits 16..22, prev in bits 24..30 of the payload word
        const uint32_t sg1 = static_cast<uint32_t>((key & pay_mask) >> off_bits);
        const uint32_t sg2 = multi ? static_cast<uint32_t>((k1 & pay_mask) >> off_bits) : 0u;
        post_pos[at] = static_cast<uint32_t>((pos < kPosUnknown ? pos : kPosUnknown) | (multi ? kPosMulti : 0)) |
                       ((sg1 & next_mask) << 16) | ((sg1 >> next_bits) << 24);
        post_pos2[at] = static_cast<uint32_t>((pos2 < kPosUnknown ? pos2 : kPosUnknown) | (third ? kPosMulti : 0)) |
                        ((sg2 & next_mask) << 16) | ((sg2 >> next_bits) << 24);
      }
      if ((heads[k] & 2u) != 0) {
        const uint64_t tat = tp + __popc(term_mask & lt_mask);
        term_keys[tat] = key >> pb;
        term_off[tat] = at;
      }
    }
    pp += __popc(pair_mask);
    tp += __popc(term_mask);
  }
}

// ------------------------------------------------------------------ wide keys (n-gram sizes 4..10)
__global__ void iota_u32_kernel(uint32_t* __restrict__ out, uint64_t n) {
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    out[i] = static_cast<uint32_t>(i);
  }
}

__global__ void gather_u64_kernel(const uint64_t* __restrict__ src, const uint32_t* __restrict__ perm,
                                  uint64_t* __restrict__ dst, uint64_t n) {
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    dst[i] = src[perm[i]];
  }
}

// word0[i] / perm[i]: the slots in wide-key order (stable); words: the tokenizer's arrays (word w of slot s at
// words[w * stride + s]). flags[i] = 1 where a new n-gram starts (placeholder slots, word 0 == 0, never do).
__global__ void wide_heads_kernel(const uint64_t* __restrict__ word0, const uint32_t* __restrict__ perm,
                                  const uint64_t* __restrict__ words, uint64_t stride, int n_words, uint64_t n,
                                  uint32_t* __restrict__ flags) {
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t k = word0[i];
    uint32_t head = 0;
    if (k != 0) {
      head = (i == 0 || word0[i - 1] != k) ? 1u : 0u;
      if (head == 0) {
        const uint32_t a = perm[i - 1];
        const uint32_t b = perm[i];
        for (int w = 1; w < n_words; ++w) {
          if (words[static_cast<uint64_t>(w) * stride + a] != words[static_cast<uint64_t>(w) * stride + b]) {
            head = 1;
          }
        }
      }
    }
    flags[i] = head;
  }
}

// rank[i] = heads before i. The n-gram of slot i gets the key rank + 1 (its rank among the distinct n-grams, from 1),
// placeholders keep 0; docs follow the permutation.
__global__ void wide_assign_kernel(const uint64_t* __restrict__ word0, const uint32_t* __restrict__ perm,
                                   const uint32_t* __restrict__ flags, const uint64_t* __restrict__ rank,
                                   const uint32_t* __restrict__ docs_in, uint64_t n, uint64_t* __restrict__ keys_out,
                                   uint32_t* __restrict__ docs_out) {
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    keys_out[i] = word0[i] == 0 ? 0 : rank[i] + flags[i];
    docs_out[i] = docs_in[perm[i]];
  }
}

// the dictionary of wide keys: row (rank of the n-gram) = its words
__global__ void wide_table_kernel(const uint64_t* __restrict__ word0, const uint32_t* __restrict__ perm,
                                  const uint32_t* __restrict__ flags, const uint64_t* __restrict__ rank,
                                  const uint64_t* __restrict__ words, uint64_t stride, int n_words, uint64_t n,
                                  uint64_t* __restrict__ table) {
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    if (flags[i] != 0) {
      uint64_t* row = table + rank[i] * static_cast<uint64_t>(n_words);
      row[0] = word0[i];
      const uint32_t s = perm[i];
      for (int w = 1; w < n_words; ++w) {
        row[w] = words[static_cast<uint64_t>(w) * stride + s];
      }
    }
  }
}

__global__ void iota_keys_kernel(uint64_t* __restrict__ out, uint64_t n) {
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    out[i] = i + 1;
  }
}

__global__ void set_u64_kernel(uint64_t* p, uint64_t v) { *p = v; }

// ------------------------------------------------------------------ dense bitmaps
__global__ void dense_count_kernel(const uint64_t* __restrict__ term_off, uint64_t n_terms, uint64_t min_len,
                                   unsigned long long* __restrict__ count) {
  const uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < n_terms && term_off[t + 1] - term_off[t] >= min_len) {
    atomicAdd(count, 1ULL);
  }
}

__global__ void dense_assign_kernel(const uint64_t* __restrict__ term_off, uint64_t n_terms, uint64_t min_len,
                                    unsigned long long* __restrict__ counter, int32_t* __restrict__ term_bm,
                                    uint32_t* __restrict__ dense_terms) {
  const uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_terms) {
    return;
  }
  int32_t slot = -1;
  if (term_off[t + 1] - term_off[t] >= min_len) {
    slot = static_cast<int32_t>(atomicAdd(counter, 1ULL));
    dense_terms[slot] = static_cast<uint32_t>(t);
  }
  term_bm[t] = slot;
}

// grid.x = dense slot, grid.y = slices of the list
__global__ void __launch_bounds__(256) dense_fill_kernel(const uint64_t* __restrict__ term_off,
                                                         const uint32_t* __restrict__ postings,
                                                         const uint32_t* __restrict__ dense_terms,
                                                         uint32_t* __restrict__ bitmaps, uint64_t bm_words) {
  const uint32_t slot = blockIdx.x;
  const uint64_t t = dense_terms[slot];
  const uint64_t b = term_off[t];
  const uint64_t e = term_off[t + 1];
  uint32_t* bm = bitmaps + static_cast<uint64_t>(slot) * bm_words;
  for (uint64_t i = b + static_cast<uint64_t>(blockIdx.y) * blockDim.x + threadIdx.x; i < e;
       i += static_cast<uint64_t>(gridDim.y) * blockDim.x) {
    const uint32_t d = postings[i];
    atomicOr(&bm[d >> 5], 1u << (d & 31));
  }
}

}  // namespace

SearchScratch* Index::scratch_for(cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(pool_mu);
  for (SearchScratch* s : scratches) {
    if (s->stream == stream) {
      return s;
    }
  }
  SearchScratch* s = new SearchScratch();
  s->stream = stream;
  scratches.push_back(s);
  return s;
}

Index::~Index() {
  for (SearchScratch* s : scratches) {
    delete s;
  }
  drop_filter_columns();
}

void Index::drop_filter_columns() {
  for (FilterColumn*& c : columns) {
    delete c;
    c = nullptr;
  }
}

uint64_t Index::device_bytes() const { return resident_a.blob.bytes() + resident_b.blob.bytes() + d_bitmaps.bytes(); }

namespace {
struct TokScratch {
  unsigned long long* counters;
  uint32_t* tile_first_doc;
  uint32_t* valid_bytes;
  uint32_t* tile_cnt;
  uint64_t* scan;
  uint64_t n_tiles;
};
uint64_t flat_tiles(uint64_t n_docs, uint64_t text_bytes) {
  return n_docs > 0 ? (text_bytes + kFlatTile - 1) / kFlatTile : 0;
}
TokScratch tok_scratch(uint64_t* d_scratch, uint64_t n_docs, uint64_t text_bytes) {
  TokScratch t;
  t.n_tiles = flat_tiles(n_docs, text_bytes);
  uint64_t* p = d_scratch;
  t.counters = reinterpret_cast<unsigned long long*>(p);
  p += 8;
  t.tile_first_doc = reinterpret_cast<uint32_t*>(p);
  p += (t.n_tiles + 2) / 2 + 1;
  t.valid_bytes = reinterpret_cast<uint32_t*>(p);
  p += (n_docs + 2) / 2 + 1;
  t.tile_cnt = reinterpret_cast<uint32_t*>(p);
  p += (t.n_tiles + 2) / 2 + 1;
  t.scan = p;
  return t;
}
unsigned flat_grid(uint64_t n_tiles) {
  int sm_count = 148;
  int dev = 0;
  MGX_CUDA(cudaGetDevice(&dev));
  MGX_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  return static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(static_cast<uint64_t>(sm_count) * 4, n_tiles)));
}
}  // namespace

size_t tokenize_tile_count(uint64_t n_docs, uint64_t text_bytes) { return flat_tiles(n_docs, text_bytes); }

size_t tokenize_scratch_elems(uint64_t n_docs, uint64_t text_bytes) {
  const uint64_t n_tiles = flat_tiles(n_docs, text_bytes);
  return 8 + 2 * ((n_tiles + 2) / 2 + 1) + (n_docs + 2) / 2 + 1 + scan_scratch_elems(std::max<uint64_t>(n_tiles, 1)) + 8;
}

void tokenize_count(int ngram, int kanji, bool cross, int width, const uint8_t* d_text, const uint64_t* d_text_off,
                    uint64_t n_docs, uint64_t text_bytes, uint32_t* d_doc_len, uint64_t* d_tile_off,
                    uint64_t* d_scratch, uint64_t* n_slots, uint64_t* counters_out, cudaStream_t stream) {
  const TokScratch ts = tok_scratch(d_scratch, n_docs, text_bytes);
  MGX_CUDA(cudaMemsetAsync(ts.counters, 0, 8 * sizeof(unsigned long long), stream));
  if (n_docs > 0) {
    MGX_CUDA(cudaMemsetAsync(d_doc_len, 0, n_docs * sizeof(uint32_t), stream));
    MGX_CUDA(cudaMemsetAsync(ts.valid_bytes, 0, n_docs * sizeof(uint32_t), stream));
  }
  if (ts.n_tiles > 0) {
    flat_tile_first_doc_kernel<<<static_cast<unsigned>((ts.n_tiles + 255) / 256), 256, 0, stream>>>(
        d_text_off, n_docs, ts.n_tiles, ts.tile_first_doc);
    MGX_LAUNCH_CHECK();
    tokenize_flat_kernel<kTokCount><<<flat_grid(ts.n_tiles), kFlatThreads, 0, stream>>>(
        d_text, d_text_off, n_docs, text_bytes, ts.n_tiles, ts.tile_first_doc, ngram, kanji, cross ? 1 : 0, width, 0,
        d_doc_len, ts.valid_bytes, ts.tile_cnt, nullptr, nullptr, nullptr, wide_words_for(width), 0, 0);
    MGX_LAUNCH_CHECK();
  }
  if (n_docs > 0) {
    flat_doc_stats_kernel<<<static_cast<unsigned>(std::min<uint64_t>((n_docs + 255) / 256, 148 * 8)), 256, 0, stream>>>(
        d_text_off, d_doc_len, ts.valid_bytes, n_docs, ts.counters);
    MGX_LAUNCH_CHECK();
  }
  exclusive_scan_u32_u64(ts.tile_cnt, d_tile_off, ts.n_tiles, ts.scan, stream);
  uint64_t total = 0;
  unsigned long long counters[3] = {0, 0, 0};
  MGX_CUDA(cudaMemcpyAsync(&total, d_tile_off + ts.n_tiles, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  MGX_CUDA(cudaMemcpyAsync(counters, ts.counters, sizeof(counters), cudaMemcpyDeviceToHost, stream));
  MGX_CUDA(cudaStreamSynchronize(stream));
  *n_slots = total;
  counters_out[0] = counters[0];
  counters_out[1] = counters[1];
  counters_out[2] = counters[2];
}

void tokenize_emit(int ngram, int kanji, bool cross, int width, const uint8_t* d_text, const uint64_t* d_text_off,
                   uint64_t n_docs, uint64_t text_bytes, const uint64_t* d_tile_off, uint64_t* d_scratch,
                   uint64_t* d_keys, uint32_t* d_docs, int pos_bits, cudaStream_t stream, uint64_t wide_stride) {
  const TokScratch ts = tok_scratch(d_scratch, n_docs, text_bytes);  // tile_first_doc was filled by tokenize_count
  if (ts.n_tiles == 0) {
    return;
  }
  tokenize_flat_kernel<kTokEmit><<<flat_grid(ts.n_tiles), kFlatThreads, 0, stream>>>(
      d_text, d_text_off, n_docs, text_bytes, ts.n_tiles, ts.tile_first_doc, ngram, kanji, cross ? 1 : 0, width, pos_bits,
      nullptr, nullptr, nullptr, d_tile_off, d_keys, d_docs, wide_words_for(width), wide_stride, 0);
  MGX_LAUNCH_CHECK();
}

// Fused form used by the index build. Step 1: upper bound of the n-gram slots (scanned per tile into d_tile_off).
void tokenize_bound(const uint8_t* d_text, const uint64_t* d_text_off, uint64_t n_docs, uint64_t text_bytes,
                    uint64_t* d_tile_off, uint64_t* d_scratch, uint64_t* n_slots_bound, cudaStream_t stream) {
  const TokScratch ts = tok_scratch(d_scratch, n_docs, text_bytes);
  if (ts.n_tiles > 0) {
    flat_tile_first_doc_kernel<<<static_cast<unsigned>((ts.n_tiles + 255) / 256), 256, 0, stream>>>(
        d_text_off, n_docs, ts.n_tiles, ts.tile_first_doc);
    MGX_LAUNCH_CHECK();
    lead_count_kernel<<<flat_grid(ts.n_tiles), kFlatThreads, 0, stream>>>(d_text, text_bytes, ts.n_tiles, ts.tile_cnt);
    MGX_LAUNCH_CHECK();
  }
  exclusive_scan_u32_u64(ts.tile_cnt, d_tile_off, ts.n_tiles, ts.scan, stream);
  uint64_t total = 0;
  MGX_CUDA(cudaMemcpyAsync(&total, d_tile_off + ts.n_tiles, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  MGX_CUDA(cudaStreamSynchronize(stream));
  *n_slots_bound = total;
}

// Step 2: per-document totals AND the (key, doc) pairs in one pass; unused slots of every tile hold key 0.
void tokenize_fused(int ngram, int kanji, bool cross, int width, const uint8_t* d_text, const uint64_t* d_text_off,
                    uint64_t n_docs, uint64_t text_bytes, uint32_t* d_doc_len, const uint64_t* d_tile_off,
                    uint64_t* d_scratch, uint64_t* d_keys, uint32_t* d_docs, int pos_bits, uint64_t* counters_out,
                    cudaStream_t stream, uint64_t wide_stride = 0, int sig_cfg = 0) {
  const TokScratch ts = tok_scratch(d_scratch, n_docs, text_bytes);
  MGX_CUDA(cudaMemsetAsync(ts.counters, 0, 8 * sizeof(unsigned long long), stream));
  if (n_docs > 0) {
    MGX_CUDA(cudaMemsetAsync(d_doc_len, 0, n_docs * sizeof(uint32_t), stream));
    MGX_CUDA(cudaMemsetAsync(ts.valid_bytes, 0, n_docs * sizeof(uint32_t), stream));
  }
  if (ts.n_tiles > 0) {
    tokenize_flat_kernel<kTokFused><<<flat_grid(ts.n_tiles), kFlatThreads, 0, stream>>>(
        d_text, d_text_off, n_docs, text_bytes, ts.n_tiles, ts.tile_first_doc, ngram, kanji, cross ? 1 : 0, width, pos_bits,
        d_doc_len, ts.valid_bytes, nullptr, d_tile_off, d_keys, d_docs, wide_words_for(width), wide_stride,
        pos_bits > 0 ? sig_cfg : 0);
    MGX_LAUNCH_CHECK();
  }
  if (n_docs > 0) {
    flat_doc_stats_kernel<<<static_cast<unsigned>(std::min<uint64_t>((n_docs + 255) / 256, 148 * 8)), 256, 0, stream>>>(
        d_text_off, d_doc_len, ts.valid_bytes, n_docs, ts.counters);
    MGX_LAUNCH_CHECK();
  }
  unsigned long long counters[3] = {0, 0, 0};
  MGX_CUDA(cudaMemcpyAsync(counters, ts.counters, sizeof(counters), cudaMemcpyDeviceToHost, stream));
  MGX_CUDA(cudaStreamSynchronize(stream));
  counters_out[0] = counters[0];
  counters_out[1] = counters[1];
  counters_out[2] = counters[2];
}

// doc ids ascending: first_id + i everywhere?  (one pass over the resident copy)
// bad[1] = bytes of the longest document (saturated): sizes the offset field of the posting payload (SigLayout)
__global__ void sequential_check_kernel(const uint32_t* __restrict__ ids, const uint64_t* __restrict__ text_off,
                                        uint64_t n, unsigned int* __restrict__ bad) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i + 1 < n) {
    if (ids[i + 1] != ids[i] + 1) {
      atomicOr(bad, ids[i + 1] > ids[i] ? 1u : 2u);  // 1: gap, 2: not ascending
    }
  }
  uint64_t bytes = 0;
  if (i < n) {
    bytes = text_off[i + 1] - text_off[i];
  }
  const unsigned int longest = __reduce_max_sync(0xffffffffu, static_cast<unsigned int>(umin_u64(bytes, 0xFFFFFFFFu)));
  if ((threadIdx.x & 31u) == 0 && longest != 0) {
    atomicMax(bad + 1, longest);
  }
}

// tile_first_doc[t] = the document that holds byte t * kTextTileBytes of the arena: the largest d with
// text_off[d] <= that byte (documents are stored back to back; empty documents share an offset).
__global__ void tile_first_doc_kernel(const uint64_t* __restrict__ text_off, uint64_t n_docs, uint64_t n_tiles,
                                      uint32_t* __restrict__ out) {
  const uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_tiles) {
    return;
  }
  const uint64_t byte = t * kTextTileBytes;
  uint64_t lo = 0;  // invariant: text_off[lo] <= byte
  uint64_t hi = n_docs;
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (text_off[mid] <= byte) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  out[t] = static_cast<uint32_t>(lo);
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
PhaseTrace::PhaseTrace(cudaStream_t s) : on(std::getenv("MGX_BUILD_TRACE") != nullptr), stream(s), t0(now_ms()) {}
void PhaseTrace::mark(const char* name) {
  if (!on) {
    return;
  }
  cudaStreamSynchronize(stream);
  const double t1 = now_ms();
  std::fprintf(stderr, "[mgx build] %-34s %8.2f ms\n", name, t1 - t0);
  t0 = t1;
}

namespace {
struct ResetOnFailure {
  Index& ix;
  bool armed = true;
  ~ResetOnFailure() {
    if (!armed) {
      return;
    }
    ix.n_docs = ix.text_bytes = ix.n_text_tiles = 0;
    ix.n_terms = ix.n_postings = ix.n_dense = ix.bm_words = 0;
    ix.doc_count = ix.total_doc_length = ix.n_pair_slots = 0;
    ix.has_positions = false;
    ix.sequential_ids = true;
    ix.first_id = 1;
    ix.d_doc_ids.release();
    ix.d_text.release();
    ix.d_text_off.release();
    ix.d_doc_len.release();
    ix.d_tile_first_doc.release();
    ix.d_term_keys.release();
    ix.d_wide_keys.release();
    ix.d_term_off.release();
    ix.d_postings.release();
    ix.d_post_pos.release();
    ix.d_post_pos2.release();
    ix.d_term_bm.release();
  }
};

// Bitmaps for the dense lists of a finished CSR (d_dense_terms: n_terms uint32 of scratch, d_count: one counter).
void build_dense_bitmaps(Index& ix, uint64_t n_docs, uint32_t* d_dense_terms, unsigned long long* d_count,
                         cudaStream_t stream) {
  // ---- dense bitmaps
  ix.bm_words = (n_docs + 31) / 32;
  const double thr = ix.cfg.dense_threshold > 0.0 ? ix.cfg.dense_threshold : 1.0 / 128.0;
  uint64_t min_len = std::max<uint64_t>(1, static_cast<uint64_t>(thr * static_cast<double>(n_docs)));
  // a bitmap only pays for lists long enough that probing beats searching
  min_len = std::max<uint64_t>(min_len, 1024);
  const uint64_t max_bytes = ix.cfg.max_dense_bytes != 0 ? ix.cfg.max_dense_bytes : (8ULL << 30);
  unsigned long long n_dense = 0;
  const unsigned term_grid = static_cast<unsigned>((ix.n_terms + 255) / 256);
  if (ix.n_terms > 0) {
    for (;;) {
      MGX_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), stream));
      dense_count_kernel<<<term_grid, 256, 0, stream>>>(ix.d_term_off.p, ix.n_terms, min_len, d_count);
      MGX_LAUNCH_CHECK();
      MGX_CUDA(cudaMemcpyAsync(&n_dense, d_count, sizeof(n_dense), cudaMemcpyDeviceToHost, stream));
      MGX_CUDA(cudaStreamSynchronize(stream));
      if (n_dense * ix.bm_words * 4 <= max_bytes) {
        break;
      }
      min_len *= 2;
    }
  }
  ix.n_dense = n_dense;
  ix.dense_min_len = min_len;
  ix.d_bitmaps.reserve(ix.n_dense * ix.bm_words);
  if (ix.n_terms > 0) {
    MGX_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), stream));
    if (ix.n_dense > 0) {
      MGX_CUDA(cudaMemsetAsync(ix.d_bitmaps.p, 0, ix.d_bitmaps.bytes(), stream));
    }
    dense_assign_kernel<<<term_grid, 256, 0, stream>>>(ix.d_term_off.p, ix.n_terms, min_len, d_count, ix.d_term_bm.p,
                                                       d_dense_terms);
    MGX_LAUNCH_CHECK();
    if (ix.n_dense > 0) {
      const unsigned slices = static_cast<unsigned>(std::min<uint64_t>(64, (n_docs + 65535) / 65536 + 1));
      dense_fill_kernel<<<dim3(static_cast<unsigned>(ix.n_dense), slices), 256, 0, stream>>>(
          ix.d_term_off.p, ix.d_postings.p, d_dense_terms, ix.d_bitmaps.p, ix.bm_words);
      MGX_LAUNCH_CHECK();
    }
  }
}
}  // namespace

void build_index_device(Index& ix, const uint32_t* d_doc_ids_in, const uint8_t* d_text_in,
                        const uint64_t* d_text_off_in, uint64_t n_docs, uint64_t text_bytes, cudaStream_t stream) {
  PhaseTrace trace(stream);
  uint32_t first_id = 1;
  bool sequential_ids = true;
  struct Events {
    cudaEvent_t a = nullptr;
    cudaEvent_t b = nullptr;
    ~Events() {
      if (a != nullptr) {
        cudaEventDestroy(a);
      }
      if (b != nullptr) {
        cudaEventDestroy(b);
      }
    }
  } events;
  MGX_CUDA(cudaEventCreate(&events.a));
  MGX_CUDA(cudaEventCreate(&events.b));
  cudaEvent_t ev0 = events.a;
  cudaEvent_t ev1 = events.b;
  // The arrays of the previous index are released below, before the first step that can fail (bad ids, too many
  // n-gram occurrences, a CUDA error): whatever fails, the handle must not keep the OLD sizes next to released or
  // half-written arrays. A failed build therefore leaves an EMPTY index (every query answers empty, the next build
  // starts clean); the caller gets the error.
  ResetOnFailure reset_on_failure{ix};

  // ---- resident arena A: the device mirror of DocumentStore's normalised text + ids + lengths.
  // The text arena is padded so 16-byte tile loads never leave the allocation.
  ix.n_docs = n_docs;
  ix.text_bytes = text_bytes;
  ix.text_less = false;
  ix.drop_filter_columns();  // rows of the old corpus
  ix.d_doc_ids.release();
  ix.d_text.release();
  ix.d_text_off.release();
  ix.d_doc_len.release();
  ix.d_tile_first_doc.release();
  ix.d_term_keys.release();
  ix.d_wide_keys.release();
  ix.d_term_off.release();
  ix.d_postings.release();
  ix.d_post_pos.release();
  ix.d_post_pos2.release();
  ix.d_term_bm.release();
  // resident_b and the bitmaps keep their allocations (grow-only): a rebuild of a shard of similar size -- the
  // journal commits of the mutation path -- then pays no cudaMalloc / cudaFree for them
  ix.n_text_tiles = n_docs > 0 ? (text_bytes + kTextTileBytes - 1) / kTextTileBytes : 0;
  ix.resident_a.reserve(DevArena::padded(n_docs * 4 + 4) + DevArena::padded(text_bytes + 64) +
                        DevArena::padded((n_docs + 1) * 8) + DevArena::padded(n_docs * 4 + 4) +
                        DevArena::padded(ix.n_text_tiles * 4 + 4));
  ix.d_text.borrow(ix.resident_a.take<uint8_t>(text_bytes + 64), text_bytes + 64);
  ix.d_text_off.borrow(ix.resident_a.take<uint64_t>(n_docs + 1), n_docs + 1);
  ix.d_doc_ids.borrow(ix.resident_a.take<uint32_t>(n_docs), n_docs);
  ix.d_doc_len.borrow(ix.resident_a.take<uint32_t>(n_docs), n_docs);
  ix.d_tile_first_doc.borrow(ix.resident_a.take<uint32_t>(ix.n_text_tiles), ix.n_text_tiles);
  MGX_CUDA(cudaEventRecord(ev0, stream));
  MGX_CUDA(cudaMemcpyAsync(ix.d_doc_ids.p, d_doc_ids_in, n_docs * sizeof(uint32_t), cudaMemcpyDefault, stream));
  MGX_CUDA(cudaMemcpyAsync(ix.d_text.p, d_text_in, text_bytes, cudaMemcpyDefault, stream));
  MGX_CUDA(cudaMemsetAsync(ix.d_text.p + text_bytes, 0, 64, stream));
  MGX_CUDA(cudaMemcpyAsync(ix.d_text_off.p, d_text_off_in, (n_docs + 1) * sizeof(uint64_t), cudaMemcpyDefault, stream));

  if (ix.n_text_tiles > 0) {
    tile_first_doc_kernel<<<static_cast<unsigned>((ix.n_text_tiles + 255) / 256), 256, 0, stream>>>(
        ix.d_text_off.p, n_docs, ix.n_text_tiles, ix.d_tile_first_doc.p);
    MGX_LAUNCH_CHECK();
  }

  // ---- temporary arena T0: counting stage. Both build workspaces stay with the index (grow-only) until
  // mgx_index_trim / destroy: cudaFree of a multi-GB block was measured at 0.1-0.9 s on some boxes, and an index
  // that takes mutations rebuilds again and again
  DevArena& t0 = ix.build_arena0;
  const size_t count_scratch = tokenize_scratch_elems(n_docs, text_bytes);
  const size_t n_tok_tiles = tokenize_tile_count(n_docs, text_bytes);
  t0.reserve(DevArena::padded((n_tok_tiles + 1) * 8) + DevArena::padded(count_scratch * 8) + 512);
  uint64_t* d_slot_off = t0.take<uint64_t>(n_tok_tiles + 1);
  uint64_t* d_count_scratch = t0.take<uint64_t>(count_scratch);
  unsigned int* d_bad = t0.take<unsigned int>(2);
  uint64_t max_doc_bytes = 0;
  if (n_docs > 0) {
    MGX_CUDA(cudaMemsetAsync(d_bad, 0, 2 * sizeof(unsigned int), stream));
    sequential_check_kernel<<<static_cast<unsigned>((n_docs + 255) / 256), 256, 0, stream>>>(
        ix.d_doc_ids.p, ix.d_text_off.p, n_docs, d_bad);
    MGX_LAUNCH_CHECK();
    unsigned int bad2[2] = {0, 0};
    MGX_CUDA(cudaMemcpyAsync(bad2, d_bad, sizeof(bad2), cudaMemcpyDeviceToHost, stream));
    MGX_CUDA(cudaMemcpyAsync(&first_id, ix.d_doc_ids.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    MGX_CUDA(cudaStreamSynchronize(stream));
    const unsigned int bad = bad2[0];
    max_doc_bytes = bad2[1];
    if (bad & 2u) {
      set_last_error("doc_ids must be strictly ascending");
      throw CudaFailure{MGX_ERR_INVALID_ARGUMENT};
    }
    sequential_ids = bad == 0;
  }
  ix.first_id = first_id;
  ix.sequential_ids = sequential_ids;
  trace.mark("resident copies + id check");

  uint64_t n_slots = 0;
  uint64_t counters[3] = {0, 0, 0};
  tokenize_bound(ix.d_text.p, ix.d_text_off.p, n_docs, text_bytes, d_slot_off, d_count_scratch, &n_slots, stream);
  trace.mark("tokenize: slot bound + scan");
  ix.n_pair_slots = n_slots;
  if (n_slots >= (1ULL << 32)) {
    set_last_error("shard too large: more than 2^32 n-gram occurrences in one shard; split by doc-id range");
    throw CudaFailure{MGX_ERR_UNSUPPORTED};
  }
  const int pb = pos_bits_for_width(ix.width);
  ix.has_positions = pb > 0;
  // split of the payload bits: offset field sized by the longest document, the rest neighbour signatures
  // (MGX_BUILD_NO_SIG: offsets only, the round-1 payload, for A/B runs)
  ix.sig = sig_layout_for(pb, max_doc_bytes);
  if (std::getenv("MGX_BUILD_NO_SIG") != nullptr) {
    ix.sig = SigLayout{};
  }
  const int sig_cfg = (ix.sig.next_bits | ix.sig.prev_bits) != 0
                          ? (ix.sig.pos_bits | (ix.sig.next_bits << 8) | (ix.sig.prev_bits << 16))
                          : 0;

  // ---- temporary arena T1: pairs (double-buffered), sort scratch, CSR block arrays
  const int W = wide_words_for(ix.width);
  ix.wide_words = W;
  const uint64_t n_blocks = (n_slots + kCsrTile - 1) / kCsrTile;
  DevArena& t1 = ix.build_arena;
  const size_t wide_extra =
      W > 0 ? DevArena::padded(static_cast<size_t>(W) * n_slots * 8 + 8) + DevArena::padded(n_slots * 4 + 4) +
                  DevArena::padded(n_slots * 4 + 4) + DevArena::padded((n_slots + 1) * 8) +
                  DevArena::padded(scan_scratch_elems(std::max<uint64_t>(n_slots, 1)) * 8 + 64)
            : 0;
  t1.reserve(2 * (DevArena::padded(n_slots * 8 + 8) + DevArena::padded(n_slots * 4 + 4)) + wide_extra +
             DevArena::padded(radix_sort_scratch_bytes(n_slots) + 512) + 2 * DevArena::padded((n_blocks + 2) * 8) + 1024);
  uint64_t* d_keys_a = t1.take<uint64_t>(n_slots);
  uint32_t* d_docs_a = t1.take<uint32_t>(n_slots);
  uint64_t* d_keys_b = t1.take<uint64_t>(n_slots);
  uint32_t* d_docs_b = t1.take<uint32_t>(n_slots);
  uint8_t* d_sort_scratch = t1.take<uint8_t>(radix_sort_scratch_bytes(n_slots) + 256);
  uint64_t* d_block_pairs = t1.take<uint64_t>(n_blocks + 2);
  uint64_t* d_block_terms = t1.take<uint64_t>(n_blocks + 2);
  uint64_t* d_totals = t1.take<uint64_t>(4);
  // wide keys only: the tokenizer's word arrays, the documents in text order, head flags and ranks
  uint64_t* d_words = W > 0 ? t1.take<uint64_t>(static_cast<size_t>(W) * n_slots) : nullptr;
  uint32_t* d_docs_text = W > 0 ? t1.take<uint32_t>(n_slots) : nullptr;
  uint32_t* d_flags = W > 0 ? t1.take<uint32_t>(n_slots) : nullptr;
  uint64_t* d_rank = W > 0 ? t1.take<uint64_t>(n_slots + 1) : nullptr;
  uint64_t* d_wide_scan = W > 0 ? t1.take<uint64_t>(scan_scratch_elems(std::max<uint64_t>(n_slots, 1)) + 8) : nullptr;
  trace.mark("alloc pair arena");
  // one pass: per-document totals + the (key, doc) pairs (placeholder key 0 in the unused slots of each tile)
  tokenize_fused(ix.ngram, ix.kanji, ix.cross, ix.width, ix.d_text.p, ix.d_text_off.p, n_docs, text_bytes, ix.d_doc_len.p,
                 d_slot_off, d_count_scratch, W > 0 ? d_words : d_keys_a, W > 0 ? d_docs_text : d_docs_a, pb, counters,
                 stream, n_slots, sig_cfg);
  ix.doc_count = counters[0];
  ix.all_valid_utf8 = counters[1] == 0;
  ix.total_doc_length = counters[2];
  trace.mark("tokenize: fused stats + emit");
  SortResult sorted{d_keys_a, d_docs_a};
  SortResult wide_order{nullptr, nullptr};  // wide keys: word 0 of every slot in key order + the slot it came from
  const unsigned wide_grid = static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>((n_slots + 255) / 256, 148 * 16)));
  if (W == 0) {
    sorted = radix_sort_pairs(d_keys_a, d_docs_a, d_keys_b, d_docs_b, n_slots, 21 * ix.width, pb, d_sort_scratch, stream);
  } else if (n_slots > 0) {
    // least significant word first, each sort stable, the permutation carried as the value: afterwards the slots are
    // in wide-key order and, within one n-gram, in text order (= ascending documents)
    SortResult cur{d_keys_a, d_docs_a};  // d_docs_* hold slot numbers until wide_assign_kernel
    SortResult alt{d_keys_b, d_docs_b};
    iota_u32_kernel<<<wide_grid, 256, 0, stream>>>(cur.vals, n_slots);
    MGX_LAUNCH_CHECK();
    for (int w = W - 1; w >= 0; --w) {
      const uint64_t* src = d_words + static_cast<size_t>(w) * n_slots;
      if (w == W - 1) {
        MGX_CUDA(cudaMemcpyAsync(cur.keys, src, n_slots * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
      } else {
        gather_u64_kernel<<<wide_grid, 256, 0, stream>>>(src, cur.vals, cur.keys, n_slots);
        MGX_LAUNCH_CHECK();
      }
      const SortResult r = radix_sort_pairs(cur.keys, cur.vals, alt.keys, alt.vals, n_slots, 63, 0, d_sort_scratch, stream);
      if (r.keys != cur.keys) {
        std::swap(cur, alt);
      }
      // cur = sorted by words w.. ; the next word is gathered over cur.keys (its content is no longer needed)
    }
    wide_order = cur;
    wide_heads_kernel<<<wide_grid, 256, 0, stream>>>(cur.keys, cur.vals, d_words, n_slots, W, n_slots, d_flags);
    MGX_LAUNCH_CHECK();
    exclusive_scan_u32_u64(d_flags, d_rank, n_slots, d_wide_scan, stream);
    wide_assign_kernel<<<wide_grid, 256, 0, stream>>>(cur.keys, cur.vals, d_flags, d_rank, d_docs_text, n_slots,
                                                     alt.keys, alt.vals);
    MGX_LAUNCH_CHECK();
    sorted = alt;
  }
  trace.mark("radix sort");

  // ---- segmented unique + compaction into CSR
  if (n_blocks > 0) {
    csr_count_kernel<<<static_cast<unsigned>(n_blocks), kCsrThreads, 0, stream>>>(sorted.keys, sorted.vals, n_slots, pb,
                                                                                   d_block_pairs, d_block_terms);
    MGX_LAUNCH_CHECK();
  }
  csr_scan_kernel<<<1, 1024, 0, stream>>>(d_block_pairs, d_block_terms, n_blocks, d_totals);
  MGX_LAUNCH_CHECK();
  uint64_t totals[2] = {0, 0};
  MGX_CUDA(cudaMemcpyAsync(totals, d_totals, sizeof(totals), cudaMemcpyDeviceToHost, stream));
  MGX_CUDA(cudaStreamSynchronize(stream));
  ix.n_postings = totals[0];
  ix.n_terms = totals[1];
  ix.resident_b.reserve(DevArena::padded(ix.n_terms * 8 + 8) + DevArena::padded((ix.n_terms + 1) * 8) +
                        DevArena::padded(ix.n_postings * 4 + 4) + 2 * DevArena::padded(ix.n_terms * 4 + 4) +
                        2 * DevArena::padded(ix.n_postings * 4 + 4) +
                        DevArena::padded(static_cast<size_t>(W) * ix.n_terms * 8 + 8) + 512);
  ix.d_term_keys.borrow(ix.resident_b.take<uint64_t>(ix.n_terms), ix.n_terms);
  if (W > 0) {
    ix.d_wide_keys.borrow(ix.resident_b.take<uint64_t>(static_cast<size_t>(W) * ix.n_terms),
                          static_cast<size_t>(W) * ix.n_terms);
    if (ix.n_terms > 0) {
      wide_table_kernel<<<wide_grid, 256, 0, stream>>>(wide_order.keys, wide_order.vals, d_flags, d_rank, d_words,
                                                      n_slots, W, n_slots, ix.d_wide_keys.p);
      MGX_LAUNCH_CHECK();
    }
  }
  ix.d_term_off.borrow(ix.resident_b.take<uint64_t>(ix.n_terms + 1), ix.n_terms + 1);
  ix.d_postings.borrow(ix.resident_b.take<uint32_t>(ix.n_postings), ix.n_postings);
  ix.d_term_bm.borrow(ix.resident_b.take<int32_t>(ix.n_terms), ix.n_terms);
  ix.d_post_pos.borrow(ix.resident_b.take<uint32_t>(ix.n_postings), ix.n_postings);
  ix.d_post_pos2.borrow(ix.resident_b.take<uint32_t>(ix.n_postings), ix.n_postings);
  uint32_t* d_dense_terms = ix.resident_b.take<uint32_t>(ix.n_terms);  // only the first n_dense entries are used
  unsigned long long* d_count = reinterpret_cast<unsigned long long*>(ix.resident_b.take<uint64_t>(2));
  if (n_blocks > 0) {
    csr_write_kernel<<<static_cast<unsigned>(n_blocks), kCsrThreads, 0, stream>>>(
        sorted.keys, sorted.vals, n_slots, pb, d_block_pairs, d_block_terms, ix.d_term_keys.p, ix.d_term_off.p,
        ix.d_postings.p, ix.d_post_pos.p, ix.d_post_pos2.p, sig_cfg);
    MGX_LAUNCH_CHECK();
  }
  set_u64_kernel<<<1, 1, 0, stream>>>(ix.d_term_off.p + ix.n_terms, ix.n_postings);
  MGX_LAUNCH_CHECK();
  trace.mark("csr");

  build_dense_bitmaps(ix, n_docs, d_dense_terms, d_count, stream);
  MGX_CUDA(cudaEventRecord(ev1, stream));
  MGX_CUDA(cudaEventSynchronize(ev1));
  trace.mark("dense bitmaps + free");
  float ms = 0.f;
  MGX_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  ix.last_build_ms = ms;
  reset_on_failure.armed = false;
}

// ------------------------------------------------------------------ MGIX stream -> device index
namespace {
__global__ void mark_docs_kernel(const uint32_t* __restrict__ postings, uint64_t n, uint32_t* __restrict__ bits) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint32_t d = postings[i];
    atomicOr(bits + (d >> 5), 1u << (d & 31));
  }
}
__global__ void word_popc_kernel(const uint32_t* __restrict__ bits, uint64_t n_words, uint32_t* __restrict__ cnt) {
  const uint64_t w = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w < n_words) {
    cnt[w] = __popc(bits[w]);
  }
}
__global__ void ids_from_bits_kernel(const uint32_t* __restrict__ bits, const uint64_t* __restrict__ rank,
                                     uint64_t n_words, uint32_t* __restrict__ ids) {
  const uint64_t w = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= n_words) {
    return;
  }
  uint32_t word = bits[w];
  uint64_t at = rank[w];
  while (word != 0) {
    const int b = __ffs(static_cast<int>(word)) - 1;
    word &= word - 1;
    ids[at++] = static_cast<uint32_t>(w * 32 + b);
  }
}
// global doc id -> index in the ascending list of the ids that occur in the stream
__global__ void localise_postings_kernel(const uint32_t* __restrict__ in, uint64_t n, const uint32_t* __restrict__ bits,
                                         const uint64_t* __restrict__ rank, uint32_t* __restrict__ out) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint32_t d = in[i];
    out[i] = static_cast<uint32_t>(rank[d >> 5]) + __popc(bits[d >> 5] & ((1u << (d & 31)) - 1u));
  }
}
}  // namespace

void load_index_device(Index& ix, const uint64_t* h_keys, const uint64_t* h_term_off, const uint32_t* h_postings,
                       uint64_t n_terms, uint64_t n_postings, cudaStream_t stream) {
  // wide keys (width > 3): h_keys holds wide_words words per term, ascending
  const int W = wide_words_for(ix.width);
  ix.wide_words = W;
  ResetOnFailure reset_on_failure{ix};
  ix.drop_filter_columns();
  ix.d_doc_ids.release();
  ix.d_text.release();
  ix.d_text_off.release();
  ix.d_doc_len.release();
  ix.d_tile_first_doc.release();
  ix.d_term_keys.release();
  ix.d_wide_keys.release();
  ix.d_term_off.release();
  ix.d_postings.release();
  ix.d_post_pos.release();
  ix.d_post_pos2.release();
  ix.d_term_bm.release();
  ix.n_terms = n_terms;
  ix.n_postings = n_postings;
  ix.has_positions = false;
  ix.n_pair_slots = 0;
  ix.text_bytes = 0;
  ix.n_text_tiles = 0;

  // ---- the documents the stream knows = the union of its lists, through a bitmap over [0, max id]
  uint32_t max_id = 0;
  for (uint64_t i = 0; i < n_postings; ++i) {
    max_id = std::max(max_id, h_postings[i]);
  }
  const uint64_t n_words = n_postings > 0 ? static_cast<uint64_t>(max_id) / 32 + 1 : 0;
  DevArena& t1 = ix.build_arena;
  t1.reserve(DevArena::padded(n_postings * 4 + 4) + DevArena::padded(n_words * 4 + 4) + DevArena::padded(n_words * 4 + 4) +
             DevArena::padded((n_words + 1) * 8) + DevArena::padded(scan_scratch_elems(n_words) * 8 + 8) + 1024);
  uint32_t* d_global = t1.take<uint32_t>(n_postings);
  uint32_t* d_bits = t1.take<uint32_t>(n_words);
  uint32_t* d_cnt = t1.take<uint32_t>(n_words);
  uint64_t* d_rank = t1.take<uint64_t>(n_words + 1);
  uint64_t* d_scan = t1.take<uint64_t>(scan_scratch_elems(n_words) + 1);
  uint64_t n_docs = 0;
  if (n_postings > 0) {
    MGX_CUDA(cudaMemcpyAsync(d_global, h_postings, n_postings * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    MGX_CUDA(cudaMemsetAsync(d_bits, 0, n_words * sizeof(uint32_t), stream));
    mark_docs_kernel<<<static_cast<unsigned>((n_postings + 255) / 256), 256, 0, stream>>>(d_global, n_postings, d_bits);
    MGX_LAUNCH_CHECK();
    word_popc_kernel<<<static_cast<unsigned>((n_words + 255) / 256), 256, 0, stream>>>(d_bits, n_words, d_cnt);
    MGX_LAUNCH_CHECK();
    exclusive_scan_u32_u64(d_cnt, d_rank, n_words, d_scan, stream);
    MGX_CUDA(cudaMemcpyAsync(&n_docs, d_rank + n_words, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    MGX_CUDA(cudaStreamSynchronize(stream));
  }
  ix.n_docs = n_docs;

  // ---- resident arena A: ids, and the (empty) text mirror every kernel expects
  ix.resident_a.reserve(DevArena::padded(n_docs * 4 + 4) + DevArena::padded(64) + DevArena::padded((n_docs + 1) * 8) +
                        DevArena::padded(n_docs * 4 + 4) + DevArena::padded(4));
  ix.d_text.borrow(ix.resident_a.take<uint8_t>(64), 64);
  ix.d_text_off.borrow(ix.resident_a.take<uint64_t>(n_docs + 1), n_docs + 1);
  ix.d_doc_ids.borrow(ix.resident_a.take<uint32_t>(n_docs), n_docs);
  ix.d_doc_len.borrow(ix.resident_a.take<uint32_t>(n_docs), n_docs);
  ix.d_tile_first_doc.borrow(ix.resident_a.take<uint32_t>(1), 0);
  MGX_CUDA(cudaMemsetAsync(ix.d_text.p, 0, 64, stream));
  MGX_CUDA(cudaMemsetAsync(ix.d_text_off.p, 0, (n_docs + 1) * sizeof(uint64_t), stream));
  if (n_docs > 0) {
    MGX_CUDA(cudaMemsetAsync(ix.d_doc_len.p, 0, n_docs * sizeof(uint32_t), stream));
    ids_from_bits_kernel<<<static_cast<unsigned>((n_words + 255) / 256), 256, 0, stream>>>(d_bits, d_rank, n_words,
                                                                                            ix.d_doc_ids.p);
    MGX_LAUNCH_CHECK();
  }
  uint32_t first_id = 1;
  bool sequential_ids = true;
  if (n_docs > 0) {
    MGX_CUDA(cudaMemcpyAsync(&first_id, ix.d_doc_ids.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    MGX_CUDA(cudaStreamSynchronize(stream));
    sequential_ids = static_cast<uint64_t>(max_id) - first_id + 1 == n_docs;
  }
  ix.first_id = first_id;
  ix.sequential_ids = sequential_ids;
  // BM25Stats belong to the DocumentStore (server_orchestrator.cpp:758-772): a stream carries none
  ix.doc_count = 0;
  ix.total_doc_length = 0;
  ix.all_valid_utf8 = true;

  // ---- resident arena B: dictionary + CSR with LOCAL doc indices
  ix.resident_b.reserve(DevArena::padded(n_terms * 8 + 8) + DevArena::padded((n_terms + 1) * 8) +
                        DevArena::padded(n_postings * 4 + 4) + 2 * DevArena::padded(n_terms * 4 + 4) +
                        DevArena::padded(static_cast<size_t>(W) * n_terms * 8 + 8) + 512);
  ix.d_term_keys.borrow(ix.resident_b.take<uint64_t>(n_terms), n_terms);
  if (W > 0) {
    ix.d_wide_keys.borrow(ix.resident_b.take<uint64_t>(static_cast<size_t>(W) * n_terms), static_cast<size_t>(W) * n_terms);
  }
  ix.d_term_off.borrow(ix.resident_b.take<uint64_t>(n_terms + 1), n_terms + 1);
  ix.d_postings.borrow(ix.resident_b.take<uint32_t>(n_postings), n_postings);
  ix.d_term_bm.borrow(ix.resident_b.take<int32_t>(n_terms), n_terms);
  uint32_t* d_dense_terms = ix.resident_b.take<uint32_t>(n_terms);
  unsigned long long* d_count = reinterpret_cast<unsigned long long*>(ix.resident_b.take<uint64_t>(2));
  if (n_terms > 0 && W == 0) {
    MGX_CUDA(cudaMemcpyAsync(ix.d_term_keys.p, h_keys, n_terms * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
  } else if (n_terms > 0) {
    MGX_CUDA(cudaMemcpyAsync(ix.d_wide_keys.p, h_keys, static_cast<size_t>(W) * n_terms * sizeof(uint64_t),
                             cudaMemcpyHostToDevice, stream));
    iota_keys_kernel<<<static_cast<unsigned>(std::min<uint64_t>((n_terms + 255) / 256, 148 * 16)), 256, 0, stream>>>(
        ix.d_term_keys.p, n_terms);
    MGX_LAUNCH_CHECK();
  }
  MGX_CUDA(cudaMemcpyAsync(ix.d_term_off.p, h_term_off, (n_terms + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
  if (n_postings > 0) {
    localise_postings_kernel<<<static_cast<unsigned>((n_postings + 255) / 256), 256, 0, stream>>>(
        d_global, n_postings, d_bits, d_rank, ix.d_postings.p);
    MGX_LAUNCH_CHECK();
  }
  build_dense_bitmaps(ix, n_docs, d_dense_terms, d_count, stream);
  MGX_CUDA(cudaStreamSynchronize(stream));
  ix.text_less = true;
  ix.last_build_ms = 0.0;
  reset_on_failure.armed = false;
}

// ------------------------------------------------------------------ reference-style representation counters
namespace {
__global__ void count_roaring_kernel(const uint64_t* __restrict__ term_off, uint64_t n_terms, double min_density_len,
                                     unsigned long long* __restrict__ count) {
  const uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  bool roaring = false;
  if (t < n_terms) {
    const uint64_t size = term_off[t + 1] - term_off[t];
    // > 4096 entries: Roaring since insertion (posting_list.cpp:21,917-922); otherwise only after Optimize, by density
    roaring = size > 4096 || (min_density_len > 0.0 && static_cast<double>(size) >= min_density_len);
  }
  const unsigned n = __popc(__ballot_sync(0xffffffffu, roaring));
  if ((threadIdx.x & 31u) == 0 && n != 0) {
    atomicAdd(count, static_cast<unsigned long long>(n));
  }
}
}  // namespace

uint64_t count_roaring_lists(const Index& ix, double roaring_threshold, uint64_t optimized_total_docs,
                             cudaStream_t stream) {
  if (ix.n_terms == 0) {
    return 0;
  }
  DevBuf<unsigned long long> d_count;
  d_count.alloc(1);
  MGX_CUDA(cudaMemsetAsync(d_count.p, 0, sizeof(unsigned long long), stream));
  // density = size / total_docs >= threshold  <=>  size >= threshold * total_docs (both sides exact in double here)
  const double min_len = optimized_total_docs > 0 ? roaring_threshold * static_cast<double>(optimized_total_docs) : 0.0;
  count_roaring_kernel<<<static_cast<unsigned>((ix.n_terms + 255) / 256), 256, 0, stream>>>(ix.d_term_off.p, ix.n_terms,
                                                                                            min_len, d_count.p);
  MGX_LAUNCH_CHECK();
  unsigned long long h = 0;
  MGX_CUDA(cudaMemcpyAsync(&h, d_count.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
  MGX_CUDA(cudaStreamSynchronize(stream));
  return h;
}

// ------------------------------------------------------------------ incremental mutations (journal -> rebuild)
// Index::AddDocument / UpdateDocument / RemoveDocument (index.cpp:39-197) are journaled on the host and folded in
// here before the next read: the resident corpus (doc ids, text) is merged ON THE DEVICE with the journal's
// documents (sorted by id; an id in the journal replaces or removes the resident document of that id), and the
// shard is rebuilt from the merged corpus by the same tokenise / sort / CSR pipeline as a bulk build.
namespace {

__device__ __forceinline__ uint64_t lower_bound_ids(const uint32_t* __restrict__ p, uint64_t n, uint32_t v) {
  uint64_t lo = 0;
  uint64_t hi = n;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (p[mid] < v) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return lo;
}

// keep[i] = 1 when resident document i is untouched by the journal
__global__ void journal_mark_kernel(const uint32_t* __restrict__ old_ids, uint64_t n_old,
                                    const uint32_t* __restrict__ j_ids, uint64_t n_j, uint32_t* __restrict__ keep) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_old) {
    return;
  }
  const uint64_t pos = lower_bound_ids(j_ids, n_j, old_ids[i]);
  keep[i] = (pos < n_j && j_ids[pos] == old_ids[i]) ? 0u : 1u;
}

__global__ void journal_live_kernel(const uint8_t* __restrict__ removed, uint64_t n_j, uint32_t* __restrict__ live) {
  const uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j < n_j) {
    live[j] = removed[j] != 0 ? 0u : 1u;
  }
}

// position of every surviving document in the merged corpus (two sorted runs, ranks by binary search)
__global__ void journal_place_old_kernel(const uint32_t* __restrict__ old_ids, const uint64_t* __restrict__ old_off,
                                         uint64_t n_old, const uint32_t* __restrict__ keep,
                                         const uint64_t* __restrict__ keep_rank, const uint32_t* __restrict__ j_ids,
                                         uint64_t n_j, const uint64_t* __restrict__ live_rank,
                                         uint32_t* __restrict__ new_ids, uint32_t* __restrict__ new_len,
                                         uint32_t* __restrict__ src) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_old || keep[i] == 0) {
    return;
  }
  const uint64_t dest = keep_rank[i] + live_rank[lower_bound_ids(j_ids, n_j, old_ids[i])];
  new_ids[dest] = old_ids[i];
  new_len[dest] = static_cast<uint32_t>(old_off[i + 1] - old_off[i]);
  src[dest] = static_cast<uint32_t>(i);
}

__global__ void journal_place_new_kernel(const uint32_t* __restrict__ j_ids, const uint64_t* __restrict__ j_off,
                                         uint64_t n_j, const uint32_t* __restrict__ live,
                                         const uint64_t* __restrict__ live_rank, const uint32_t* __restrict__ old_ids,
                                         uint64_t n_old, const uint64_t* __restrict__ keep_rank,
                                         uint32_t* __restrict__ new_ids, uint32_t* __restrict__ new_len,
                                         uint32_t* __restrict__ src) {
  const uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n_j || live[j] == 0) {
    return;
  }
  // resident documents with a smaller id that survive (the resident document of the SAME id never survives)
  const uint64_t dest = live_rank[j] + keep_rank[lower_bound_ids(old_ids, n_old, j_ids[j])];
  new_ids[dest] = j_ids[j];
  new_len[dest] = static_cast<uint32_t>(j_off[j + 1] - j_off[j]);
  src[dest] = static_cast<uint32_t>(j) | 0x80000000u;
}

// filter columns follow the documents: a surviving document keeps its row, a document from the journal is NULL
__global__ void journal_remap_column_kernel(const uint32_t* __restrict__ src, uint64_t n_new,
                                            const uint64_t* __restrict__ old_values, const uint8_t* __restrict__ old_nulls,
                                            uint64_t* __restrict__ new_values, uint8_t* __restrict__ new_nulls) {
  const uint64_t d = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (d >= n_new) {
    return;
  }
  const uint32_t s = src[d];
  const bool from_journal = (s & 0x80000000u) != 0;
  new_values[d] = from_journal ? 0ULL : old_values[s];
  new_nulls[d] = from_journal ? static_cast<uint8_t>(1) : old_nulls[s];
}

// one warp per merged document: bytes from the resident arena or from the journal arena
__global__ void __launch_bounds__(256) journal_copy_kernel(const uint32_t* __restrict__ src, uint64_t n_new,
                                                           const uint64_t* __restrict__ new_off,
                                                           const uint8_t* __restrict__ old_text,
                                                           const uint64_t* __restrict__ old_off,
                                                           const uint8_t* __restrict__ j_text,
                                                           const uint64_t* __restrict__ j_off,
                                                           uint8_t* __restrict__ out) {
  const uint64_t d = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (d >= n_new) {
    return;
  }
  const uint32_t s = src[d];
  const uint8_t* from = (s & 0x80000000u) != 0 ? j_text + j_off[s & 0x7FFFFFFFu] : old_text + old_off[s];
  const uint64_t n = new_off[d + 1] - new_off[d];
  uint8_t* to = out + new_off[d];
  for (uint64_t i = threadIdx.x & 31u; i < n; i += 32) {
    to[i] = from[i];
  }
}

}  // namespace

void copy_index_config(Index& to, const Index& from) {
  to.cfg = from.cfg;
  to.ngram = from.ngram;
  to.kanji = from.kanji;
  to.cross = from.cross;
  to.width = from.width;
  to.device = from.device;
  to.wide_words = from.wide_words;
}

void swap_generation(Index& a, Index& b) {
  using std::swap;
  swap(a.n_docs, b.n_docs);
  swap(a.sequential_ids, b.sequential_ids);
  swap(a.first_id, b.first_id);
  swap(a.d_doc_ids, b.d_doc_ids);
  swap(a.d_text, b.d_text);
  swap(a.d_text_off, b.d_text_off);
  swap(a.d_doc_len, b.d_doc_len);
  swap(a.text_bytes, b.text_bytes);
  swap(a.d_tile_first_doc, b.d_tile_first_doc);
  swap(a.n_text_tiles, b.n_text_tiles);
  swap(a.n_terms, b.n_terms);
  swap(a.n_postings, b.n_postings);
  swap(a.d_term_keys, b.d_term_keys);
  swap(a.wide_words, b.wide_words);
  swap(a.d_wide_keys, b.d_wide_keys);
  swap(a.d_term_off, b.d_term_off);
  swap(a.d_postings, b.d_postings);
  swap(a.d_post_pos, b.d_post_pos);
  swap(a.d_post_pos2, b.d_post_pos2);
  swap(a.has_positions, b.has_positions);
  swap(a.sig, b.sig);
  swap(a.text_less, b.text_less);
  swap(a.d_term_bm, b.d_term_bm);
  swap(a.d_bitmaps, b.d_bitmaps);
  swap(a.resident_a, b.resident_a);
  swap(a.resident_b, b.resident_b);
  swap(a.n_dense, b.n_dense);
  swap(a.bm_words, b.bm_words);
  swap(a.dense_min_len, b.dense_min_len);
  swap(a.total_doc_length, b.total_doc_length);
  swap(a.doc_count, b.doc_count);
  swap(a.all_valid_utf8, b.all_valid_utf8);
  swap(a.n_pair_slots, b.n_pair_slots);
  swap(a.last_build_ms, b.last_build_ms);
  swap(a.columns, b.columns);
}

void apply_journal_device(const Index& ix, Index& next, const uint32_t* h_ids, const uint8_t* h_removed,
                          const uint8_t* h_text, const uint64_t* h_off, uint64_t n_j, cudaStream_t stream) {
  const uint64_t n_old = ix.n_docs;
  const uint64_t j_bytes = h_off[n_j];
  DevBuf<uint32_t> d_jids, d_keep, d_live, d_new_ids, d_new_len, d_src;
  DevBuf<uint8_t> d_jrem, d_jtext, d_new_text;
  DevBuf<uint64_t> d_joff, d_keep_rank, d_live_rank, d_new_off, d_scan;
  d_jids.alloc(n_j);
  d_jrem.alloc(n_j);
  d_jtext.alloc(j_bytes + 64);
  d_joff.alloc(n_j + 1);
  d_keep.alloc(n_old + 1);
  d_live.alloc(n_j + 1);
  d_keep_rank.alloc(n_old + 1);
  d_live_rank.alloc(n_j + 1);
  d_scan.alloc(scan_scratch_elems(std::max<uint64_t>(n_old + n_j, 1)) + 8);
  MGX_CUDA(cudaMemcpyAsync(d_jids.p, h_ids, n_j * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  MGX_CUDA(cudaMemcpyAsync(d_jrem.p, h_removed, n_j, cudaMemcpyHostToDevice, stream));
  if (j_bytes > 0) {
    MGX_CUDA(cudaMemcpyAsync(d_jtext.p, h_text, j_bytes, cudaMemcpyHostToDevice, stream));
  }
  MGX_CUDA(cudaMemcpyAsync(d_joff.p, h_off, (n_j + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
  const unsigned g_old = static_cast<unsigned>((n_old + 255) / 256);
  const unsigned g_j = static_cast<unsigned>((n_j + 255) / 256);
  if (n_old > 0) {
    journal_mark_kernel<<<g_old, 256, 0, stream>>>(ix.d_doc_ids.p, n_old, d_jids.p, n_j, d_keep.p);
    MGX_LAUNCH_CHECK();
  }
  journal_live_kernel<<<g_j, 256, 0, stream>>>(d_jrem.p, n_j, d_live.p);
  MGX_LAUNCH_CHECK();
  exclusive_scan_u32_u64(d_keep.p, d_keep_rank.p, n_old, d_scan.p, stream);
  exclusive_scan_u32_u64(d_live.p, d_live_rank.p, n_j, d_scan.p, stream);
  uint64_t n_keep = 0;
  uint64_t n_live = 0;
  MGX_CUDA(cudaMemcpyAsync(&n_keep, d_keep_rank.p + n_old, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  MGX_CUDA(cudaMemcpyAsync(&n_live, d_live_rank.p + n_j, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  MGX_CUDA(cudaStreamSynchronize(stream));
  const uint64_t n_new = n_keep + n_live;
  d_new_ids.alloc(n_new + 1);
  d_new_len.alloc(n_new + 1);
  d_src.alloc(n_new + 1);
  d_new_off.alloc(n_new + 1);
  if (n_old > 0) {
    journal_place_old_kernel<<<g_old, 256, 0, stream>>>(ix.d_doc_ids.p, ix.d_text_off.p, n_old, d_keep.p, d_keep_rank.p,
                                                        d_jids.p, n_j, d_live_rank.p, d_new_ids.p, d_new_len.p, d_src.p);
    MGX_LAUNCH_CHECK();
  }
  journal_place_new_kernel<<<g_j, 256, 0, stream>>>(d_jids.p, d_joff.p, n_j, d_live.p, d_live_rank.p, ix.d_doc_ids.p,
                                                    n_old, d_keep_rank.p, d_new_ids.p, d_new_len.p, d_src.p);
  MGX_LAUNCH_CHECK();
  exclusive_scan_u32_u64(d_new_len.p, d_new_off.p, n_new, d_scan.p, stream);
  uint64_t new_bytes = 0;
  MGX_CUDA(cudaMemcpyAsync(&new_bytes, d_new_off.p + n_new, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  MGX_CUDA(cudaStreamSynchronize(stream));
  d_new_text.alloc(new_bytes + 64);
  if (n_new > 0) {
    journal_copy_kernel<<<static_cast<unsigned>((n_new * 32 + 255) / 256), 256, 0, stream>>>(
        d_src.p, n_new, d_new_off.p, ix.d_text.p, ix.d_text_off.p, d_jtext.p, d_joff.p, d_new_text.p);
    MGX_LAUNCH_CHECK();
  }
  // the device mirror of the filter columns (mgx_index_set_filter_column) follows the documents through the rebuild
  std::vector<FilterColumn*> carried(kMaxFilterColumns, nullptr);
  for (uint32_t c = 0; c < kMaxFilterColumns; ++c) {
    FilterColumn* old_col = ix.columns[c];
    if (old_col == nullptr || old_col->n_docs != n_old) {
      continue;
    }
    FilterColumn* col = new FilterColumn();
    carried[c] = col;
    col->cls = old_col->cls;
    col->n_docs = n_new;
    col->dict = old_col->dict;
    col->values.alloc(n_new);
    col->nulls.alloc(n_new);
    if (n_new > 0) {
      journal_remap_column_kernel<<<static_cast<unsigned>((n_new + 255) / 256), 256, 0, stream>>>(
          d_src.p, n_new, old_col->values.p, old_col->nulls.p, col->values.p, col->nulls.p);
      MGX_LAUNCH_CHECK();
    }
  }
  MGX_CUDA(cudaStreamSynchronize(stream));
  try {
    build_index_device(next, d_new_ids.p, d_new_text.p, d_new_off.p, n_new, new_bytes, stream);  // drops next's columns
  } catch (...) {
    for (FilterColumn* c : carried) {
      delete c;
    }
    throw;
  }
  for (uint32_t c = 0; c < kMaxFilterColumns; ++c) {
    next.columns[c] = carried[c];
  }
}

}  // namespace mgx
