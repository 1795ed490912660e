// build.cu — index construction on the device.
//
// Replaces Index::AddDocumentBatch (reference src/index/index.cpp:76-119) for a
// whole shard: GenerateHybridNgrams (utils/string_utils.cpp:452-509) per
// document, per-document sort+unique (index.cpp:88-91), term -> ascending doc
// ids (index.cpp:104-118, posting_list.cpp:291-324).
//
//   K1 tokenize<COUNT>  text -> code-point count per doc (= CountCodePoints,
//                       string_utils.cpp:655-669) ; 128-bit loads, warp per doc
//      scan             slot offsets (one pair slot per code point)
//   K2 tokenize<EMIT>   text -> (packed n-gram key, doc) pairs, decoded through a
//                       per-warp shared-memory window; non-emitting positions
//                       get kInvalidKey so no second count pass is needed
//   K3 radix sort       stable by key => docs ascending inside a key (primitives.cu)
//   K4 csr              segmented unique (drops duplicate (key, doc) = the
//                       per-document unique) + compaction into CSR
//   K5 bitmaps          doc bitmaps for lists with density >= dense_threshold
#include <algorithm>
#include <chrono>
#include <cstdlib>

#include "mgx_internal.cuh"

namespace mgx {

namespace {

constexpr int kTokThreads = 256;
constexpr int kTokWarps = kTokThreads / 32;
constexpr int kTokTileBytes = 512;                 // bytes per warp iteration (32 lanes x 16 B)
constexpr int kTokBuf = kTokTileBytes + 4;         // code points per warp window (+ carry)

__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[5], int j) {
  return (w[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
}

__device__ __forceinline__ uint4 ld_stream_16(const uint8_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// One warp walks one document in 512-byte tiles. Strict UTF-8 decoding is a
// LOCAL predicate per byte: every non-continuation byte is visited by the
// reference's "skip one byte and retry" scan (string_utils.cpp:206-216), because
// a valid multi-byte character only ever covers continuation bytes; so a code
// point starts at byte i iff TryParseUtf8Char(i) succeeds with the bytes that
// remain in the document.
template <bool EMIT>
__global__ void __launch_bounds__(kTokThreads)
tokenize_kernel(const uint8_t* __restrict__ text, const uint64_t* __restrict__ text_off, uint64_t n_docs, int ngram,
                int kanji, int cross, int width, uint32_t* __restrict__ doc_len,
                const uint64_t* __restrict__ slot_off, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ docs_out,
                unsigned long long* __restrict__ counters /* [0]=non-empty docs, [1]=invalid bytes seen */) {
  __shared__ uint32_t cp_buf[EMIT ? kTokWarps : 1][EMIT ? kTokBuf : 1];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp_in_cta = threadIdx.x >> 5;
  const uint64_t warp_global = static_cast<uint64_t>(blockIdx.x) * kTokWarps + warp_in_cta;
  const uint64_t warp_stride = static_cast<uint64_t>(gridDim.x) * kTokWarps;
  unsigned long long nonempty = 0;
  unsigned long long invalid = 0;

  for (uint64_t d = warp_global; d < n_docs; d += warp_stride) {
    const uint64_t b = text_off[d];
    const uint64_t e = text_off[d + 1];
    if (e <= b) {
      if (!EMIT && lane == 0) {
        doc_len[d] = 0;
      }
      continue;
    }
    nonempty += (lane == 0);
    uint64_t cps_done = 0;   // code points of this doc already finalised (EMIT: emitted)
    uint32_t carry_n = 0;    // EMIT: undecided code points kept at the front of cp_buf
    uint64_t valid_bytes = 0;
    const uint64_t slot_base = EMIT ? slot_off[d] : 0;

    for (uint64_t base = b & ~15ULL; base < e; base += kTokTileBytes) {
      const uint64_t my = base + static_cast<uint64_t>(lane) * 16;
      uint32_t w[5] = {0, 0, 0, 0, 0};
      if (my < e) {
        const uint4 v = ld_stream_16(text + my);  // arena is padded, 16-B aligned
        w[0] = v.x;
        w[1] = v.y;
        w[2] = v.z;
        w[3] = v.w;
      }
      uint32_t nxt = __shfl_down_sync(0xffffffffu, w[0], 1);
      if (lane == 31) {
        nxt = (my + 16 < e) ? *reinterpret_cast<const uint32_t*>(text + my + 16) : 0u;
      }
      w[4] = nxt;

      uint32_t flags = 0;
      uint32_t cps[16];
      uint32_t len_sum = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint64_t pos = my + j;
        cps[j] = 0;
        if (pos >= b && pos < e) {
          uint32_t cp = 0;
          const int len = parse_utf8(byte_of(w, j), byte_of(w, j + 1), byte_of(w, j + 2), byte_of(w, j + 3), e - pos, &cp);
          if (len > 0) {
            flags |= 1u << j;
            cps[j] = cp;
            len_sum += static_cast<uint32_t>(len);
          }
        }
      }
      // warp exclusive scan of per-lane code point counts
      const uint32_t cnt = __popc(flags);
      uint32_t inc = cnt;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= static_cast<unsigned>(s)) {
          inc += o;
        }
      }
      const uint32_t tile_total = __shfl_sync(0xffffffffu, inc, 31);
      uint32_t bytes_inc = len_sum;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        bytes_inc += __shfl_xor_sync(0xffffffffu, bytes_inc, s);
      }
      valid_bytes += bytes_inc;

      if (EMIT) {
        uint32_t* buf = cp_buf[warp_in_cta];
        uint32_t wpos = carry_n + inc - cnt;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (flags & (1u << j)) {
            buf[wpos++] = cps[j];
          }
        }
        __syncwarp();
        const uint32_t m = carry_n + tile_total;  // code points available in the window
        const bool last_tile = base + kTokTileBytes >= e;
        const uint32_t keep = last_tile ? 0u : min(m, static_cast<uint32_t>(width - 1));
        const uint32_t emit_n = m - keep;
        for (uint32_t p = lane; p < emit_n; p += 32) {
          const uint32_t c0 = buf[p];
          const bool cjk = is_cjk_ideograph(c0);
          const int size = cjk ? kanji : ngram;  // string_utils.cpp:484-485: chosen by the START code point
          uint64_t key = kInvalidKey;
          if (p + static_cast<uint32_t>(size) <= m) {  // :487 (on the last tile m is the document's end)
            bool ok = true;
            uint64_t k = static_cast<uint64_t>(c0) + 1;
            for (int j = 1; j < width; ++j) {
              uint64_t field = 0;
              if (j < size) {
                const uint32_t cj = buf[p + j];
                if (!cross && is_cjk_ideograph(cj) != cjk) {  // :491-503 legacy boundary rejection
                  ok = false;
                }
                field = static_cast<uint64_t>(cj) + 1;
              }
              k = (k << 21) | field;
            }
            if (ok) {
              key = k;
            }
          }
          const uint64_t slot = slot_base + cps_done + p;
          keys_out[slot] = key;
          docs_out[slot] = static_cast<uint32_t>(d);
        }
        __syncwarp();
        uint32_t carried = 0;
        if (lane < keep) {
          carried = buf[emit_n + lane];
        }
        __syncwarp();
        if (lane < keep) {
          buf[lane] = carried;
        }
        __syncwarp();
        cps_done += emit_n;
        carry_n = keep;
      } else {
        cps_done += tile_total;
      }
    }
    if (!EMIT && lane == 0) {
      doc_len[d] = static_cast<uint32_t>(cps_done);
    }
    if (valid_bytes != e - b) {
      invalid += (lane == 0);
    }
  }
  if (!EMIT && lane == 0) {
    if (nonempty != 0) {
      atomicAdd(&counters[0], nonempty);
    }
    if (invalid != 0) {
      atomicAdd(&counters[1], invalid);
    }
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

__device__ __forceinline__ Heads head_flags(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ docs,
                                            uint64_t i, uint64_t n, uint64_t* key_out) {
  Heads h{0, 0};
  if (i >= n) {
    return h;
  }
  const uint64_t k = keys[i];
  *key_out = k;
  if (k == kInvalidKey) {
    return h;
  }
  const bool term_head = (i == 0) || keys[i - 1] != k;
  const bool pair_head = term_head || docs[i - 1] != docs[i];
  h.terms = term_head;
  h.pairs = pair_head;
  return h;
}

__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, s);
  }
  return v;
}

__global__ void __launch_bounds__(kCsrThreads) csr_count_kernel(const uint64_t* __restrict__ keys,
                                                                const uint32_t* __restrict__ docs, uint64_t n,
                                                                uint64_t* __restrict__ block_pairs,
                                                                uint64_t* __restrict__ block_terms) {
  __shared__ uint32_t sp[kCsrThreads / 32];
  __shared__ uint32_t st[kCsrThreads / 32];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kCsrTile;
  uint32_t pairs = 0;
  uint32_t terms = 0;
#pragma unroll
  for (int k = 0; k < kCsrItems; ++k) {
    const uint64_t i = base + static_cast<uint64_t>(k) * kCsrThreads + threadIdx.x;
    uint64_t key;
    const Heads h = head_flags(keys, docs, i, n, &key);
    pairs += h.pairs;
    terms += h.terms;
  }
  pairs = warp_sum_u32(pairs);
  terms = warp_sum_u32(terms);
  if ((threadIdx.x & 31) == 0) {
    sp[threadIdx.x >> 5] = pairs;
    st[threadIdx.x >> 5] = terms;
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
                                                                const uint32_t* __restrict__ docs, uint64_t n,
                                                                const uint64_t* __restrict__ block_pairs,
                                                                const uint64_t* __restrict__ block_terms,
                                                                uint64_t* __restrict__ term_keys,
                                                                uint64_t* __restrict__ term_off,
                                                                uint32_t* __restrict__ postings) {
  __shared__ uint32_t sp[kCsrThreads / 32];
  __shared__ uint32_t st[kCsrThreads / 32];
  // thread t owns kCsrItems consecutive items so scan order == array order
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kCsrTile + static_cast<uint64_t>(threadIdx.x) * kCsrItems;
  Heads h[kCsrItems];
  uint64_t key[kCsrItems];
  uint32_t pairs = 0;
  uint32_t terms = 0;
#pragma unroll
  for (int k = 0; k < kCsrItems; ++k) {
    key[k] = 0;
    h[k] = head_flags(keys, docs, base + k, n, &key[k]);
    pairs += h[k].pairs;
    terms += h[k].terms;
  }
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  uint32_t ip = pairs;
  uint32_t it = terms;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const uint32_t op = __shfl_up_sync(0xffffffffu, ip, s);
    const uint32_t ot = __shfl_up_sync(0xffffffffu, it, s);
    if (lane >= static_cast<unsigned>(s)) {
      ip += op;
      it += ot;
    }
  }
  if (lane == 31) {
    sp[warp] = ip;
    st[warp] = it;
  }
  __syncthreads();
  uint64_t pp = block_pairs[blockIdx.x] + ip - pairs;
  uint64_t tp = block_terms[blockIdx.x] + it - terms;
  for (unsigned w = 0; w < warp; ++w) {
    pp += sp[w];
    tp += st[w];
  }
#pragma unroll
  for (int k = 0; k < kCsrItems; ++k) {
    if (h[k].pairs) {
      postings[pp] = docs[base + k];
      if (h[k].terms) {
        term_keys[tp] = key[k];
        term_off[tp] = pp;
        ++tp;
      }
      ++pp;
    }
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

uint64_t Index::device_bytes() const {
  return d_doc_ids.bytes() + d_text.bytes() + d_text_off.bytes() + d_doc_len.bytes() + d_term_keys.bytes() +
         d_term_off.bytes() + d_postings.bytes() + d_term_bm.bytes() + d_bitmaps.bytes();
}

void tokenize_device(int ngram, int kanji, bool cross, int width, const uint8_t* d_text, const uint64_t* d_text_off,
                     uint64_t n_docs, DevBuf<uint32_t>& d_doc_len, DevBuf<uint64_t>& d_slot_off, DevBuf<uint64_t>& d_keys,
                     DevBuf<uint32_t>& d_docs, uint64_t* n_slots, uint64_t* counters_out, cudaStream_t stream) {
  int sm_count = 148;
  int dev = 0;
  MGX_CUDA(cudaGetDevice(&dev));
  MGX_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  // persistent-style grid: 8 CTAs of 8 warps per SM, warps stride over documents
  const unsigned grid = static_cast<unsigned>(
      std::max<uint64_t>(1, std::min<uint64_t>(static_cast<uint64_t>(sm_count) * 8, (n_docs + kTokWarps - 1) / kTokWarps)));
  d_doc_len.alloc(n_docs);
  d_slot_off.alloc(n_docs + 1);
  DevBuf<unsigned long long> d_counters;
  d_counters.alloc(2);
  MGX_CUDA(cudaMemsetAsync(d_counters.p, 0, 2 * sizeof(unsigned long long), stream));
  tokenize_kernel<false><<<grid, kTokThreads, 0, stream>>>(d_text, d_text_off, n_docs, ngram, kanji, cross ? 1 : 0,
                                                           width, d_doc_len.p, nullptr, nullptr, nullptr, d_counters.p);
  MGX_LAUNCH_CHECK();
  exclusive_scan_u32_u64(d_doc_len.p, d_slot_off.p, n_docs, stream);
  uint64_t total = 0;
  MGX_CUDA(cudaMemcpyAsync(&total, d_slot_off.p + n_docs, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
  unsigned long long counters[2] = {0, 0};
  MGX_CUDA(cudaMemcpyAsync(counters, d_counters.p, sizeof(counters), cudaMemcpyDeviceToHost, stream));
  MGX_CUDA(cudaStreamSynchronize(stream));
  *n_slots = total;
  counters_out[0] = counters[0];
  counters_out[1] = counters[1];
  d_keys.alloc(total);
  d_docs.alloc(total);
  if (total > 0) {
    tokenize_kernel<true><<<grid, kTokThreads, 0, stream>>>(d_text, d_text_off, n_docs, ngram, kanji, cross ? 1 : 0,
                                                            width, nullptr, d_slot_off.p, d_keys.p, d_docs.p, nullptr);
    MGX_LAUNCH_CHECK();
  }
}

// doc ids ascending: first_id + i everywhere?  (one pass over the resident copy)
__global__ void sequential_check_kernel(const uint32_t* __restrict__ ids, uint64_t n, unsigned int* __restrict__ bad) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i + 1 < n) {
    if (ids[i + 1] != ids[i] + 1) {
      atomicOr(bad, ids[i + 1] > ids[i] ? 1u : 2u);  // 1: gap, 2: not ascending
    }
  }
}

namespace {
// MGX_BUILD_TRACE=1 prints the host-side wall time of each build phase (stream synchronised) to stderr.
struct PhaseTrace {
  bool on;
  cudaStream_t stream;
  std::chrono::steady_clock::time_point t0;
  PhaseTrace(cudaStream_t s) : on(std::getenv("MGX_BUILD_TRACE") != nullptr), stream(s), t0(std::chrono::steady_clock::now()) {}
  void mark(const char* name) {
    if (!on) {
      return;
    }
    cudaStreamSynchronize(stream);
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[mgx build] %-28s %8.2f ms\n", name, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};
}  // namespace

void build_index_device(Index& ix, const uint32_t* d_doc_ids_in, const uint8_t* d_text_in,
                        const uint64_t* d_text_off_in, uint64_t n_docs, uint64_t text_bytes, cudaStream_t stream) {
  PhaseTrace trace(stream);
  uint32_t first_id = 1;
  bool sequential_ids = true;
  cudaEvent_t ev0;
  cudaEvent_t ev1;
  MGX_CUDA(cudaEventCreate(&ev0));
  MGX_CUDA(cudaEventCreate(&ev1));
  MGX_CUDA(cudaEventRecord(ev0, stream));

  // resident copies (the device mirror of DocumentStore's normalised text); the
  // text arena is padded so 16-byte tile loads never leave the allocation
  ix.n_docs = n_docs;
  ix.text_bytes = text_bytes;
  ix.first_id = first_id;
  ix.sequential_ids = sequential_ids;
  ix.d_doc_ids.alloc(n_docs);
  ix.d_text.alloc(text_bytes + 64);
  ix.d_text_off.alloc(n_docs + 1);
  MGX_CUDA(cudaMemcpyAsync(ix.d_doc_ids.p, d_doc_ids_in, n_docs * sizeof(uint32_t), cudaMemcpyDefault, stream));
  MGX_CUDA(cudaMemcpyAsync(ix.d_text.p, d_text_in, text_bytes, cudaMemcpyDefault, stream));
  MGX_CUDA(cudaMemsetAsync(ix.d_text.p + text_bytes, 0, 64, stream));
  MGX_CUDA(cudaMemcpyAsync(ix.d_text_off.p, d_text_off_in, (n_docs + 1) * sizeof(uint64_t), cudaMemcpyDefault, stream));
  if (n_docs > 0) {
    DevBuf<unsigned int> d_bad;
    d_bad.alloc(1);
    MGX_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(unsigned int), stream));
    sequential_check_kernel<<<static_cast<unsigned>((n_docs + 255) / 256), 256, 0, stream>>>(ix.d_doc_ids.p, n_docs,
                                                                                             d_bad.p);
    MGX_LAUNCH_CHECK();
    unsigned int bad = 0;
    MGX_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(bad), cudaMemcpyDeviceToHost, stream));
    MGX_CUDA(cudaMemcpyAsync(&first_id, ix.d_doc_ids.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    MGX_CUDA(cudaStreamSynchronize(stream));
    if (bad & 2u) {
      set_last_error("doc_ids must be strictly ascending");
      throw CudaFailure{MGX_ERR_INVALID_ARGUMENT};
    }
    sequential_ids = bad == 0;
  }
  ix.first_id = first_id;
  ix.sequential_ids = sequential_ids;

  trace.mark("resident copies + id check");
  DevBuf<uint64_t> d_slot_off;
  DevBuf<uint64_t> d_keys_a;
  DevBuf<uint32_t> d_docs_a;
  uint64_t n_slots = 0;
  uint64_t counters[2] = {0, 0};
  tokenize_device(ix.ngram, ix.kanji, ix.cross, ix.width, ix.d_text.p, ix.d_text_off.p, n_docs, ix.d_doc_len,
                  d_slot_off, d_keys_a, d_docs_a, &n_slots, counters, stream);
  d_slot_off.release();
  trace.mark("tokenize (count+scan+emit)");
  ix.n_pair_slots = n_slots;
  ix.total_doc_length = n_slots;  // one slot per code point
  ix.doc_count = counters[0];
  ix.all_valid_utf8 = counters[1] == 0;
  if (n_slots >= (1ULL << 32)) {
    set_last_error("shard too large: more than 2^32 code points in one shard; split by doc-id range");
    throw CudaFailure{MGX_ERR_UNSUPPORTED};
  }

  DevBuf<uint64_t> d_keys_b;
  DevBuf<uint32_t> d_docs_b;
  d_keys_b.alloc(n_slots);
  d_docs_b.alloc(n_slots);
  trace.mark("alloc sort buffers");
  const SortResult sorted =
      radix_sort_pairs(d_keys_a.p, d_docs_a.p, d_keys_b.p, d_docs_b.p, n_slots, 21 * ix.width, stream);
  trace.mark("radix sort");

  // segmented unique + compaction
  const uint64_t n_blocks = (n_slots + kCsrTile - 1) / kCsrTile;
  DevBuf<uint64_t> d_block_pairs;
  DevBuf<uint64_t> d_block_terms;
  DevBuf<uint64_t> d_totals;
  d_block_pairs.alloc(n_blocks + 1);
  d_block_terms.alloc(n_blocks + 1);
  d_totals.alloc(2);
  if (n_blocks > 0) {
    csr_count_kernel<<<static_cast<unsigned>(n_blocks), kCsrThreads, 0, stream>>>(sorted.keys, sorted.vals, n_slots,
                                                                                   d_block_pairs.p, d_block_terms.p);
    MGX_LAUNCH_CHECK();
  }
  csr_scan_kernel<<<1, 1024, 0, stream>>>(d_block_pairs.p, d_block_terms.p, n_blocks, d_totals.p);
  MGX_LAUNCH_CHECK();
  uint64_t totals[2] = {0, 0};
  MGX_CUDA(cudaMemcpyAsync(totals, d_totals.p, sizeof(totals), cudaMemcpyDeviceToHost, stream));
  MGX_CUDA(cudaStreamSynchronize(stream));
  ix.n_postings = totals[0];
  ix.n_terms = totals[1];
  ix.d_term_keys.alloc(ix.n_terms);
  ix.d_term_off.alloc(ix.n_terms + 1);
  ix.d_postings.alloc(ix.n_postings);
  if (n_blocks > 0) {
    csr_write_kernel<<<static_cast<unsigned>(n_blocks), kCsrThreads, 0, stream>>>(
        sorted.keys, sorted.vals, n_slots, d_block_pairs.p, d_block_terms.p, ix.d_term_keys.p, ix.d_term_off.p,
        ix.d_postings.p);
    MGX_LAUNCH_CHECK();
  }
  set_u64_kernel<<<1, 1, 0, stream>>>(ix.d_term_off.p + ix.n_terms, ix.n_postings);
  MGX_LAUNCH_CHECK();
  MGX_CUDA(cudaStreamSynchronize(stream));
  trace.mark("csr");
  d_keys_a.release();
  d_docs_a.release();
  d_keys_b.release();
  d_docs_b.release();
  trace.mark("free sort buffers");

  // dense bitmaps
  ix.bm_words = (n_docs + 31) / 32;
  const double thr = ix.cfg.dense_threshold > 0.0 ? ix.cfg.dense_threshold : 1.0 / 32.0;
  uint64_t min_len = std::max<uint64_t>(1, static_cast<uint64_t>(thr * static_cast<double>(n_docs)));
  // a bitmap only pays for lists long enough that probing beats searching
  min_len = std::max<uint64_t>(min_len, 1024);
  const uint64_t max_bytes = ix.cfg.max_dense_bytes != 0 ? ix.cfg.max_dense_bytes : (8ULL << 30);
  ix.d_term_bm.alloc(ix.n_terms);
  DevBuf<unsigned long long> d_count;
  d_count.alloc(1);
  unsigned long long n_dense = 0;
  const unsigned term_grid = static_cast<unsigned>((ix.n_terms + 255) / 256);
  if (ix.n_terms > 0) {
    for (;;) {
      MGX_CUDA(cudaMemsetAsync(d_count.p, 0, sizeof(unsigned long long), stream));
      dense_count_kernel<<<term_grid, 256, 0, stream>>>(ix.d_term_off.p, ix.n_terms, min_len, d_count.p);
      MGX_LAUNCH_CHECK();
      MGX_CUDA(cudaMemcpyAsync(&n_dense, d_count.p, sizeof(n_dense), cudaMemcpyDeviceToHost, stream));
      MGX_CUDA(cudaStreamSynchronize(stream));
      if (n_dense * ix.bm_words * 4 <= max_bytes) {
        break;
      }
      min_len *= 2;
    }
  }
  ix.n_dense = n_dense;
  ix.dense_min_len = min_len;
  ix.d_bitmaps.alloc(ix.n_dense * ix.bm_words);
  if (ix.n_terms > 0) {
    DevBuf<uint32_t> d_dense_terms;
    d_dense_terms.alloc(ix.n_dense);
    MGX_CUDA(cudaMemsetAsync(d_count.p, 0, sizeof(unsigned long long), stream));
    if (ix.n_dense > 0) {
      MGX_CUDA(cudaMemsetAsync(ix.d_bitmaps.p, 0, ix.d_bitmaps.bytes(), stream));
    }
    dense_assign_kernel<<<term_grid, 256, 0, stream>>>(ix.d_term_off.p, ix.n_terms, min_len, d_count.p, ix.d_term_bm.p,
                                                       d_dense_terms.p);
    MGX_LAUNCH_CHECK();
    if (ix.n_dense > 0) {
      const unsigned slices = static_cast<unsigned>(std::min<uint64_t>(64, (n_docs + 65535) / 65536 + 1));
      dense_fill_kernel<<<dim3(static_cast<unsigned>(ix.n_dense), slices), 256, 0, stream>>>(
          ix.d_term_off.p, ix.d_postings.p, d_dense_terms.p, ix.d_bitmaps.p, ix.bm_words);
      MGX_LAUNCH_CHECK();
    }
    MGX_CUDA(cudaStreamSynchronize(stream));
  }

  trace.mark("dense bitmaps");
  MGX_CUDA(cudaEventRecord(ev1, stream));
  MGX_CUDA(cudaEventSynchronize(ev1));
  float ms = 0.f;
  MGX_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  ix.last_build_ms = ms;
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
}

}  // namespace mgx
