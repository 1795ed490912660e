// query.cu — batched query execution on the device.
//
// Reference path replaced (all paths relative to the reference tree):
//   GenerateTermInfos / PopulateTermDocumentFrequency   server/search_pipeline.cpp:569-603, 542-565
//   Execute (AND smallest-first, early exit)             server/search_pipeline.cpp:795-869
//   ApplyNotFilter                                       server/search_pipeline.cpp:871-932
//   PostFilterByText (verify_text / hybrid fragments)    server/search_pipeline.cpp:1239-1266, 858-866
//   Index::SearchAnd / FilterByNgrams / SearchOr / Not   index/index.cpp:199-486
//   BM25Scorer::ScoreDocuments                           index/bm25_scorer.cpp:47-99
//   ResultSorter::SortByScore                            query/result_sorter.cpp:661-716
//
// Execution model. A batch is compiled on the host into unique terms (bytes +
// packed n-gram keys) and queries (term ids). On the device:
//   lookup        keys -> dictionary slots (binary search over sorted packed keys)
//   term_plan     per term: lists by ascending length, estimated size, df work
//   df_tile       per 1024-entry tile of a term's shortest list: membership in the
//                 term's other lists, then substring check of the survivors' text
//   query_plan    per query: terms by estimated size, merged list table, driver
//   and_tile      per 1024-entry tile of the query's shortest list ("driver"):
//                 membership in every other list (bit probe for dense lists,
//                 binary search for sparse ones), NOT groups, then the fused
//                 epilogue: text scan -> tf -> BM25 (FP64) -> (doc, score) records
//   topk          per query: count, then block-level top-k (shared-memory bitonic
//                 for small sets, MSB radix select for large ones)
// The driver formulation reads each query's smallest list once and probes the
// rest, so the bytes moved are <= the SURVEY §8(d) algorithmic bytes.
#include <algorithm>
#include <charconv>
#include <cstdlib>
#include <cstring>

#include "query.cuh"

namespace mgx {

namespace {

struct BatchView {
  // terms
  const uint8_t* term_bytes;
  const uint32_t* term_boff;
  const uint32_t* term_koff;
  uint32_t* key_list;
  uint32_t* key_len;
  uint32_t* key_toff;  // low 16 bits: offset | count (kNoTermOffset), high 16: neighbour hashes (mgx_internal.cuh)
  uint64_t* t_est;
  uint32_t* t_df_tiles;
  uint64_t* t_df_tile_off;
  uint64_t* t_df;
  uint32_t n_terms;
  // queries
  const uint32_t* q_toff;
  uint32_t* q_tids;
  const uint32_t* q_noff;
  const uint32_t* q_ntids;
  const uint32_t* q_loff;
  uint32_t* q_list;
  uint32_t* q_list_len;
  uint32_t* q_nlists;
  uint32_t* q_flags;
  const uint32_t* q_host_flags;
  const uint32_t* q_threshold;
  const uint32_t* q_poff;
  const uint8_t* prog_op;
  const uint32_t* prog_arg;
  const uint32_t* q_coff;
  const uint32_t* q_conj;
  const uint32_t* q_foff;
  const FilterPred* filters;
  uint32_t* q_driver_len;
  uint32_t* q_ntiles;
  uint64_t* q_tile_off;
  uint64_t* q_group_off;           // [Q+1] top-k pre-reduction groups per query (streamed batches)
  const double* q_idf;
  uint32_t n_queries;
  // explicit driver (query 0)
  const uint32_t* explicit_ids;
  uint32_t explicit_n;
  unsigned long long* stats;  // StatSlot counters
  const DfTileDesc* df_tile_desc;  // [df tiles]
  KeyRef* key_ref;                 // [K]
  const uint32_t* tile_query;    // [and tiles] query index of each intersect tile
  // streaming df pass
  const uint32_t* stream_slots;        // bucket table: (first entry << 8) | count
  const StreamEntry* stream_entries;   // entries in bucket order
  const uint32_t* stream_bloom;        // [kStreamBloomWords]
  uint32_t stream_len8_mask;           // stage-1 classes present (bit m: min(len, 8) == m)
  uint32_t stream_len12_mask;          // stage-2 classes present (bit m: min(len, 12) == m)
  uint32_t stream_slot_mask;           // n_slots - 1
  uint32_t* df_mode;                   // [0] = 1: streaming pass chosen for this batch
  uint32_t* launch;                    // LaunchSlot block of a streamed batch, nullptr otherwise
  const uint8_t* term_flags;           // [T] bit0 raw, bit1 exact_single, bit2 stream-eligible, bit3 tf from the posting payload
  const uint32_t* q_tids0;             // [sum] search terms in QUERY order (q_tids is re-ordered by the planner)
  uint64_t* key_glen;                  // [K] posting size per key in UPLOAD order; summed over the shards by the df
                                       // exchange of the sharded pipeline (global term order, see global_order_kernel)
  unsigned long long* q_thr;           // [Q] running score threshold of the per-tile top-k pruning (and_tile_body)
};

struct ScoreParams {
  double k1;
  double b;
  double avgdl_clamped;  // max(avg_doc_length, 1.0), bm25_scorer.cpp:77
  int compute_score;
  int descending;
};

// ------------------------------------------------------------------ small device helpers
__host__ __device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* __restrict__ p, uint32_t n, uint32_t v) {
  uint32_t lo = 0;
  uint32_t hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(p + mid) < v) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return lo;
}

struct ListRef {
  const uint32_t* p;   // sorted local doc indices
  uint32_t len;
  const uint32_t* bm;  // dense bitmap or nullptr
};

__device__ __forceinline__ ListRef make_list(const IndexView& iv, uint32_t dict_term, uint32_t len) {
  ListRef r;
  r.p = iv.postings + iv.term_off[dict_term];
  r.len = len;
  const int32_t slot = iv.term_bm[dict_term];
  r.bm = slot >= 0 ? iv.bitmaps + static_cast<uint64_t>(slot) * iv.bm_words : nullptr;
  return r;
}

__device__ __forceinline__ bool list_contains(const ListRef& l, uint32_t doc) {
  if (l.bm != nullptr) {
    return (__ldg(l.bm + (doc >> 5)) >> (doc & 31)) & 1u;
  }
  const uint32_t pos = lower_bound_u32(l.p, l.len, doc);
  return pos < l.len && __ldg(l.p + pos) == doc;
}

// ------------------------------------------------------------------ text scanning (one warp per document)
constexpr uint32_t kTextChunk = 1024;                         // start positions per staged chunk
constexpr uint32_t kStageMax = kTextChunk + kMaxTermBytes;    // bytes staged at once
constexpr uint32_t kStageBuf = kStageMax + 32;                // + alignment slack, multiple of 16

__device__ __forceinline__ uint4 ld16(const uint8_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// Copies text[gb, gb+n) into the warp's shared buffer; returns the offset of
// text[gb] inside it. The arena is padded by 64 bytes, so whole 16-byte vectors
// can always be read.
__device__ __forceinline__ uint32_t warp_stage(const uint8_t* __restrict__ text, uint64_t gb, uint32_t n, uint8_t* buf) {
  const unsigned lane = threadIdx.x & 31u;
  const uint64_t a = gb & ~15ULL;
  const uint32_t shift = static_cast<uint32_t>(gb - a);
  const uint32_t nvec = (shift + n + 15u) >> 4;
  for (uint32_t v = lane; v < nvec; v += 32) {
    *reinterpret_cast<uint4*>(buf + v * 16u) = ld16(text + a + static_cast<uint64_t>(v) * 16u);
  }
  __syncwarp();
  return shift;
}

// Non-overlapping, left-to-right occurrence count of `term` among the start
// positions [0, npos) of the staged bytes `s` (BM25Scorer::CountTermOccurrences,
// bm25_scorer.cpp:27-45). `next_ok`/`count` carry the greedy state across chunks;
// positions are relative to `pos0`. Warp-uniform result.
__device__ __forceinline__ void warp_count_staged(const uint8_t* s, uint32_t npos, const uint8_t* __restrict__ term,
                                                  uint32_t tl, uint64_t pos0, uint64_t* next_ok, uint32_t* count,
                                                  bool exists_only) {
  const unsigned lane = threadIdx.x & 31u;
  const uint8_t t0 = __ldg(term);
  for (uint32_t g0 = 0; g0 < npos; g0 += 32) {
    const uint32_t p = g0 + lane;
    bool m = false;
    if (p < npos && s[p] == t0) {
      m = true;
      for (uint32_t i = 1; i < tl; ++i) {
        if (s[p + i] != __ldg(term + i)) {
          m = false;
          break;
        }
      }
    }
    unsigned mask = __ballot_sync(0xffffffffu, m);
    while (mask != 0) {
      const uint32_t bit = static_cast<uint32_t>(__ffs(static_cast<int>(mask))) - 1u;
      const uint64_t pos = pos0 + g0 + bit;
      if (pos >= *next_ok) {
        *count += 1;
        *next_ok = pos + tl;
      }
      mask &= mask - 1;
    }
    if (exists_only && *count != 0) {
      return;
    }
  }
}

struct DocText {
  const uint8_t* text;  // arena
  uint64_t b;           // first byte of the document
  uint32_t len;         // bytes
  uint8_t* buf;         // warp staging buffer (kStageBuf bytes)
  uint32_t shift;       // valid when whole == true
  bool whole;           // the whole document is staged
};

__device__ __forceinline__ DocText doc_open(const IndexView& iv, uint32_t doc, uint8_t* buf) {
  DocText d;
  d.text = iv.text;
  d.b = iv.text_off[doc];
  const uint64_t e = iv.text_off[doc + 1];
  d.len = static_cast<uint32_t>(e - d.b);
  d.buf = buf;
  d.whole = d.len <= kStageMax;
  d.shift = 0;
  if (d.whole && d.len > 0) {
    d.shift = warp_stage(d.text, d.b, d.len, buf);
  }
  return d;
}

__device__ __forceinline__ uint32_t doc_count_term(DocText& d, const uint8_t* __restrict__ term, uint32_t tl,
                                                   bool exists_only) {
  if (tl == 0 || tl > d.len) {
    return 0;
  }
  uint32_t count = 0;
  uint64_t next_ok = 0;
  if (d.whole) {
    warp_count_staged(d.buf + d.shift, d.len - tl + 1, term, tl, 0, &next_ok, &count, exists_only);
    return count;
  }
  for (uint64_t c0 = 0; c0 + tl <= d.len; c0 += kTextChunk) {
    const uint32_t n = static_cast<uint32_t>(min(static_cast<uint64_t>(d.len) - c0, static_cast<uint64_t>(kTextChunk + tl - 1)));
    __syncwarp();
    const uint32_t shift = warp_stage(d.text, d.b + c0, n, d.buf);
    const uint32_t npos = min(kTextChunk, n - tl + 1);
    warp_count_staged(d.buf + shift, npos, term, tl, c0, &next_ok, &count, exists_only);
    if (exists_only && count != 0) {
      break;
    }
  }
  __syncwarp();
  return count;
}

// ------------------------------------------------------------------ lookup + term planning
// wide_keys != nullptr: the dictionary is the table of wide keys (n_words words per term, ascending word by word) and
// key i of the batch is qwide[i * n_words ..]; keys[i] only tells "cannot exist" (kInvalidKey).
__device__ __forceinline__ int wide_cmp(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b, int n_words) {
  for (int w = 0; w < n_words; ++w) {
    const uint64_t x = a[w];
    const uint64_t y = b[w];
    if (x != y) {
      return x < y ? -1 : 1;
    }
  }
  return 0;
}

__device__ __forceinline__ void lookup_one(const uint64_t* __restrict__ term_keys, const uint64_t* __restrict__ term_off,
                                           uint64_t n_dict, const uint64_t* __restrict__ keys, uint32_t i,
                                           uint32_t* __restrict__ key_list, uint32_t* __restrict__ key_len,
                                           uint64_t* __restrict__ key_glen, const uint64_t* __restrict__ wide_keys,
                                           int n_words, const uint64_t* __restrict__ qwide) {
  const uint64_t key = keys[i];
  uint64_t lo = 0;
  uint64_t hi = n_dict;
  bool found = false;
  if (wide_keys != nullptr) {
    if (key != kInvalidKey) {
      const uint64_t* q = qwide + static_cast<uint64_t>(i) * n_words;
      while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (wide_cmp(wide_keys + mid * n_words, q, n_words) < 0) {
          lo = mid + 1;
        } else {
          hi = mid;
        }
      }
      found = lo < n_dict && wide_cmp(wide_keys + lo * n_words, q, n_words) == 0;
    }
  } else {
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (term_keys[mid] < key) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    found = lo < n_dict && term_keys[lo] == key;
  }
  if (found) {
    key_list[i] = static_cast<uint32_t>(lo);
    key_len[i] = static_cast<uint32_t>(term_off[lo + 1] - term_off[lo]);
  } else {
    key_list[i] = kNone;
    key_len[i] = 0;
  }
  if (key_glen != nullptr) {
    key_glen[i] = key_len[i];  // upload order: the same position on every shard (term_plan sorts the other arrays)
  }
}

__global__ void lookup_kernel(const uint64_t* __restrict__ term_keys, const uint64_t* __restrict__ term_off,
                              uint64_t n_dict, const uint64_t* __restrict__ keys, uint32_t n_keys,
                              uint32_t* __restrict__ key_list, uint32_t* __restrict__ key_len,
                              uint64_t* __restrict__ key_glen, const uint64_t* __restrict__ wide_keys, int n_words,
                              const uint64_t* __restrict__ qwide) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_keys) {
    lookup_one(term_keys, term_off, n_dict, keys, i, key_list, key_len, key_glen, wide_keys, n_words, qwide);
  }
}

// df_unit: driver entries per df work item (kTile with tile descriptors, kDfUnit in a streamed batch)
__device__ __forceinline__ void term_plan_one(const BatchView& bv, uint32_t t, int compute_df, int all_valid_utf8,
                                              const uint8_t* __restrict__ raw_flags, uint64_t bitmap_bytes,
                                              uint32_t df_unit) {
  const uint32_t k0 = bv.term_koff[t];
  const uint32_t k1 = bv.term_koff[t + 1];
  uint64_t est = kEstNone;
  if (k1 > k0) {
    for (uint32_t i = k0 + 1; i < k1; ++i) {  // insertion sort by length (missing lists have length 0)
      const uint32_t li = bv.key_list[i];
      const uint32_t ln = bv.key_len[i];
      const uint32_t to = bv.key_toff[i];
      uint32_t j = i;
      while (j > k0 && bv.key_len[j - 1] > ln) {
        bv.key_list[j] = bv.key_list[j - 1];
        bv.key_len[j] = bv.key_len[j - 1];
        bv.key_toff[j] = bv.key_toff[j - 1];
        --j;
      }
      bv.key_list[j] = li;
      bv.key_len[j] = ln;
      bv.key_toff[j] = to;
    }
    est = bv.key_len[k0];  // min posting size, 0 if any n-gram is missing (search_pipeline.cpp:583-593)
  }
  bv.t_est[t] = est;
  uint64_t df = 0;
  uint32_t tiles = 0;
  if (compute_df && k1 > k0 && est != 0 && (raw_flags[t] & 1) == 0) {
    if (k1 - k0 == 1 && all_valid_utf8 && (raw_flags[t] & 2) != 0) {
      df = est;  // the term is exactly its one n-gram: every posting contains it as a substring
    } else {
      tiles = static_cast<uint32_t>((est + df_unit - 1) / df_unit);
      unsigned long long bytes = 0;
      for (uint32_t i = k0; i < k1; ++i) {
        bytes += umin64(4ULL * bv.key_len[i], bitmap_bytes);
      }
      atomicAdd(bv.stats + kStatDfLists * kStatStripes + (t & (kStatStripes - 1)), bytes);
      if ((raw_flags[t] & 4) != 0) {
        atomicAdd(bv.stats + kStatStreamEntries * kStatStripes + (t & (kStatStripes - 1)),
                  static_cast<unsigned long long>(est));
      }
    }
  }
  bv.t_df[t] = df;
  bv.t_df_tiles[t] = tiles;
}

__global__ void term_plan_kernel(BatchView bv, int compute_df, int all_valid_utf8, const uint8_t* __restrict__ raw_flags,
                                 uint64_t bitmap_bytes) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < bv.n_terms) {
    term_plan_one(bv, t, compute_df, all_valid_utf8, raw_flags, bitmap_bytes, kTile);
  }
}

// Chooses, per batch, how the verified document frequencies of the stream-eligible terms are computed:
// per-term candidate tiles (df_tile_kernel: cost ~ entries of the terms' shortest lists) or ONE pass over the
// text arena matching all of them at once (df_stream_kernel: cost ~ text bytes). force: 0 auto, 1 tiles, 2 stream.
constexpr unsigned long long kStreamCostRatio = 6;  // measured: ~25 ps per list entry of candidate work, ~4 ps per streamed byte
__device__ __forceinline__ void df_mode_one(const BatchView& bv, uint32_t t, const uint8_t* __restrict__ raw_flags,
                                            uint64_t text_bytes, int force, int have_table) {
  unsigned long long entries = 0;
  for (int i = 0; i < kStatStripes; ++i) {
    entries += bv.stats[kStatStreamEntries * kStatStripes + i];
  }
  bool stream = have_table != 0 && entries > 0;
  if (force == 1) {
    stream = false;
  } else if (force == 0) {
    stream = stream && entries * kStreamCostRatio >= text_bytes;
  }
  if (stream && (raw_flags[t] & 4) != 0) {
    bv.t_df_tiles[t] = 0;  // counted by the streaming pass instead
  }
  if (t == 0) {
    bv.df_mode[0] = stream ? 1u : 0u;
  }
}

__global__ void df_mode_kernel(BatchView bv, const uint8_t* __restrict__ raw_flags, uint64_t text_bytes, int force,
                               int have_table) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < bv.n_terms) {
    df_mode_one(bv, t, raw_flags, text_bytes, force, have_table);
  }
}

// ------------------------------------------------------------------ tile membership (block-cooperative)
// Lower bound of v in p[0, n) by a whole warp: every step probes 32 evenly spaced positions, so a 300k-entry list
// needs 4 dependent rounds instead of 18.
__device__ __forceinline__ uint32_t warp_lower_bound(const uint32_t* __restrict__ p, uint32_t n, uint32_t v) {
  const unsigned lane = threadIdx.x & 31u;
  uint32_t lo = 0;
  uint32_t hi = n;
  while (hi - lo > 32) {
    const uint32_t size = hi - lo;
    const uint32_t idx = lo + static_cast<uint32_t>((static_cast<uint64_t>(lane + 1) * size) / 33);
    const bool less = __ldg(p + idx) < v;
    const unsigned mask = __ballot_sync(0xffffffffu, less);
    const int c = __popc(mask);  // `less` is monotone in the lane index
    const uint32_t idx_prev = __shfl_sync(0xffffffffu, idx, c > 0 ? c - 1 : 0);
    const uint32_t idx_next = __shfl_sync(0xffffffffu, idx, c < 32 ? c : 31);
    if (c > 0) {
      lo = idx_prev + 1;
    }
    if (c < 32) {
      hi = idx_next;
    }
  }
  const uint32_t i = lo + lane;
  const bool less = i < hi && __ldg(p + i) < v;
  return lo + __popc(__ballot_sync(0xffffffffu, less));
}

constexpr uint32_t kStageCap = 4096;  // posting entries of one list staged per tile (16 KB)

__device__ __forceinline__ bool sorted_contains(const uint32_t* p, uint32_t n, uint32_t v) {
  uint32_t lo = 0;
  uint32_t hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (p[mid] < v) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return lo < n && p[lo] == v;
}

// Keeps, among the thread's kTileItems driver entries, those present in list `l`.
// Dense list: one bit probe per entry. Sparse list: the tile's doc range [dmin, dmax] selects a sub-range of the
// list (two warp-cooperative bound searches); if it fits, it is staged in shared memory with coalesced loads and
// searched there, otherwise searched in place. Must be called by every thread of the CTA.
// `pre` (optional): the sub-range [pre[0], pre[1]) found beforehand (and_tile_kernel searches the bounds of the first
// lists of a query with all its warps at once, before walking the lists one after the other).
__device__ __forceinline__ void tile_filter_list(const ListRef& l, uint32_t dmin, uint32_t dmax, bool narrow,
                                                 uint32_t* s_stage, uint32_t* s_range,
                                                 const uint32_t (&my_doc)[kTileItems], uint32_t* alive_mask,
                                                 const uint32_t* pre = nullptr) {
  if (l.bm != nullptr) {
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
      if ((*alive_mask >> k) & 1u) {
        const uint32_t d = my_doc[k];
        if (d == kNone || ((__ldg(l.bm + (d >> 5)) >> (d & 31)) & 1u) == 0) {
          *alive_mask &= ~(1u << k);
        }
      }
    }
    return;
  }
  if (pre == nullptr) {
    const unsigned w = threadIdx.x >> 5;
    if (w < 2) {  // the two bounds by two warps, side by side
      const uint32_t r = narrow ? warp_lower_bound(l.p, l.len, w == 0 ? dmin : dmax + 1u) : (w == 0 ? 0u : l.len);
      if ((threadIdx.x & 31u) == 0) {
        s_range[w] = r;
      }
    }
    __syncthreads();
    pre = s_range;
  }
  const uint32_t lo = pre[0];
  const uint32_t cnt = pre[1] - lo;
  if (cnt == 0) {
    *alive_mask = 0;
  } else if (cnt <= kStageCap) {
    for (uint32_t i = threadIdx.x; i < cnt; i += kTileThreads) {
      s_stage[i] = __ldg(l.p + lo + i);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
      if ((*alive_mask >> k) & 1u) {
        if (my_doc[k] == kNone || !sorted_contains(s_stage, cnt, my_doc[k])) {
          *alive_mask &= ~(1u << k);
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
      if ((*alive_mask >> k) & 1u) {
        const uint32_t d = my_doc[k];
        const uint32_t pos = d == kNone ? cnt : lower_bound_u32(l.p + lo, cnt, d);
        if (pos >= cnt || __ldg(l.p + lo + pos) != d) {
          *alive_mask &= ~(1u << k);
        }
      }
    }
  }
  __syncthreads();  // s_range / s_stage are reused by the next list
}

// ------------------------------------------------------------------ text scanning, one THREAD per document
// Short terms (<= 16 bytes) and ordinary documents: every thread streams its own document through two 16-byte
// registers, so a warp keeps 32 documents in flight. The first 4 bytes of the term are compared at all 16 byte
// offsets of a chunk with funnel shifts; the rare candidates are confirmed byte by byte (L1-resident).
#ifndef MGX_AND_OCC
#define MGX_AND_OCC 5  // resident CTAs per SM the register allocation of and_tile_kernel is sized for
#endif
constexpr int kGroupScanLanes = 8;          // lanes per document in the epilogue of a tile with few survivors
constexpr uint32_t kPairScanMaxItems = 64;  // (document, term) pairs counted side by side, kGroupScanLanes lanes each
constexpr uint32_t kGroupScanMaxDocs = 64;  // "few": at most two rounds of kTileThreads / kGroupScanLanes documents
constexpr uint32_t kThreadScanMaxTerm = 16;
constexpr uint32_t kThreadScanMaxDoc = 4096;

// The first 12 bytes of a term held in registers (3 little-endian words + byte masks).
struct TermRegs {
  uint32_t w[3];
  uint32_t m[3];
  uint32_t nw;  // words in use: ceil(min(tl, 12) / 4)
};

__device__ __forceinline__ TermRegs load_term_regs(const uint8_t* __restrict__ term, uint32_t tl) {
  TermRegs t;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    t.w[k] = 0;
    t.m[k] = 0;
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    if (static_cast<uint32_t>(i) < tl) {
      t.w[i >> 2] |= static_cast<uint32_t>(__ldg(term + i)) << (8 * (i & 3));
      t.m[i >> 2] |= 0xFFu << (8 * (i & 3));
    }
  }
  t.nw = tl >= 9 ? 3u : (tl >= 5 ? 2u : 1u);
  return t;
}

// 4 bytes of the 32-byte window w[0..7] starting at byte offset OFF (compile-time)
template <int OFF>
__device__ __forceinline__ uint32_t window_word(const uint32_t (&w)[8]) {
  if constexpr ((OFF & 3) == 0) {
    return w[OFF >> 2];
  } else {
    return __funnelshift_r(w[OFF >> 2], w[(OFF >> 2) + 1], (OFF & 3) * 8);
  }
}

// Bit j of the result is set iff the first min(tl, 12) bytes of the term match at byte offset j of the window.
template <int J>
__device__ __forceinline__ void match_at(const uint32_t (&w)[8], const TermRegs& t, uint32_t* cand) {
  const uint32_t x0 = window_word<J>(w);
  if (((x0 ^ t.w[0]) & t.m[0]) == 0) {
    bool ok = true;
    if (t.nw > 1) {
      ok = ((window_word<J + 4>(w) ^ t.w[1]) & t.m[1]) == 0;
    }
    if (ok && t.nw > 2) {
      ok = ((window_word<J + 8>(w) ^ t.w[2]) & t.m[2]) == 0;
    }
    if (ok) {
      *cand |= 1u << J;
    }
  }
}

__device__ __forceinline__ uint32_t chunk_candidates(const uint4& cur, const uint4& nxt, const TermRegs& t) {
  const uint32_t w[8] = {cur.x, cur.y, cur.z, cur.w, nxt.x, nxt.y, nxt.z, nxt.w};
  uint32_t cand = 0;
  match_at<0>(w, t, &cand);
  match_at<1>(w, t, &cand);
  match_at<2>(w, t, &cand);
  match_at<3>(w, t, &cand);
  match_at<4>(w, t, &cand);
  match_at<5>(w, t, &cand);
  match_at<6>(w, t, &cand);
  match_at<7>(w, t, &cand);
  match_at<8>(w, t, &cand);
  match_at<9>(w, t, &cand);
  match_at<10>(w, t, &cand);
  match_at<11>(w, t, &cand);
  match_at<12>(w, t, &cand);
  match_at<13>(w, t, &cand);
  match_at<14>(w, t, &cand);
  match_at<15>(w, t, &cand);
  return cand;
}

// bytes 12.. of a longer term (13..16 bytes), confirmed from memory (rare)
__device__ __forceinline__ bool term_tail_matches(const uint8_t* __restrict__ text, uint64_t pos,
                                                  const uint8_t* __restrict__ term, uint32_t tl) {
  for (uint32_t i = 12; i < tl; ++i) {
    if (__ldg(text + pos + i) != __ldg(term + i)) {
      return false;
    }
  }
  return true;
}

// BM25Scorer::CountTermOccurrences (bm25_scorer.cpp:27-45): non-overlapping, left to right, on bytes.
__device__ __forceinline__ uint32_t thread_count_term(const uint8_t* __restrict__ text, uint64_t b, uint32_t len,
                                                      const uint8_t* __restrict__ term, uint32_t tl,
                                                      const TermRegs& t, bool exists_only) {
  if (tl == 0 || tl > len) {
    return 0;
  }
  const uint64_t last = b + len - tl;  // last admissible start
  uint64_t next_ok = b;
  uint32_t count = 0;
  uint64_t a = b & ~15ULL;
  uint4 cur = ld16(text + a);
  for (; a <= last; a += 16) {
    const uint4 nxt = ld16(text + a + 16);  // the arena is padded by 64 bytes
    uint32_t cand = chunk_candidates(cur, nxt, t);
    while (cand != 0) {
      const uint32_t j = static_cast<uint32_t>(__ffs(static_cast<int>(cand))) - 1u;
      cand &= cand - 1;
      const uint64_t pos = a + j;
      if (pos < next_ok || pos > last) {
        continue;
      }
      if (tl > 12 && !term_tail_matches(text, pos, term, tl)) {
        continue;
      }
      ++count;
      next_ok = pos + tl;
      if (exists_only) {
        return 1;
      }
    }
    cur = nxt;
  }
  return count;
}

// thread_count_term with G lanes per document: in every round the lanes of a group load G consecutive 16-byte chunks
// (G loads in flight instead of one), then all of them replay the left-to-right, non-overlapping count over the
// round's candidate masks, so the result is uniform inside the group. For tiles with few surviving documents, where
// one thread per document leaves the CTA waiting on a chain of dependent loads. Every lane of an aligned group of G
// must call this with the same arguments; shuffles use the group's own mask.
template <int G>
__device__ __forceinline__ uint32_t group_count_term(const uint8_t* __restrict__ text, uint64_t b, uint32_t len,
                                                     const uint8_t* __restrict__ term, uint32_t tl, const TermRegs& t,
                                                     bool exists_only) {
  if (tl == 0 || tl > len) {
    return 0;
  }
  const unsigned lane = threadIdx.x & 31u;
  const unsigned gl = lane & (G - 1);
  const unsigned gbase = lane & ~static_cast<unsigned>(G - 1);
  const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << gbase;
  const uint64_t last = b + len - tl;  // last admissible start
  const uint64_t a0 = b & ~15ULL;
  const uint32_t nchunks = static_cast<uint32_t>((last - a0) >> 4) + 1u;
  uint64_t next_ok = b;
  uint32_t count = 0;
  for (uint32_t c0 = 0; c0 < nchunks; c0 += G) {
    const uint32_t c = c0 + gl;
    uint32_t cand = 0;
    if (c < nchunks) {
      const uint64_t a = a0 + static_cast<uint64_t>(c) * 16;
      const uint4 cur = ld16(text + a);
      const uint4 nxt = ld16(text + a + 16);  // the arena is padded by 64 bytes
      cand = chunk_candidates(cur, nxt, t);
      if (tl > 12) {
        uint32_t m = cand;
        while (m != 0) {
          const uint32_t j = static_cast<uint32_t>(__ffs(static_cast<int>(m))) - 1u;
          m &= m - 1;
          const uint64_t pos = a + j;
          if (pos < b || pos > last || !term_tail_matches(text, pos, term, tl)) {
            cand &= ~(1u << j);
          }
        }
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      uint32_t m = __shfl_sync(gmask, cand, static_cast<int>(gbase) + g);
      const uint64_t ag = a0 + static_cast<uint64_t>(c0 + g) * 16;
      while (m != 0) {
        const uint32_t j = static_cast<uint32_t>(__ffs(static_cast<int>(m))) - 1u;
        m &= m - 1;
        const uint64_t pos = ag + j;
        if (pos < next_ok || pos > last) {
          continue;
        }
        ++count;
        next_ok = pos + tl;
        if (exists_only) {
          return 1;
        }
      }
    }
  }
  return count;
}

// Existence test with G lanes per document: lane g of the group takes the 16-byte chunks c = g, g+G, ... so one
// document keeps G loads in flight and short survivor lists still fill the warp. All lanes of the warp must call
// this (with their group's document, or len = 0 for idle groups); the result is uniform inside a group.
template <int G>
__device__ __forceinline__ bool group_contains_term(const uint8_t* __restrict__ text, uint64_t b, uint32_t len,
                                                    const uint8_t* __restrict__ term, uint32_t tl, const TermRegs& t) {
  const unsigned gl = threadIdx.x & (G - 1);
  bool found = false;
  if (tl != 0 && tl <= len) {
    const uint64_t last = b + len - tl;
    const uint64_t a0 = b & ~15ULL;
    const uint32_t nchunks = static_cast<uint32_t>((last - a0) >> 4) + 1u;
    for (uint32_t c = gl; c < nchunks && !found; c += G) {
      const uint64_t a = a0 + static_cast<uint64_t>(c) * 16;
      const uint4 cur = ld16(text + a);
      const uint4 nxt = ld16(text + a + 16);
      uint32_t cand = chunk_candidates(cur, nxt, t);
      // admissible starts of this chunk: b <= a + j <= last
      if (a < b) {
        cand &= ~((1u << static_cast<uint32_t>(b - a)) - 1u);
      }
      if (last - a < 15) {
        cand &= (2u << static_cast<uint32_t>(last - a)) - 1u;
      }
      if (tl > 12) {
        while (cand != 0 && !found) {
          const uint32_t j = static_cast<uint32_t>(__ffs(static_cast<int>(cand))) - 1u;
          cand &= cand - 1;
          found = term_tail_matches(text, a + j, term, tl);
        }
      } else {
        found = cand != 0;
      }
    }
  }
  // OR over the group (xor-shuffles stay inside an aligned group of G lanes)
#pragma unroll
  for (int s = G / 2; s > 0; s >>= 1) {
    found = (__shfl_xor_sync(0xffffffffu, found ? 1 : 0, s) != 0) || found;
  }
  return found;
}

// key -> resolved list, after term_plan_kernel has ordered each term's keys by list length
__device__ __forceinline__ void key_ref_one(const IndexView& iv, const uint32_t* __restrict__ key_list,
                                            const uint32_t* __restrict__ key_len, uint32_t i, KeyRef* __restrict__ out) {
  KeyRef r{};
  const uint32_t dict = key_list[i];
  if (dict != kNone) {
    r.p = iv.postings + iv.term_off[dict];
    const int32_t slot = iv.term_bm[dict];
    r.bm = slot >= 0 ? iv.bitmaps + static_cast<uint64_t>(slot) * iv.bm_words : nullptr;
    r.len = key_len[i];
  }
  out[i] = r;
}

__global__ void key_ref_kernel(IndexView iv, const uint32_t* __restrict__ key_list, const uint32_t* __restrict__ key_len,
                               uint32_t n_keys, KeyRef* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_keys) {
    key_ref_one(iv, key_list, key_len, i, out);
  }
}

// Streamed planning, stage 1: everything that concerns ONE term in one thread -- dictionary lookup of its keys, lists
// by ascending length, estimate, df work units, resolved list references (lookup + term_plan + key_ref in one launch).
__global__ void plan_terms_kernel(IndexView iv, BatchView bv, const uint64_t* __restrict__ keys, int compute_df,
                                  int all_valid_utf8, const uint8_t* __restrict__ raw_flags, uint64_t bitmap_bytes,
                                  const uint64_t* __restrict__ qwide) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= bv.n_terms) {
    return;
  }
  const uint32_t k0 = bv.term_koff[t];
  const uint32_t k1 = bv.term_koff[t + 1];
  for (uint32_t i = k0; i < k1; ++i) {
    lookup_one(iv.term_keys, iv.term_off, iv.n_terms, keys, i, bv.key_list, bv.key_len, bv.key_glen, iv.wide_keys,
               iv.wide_words, qwide);
  }
  term_plan_one(bv, t, compute_df, all_valid_utf8, raw_flags, bitmap_bytes, kDfUnit);
  for (uint32_t i = k0; i < k1; ++i) {
    key_ref_one(iv, bv.key_list, bv.key_len, i, bv.key_ref);
  }
}

// df tile descriptors: term t owns tiles [off[t], off[t+1]); one warp per term.
__global__ void fill_df_desc_kernel(const uint64_t* __restrict__ off, const uint32_t* __restrict__ term_koff,
                                    const KeyRef* __restrict__ key_ref, uint32_t n_terms, DfTileDesc* __restrict__ out) {
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= n_terms) {
    return;
  }
  const uint64_t b = off[t];
  const uint64_t e = off[t + 1];
  if (e == b) {
    return;
  }
  DfTileDesc d{};
  d.k0 = term_koff[t];
  d.k1 = term_koff[t + 1];
  d.term = t;
  d.drv_p = key_ref[d.k0].p;
  d.drv_len = key_ref[d.k0].len;
  for (uint64_t i = b + (threadIdx.x & 31u); i < e; i += 32) {
    d.tile = static_cast<uint32_t>(i - b);
    out[i] = d;
  }
}

// Two lower bounds in one list by a whole warp, the probes of both searches in flight together.
// `start`: a position all elements before which are known to be < v1 (0 = no knowledge). Consecutive pieces of a
// driver list ask for consecutive doc ranges, so the bounds of one piece start where the previous piece's ended: one
// probe step over the next ~1000 entries usually brackets both bounds (two dependent loads instead of log33(n) + 1).
__device__ __forceinline__ void warp_lower_bound_pair(const uint32_t* __restrict__ p, uint32_t n, uint32_t v1, uint32_t v2,
                                                      uint32_t* out1, uint32_t* out2, uint32_t start = 0) {
  const unsigned lane = threadIdx.x & 31u;
  uint32_t lo1 = start, hi1 = n, lo2 = start, hi2 = n;
  if (start != 0 && n - start > 32) {
    const uint32_t w = min(n - start, 1056u);
    const uint32_t idx = start + static_cast<uint32_t>((static_cast<uint64_t>(lane + 1) * w) / 33);
    const uint32_t x = __ldg(p + idx);  // one load serves both bounds
    {
      const int c = __popc(__ballot_sync(0xffffffffu, x < v1));
      const uint32_t prev = __shfl_sync(0xffffffffu, idx, c > 0 ? c - 1 : 0);
      const uint32_t next = __shfl_sync(0xffffffffu, idx, c < 32 ? c : 31);
      if (c > 0) lo1 = prev + 1;
      if (c < 32) hi1 = next;  // else: beyond the window, the rest of the list stays open
    }
    {
      const int c = __popc(__ballot_sync(0xffffffffu, x < v2));
      const uint32_t prev = __shfl_sync(0xffffffffu, idx, c > 0 ? c - 1 : 0);
      const uint32_t next = __shfl_sync(0xffffffffu, idx, c < 32 ? c : 31);
      if (c > 0) lo2 = prev + 1;
      if (c < 32) hi2 = next;
    }
  }
  while (hi1 - lo1 > 32 || hi2 - lo2 > 32) {
    const bool a1 = hi1 - lo1 > 32;
    const bool a2 = hi2 - lo2 > 32;
    const uint32_t idx1 = lo1 + static_cast<uint32_t>((static_cast<uint64_t>(lane + 1) * (hi1 - lo1)) / 33);
    const uint32_t idx2 = lo2 + static_cast<uint32_t>((static_cast<uint64_t>(lane + 1) * (hi2 - lo2)) / 33);
    const uint32_t x1 = a1 ? __ldg(p + idx1) : 0u;
    const uint32_t x2 = a2 ? __ldg(p + idx2) : 0u;
    if (a1) {
      const int c = __popc(__ballot_sync(0xffffffffu, x1 < v1));  // monotone in the lane index
      const uint32_t prev = __shfl_sync(0xffffffffu, idx1, c > 0 ? c - 1 : 0);
      const uint32_t next = __shfl_sync(0xffffffffu, idx1, c < 32 ? c : 31);
      if (c > 0) lo1 = prev + 1;
      if (c < 32) hi1 = next;
    }
    if (a2) {
      const int c = __popc(__ballot_sync(0xffffffffu, x2 < v2));
      const uint32_t prev = __shfl_sync(0xffffffffu, idx2, c > 0 ? c - 1 : 0);
      const uint32_t next = __shfl_sync(0xffffffffu, idx2, c < 32 ? c : 31);
      if (c > 0) lo2 = prev + 1;
      if (c < 32) hi2 = next;
    }
  }
  const uint32_t i1 = lo1 + lane;
  const uint32_t i2 = lo2 + lane;
  const uint32_t y1 = i1 < hi1 ? __ldg(p + i1) : 0xFFFFFFFFu;
  const uint32_t y2 = i2 < hi2 ? __ldg(p + i2) : 0xFFFFFFFFu;
  *out1 = lo1 + __popc(__ballot_sync(0xffffffffu, i1 < hi1 && y1 < v1));
  *out2 = lo2 + __popc(__ballot_sync(0xffffffffu, i2 < hi2 && y2 < v2));
}

// tile -> segment map: segment g owns tiles [off[g], off[g+1]); one warp per segment.
__global__ void fill_tile_map_kernel(const uint64_t* __restrict__ off, uint32_t n_segments, uint32_t* __restrict__ out) {
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= n_segments) {
    return;
  }
  const uint64_t b = off[g];
  const uint64_t e = off[g + 1];
  for (uint64_t i = b + (threadIdx.x & 31u); i < e; i += 32) {
    out[i] = g;
  }
}

// ------------------------------------------------------------------ df tiles
// Verified document frequency of a term (PopulateTermDocumentFrequency, search_pipeline.cpp:542-565): the
// documents of SearchAnd(term n-grams) whose text contains the term. A CTA covers one 1024-entry tile of the
// term's shortest list, but every WARP owns its own 128 entries and runs to completion without any block barrier:
// narrow the other lists to the warp's doc range (32-ary bound search), stage that sub-range in the warp's
// shared-memory slice with coalesced loads, search it there, then scan the survivors' text one thread per document.
#ifndef MGX_DF_OCC
#define MGX_DF_OCC 3  // resident CTAs per SM the register allocation of the df kernels is sized for (80 registers: measured best)
#endif
constexpr int kWarpItems = 4;                       // driver entries per lane
constexpr int kWarpTile = 32 * kWarpItems;          // 128 entries per warp
constexpr uint32_t kWarpStageCap = 512;             // staged entries of one other list per warp (2 KB)
constexpr uint32_t kDfHintLists = 2;                // other lists whose last sub-range end a warp carries from piece to piece
static_assert(kWarpTile * (kTileThreads / 32) == kTile, "warps must tile the CTA tile exactly");

__device__ __forceinline__ void stat_add(const BatchView& bv, int slot, unsigned long long v) {
  atomicAdd(bv.stats + slot * kStatStripes + (blockIdx.x & (kStatStripes - 1)), v);
}
__device__ __forceinline__ void stat_add_at(const BatchView& bv, int slot, uint32_t stripe, unsigned long long v) {
  atomicAdd(bv.stats + slot * kStatStripes + (stripe & (kStatStripes - 1)), v);
}

// 12 bytes of text at an arbitrary byte address as three little-endian words (the arena is padded by 64 bytes)
__device__ __forceinline__ void load_text_words(const uint8_t* __restrict__ text, uint64_t at, uint32_t (&x)[3]) {
  const uint64_t base = at & ~3ULL;
  const uint32_t sh = static_cast<uint32_t>(at - base) * 8u;
  const uint32_t a0 = __ldg(reinterpret_cast<const uint32_t*>(text + base));
  const uint32_t a1 = __ldg(reinterpret_cast<const uint32_t*>(text + base + 4));
  const uint32_t a2 = __ldg(reinterpret_cast<const uint32_t*>(text + base + 8));
  const uint32_t a3 = __ldg(reinterpret_cast<const uint32_t*>(text + base + 12));
  x[0] = __funnelshift_r(a0, a1, sh);
  x[1] = __funnelshift_r(a1, a2, sh);
  x[2] = __funnelshift_r(a2, a3, sh);
}

static_assert(kWarpStageCap * sizeof(uint32_t) >= kStageBuf, "the text staging buffer aliases the list staging buffer");
constexpr uint32_t kDfFewMembers = 16;         // at most this many candidates of a chunk: searched one by one
#ifndef MGX_DF_PIECE_ITEMS
#define MGX_DF_PIECE_ITEMS 4  // payload words a lane loads per piece (measured: 2 -> 1.50, 4 -> 1.29, 8 -> 1.37, 16 -> 1.67 ms per C2 batch)
#endif
constexpr int kDfPieceItems = MGX_DF_PIECE_ITEMS;
constexpr uint32_t kDfPiece = 32 * kDfPieceItems;      // entries a warp filters at a time
constexpr uint32_t kDfSetCap = kWarpTile + kDfPiece;   // collected candidates per warp: fewer than 128 waiting + one piece

// What is the same for every piece of one (term, driver list) pair.
struct DfTerm {
  const uint32_t* drv_pos;   // payload of the driver n-gram in each document: first / second occurrence + neighbour
  const uint32_t* drv_pos2;  // signatures (Index::d_post_pos), nullptr when the index carries none
  uint32_t o1;               // byte offset of the driver n-gram's first occurrence in the term
  uint32_t m_term;           // its occurrences in the term: 1, 2, or 3 = "three or more"
  uint32_t tl;               // term bytes
  uint32_t sig_want;         // neighbour signatures the term asks for, in payload position
  uint32_t sig_mask;
  bool toff_ok;
  bool prefilter;
};

__device__ __forceinline__ DfTerm df_term_consts(const IndexView& iv, const BatchView& bv, const ListRef& drv, uint32_t t,
                                                 uint32_t k0) {
  DfTerm c;
  c.drv_pos = iv.post_pos != nullptr ? iv.post_pos + (drv.p - iv.postings) : nullptr;
  c.drv_pos2 = iv.post_pos2 != nullptr ? iv.post_pos2 + (drv.p - iv.postings) : nullptr;
  const uint32_t toff_raw = bv.key_toff[k0];
  c.toff_ok = (toff_raw & 0xFFFFu) != kNoTermOffset;
  c.o1 = toff_raw & kTermOffsetMask;
  c.m_term = (toff_raw & 0xFFFFu) >> kTermCountShift;
  c.tl = bv.term_boff[t + 1] - bv.term_boff[t];
  // Signature pre-filter (mgx_internal.cuh "Neighbour signatures"): the character of the term right after / before
  // the driver n-gram must be what the index recorded next to one of the document's occurrences. It runs on the
  // payload alone, BEFORE the probes of the other lists and the text comparison, and leaves them the true matches
  // plus a 2^-bits share of the rest.
  c.sig_want = 0;
  c.sig_mask = 0;
  if (c.toff_ok && c.tl != 0) {
    if ((toff_raw & kToffHasNext) != 0 && iv.sig_next_bits > 0) {
      c.sig_want |= sig_of_hash((toff_raw >> kToffNextShift) & 0x7Fu, iv.sig_next_bits) << 16;
      c.sig_mask |= 0x7Fu << 16;
    }
    if ((toff_raw & kToffHasPrev) != 0 && iv.sig_prev_bits > 0) {
      c.sig_want |= sig_of_hash((toff_raw >> kToffPrevShift) & 0x7Fu, iv.sig_prev_bits) << 24;
      c.sig_mask |= 0x7Fu << 24;
    }
  }
  // without signatures in the payload (or with MGX_DF_NO_SIG) the stage runs as it did before them: every entry goes
  // through the membership stage, so its candidates are exactly the documents of SearchAnd(n-grams) -- the set the
  // reference scans, which is what the accounting of B_df is defined on
  c.prefilter = c.toff_ok && c.tl != 0 && c.drv_pos != nullptr && (iv.sig_next_bits | iv.sig_prev_bits) != 0;
  return c;
}

// per-lane partial sums of one unit, reduced once at its end
struct DfAcc {
  uint32_t hits = 0;
  uint32_t cand = 0;     // lane 0 only
  uint32_t scanned = 0;  // lane 0 only
  unsigned long long text_bytes = 0;
};

// One warp, one piece of kDfPiece entries of a term's shortest list (entries [pe0, pe0 + kDfPiece) of drv, pe0 =
// e0 + rel0): the entries the payload cannot rule out are appended, in order, to `cand` at n_set -- as their distance
// from e0, the documents and payloads are fetched again (from L2) for the few that stay. Lane l takes the entries
// l, l + 32, ...: every load is one coalesced line, all of them are in flight before the first is looked at, and the
// second-occurrence words (needed by roughly one entry in ten) follow in a second wave.
__device__ __forceinline__ uint32_t df_collect_piece(const ListRef& drv, const DfTerm& tc, uint64_t e0, uint32_t rel0,
                                                     uint64_t unit_end, uint32_t* cand, uint32_t n_set) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint64_t pe0 = e0 + rel0;
  const uint32_t tile_n = static_cast<uint32_t>(umin64(kDfPiece, unit_end - pe0));
  uint32_t keep = 0;
  if (tc.drv_pos == nullptr) {
#pragma unroll
    for (int k = 0; k < kDfPieceItems; ++k) {
      keep |= (k * 32 + lane < tile_n) ? (1u << k) : 0u;
    }
  } else {
    uint32_t valid = (1u << kDfPieceItems) - 1u;  // entries of this lane inside the piece
    if (tile_n < kDfPiece) {
      valid = 0;
#pragma unroll
      for (int k = 0; k < kDfPieceItems; ++k) {
        valid |= (k * 32 + lane < tile_n) ? (1u << k) : 0u;
      }
    }
    uint32_t pay1[kDfPieceItems];
#pragma unroll
    for (int k = 0; k < kDfPieceItems; ++k) {
      pay1[k] = ((valid >> k) & 1u) ? __ldg(tc.drv_pos + pe0 + k * 32 + lane) : 0u;
    }
    if (!tc.prefilter) {
      keep = valid;
    } else {
      // one recorded occurrence (the common entry): kept iff the term has the n-gram once, the occurrence leaves room
      // for the bytes of the term before it and the signatures agree -- two logic operations and two compares;
      // everything else (a second occurrence, an unrecorded offset) is sorted out below
      const uint32_t lim = kPosUnknown - tc.o1;           // (p - o1) < lim  <=>  o1 <= p < 0x7FFF
      const uint32_t fast_mask = tc.sig_mask | kPosMulti;  // single occurrence + signatures
      const bool once = tc.m_term <= 1;
      uint32_t other = 0;
#pragma unroll
      for (int k = 0; k < kDfPieceItems; ++k) {
        const uint32_t p = pay1[k] & kPosUnknown;
        const bool in_range = p - tc.o1 < lim;
        const bool sig_ok = ((pay1[k] ^ tc.sig_want) & fast_mask) == 0;
        keep |= (once && in_range && sig_ok) ? (1u << k) : 0u;
        other |= ((pay1[k] & kPosMulti) != 0 || p == kPosUnknown) ? (1u << k) : 0u;
      }
      keep &= valid;
      other &= valid;
      if (__any_sync(0xffffffffu, other != 0)) {
        uint32_t pay2[kDfPieceItems];
#pragma unroll
        for (int k = 0; k < kDfPieceItems; ++k) {
          const bool multi = ((other >> k) & 1u) != 0 && (pay1[k] & kPosMulti) != 0;
          pay2[k] = multi ? __ldg(tc.drv_pos2 + pe0 + k * 32 + lane) : 0u;
        }
#pragma unroll
        for (int k = 0; k < kDfPieceItems; ++k) {
          if ((other >> k) & 1u) {
            const uint32_t p1 = pay1[k] & kPosUnknown;
            const uint32_t p2 = pay2[k] & kPosUnknown;
            if (p1 == kPosUnknown || (pay2[k] & kPosMulti) != 0 || p2 == kPosUnknown) {
              keep |= 1u << k;  // unrecorded, or three and more occurrences: the scanning path decides
            } else {
              const bool good1 = p1 - tc.o1 < lim && ((pay1[k] ^ tc.sig_want) & tc.sig_mask) == 0;
              const bool good2 = p2 - tc.o1 < lim && ((pay2[k] ^ tc.sig_want) & tc.sig_mask) == 0;
              keep |= (tc.m_term <= 2 && (good1 || good2)) ? (1u << k) : 0u;  // exactly two, both recorded
            }
          }
        }
      }
    }
  }
  uint32_t n = 0;
#pragma unroll
  for (int k = 0; k < kDfPieceItems; ++k) {
    const unsigned m = __ballot_sync(0xffffffffu, (keep >> k) & 1u);
    if ((keep >> k) & 1u) {
      cand[n_set + n + __popc(m & lt_mask)] = rel0 + k * 32 + lane;
    }
    n += __popc(m);
  }
  __syncwarp();
  return n;
}

// four lower bounds in one ascending global array, their probes in flight together
__device__ __forceinline__ void lower_bound_x4(const uint32_t* __restrict__ p, uint32_t n, const uint32_t (&v)[kWarpItems],
                                               uint32_t mask, uint32_t (&pos)[kWarpItems]) {
  uint32_t base[kWarpItems];
#pragma unroll
  for (int k = 0; k < kWarpItems; ++k) {
    base[k] = 0;
  }
  uint32_t len = n;
  while (len > 1) {
    const uint32_t half = len >> 1;
#pragma unroll
    for (int k = 0; k < kWarpItems; ++k) {
      if ((mask >> k) & 1u) {
        base[k] = __ldg(p + base[k] + half - 1) < v[k] ? base[k] + half : base[k];
      }
    }
    len -= half;
  }
#pragma unroll
  for (int k = 0; k < kWarpItems; ++k) {
    pos[k] = base[k];
    if (((mask >> k) & 1u) && n != 0 && __ldg(p + base[k]) < v[k]) {
      pos[k] = base[k] + 1;
    }
  }
}

// One warp, the first n (<= 128) collected candidates, ascending in surv / spos: membership in the term's other
// lists, then the verification of what is left. stage is the calling warp's own shared-memory slice.
__device__ __forceinline__ void df_process_chunk(const IndexView& iv, const BatchView& bv, const ListRef& drv,
                                                 const DfTerm& tc, uint32_t t, uint32_t k0, uint32_t k1, uint64_t e0,
                                                 uint32_t* surv, uint32_t* spos, uint32_t n_in, uint32_t* stage,
                                                 uint32_t (&hint)[kDfHintLists], DfAcc& acc) {
  const unsigned lane = threadIdx.x & 31u;
  uint32_t my_doc[kWarpItems];
  uint32_t my_pos[kWarpItems];
  uint32_t alive = 0;   // candidates that go through the membership stage
  uint32_t direct = 0;  // candidates that do not
#pragma unroll
  for (int k = 0; k < kWarpItems; ++k) {
    const uint32_t i = lane * kWarpItems + k;
    my_doc[k] = kNone;
    my_pos[k] = kPosUnknown;
    if (i < n_in) {
      const uint64_t e = e0 + surv[i];  // surv holds the candidates' distances from e0 on entry
      my_doc[k] = __ldg(drv.p + e);
      if (tc.drv_pos != nullptr) {
        const uint32_t pay1 = __ldg(tc.drv_pos + e);
        const uint32_t pay2 = (pay1 & kPosMulti) != 0 ? __ldg(tc.drv_pos2 + e) : 0u;
        my_pos[k] = (pay1 & 0xFFFFu) | (pay2 << 16);
      }
      // Candidates whose occurrences of the driver n-gram are all recorded (one or two) are settled by comparing
      // the text at those places: a document that holds the term there holds every n-gram of the term (tokenizer
      // agreement, the precondition of toff_ok), so looking them up in the other lists first would only be a
      // filter -- and the signatures have filtered better already. They skip the membership stage.
      const uint32_t p1 = my_pos[k] & 0xFFFFu;
      const uint32_t p2 = my_pos[k] >> 16;
      const bool known = (p1 & kPosUnknown) != kPosUnknown &&
                         ((p1 & kPosMulti) == 0 || ((p2 & kPosMulti) == 0 && (p2 & kPosUnknown) != kPosUnknown));
      if (tc.prefilter && known) {
        direct |= 1u << k;
      } else {
        alive |= 1u << k;
      }
    }
  }
  const uint32_t n_member = __reduce_add_sync(0xffffffffu, __popc(alive));
  // document range of the candidates that take the membership stage (ascending in the lane-major order)
  uint32_t dmin = 0xFFFFFFFFu;
  uint32_t dmax = 0;
#pragma unroll
  for (int k = 0; k < kWarpItems; ++k) {
    if ((alive >> k) & 1u) {
      dmin = min(dmin, my_doc[k]);
      dmax = max(dmax, my_doc[k]);
    }
  }
  dmin = __reduce_min_sync(0xffffffffu, dmin);
  dmax = __reduce_max_sync(0xffffffffu, dmax);
  for (uint32_t j = k0 + 1; j < k1 && n_member != 0; ++j) {
    const uint4* rp = reinterpret_cast<const uint4*>(bv.key_ref + j);
    const uint4 r0 = __ldg(rp);
    const uint4 r1 = __ldg(rp + 1);
    ListRef l;
    l.p = reinterpret_cast<const uint32_t*>(static_cast<uintptr_t>(r0.x) | (static_cast<uintptr_t>(r0.y) << 32));
    l.bm = reinterpret_cast<const uint32_t*>(static_cast<uintptr_t>(r0.z) | (static_cast<uintptr_t>(r0.w) << 32));
    l.len = r1.x;
    if (l.bm != nullptr) {
#pragma unroll
      for (int k = 0; k < kWarpItems; ++k) {
        if (((alive >> k) & 1u) && ((__ldg(l.bm + (my_doc[k] >> 5)) >> (my_doc[k] & 31)) & 1u) == 0) {
          alive &= ~(1u << k);
        }
      }
    } else if (n_member <= kDfFewMembers) {
      // a handful of candidates: one search of the whole list each, no narrowing, no staging
      uint32_t pos[kWarpItems];
      lower_bound_x4(l.p, l.len, my_doc, alive, pos);
#pragma unroll
      for (int k = 0; k < kWarpItems; ++k) {
        if (((alive >> k) & 1u) && (pos[k] >= l.len || __ldg(l.p + pos[k]) != my_doc[k])) {
          alive &= ~(1u << k);
        }
      }
    } else {
      uint32_t lo = 0;
      uint32_t hi = 0;
      const uint32_t hj = j - (k0 + 1);  // the first kDfHintLists other lists remember where the last chunk ended
      warp_lower_bound_pair(l.p, l.len, dmin, dmax + 1u, &lo, &hi, hj < kDfHintLists ? hint[hj] : 0u);
      if (hj < kDfHintLists) {
        hint[hj] = hi;
      }
      const uint32_t cnt = hi - lo;
      if (cnt == 0) {
        alive = 0;
      } else if (cnt <= kWarpStageCap && cnt <= 8 * n_in) {
        // staging pays when the sub-range is not much longer than the candidates are many
        for (uint32_t i = lane; i < cnt; i += 32) {
          stage[i] = __ldg(l.p + lo + i);
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < kWarpItems; ++k) {
          if (((alive >> k) & 1u) && !sorted_contains(stage, cnt, my_doc[k])) {
            alive &= ~(1u << k);
          }
        }
        __syncwarp();
      } else {
        uint32_t pos[kWarpItems];
        lower_bound_x4(l.p + lo, cnt, my_doc, alive, pos);
#pragma unroll
        for (int k = 0; k < kWarpItems; ++k) {
          if (((alive >> k) & 1u) && (pos[k] >= cnt || __ldg(l.p + lo + pos[k]) != my_doc[k])) {
            alive &= ~(1u << k);
          }
        }
      }
    }
    if (__ballot_sync(0xffffffffu, alive != 0) == 0) {
      break;
    }
  }
  alive |= direct;
  if (__ballot_sync(0xffffffffu, alive != 0) == 0) {
    return;
  }
  // survivors -> the front of the warp's list (order irrelevant for a count)
  const uint32_t mine = __popc(alive);
  uint32_t inc = mine;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, s);
    if (lane >= static_cast<unsigned>(s)) {
      inc += o;
    }
  }
  const uint32_t n = __shfl_sync(0xffffffffu, inc, 31);
  uint32_t w = inc - mine;
  __syncwarp();  // every lane holds its entries in registers
#pragma unroll
  for (int k = 0; k < kWarpItems; ++k) {
    if ((alive >> k) & 1u) {
      surv[w] = my_doc[k];
      spos[w] = my_pos[k];
      ++w;
    }
  }
  __syncwarp();
  const uint32_t tl = tc.tl;
  const bool toff_ok = tc.toff_ok;
  const uint32_t o1 = tc.o1;
  const uint32_t m_term = tc.m_term;
  const uint8_t* term = bv.term_bytes + bv.term_boff[t];
  const TermRegs tregs = load_term_regs(term, tl);

  // ---- pass 1, one lane per candidate. The driver n-gram occurs m times in the term (first at byte offset o1) and
  // c times in the document (positions recorded for c <= 2). An occurrence of the term puts m occurrences of the
  // n-gram into the document, the first of them at (term start + o1): so m > c rules the document out, and for
  // c <= 2 the term can only start at (p_a - o1) for a recorded position p_a -- one or two comparisons instead of a
  // scan. Documents with three or more occurrences (or no recorded position) are compacted to the front of the
  // list for the scanning pass.
  // Every lane takes up to kWarpItems candidates and keeps their loads in flight together: all offsets first, then all
  // first-occurrence comparisons (two dependent memory round trips per chunk instead of per 32 candidates).
  uint32_t n_scan = 0;
  {
    uint32_t c_doc_[kWarpItems];
    uint32_t pp_[kWarpItems];
    uint64_t b_[kWarpItems];
    uint32_t len_[kWarpItems];
    uint32_t state = 0;  // per candidate 2 bits: 0 nothing to do, 1 compare, 2 scan
#pragma unroll
    for (int r = 0; r < kWarpItems; ++r) {
      const uint32_t sidx = r * 32 + lane;
      c_doc_[r] = 0;
      pp_[r] = 0;
      b_[r] = 0;
      len_[r] = 0;
      if (sidx < n) {
        c_doc_[r] = surv[sidx];
        pp_[r] = spos[sidx];
        b_[r] = iv.text_off[c_doc_[r]];
        len_[r] = static_cast<uint32_t>(iv.text_off[c_doc_[r] + 1] - b_[r]);
      }
    }
    __syncwarp();  // surv is rewritten below
    uint32_t x_[kWarpItems][3];
#pragma unroll
    for (int r = 0; r < kWarpItems; ++r) {
      const uint32_t sidx = r * 32 + lane;
      x_[r][0] = x_[r][1] = x_[r][2] = 0;
      if (sidx < n) {
        acc.text_bytes += len_[r];
        const uint32_t p1 = pp_[r] & 0xFFFFu;
        const uint32_t p2 = pp_[r] >> 16;
        const bool two = (p1 & kPosMulti) != 0;
        const uint32_t c_doc = !two ? 1u : ((p2 & kPosMulti) == 0 ? 2u : 3u);
        if (!toff_ok || tl == 0 || (p1 & kPosUnknown) == kPosUnknown || (two && (p2 & kPosUnknown) == kPosUnknown)) {
          state |= 2u << (2 * r);
        } else if (m_term <= c_doc) {
          // three or more occurrences (state 3): the two recorded ones are compared like the others; only when
          // neither holds the term is the text behind the second one scanned
          state |= (c_doc == 3 ? 3u : 1u) << (2 * r);
          const uint32_t at = p1 & kPosUnknown;
          if (at >= o1 && at - o1 + tl <= len_[r]) {
            load_text_words(iv.text, b_[r] + (at - o1), x_[r]);
          } else {
            state |= 0x100u << r;  // the first occurrence cannot hold the term
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kWarpItems; ++r) {
      bool scan = ((state >> (2 * r)) & 3u) == 2u;
      uint32_t scan_from = 0;  // first start position the scan has to look at
      if ((((state >> (2 * r)) & 3u) & 1u) != 0) {
        const uint32_t p1 = pp_[r] & 0xFFFFu;
        const uint32_t p2 = pp_[r] >> 16;
        bool found = false;
        if (((state >> (8 + r)) & 1u) == 0) {
          bool ok = ((x_[r][0] ^ tregs.w[0]) & tregs.m[0]) == 0;
          if (tregs.nw > 1) {
            ok = ok && ((x_[r][1] ^ tregs.w[1]) & tregs.m[1]) == 0;
          }
          if (tregs.nw > 2) {
            ok = ok && ((x_[r][2] ^ tregs.w[2]) & tregs.m[2]) == 0;
          }
          if (ok && tl > 12) {
            ok = term_tail_matches(iv.text, b_[r] + ((p1 & kPosUnknown) - o1), term, tl);
          }
          found = ok;
        }
        if (!found && (p1 & kPosMulti) != 0) {  // the second recorded occurrence
          const uint32_t at = p2 & kPosUnknown;
          if (at >= o1 && at - o1 + tl <= len_[r]) {
            const uint64_t start = b_[r] + (at - o1);
            uint32_t x[3];
            load_text_words(iv.text, start, x);
            bool ok = ((x[0] ^ tregs.w[0]) & tregs.m[0]) == 0;
            if (tregs.nw > 1) {
              ok = ok && ((x[1] ^ tregs.w[1]) & tregs.m[1]) == 0;
            }
            if (tregs.nw > 2) {
              ok = ok && ((x[2] ^ tregs.w[2]) & tregs.m[2]) == 0;
            }
            if (ok && tl > 12) {
              ok = term_tail_matches(iv.text, start, term, tl);
            }
            found = ok;
          }
        }
        acc.hits += found ? 1u : 0u;
        if (!found && ((state >> (2 * r)) & 3u) == 3u) {
          // a further occurrence of the term puts the driver n-gram behind its second recorded occurrence
          scan = true;
          scan_from = (p2 & kPosUnknown) >= o1 ? (p2 & kPosUnknown) - o1 + 1u : 0u;
        }
      }
      const unsigned scan_mask = __ballot_sync(0xffffffffu, scan);
      if (scan) {
        const uint32_t at = n_scan + __popc(scan_mask & ((1u << lane) - 1u));
        surv[at] = c_doc_[r];
        spos[at] = scan_from;
      }
      n_scan += __popc(scan_mask);
    }
    __syncwarp();
  }

  // ---- pass 2: the remaining candidates are scanned, 4 lanes per document
  constexpr int kGroup = 4;                 // lanes per document
  constexpr int kDocsPerIter = 32 / kGroup; // documents per warp iteration
  const unsigned group = lane / kGroup;
  for (uint32_t s0 = 0; s0 < n_scan; s0 += kDocsPerIter) {
    const uint32_t s = s0 + group;
    uint64_t b = 0;
    uint32_t len = 0;
    bool slow = false;
    if (s < n_scan) {
      const uint32_t doc = surv[s];
      b = iv.text_off[doc];
      len = static_cast<uint32_t>(iv.text_off[doc + 1] - b);
      if (tl > kThreadScanMaxTerm || len > kThreadScanMaxDoc) {
        slow = true;
        len = 0;  // handled below
      } else {
        const uint32_t from = min(spos[s], len);  // start positions before it are settled
        b += from;
        len -= from;
      }
    }
    const bool found = group_contains_term<kGroup>(iv.text, b, len, term, tl, tregs);
    if (found && (lane & (kGroup - 1)) == 0) {
      ++acc.hits;
    }
    // rare: long terms / very long documents go through the warp-cooperative scanner, one document at a time
    unsigned slow_mask = __ballot_sync(0xffffffffu, slow && (lane & (kGroup - 1)) == 0);
    while (slow_mask != 0) {
      const uint32_t src = static_cast<uint32_t>(__ffs(static_cast<int>(slow_mask))) - 1u;
      slow_mask &= slow_mask - 1;
      // the warp's list staging buffer is free after the membership phase: it stages the text now
      DocText d = doc_open(iv, surv[s0 + src / kGroup], reinterpret_cast<uint8_t*>(stage));
      const uint32_t c = doc_count_term(d, term, tl, true);
      __syncwarp();
      if (lane == src) {
        acc.hits += c != 0 ? 1u : 0u;
      }
    }
  }
  if (lane == 0) {
    acc.cand += n;
    acc.scanned += n_scan;
  }
}

// One warp, n_entries entries of a term's shortest list from entry e0 on: the whole df work of that range.
// The pieces' candidates are collected first (with signatures in the payload most entries never leave the collecting
// loop) and go through membership and verification 128 at a time, so the cost of narrowing the other lists is paid
// per 128 CANDIDATES instead of per 128 entries. stage / surv / spos are the calling warp's own shared-memory
// slices; stat_stripe spreads the accounting atomics.
__device__ __forceinline__ void df_warp_unit(const IndexView& iv, const BatchView& bv, const ListRef& drv, uint32_t t,
                                             uint32_t k0, uint32_t k1, uint64_t e0, uint32_t n_entries, uint32_t* stage,
                                             uint32_t* surv, uint32_t* spos, uint32_t stat_stripe) {
  const unsigned lane = threadIdx.x & 31u;
  if (e0 >= drv.len) {
    return;  // no block-wide barrier is used below, so a warp may leave early
  }
  const uint64_t unit_end = umin64(drv.len, e0 + n_entries);
  const uint32_t n_pieces = static_cast<uint32_t>((unit_end - e0 + kDfPiece - 1) / kDfPiece);
  const DfTerm tc = df_term_consts(iv, bv, drv, t, k0);
  uint32_t hint[kDfHintLists] = {0, 0};
  DfAcc acc;
  uint32_t n_set = 0;
#pragma unroll 1
  for (uint32_t piece = 0; piece < n_pieces; ++piece) {
    n_set += df_collect_piece(drv, tc, e0, piece * kDfPiece, unit_end, surv, n_set);
    const bool last = piece + 1 == n_pieces;
#pragma unroll 1
    while (n_set >= static_cast<uint32_t>(kWarpTile) || (last && n_set > 0)) {
      const uint32_t c = min(n_set, static_cast<uint32_t>(kWarpTile));
      df_process_chunk(iv, bv, drv, tc, t, k0, k1, e0, surv, spos, c, stage, hint, acc);
#ifdef MGX_DF_NO_HINT
      hint[0] = hint[1] = 0;
#endif
      __syncwarp();
      const uint32_t rest = n_set - c;  // moves down in ascending order: a lane never overwrites what another still reads
      for (uint32_t i0 = 0; i0 < rest; i0 += 32) {
        const uint32_t v = i0 + lane < rest ? surv[c + i0 + lane] : 0u;
        __syncwarp();
        if (i0 + lane < rest) {
          surv[i0 + lane] = v;
        }
      }
      __syncwarp();
      n_set = rest;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    acc.hits += __shfl_xor_sync(0xffffffffu, acc.hits, s);
    acc.text_bytes += __shfl_xor_sync(0xffffffffu, acc.text_bytes, s);
  }
  if (lane == 0) {
    if (acc.hits != 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(bv.t_df + t), static_cast<unsigned long long>(acc.hits));
    }
    stat_add_at(bv, kStatDfBytes, stat_stripe, acc.text_bytes);
    stat_add_at(bv, kStatDfCandidates, stat_stripe, acc.cand);
    if (acc.scanned != 0) {
      stat_add_at(bv, kStatDfScanned, stat_stripe, acc.scanned);
    }
  }
}

__global__ void __launch_bounds__(kTileThreads, MGX_DF_OCC) df_tile_kernel(IndexView iv, BatchView bv) {
  __shared__ __align__(16) uint32_t s_stage[kTileThreads / 32][kWarpStageCap];
  __shared__ uint32_t s_surv[kTileThreads / 32][kDfSetCap];
  __shared__ uint32_t s_spos[kTileThreads / 32][kWarpTile];
  const unsigned warp = threadIdx.x >> 5;
  // one 32-byte descriptor instead of tile -> term -> keys -> dictionary -> offsets
  const uint4* dp = reinterpret_cast<const uint4*>(bv.df_tile_desc + blockIdx.x);
  const uint4 d0 = __ldg(dp);
  const uint4 d1 = __ldg(dp + 1);
  ListRef drv;
  drv.p = reinterpret_cast<const uint32_t*>(static_cast<uintptr_t>(d0.x) | (static_cast<uintptr_t>(d0.y) << 32));
  drv.len = d0.z;
  drv.bm = nullptr;
  const uint64_t tile = d1.z;
  df_warp_unit(iv, bv, drv, d0.w, d1.x, d1.y, tile * kTile + static_cast<uint64_t>(warp) * kWarpTile, kWarpTile, s_stage[warp],
               s_surv[warp], s_spos[warp], blockIdx.x);
}

// Streamed form (no host read-back of the tile count): persistent warps pull UNITS of kDfUnit entries of the terms'
// shortest lists from a device-side counter. unit -> term by a 32-ary search over the units' prefix sums
// (t_df_tile_off, written by the planning tail); the next unit is fetched before the current one is processed, so
// the atomic's latency hides behind the work.
__device__ __forceinline__ uint32_t warp_upper_bound_u64(const uint64_t* __restrict__ p, uint32_t n, uint64_t v) {
  // number of elements <= v in the ascending array p[0, n)
  const unsigned lane = threadIdx.x & 31u;
  uint32_t lo = 0;
  uint32_t hi = n;
  while (hi - lo > 32) {
    const uint32_t size = hi - lo;
    const uint32_t idx = lo + static_cast<uint32_t>((static_cast<uint64_t>(lane + 1) * size) / 33);
    const bool le = p[idx] <= v;
    const int c = __popc(__ballot_sync(0xffffffffu, le));  // monotone in the lane index
    const uint32_t idx_prev = __shfl_sync(0xffffffffu, idx, c > 0 ? c - 1 : 0);
    const uint32_t idx_next = __shfl_sync(0xffffffffu, idx, c < 32 ? c : 31);
    if (c > 0) {
      lo = idx_prev + 1;
    }
    if (c < 32) {
      hi = idx_next;
    }
  }
  const uint32_t i = lo + lane;
  const bool le = i < hi && p[i] <= v;
  return lo + __popc(__ballot_sync(0xffffffffu, le));
}

__global__ void __launch_bounds__(kTileThreads, MGX_DF_OCC) df_units_kernel(IndexView iv, BatchView bv) {
  __shared__ __align__(16) uint32_t s_stage[kTileThreads / 32][kWarpStageCap];
  __shared__ uint32_t s_surv[kTileThreads / 32][kDfSetCap];
  __shared__ uint32_t s_spos[kTileThreads / 32][kWarpTile];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t n_units = bv.launch[kLaunchDfUnits];
  uint32_t next = 0;
  if (lane == 0) {
    next = atomicAdd(bv.launch + kLaunchDfNext, 1u);
  }
  for (;;) {
    const uint32_t unit = __shfl_sync(0xffffffffu, next, 0);
    if (unit >= n_units) {
      return;
    }
    if (lane == 0) {
      next = atomicAdd(bv.launch + kLaunchDfNext, 1u);  // consumed in the next round
    }
    // largest t with off[t] <= unit (terms without units repeat the offset of their successor)
    const uint32_t t = warp_upper_bound_u64(bv.t_df_tile_off, bv.n_terms + 1, unit) - 1u;
    const uint32_t k0 = bv.term_koff[t];
    const uint32_t k1 = bv.term_koff[t + 1];
    const uint4* rp = reinterpret_cast<const uint4*>(bv.key_ref + k0);
    const uint4 r0 = __ldg(rp);
    const uint4 r1 = __ldg(rp + 1);
    ListRef drv;
    drv.p = reinterpret_cast<const uint32_t*>(static_cast<uintptr_t>(r0.x) | (static_cast<uintptr_t>(r0.y) << 32));
    drv.len = r1.x;
    drv.bm = nullptr;
    const uint64_t e0 = (static_cast<uint64_t>(unit) - bv.t_df_tile_off[t]) * kDfUnit;
    df_warp_unit(iv, bv, drv, t, k0, k1, e0, kDfUnit, s_stage[warp], s_surv[warp], s_spos[warp], unit);
    __syncwarp();
  }
}

// ------------------------------------------------------------------ streaming df pass
// df(term) = number of documents whose text contains the term. For a term that is valid UTF-8, in an index whose
// documents are all valid UTF-8 and whose tokeniser agrees with the query side, this equals the reference's
// definition (documents of SearchAnd(term n-grams) whose text contains the term, search_pipeline.cpp:554-564):
// a document that contains the term contains every n-gram window of it. So ONE coalesced pass over the arena
// serves every eligible term of the batch (see query.cuh for the two-level table). Its cost is ~4 ps per text
// byte whatever the batch holds, so it only pays for very large batches (df_mode_kernel decides).
constexpr int kStreamThreads = 256;
constexpr int kStreamRounds = kTextTileBytes / (kStreamThreads * 16);
constexpr uint32_t kStreamPosCap = 2048;    // stage-1 survivors of one tile
constexpr uint32_t kStreamSetSlots = 2048;  // (term, document) -> first position, open addressing
constexpr uint32_t kStreamSetFill = 1400;   // stop inserting beyond this many pairs (tile handled stand-alone)
constexpr uint32_t kStreamDocCap = 512;     // documents of one tile whose offsets are staged
constexpr unsigned long long kStreamSetEmpty = ~0ULL;
static_assert(kStreamRounds * kStreamThreads * 16 == kTextTileBytes, "tile must be whole rounds");
static_assert(kTextTileBytes <= (1u << 13), "records keep 13 bits of tile position");

// any occurrence of term starting in [from, to)?  (bytes are read past `to` when an occurrence straddles it)
__device__ __forceinline__ bool occurs_between(const uint8_t* __restrict__ text, uint64_t from, uint64_t to,
                                               const uint8_t* __restrict__ term, uint32_t tl) {
  const uint8_t t0 = __ldg(term);
  for (uint64_t j = from; j < to; ++j) {
    if (__ldg(text + j) != t0) {
      continue;
    }
    uint32_t i = 1;
    while (i < tl && __ldg(text + j + i) == __ldg(term + i)) {
      ++i;
    }
    if (i == tl) {
      return true;
    }
  }
  return false;
}

__device__ __forceinline__ bool bytes_equal_from(const uint8_t* __restrict__ text, uint64_t pos,
                                                 const uint8_t* __restrict__ term, uint32_t from, uint32_t tl) {
  for (uint32_t i = from; i < tl; ++i) {
    if (__ldg(text + pos + i) != __ldg(term + i)) {
      return false;
    }
  }
  return true;
}

// Stage 2 at one text position: every table entry whose key bytes equal the text there -> hit(entry, term).
template <typename Hit>
__device__ __forceinline__ void stream_probe_position(const uint8_t* __restrict__ text, uint64_t gpos,
                                                      const BatchView& bv, Hit&& hit) {
  const uint64_t base = gpos & ~3ULL;  // the arena is padded by 64 bytes
  const uint32_t sh = static_cast<uint32_t>(gpos - base) * 8u;
  const uint32_t a0 = __ldg(reinterpret_cast<const uint32_t*>(text + base));
  const uint32_t a1 = __ldg(reinterpret_cast<const uint32_t*>(text + base + 4));
  const uint32_t a2 = __ldg(reinterpret_cast<const uint32_t*>(text + base + 8));
  const uint32_t a3 = __ldg(reinterpret_cast<const uint32_t*>(text + base + 12));
  const uint32_t t0 = __funnelshift_r(a0, a1, sh);
  const uint32_t t1 = __funnelshift_r(a1, a2, sh);
  const uint32_t t2 = __funnelshift_r(a2, a3, sh);
  uint32_t classes = bv.stream_len12_mask;
  while (classes != 0) {
    const uint32_t len = static_cast<uint32_t>(__ffs(static_cast<int>(classes))) - 1u;
    classes &= classes - 1u;
    const uint32_t w0 = t0 & low_bytes_mask(len);
    const uint32_t w1 = len > 4 ? (t1 & low_bytes_mask(len - 4)) : 0u;
    const uint32_t w2 = len > 8 ? (t2 & low_bytes_mask(len - 8)) : 0u;
    const uint32_t slot = __ldg(bv.stream_slots + (stream_hash(w0, w1, w2, len) & bv.stream_slot_mask));
    const uint32_t cnt = slot & 0xFFu;
    const uint32_t first = slot >> 8;
    for (uint32_t c = 0; c < cnt; ++c) {
      const uint4 e = __ldg(reinterpret_cast<const uint4*>(bv.stream_entries + first + c));
      if (e.x == w0 && e.y == w1 && e.z == w2 && (e.w & 0xFFu) == len) {
        hit(first + c, e.w >> 8);
      }
    }
  }
}

// Exact stand-alone handling of one key match (used when a tile's shared-memory structures overflow).
__device__ void stream_slow_hit(const IndexView& iv, const BatchView& bv, uint64_t gpos, uint32_t tid) {
  const uint8_t* term = bv.term_bytes + bv.term_boff[tid];
  const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
  uint64_t lo = 0;  // largest d with text_off[d] <= gpos
  uint64_t hi = iv.n_docs;
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (iv.text_off[mid] <= gpos) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  const uint64_t doc_b = iv.text_off[lo];
  const uint64_t doc_e = iv.text_off[lo + 1];
  if (gpos + tl > doc_e || (tl > kStreamKeyBytes && !bytes_equal_from(iv.text, gpos, term, kStreamKeyBytes, tl))) {
    return;
  }
  if (occurs_between(iv.text, doc_b, gpos, term, tl)) {
    return;  // not the first occurrence in this document
  }
  atomicAdd(reinterpret_cast<unsigned long long*>(bv.t_df + tid), 1ULL);
}

// Stage 1 over the 16 start positions of one 16-byte chunk: pass(pos_in_chunk) for every character start whose
// first bytes hit the shared-memory filter. The window is kept aligned to the current position by funnel shifts.
template <typename Pass>
__device__ __forceinline__ void stream_filter_chunk(const uint8_t* __restrict__ text, uint64_t byte0, uint32_t npos,
                                                    uint32_t len8_mask, const uint32_t* s_bloom, Pass&& pass) {
  const uint4 a = ld16(text + byte0);
  const uint4 b = ld16(text + byte0 + 16);  // the arena is padded by 64 bytes
  uint32_t w0 = a.x, w1 = a.y, w2 = a.z, w3 = a.w, w4 = b.x, w5 = b.y;  // bytes 0..23: position 15 needs 15..22
  uint32_t lead = 0;  // bit i: byte i of the chunk starts a character (is not a continuation byte)
  {
    const uint32_t c0 = __vcmpeq4(a.x & 0xC0C0C0C0u, 0x80808080u);
    const uint32_t c1 = __vcmpeq4(a.y & 0xC0C0C0C0u, 0x80808080u);
    const uint32_t c2 = __vcmpeq4(a.z & 0xC0C0C0C0u, 0x80808080u);
    const uint32_t c3 = __vcmpeq4(a.w & 0xC0C0C0C0u, 0x80808080u);
    const uint32_t n0 = ((((c0 >> 7) & 0x01010101u) * 0x00204081u) >> 21) & 0xFu;
    const uint32_t n1 = ((((c1 >> 7) & 0x01010101u) * 0x00204081u) >> 21) & 0xFu;
    const uint32_t n2 = ((((c2 >> 7) & 0x01010101u) * 0x00204081u) >> 21) & 0xFu;
    const uint32_t n3 = ((((c3 >> 7) & 0x01010101u) * 0x00204081u) >> 21) & 0xFu;
    lead = ~(n0 | (n1 << 4) | (n2 << 8) | (n3 << 12)) & ((npos >= 16 ? 0x10000u : (1u << npos)) - 1u);
  }
  uint32_t cur = 0;
  while (lead != 0) {
    const uint32_t z = static_cast<uint32_t>(__ffs(static_cast<int>(lead))) - 1u;
    lead &= lead - 1u;
    uint32_t d = z - cur;
    cur = z;
    while (d > 4) {  // only after a run of more than three continuation bytes (never in valid UTF-8)
      w0 = w1, w1 = w2, w2 = w3, w3 = w4, w4 = w5, w5 = 0;
      d -= 4;
    }
    const uint32_t sh = 8 * d;  // clamp mode: a shift of 32 moves whole words
    w0 = __funnelshift_rc(w0, w1, sh);
    w1 = __funnelshift_rc(w1, w2, sh);
    w2 = __funnelshift_rc(w2, w3, sh);
    w3 = __funnelshift_rc(w3, w4, sh);
    w4 = __funnelshift_rc(w4, w5, sh);
    w5 = __funnelshift_rc(w5, 0u, sh);
    bool hit = false;
    uint32_t classes = len8_mask;
    do {  // warp-uniform: usually one class
      const uint32_t len8 = static_cast<uint32_t>(__ffs(static_cast<int>(classes))) - 1u;
      classes &= classes - 1u;
      const uint32_t k0 = w0 & low_bytes_mask(len8);
      const uint32_t k1 = len8 > 4 ? (w1 & low_bytes_mask(len8 - 4)) : 0u;
      const uint32_t bit = stream_filter_bit(k0, k1, len8);
      hit = hit || ((s_bloom[bit >> 5] >> (bit & 31)) & 1u) != 0;
    } while (classes != 0);
    if (hit) {
      pass(z);
    }
  }
}

__global__ void __launch_bounds__(kStreamThreads) df_stream_kernel(IndexView iv, BatchView bv) {
  __shared__ uint16_t s_pos[kStreamPosCap];                 // stage-1 survivors (position in the tile)
  __shared__ unsigned long long s_set[kStreamSetSlots];     // entry:19 | doc in tile:10 | first pos:13
  __shared__ int32_t s_docrel[kStreamDocCap + 1];           // start of document (first_doc + i) relative to the tile
  __shared__ uint32_t s_bloom[kStreamBloomWords];
  __shared__ uint32_t s_npos;
  __shared__ uint32_t s_nset;
  __shared__ uint32_t s_overflow;
  unsigned long long hits = 0;
  if (bv.launch != nullptr && (bv.df_mode[0] == 0 || bv.launch[kLaunchOverflow] != 0)) {
    return;  // streamed batch: launched before the choice was known; the candidate tiles were chosen (or nothing runs)
  }
  const uint32_t len8_mask = bv.stream_len8_mask;
  for (uint32_t i = threadIdx.x; i < kStreamBloomWords; i += kStreamThreads) {
    s_bloom[i] = __ldg(bv.stream_bloom + i);
  }
  for (uint64_t tile = blockIdx.x; tile < iv.n_text_tiles; tile += gridDim.x) {
    const uint64_t tile_b = tile * kTextTileBytes;
    const uint32_t tile_len = static_cast<uint32_t>(umin64(kTextTileBytes, iv.text_bytes - tile_b));
    const uint32_t first_doc = __ldg(iv.tile_first_doc + tile);
    // documents that start inside the tile: up to the first document of the next tile (+1 entry for the end)
    const uint32_t last_doc =
        tile + 1 < iv.n_text_tiles ? __ldg(iv.tile_first_doc + tile + 1) : static_cast<uint32_t>(iv.n_docs - 1);
    const uint32_t n_tab = min(kStreamDocCap, last_doc - first_doc + 1u) + 1u;  // entries [0, n_tab) are loaded
    for (uint32_t i = threadIdx.x; i < kStreamSetSlots; i += kStreamThreads) {
      s_set[i] = kStreamSetEmpty;
    }
    for (uint32_t i = threadIdx.x; i <= kStreamDocCap; i += kStreamThreads) {
      const uint64_t d = static_cast<uint64_t>(first_doc) + i;
      const uint64_t off = (i < n_tab && d <= iv.n_docs) ? iv.text_off[d] : ~0ULL;
      // relative start, saturated: far before the tile -> -2^30, not loaded / beyond the arena -> INT_MAX
      int32_t rel;
      if (off == ~0ULL || off >= tile_b + 0x40000000ULL) {
        rel = 0x7FFFFFFF;
      } else if (off + 0x40000000ULL < tile_b) {
        rel = -0x40000000;
      } else {
        rel = static_cast<int32_t>(static_cast<int64_t>(off) - static_cast<int64_t>(tile_b));
      }
      s_docrel[i] = rel;
    }
    if (threadIdx.x == 0) {
      s_npos = 0;
      s_nset = 0;
      s_overflow = 0;
    }
    __syncthreads();
    // ---- stage 1: filter every character start of the tile
#pragma unroll 1
    for (int r = 0; r < kStreamRounds; ++r) {
      const uint32_t c0 = (static_cast<uint32_t>(r) * kStreamThreads + threadIdx.x) * 16u;
      if (c0 < tile_len) {
        stream_filter_chunk(iv.text, tile_b + c0, min(16u, tile_len - c0), len8_mask, s_bloom, [&](uint32_t pos) {
          const uint32_t at = atomicAdd(&s_npos, 1u);
          if (at < kStreamPosCap) {
            s_pos[at] = static_cast<uint16_t>(c0 + pos);
          } else {
            s_overflow = 1;
          }
        });
      }
    }
    __syncthreads();
    // the staged offsets cover the tile when the entry after the last document starting in it was loaded
    bool standalone = s_overflow != 0 || last_doc - first_doc + 1u > kStreamDocCap;
    const uint32_t n_pos = min(s_npos, kStreamPosCap);
    if (!standalone) {
      // ---- stage 2 (dense): table probe, rest of the term, document, first position per (term, document)
      for (uint32_t i = threadIdx.x; i < n_pos; i += kStreamThreads) {
        const uint32_t pos = s_pos[i];
        stream_probe_position(iv.text, tile_b + pos, bv, [&](uint32_t entry, uint32_t tid) {
          const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
          uint32_t lo = 0;  // largest j with s_docrel[j] <= pos  (s_docrel[0] <= 0 <= pos)
          uint32_t hi = kStreamDocCap;  // s_docrel[hi] >= tile_len > pos
          while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_docrel[mid] <= static_cast<int32_t>(pos)) {
              lo = mid;
            } else {
              hi = mid;
            }
          }
          if (static_cast<int64_t>(pos) + tl > static_cast<int64_t>(s_docrel[lo + 1])) {
            return;  // would run past the end of its document
          }
          if (tl > kStreamKeyBytes &&
              !bytes_equal_from(iv.text, tile_b + pos, bv.term_bytes + bv.term_boff[tid], kStreamKeyBytes, tl)) {
            return;
          }
          if (*reinterpret_cast<volatile uint32_t*>(&s_overflow) != 0) {
            return;
          }
          const unsigned long long hi_key =
              (static_cast<unsigned long long>(entry) << 23) | (static_cast<unsigned long long>(lo) << 13);
          const unsigned long long val = hi_key | pos;
          uint32_t slot = (entry * 0x9E3779B1u + lo * 0x85EBCA77u) & (kStreamSetSlots - 1);
          for (;;) {
            const unsigned long long cur = atomicCAS(&s_set[slot], kStreamSetEmpty, val);
            if (cur == kStreamSetEmpty) {
              if (atomicAdd(&s_nset, 1u) >= kStreamSetFill) {
                s_overflow = 1;
              }
              break;
            }
            if ((cur >> 13) == (hi_key >> 13)) {
              atomicMin(&s_set[slot], val);
              break;
            }
            slot = (slot + 1) & (kStreamSetSlots - 1);
          }
        });
      }
      __syncthreads();
      standalone = s_overflow != 0;
    }
    if (standalone) {
      // rare: too many survivors / pairs or too many tiny documents in this tile -> exact stand-alone handling
#pragma unroll 1
      for (int r = 0; r < kStreamRounds; ++r) {
        const uint32_t c0 = (static_cast<uint32_t>(r) * kStreamThreads + threadIdx.x) * 16u;
        if (c0 < tile_len) {
          stream_filter_chunk(iv.text, tile_b + c0, min(16u, tile_len - c0), len8_mask, s_bloom, [&](uint32_t pos) {
            const uint64_t gpos = tile_b + c0 + pos;
            stream_probe_position(iv.text, gpos, bv,
                                  [&](uint32_t, uint32_t tid) { stream_slow_hit(iv, bv, gpos, tid); });
          });
        }
      }
      __syncthreads();
      continue;
    }
    // ---- count: every occupied slot is the first occurrence of its term in its document within this tile
    for (uint32_t i = threadIdx.x; i < kStreamSetSlots; i += kStreamThreads) {
      const unsigned long long v = s_set[i];
      if (v == kStreamSetEmpty) {
        continue;
      }
      const uint32_t entry = static_cast<uint32_t>(v >> 23);
      const uint32_t j = static_cast<uint32_t>(v >> 13) & 0x3FFu;
      const uint32_t tid = bv.stream_entries[entry].len_term >> 8;
      if (s_docrel[j] < 0) {
        // the document began in an earlier tile: look for an occurrence that starts before this tile
        const uint64_t doc_b = iv.text_off[static_cast<uint64_t>(first_doc) + j];
        const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
        if (occurs_between(iv.text, doc_b, tile_b, bv.term_bytes + bv.term_boff[tid], tl)) {
          continue;
        }
      }
      atomicAdd(reinterpret_cast<unsigned long long*>(bv.t_df + tid), 1ULL);
      ++hits;
    }
    __syncthreads();
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    hits += __shfl_xor_sync(0xffffffffu, hits, s);
  }
  if ((threadIdx.x & 31u) == 0 && hits != 0) {
    stat_add(bv, kStatStreamHits, hits);
  }
}

__global__ void df_to_slots_kernel(const uint64_t* __restrict__ t_df, const uint32_t* __restrict__ slot_tid,
                                   uint32_t n_slots, uint64_t* __restrict__ out) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_slots) {
    out[s] = t_df[slot_tid[s]];
  }
}

__global__ void slots_to_df_kernel(const uint64_t* __restrict__ in, const uint32_t* __restrict__ slot_tid,
                                   uint32_t n_slots, uint64_t* __restrict__ t_df) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_slots) {
    t_df[slot_tid[s]] = in[s];  // all slots of one term carry the same value
  }
}

// ------------------------------------------------------------------ query planning
__device__ __forceinline__ void query_plan_one(const IndexView& iv, const BatchView& bv, uint32_t q) {
  uint32_t flags = bv.q_host_flags[q];
  if ((flags & kQProgram) != 0) {
    // Boolean program (QueryNode::Evaluate, query_ast.cpp:67-161). Every result satisfies each conjunct term, so the
    // shortest list of the conjunct with the smallest estimate drives; without a usable conjunct (OR / NOT at the
    // root, text-only terms) every document of the shard is evaluated.
    const uint32_t l0p = bv.q_loff[q];
    uint64_t best = kEstNone;
    uint32_t best_k = kNone;
    bool empty = bv.q_poff[q + 1] == bv.q_poff[q];
    for (uint32_t i = bv.q_coff[q]; i < bv.q_coff[q + 1]; ++i) {
      const uint32_t tid = bv.q_conj[i];
      const uint32_t k0 = bv.term_koff[tid];
      const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
      if (bv.term_koff[tid + 1] == k0) {
        empty = empty || tl == 0;  // an empty term matches nothing (substring_search.h:27-29)
        continue;                  // text-only term: cannot drive
      }
      const uint64_t est = bv.t_est[tid];
      if (est == 0) {
        empty = true;  // an n-gram of a required term is not in the index
      } else if (est < best) {
        best = est;
        best_k = k0;  // lists of a term are sorted by length: the first is the shortest
      }
    }
    uint32_t driver_len = 0;
    uint32_t n = 0;
    if (!empty) {
      if (best_k != kNone) {
        bv.q_list[l0p] = bv.key_list[best_k];
        bv.q_list_len[l0p] = bv.key_len[best_k];
        n = 1;
        driver_len = bv.key_len[best_k];
      } else {
        flags |= kQDriverAll;
        driver_len = static_cast<uint32_t>(iv.n_docs);
      }
    } else {
      flags |= kQEmpty;
    }
    bv.q_nlists[q] = n;
    bv.q_flags[q] = flags;
    bv.q_driver_len[q] = driver_len;
    bv.q_ntiles[q] = (driver_len + kTile - 1) / kTile;
    return;
  }
  const bool any_mode = (flags & kQAnyMode) != 0;
  const uint32_t t0 = bv.q_toff[q];
  const uint32_t t1 = bv.q_toff[q + 1];
  // terms by ascending estimated size, equal sizes keep query order
  // (std::sort on <= 16 elements is an insertion sort, search_pipeline.cpp:2012-2014)
  for (uint32_t i = t0 + 1; i < t1; ++i) {
    const uint32_t tid = bv.q_tids[i];
    const uint64_t est = bv.t_est[tid];
    uint32_t j = i;
    while (j > t0 && bv.t_est[bv.q_tids[j - 1]] > est) {
      bv.q_tids[j] = bv.q_tids[j - 1];
      --j;
    }
    bv.q_tids[j] = tid;
  }
  bool empty = false;
  bool text_only = false;
  const uint32_t l0 = bv.q_loff[q];
  uint32_t n = 0;
  for (uint32_t i = t0; i < t1; ++i) {
    const uint32_t tid = bv.q_tids[i];
    const uint64_t est = bv.t_est[tid];
    const uint32_t nk = bv.term_koff[tid + 1] - bv.term_koff[tid];
    const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
    if (!any_mode && (est == 0 || est == kEstNone) && (nk > 0 || tl == 0)) {
      empty = true;  // Execute :804-810
    }
    if (nk == 0 && tl > 0) {
      text_only = true;  // SearchNormalizedSubstring fallback, query/substring_search.h:24-42
    }
    for (uint32_t k = bv.term_koff[tid]; k < bv.term_koff[tid + 1]; ++k) {
      const uint32_t li = bv.key_list[k];
      if (li == kNone) {
        continue;
      }
      bool dup = false;
      for (uint32_t x = 0; x < n; ++x) {
        dup |= bv.q_list[l0 + x] == li;
      }
      if (!dup) {
        bv.q_list[l0 + n] = li;
        bv.q_list_len[l0 + n] = bv.key_len[k];
        ++n;
      }
    }
  }
  for (uint32_t i = 1; i < n; ++i) {  // lists by ascending length
    const uint32_t li = bv.q_list[l0 + i];
    const uint32_t ln = bv.q_list_len[l0 + i];
    uint32_t j = i;
    while (j > 0 && bv.q_list_len[l0 + j - 1] > ln) {
      bv.q_list[l0 + j] = bv.q_list[l0 + j - 1];
      bv.q_list_len[l0 + j] = bv.q_list_len[l0 + j - 1];
      --j;
    }
    bv.q_list[l0 + j] = li;
    bv.q_list_len[l0 + j] = ln;
  }
  uint32_t driver_len = 0;
  if (flags & kQDriverExplicit) {
    driver_len = bv.explicit_n;
  } else if (t1 == t0) {
    empty = true;  // no search terms: Execute leaves the result empty (:813)
  } else if (any_mode) {
    if (n == 0 || n < bv.q_threshold[q]) {  // fewer existing lists than the threshold (index.cpp:520-523)
      empty = true;
    } else {
      flags |= kQDriverAll;
      driver_len = static_cast<uint32_t>(iv.n_docs);
    }
  } else if (n > 0) {
    driver_len = bv.q_list_len[l0];
  } else if (text_only) {
    flags |= kQDriverAll;
    driver_len = static_cast<uint32_t>(iv.n_docs);
  } else {
    empty = true;
  }
  if (empty) {
    flags |= kQEmpty;
    driver_len = 0;
  }
  if (!empty) {
    const uint64_t bitmap_bytes = (iv.n_docs + 7) / 8;
    unsigned long long bytes = 0;
    for (uint32_t i = 0; i < n; ++i) {
      bytes += umin64(4ULL * bv.q_list_len[l0 + i], bitmap_bytes);
    }
    atomicAdd(bv.stats + kStatIntersectLists * kStatStripes + (q & (kStatStripes - 1)), bytes);
  }
  bv.q_nlists[q] = n;
  bv.q_flags[q] = flags;
  bv.q_driver_len[q] = driver_len;
  bv.q_ntiles[q] = (driver_len + kTile - 1) / kTile;
  if (driver_len != 0) {
    atomicAdd(bv.stats + kStatDriverEntries * kStatStripes + (q & (kStatStripes - 1)),
              static_cast<unsigned long long>(driver_len));
  }
}

__global__ void query_plan_kernel(IndexView iv, BatchView bv) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < bv.n_queries) {
    query_plan_one(iv, bv, q);
  }
}

// Exclusive prefix sums of in(0..n) into out[0..n] (out[n] = total) by one 256-thread CTA; values written by other
// CTAs of the same launch are read around L1.
struct ScanSmem {
  uint64_t warp_sums[8];
  uint64_t carry;
};
template <typename In>
__device__ __forceinline__ uint64_t block_exclusive_scan(In in, uint32_t n, uint64_t* __restrict__ out, ScanSmem& sm) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    sm.carry = 0;
  }
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 256u * 4u) {
    const uint32_t i0 = base + threadIdx.x * 4u;
    uint32_t v[4];
    uint64_t sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = i0 + j < n ? in(i0 + j) : 0u;
      sum += v[j];
    }
    uint64_t inc = sum;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const uint64_t o = __shfl_up_sync(0xffffffffu, inc, s);
      if (lane >= static_cast<unsigned>(s)) {
        inc += o;
      }
    }
    if (lane == 31) {
      sm.warp_sums[warp] = inc;
    }
    __syncthreads();
    uint64_t prefix = sm.carry;
    uint64_t tile_total = 0;
#pragma unroll
    for (unsigned w = 0; w < 8; ++w) {
      if (w < warp) {
        prefix += sm.warp_sums[w];
      }
      tile_total += sm.warp_sums[w];
    }
    uint64_t run = prefix + inc - sum;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j < n) {
        out[i0 + j] = run;
        run += v[j];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      sm.carry += tile_total;
    }
    __syncthreads();
  }
  const uint64_t total = sm.carry;
  if (threadIdx.x == 0) {
    out[n] = total;
  }
  __syncthreads();
  return total;
}

// Streamed planning, stage 2: the batch-wide df decision and the plan of every query (one thread each), and -- in the
// CTA that finishes last -- the prefix sums that place the work items, the work sizes, and the capacity check. After
// this kernel the batch is fully planned and nothing has visited the host.
//   cap_tiles: intersect tiles the stream's workspace holds (tile counters and record slots)
//   group_tiles: tiles per top-k pre-reduction group, 0 = no pre-reduction
__global__ void __launch_bounds__(256)
plan_queries_kernel(IndexView iv, BatchView bv, const uint8_t* __restrict__ raw_flags, int df_choice, int force,
                    int have_table, uint32_t cap_tiles, uint32_t group_tiles) {
  __shared__ ScanSmem s_scan;
  __shared__ uint32_t s_last;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (df_choice != 0 && i < bv.n_terms) {
    df_mode_one(bv, i, raw_flags, iv.text_bytes, force, have_table);
  }
  if (i < bv.n_queries) {
    query_plan_one(iv, bv, i);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    s_last = atomicAdd(bv.launch + kLaunchTicket, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last == 0) {
    return;
  }
  __threadfence();
  const uint32_t* t_units = bv.t_df_tiles;
  const uint32_t* q_ntiles = bv.q_ntiles;
  const uint64_t n_units =
      block_exclusive_scan([&](uint32_t k) { return __ldcg(t_units + k); }, bv.n_terms, bv.t_df_tile_off, s_scan);
  const uint64_t n_tiles =
      block_exclusive_scan([&](uint32_t k) { return __ldcg(q_ntiles + k); }, bv.n_queries, bv.q_tile_off, s_scan);
  const uint64_t n_groups = block_exclusive_scan(
      [&](uint32_t k) {
        const uint32_t nt = __ldcg(q_ntiles + k);
        return group_tiles != 0 && nt >= 2 * group_tiles ? (nt + group_tiles - 1) / group_tiles : 0u;
      },
      bv.n_queries, bv.q_group_off, s_scan);
  if (threadIdx.x == 0) {
    const bool overflow = n_tiles > cap_tiles || n_units >= 0x7FFFFFFFULL || n_groups >= 0x7FFFFFFFULL;
    bv.launch[kLaunchNeedTiles] = static_cast<uint32_t>(n_tiles > 0xFFFFFFFFULL ? 0xFFFFFFFFULL : n_tiles);
    bv.launch[kLaunchOverflow] = overflow ? 1u : 0u;
    bv.launch[kLaunchDfUnits] = overflow ? 0u : static_cast<uint32_t>(n_units);
    bv.launch[kLaunchAndTiles] = overflow ? 0u : static_cast<uint32_t>(n_tiles);
    bv.launch[kLaunchGroups] = overflow ? 0u : static_cast<uint32_t>(n_groups);
  }
}

// BM25Scorer::ComputeIDF, bm25_scorer.cpp:14-25 — one value per (query, term) in planner order.
__global__ void idf_kernel(const uint32_t* __restrict__ q_tids, uint32_t n, const uint64_t* __restrict__ t_df,
                           uint64_t total_docs, double* __restrict__ idf) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  double v = 0.0;
  if (total_docs != 0) {
    uint64_t df = t_df[q_tids[i]];
    if (df > total_docs) {
      df = total_docs;
    }
    const double nn = static_cast<double>(total_docs);
    const double dd = static_cast<double>(df);
    const double num = __dadd_rn(__dsub_rn(nn, dd), 0.5);
    const double den = __dadd_rn(dd, 0.5);
    v = log(__dadd_rn(__ddiv_rn(num, den), 1.0));
  }
  idf[i] = v;
}

// Sharded batches, after the df exchange: the reference orders a query's terms by their estimated size over the WHOLE
// index (search_pipeline.cpp:2012-2014) and BM25Scorer adds the terms' contributions in that order, so a shard must
// not add them in the order of its LOCAL sizes -- the FP64 sum would differ in the last bit between shard counts.
// key_glen holds every key's posting size summed over the shards (it travels with the df all-reduce); one thread per
// query re-derives the global estimates (min over the term's keys, 0 if a key is missing everywhere, search_pipeline.cpp:
// 583-593), orders the query's terms by them (stable: equal sizes keep query order, as the planner does), and writes
// the IDFs in that order.
__global__ void global_order_kernel(BatchView bv, uint64_t total_docs, double* __restrict__ idf) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= bv.n_queries) {
    return;
  }
  const uint32_t t0 = bv.q_toff[q];
  const uint32_t t1 = bv.q_toff[q + 1];
  auto global_est = [&](uint32_t tid) {
    const uint32_t k0 = bv.term_koff[tid];
    const uint32_t k1 = bv.term_koff[tid + 1];
    uint64_t est = kEstNone;
    for (uint32_t k = k0; k < k1; ++k) {
      est = umin64(est, bv.key_glen[k]);
    }
    return est;
  };
  for (uint32_t i = t0; i < t1; ++i) {
    const uint32_t tid = bv.q_tids0[i];
    const uint64_t est = global_est(tid);
    uint32_t j = i;
    while (j > t0 && global_est(bv.q_tids[j - 1]) > est) {
      bv.q_tids[j] = bv.q_tids[j - 1];
      --j;
    }
    bv.q_tids[j] = tid;
  }
  for (uint32_t i = t0; i < t1; ++i) {
    double v = 0.0;
    if (total_docs != 0) {
      uint64_t df = bv.t_df[bv.q_tids[i]];
      if (df > total_docs) {
        df = total_docs;
      }
      const double nn = static_cast<double>(total_docs);
      const double dd = static_cast<double>(df);
      const double num = __dadd_rn(__dsub_rn(nn, dd), 0.5);
      const double den = __dadd_rn(dd, 0.5);
      v = log(__dadd_rn(__ddiv_rn(num, den), 1.0));
    }
    idf[i] = v;
  }
}

// ------------------------------------------------------------------ intersect + score tiles
// Block-wide ordered compaction helper: each thread contributes `cnt` (<= kTileItems)
// items; returns the exclusive offset and writes the block total to *total.
__device__ __forceinline__ uint32_t block_offsets(uint32_t cnt, uint32_t* s_warp, uint32_t* total) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  uint32_t inc = cnt;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, s);
    if (lane >= static_cast<unsigned>(s)) {
      inc += o;
    }
  }
  if (lane == 31) {
    s_warp[warp] = inc;
  }
  __syncthreads();
  uint32_t prefix = 0;
  uint32_t tot = 0;
#pragma unroll
  for (int w = 0; w < kTileThreads / 32; ++w) {
    const uint32_t c = s_warp[w];
    if (static_cast<unsigned>(w) < warp) {
      prefix += c;
    }
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return prefix + inc - cnt;
}

// ------------------------------------------------------------------ top-k
// Sort key: (s, d) compared lexicographically, LARGER IS BETTER.
//   DESC: s = ord(score), d = doc         (higher score first; ties: higher doc id first)
//   ASC : s = ~ord(score), d = ~doc       (lower score first; ties: lower doc id first)
// which is ResultSorter::SortByScore's comparator (result_sorter.cpp:681-686).
__device__ __forceinline__ uint64_t ord_f64(double x) {
  const uint64_t bits = static_cast<uint64_t>(__double_as_longlong(x));
  return (bits >> 63) ? ~bits : (bits | 0x8000000000000000ULL);
}
__device__ __forceinline__ double unord_f64(uint64_t o) {
  const uint64_t bits = (o >> 63) ? (o & 0x7FFFFFFFFFFFFFFFULL) : ~o;
  return __longlong_as_double(static_cast<long long>(bits));
}

struct SortKey {
  uint64_t s;
  uint32_t d;
};
__device__ __forceinline__ bool key_greater(const SortKey& a, const SortKey& b) {
  return a.s > b.s || (a.s == b.s && a.d > b.d);
}
__device__ __forceinline__ SortKey make_sort_key(double score, uint32_t doc, bool descending) {
  SortKey k;
  const uint64_t o = ord_f64(score);
  k.s = descending ? o : ~o;
  k.d = descending ? doc : ~doc;
  return k;
}
// digit `pass` (0 = most significant) of the 96-bit key, 8 bits each
__device__ __forceinline__ uint32_t key_digit(const SortKey& k, int pass) {
  return pass < 8 ? static_cast<uint32_t>((k.s >> (56 - 8 * pass)) & 0xFF)
                  : static_cast<uint32_t>((k.d >> (24 - 8 * (pass - 8))) & 0xFF);
}
// does the key match `prefix` on its first `pass` digits
__device__ __forceinline__ bool key_has_prefix(const SortKey& k, const SortKey& prefix, int pass) {
  if (pass <= 0) {
    return true;
  }
  if (pass < 8) {
    const int sh = 64 - 8 * pass;
    return (k.s >> sh) == (prefix.s >> sh);
  }
  if (k.s != prefix.s) {
    return false;
  }
  if (pass == 8) {
    return true;
  }
  const int sh = 32 - 8 * (pass - 8);
  return (k.d >> sh) == (prefix.d >> sh);
}

constexpr uint32_t kTopkSmem = 1024;  // >= kMaxTopK: the selected keys always fit

// descending bitonic sort of n_pow2 keys in shared memory (256 threads)
__device__ void bitonic_sort_desc(SortKey* keys, uint32_t n_pow2) {
  for (uint32_t size = 2; size <= n_pow2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < n_pow2 / 2; i += blockDim.x) {
        const uint32_t lo = 2 * i - (i & (stride - 1));
        const uint32_t hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const SortKey a = keys[lo];
        const SortKey b = keys[hi];
        const bool swap = desc ? key_greater(b, a) : key_greater(a, b);
        if (swap) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

// Shared state of a block-wide windowed top-k (256 threads).
struct TopkShared {
  SortKey keys[kTopkSmem];
  uint32_t hist[256];
  uint32_t cnt;
  SortKey prefix;
  uint64_t need;
};

// The kth-best key (1-based, SortByScore order = descending SortKey order) of the records `scan` enumerates:
// MSB-first radix select, twelve 8-bit passes over the 96-bit key. scan(f) calls f(key) for every record, the
// records strided over the block's threads. Keys are unique per query (doc ids are).
template <class Scan>
__device__ SortKey block_select_kth(TopkShared& sh, Scan scan, uint64_t kth) {
  __syncthreads();
  if (threadIdx.x == 0) {
    sh.prefix.s = 0;
    sh.prefix.d = 0;
    sh.need = kth;
  }
  __syncthreads();
  for (int pass = 0; pass < 12; ++pass) {
    sh.hist[threadIdx.x] = 0;
    __syncthreads();
    const SortKey prefix = sh.prefix;
    scan([&](const SortKey k) {
      if (key_has_prefix(k, prefix, pass)) {
        atomicAdd(&sh.hist[key_digit(k, pass)], 1u);
      }
    });
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t need = sh.need;
      int digit = 255;
      for (; digit > 0; --digit) {
        if (sh.hist[digit] >= need) {
          break;
        }
        need -= sh.hist[digit];
      }
      sh.need = need;
      if (pass < 8) {
        sh.prefix.s |= static_cast<uint64_t>(digit) << (56 - 8 * pass);
      } else {
        sh.prefix.d |= static_cast<uint32_t>(digit) << (24 - 8 * (pass - 8));
      }
    }
    __syncthreads();
  }
  return sh.prefix;
}

// Records ranked [want_begin, want_begin + n_out) of n_records records in SortByScore order: emit(i, key) for the
// i-th record of the window. Up to kTopkSmem records: one shared-memory sort. A window that ends within the best
// kTopkSmem: one select, the head is sorted and cut. Anything else (ResultSorter::SortByScore takes any offset, and
// limit 0 = everything, result_sorter.cpp:689-710): the window is produced in runs of kTopkSmem ranks, each bounded
// by two selected keys (the lower bound of one run is the upper bound of the next), gathered and sorted.
template <class Scan, class Emit>
__device__ void block_topk_window(TopkShared& sh, Scan scan, uint64_t n_records, uint64_t want_begin, uint64_t n_out,
                                  Emit emit) {
  uint64_t begin = want_begin;
  const uint64_t end = want_begin + n_out;
  if (n_records <= kTopkSmem || end <= kTopkSmem) {
    begin = 0;  // one run from the best record on; the window is cut from it
  }
  SortKey hi;
  hi.s = 0;
  hi.d = 0;
  bool has_hi = false;
  if (begin > 0) {
    hi = block_select_kth(sh, scan, begin);
    has_hi = true;
  }
  for (uint64_t c0 = begin; c0 < end; c0 += kTopkSmem) {
    const uint64_t c1 = umin64(end, c0 + kTopkSmem);
    SortKey lo;
    lo.s = 0;
    lo.d = 0;
    const bool has_lo = c1 < n_records;
    if (has_lo) {
      lo = block_select_kth(sh, scan, c1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      sh.cnt = 0;
    }
    __syncthreads();
    scan([&](const SortKey k) {
      if ((!has_hi || key_greater(hi, k)) && (!has_lo || !key_greater(lo, k))) {
        const uint32_t pos = atomicAdd(&sh.cnt, 1u);
        if (pos < kTopkSmem) {
          sh.keys[pos] = k;
        }
      }
    });
    __syncthreads();
    const uint32_t n_keys = min(sh.cnt, kTopkSmem);
    uint32_t n_pow2 = 1;
    while (n_pow2 < n_keys) {
      n_pow2 <<= 1;
    }
    // pads (0, 0) sort last: ord(x) of any score has s != 0 unless x is the most negative NaN pattern
    for (uint32_t i = n_keys + threadIdx.x; i < n_pow2; i += blockDim.x) {
      sh.keys[i].s = 0;
      sh.keys[i].d = 0;
    }
    bitonic_sort_desc(sh.keys, n_pow2);
    for (uint32_t i = threadIdx.x; i < n_keys; i += blockDim.x) {
      const uint64_t rank = c0 + i;
      if (rank >= want_begin && rank < end) {
        emit(rank - want_begin, sh.keys[i]);
      }
    }
    hi = lo;
    has_hi = true;
  }
}

constexpr int kMaxCachedLists = 24;
constexpr uint32_t kPreSearchLists = kTileThreads / 64;  // lists whose sub-range bounds are searched up front (2 warps each)

// One column condition for one document: ApplyFiltersWithBitmap / ApplyFilters, search_pipeline.cpp:1098-1237.
__device__ __forceinline__ bool filter_pass(const FilterPred& f, uint32_t doc) {
  const uint32_t op = f.op_flags & 7u;
  if (f.cls == kFcNone) {
    return op == 1;  // no such column: no bitmap / no stored value -> only != holds
  }
  const bool is_null = f.nulls != nullptr && f.nulls[doc] != 0;
  const uint64_t v = f.values[doc];
  const bool valid = (f.op_flags & kFilterValid) != 0;
  if ((f.op_flags & kFilterBitmapMode) != 0) {
    // FilterIndex semantics: the value's serialisation equals that of a type interpretation of the literal
    // (BuildTypeUnionBitmap :1021-1094); NULLs are not indexed, so != keeps them (:1223-1229)
    const bool eq = !is_null && valid && (f.cls == kFcBool ? ((v != 0) == (f.c != 0)) : v == f.c);
    return op == 0 ? eq : !eq;
  }
  if (is_null) {
    return op == 1;  // :1126-1132
  }
  bool lt, eq, gt;
  if (f.cls == kFcString) {  // c = rank of the literal among the column's distinct strings (lower bound)
    const bool found = (f.op_flags & kFilterFound) != 0;
    lt = v < f.c;
    eq = found && v == f.c;
    gt = found ? v > f.c : v >= f.c;
  } else if (f.cls == kFcBool) {
    const bool a = v != 0, c = f.c != 0;
    lt = !a && c;
    eq = a == c;
    gt = a && !c;
  } else if (!valid) {
    return false;  // "Invalid number" :1151-1171
  } else if (f.cls == kFcDouble) {
    const double a = __longlong_as_double(static_cast<long long>(v));
    const double c = __longlong_as_double(static_cast<long long>(f.c));
    if (op <= 1) {  // CompareDoubleValues with kFilterValueEpsilon = 1e-9 (comparison_utils.h:56-61)
      const bool close = fabs(__dsub_rn(a, c)) < 1e-9;
      return op == 0 ? close : !close;
    }
    lt = a < c;
    eq = a == c;
    gt = a > c;
  } else if (f.cls == kFcUnsigned) {
    lt = v < f.c;
    eq = v == f.c;
    gt = v > f.c;
  } else {
    const long long a = static_cast<long long>(v), c = static_cast<long long>(f.c);
    lt = a < c;
    eq = a == c;
    gt = a > c;
  }
  switch (op) {
    case 0: return eq;
    case 1: return !eq;
    case 2: return gt;
    case 3: return gt || eq;
    case 4: return lt;
    default: return lt || eq;
  }
}

// TERM node of a boolean program: the documents of SearchAnd(n-grams of the term) (query_ast.cpp:76-93); a term
// without n-grams falls back to a substring test of the stored text (query/substring_search.h:24-42).
__device__ bool program_term_holds(const IndexView& iv, const BatchView& bv, uint32_t tid, uint32_t doc,
                                   bool optimistic) {
  const uint32_t k0 = bv.term_koff[tid];
  const uint32_t k1 = bv.term_koff[tid + 1];
  if (k1 == k0) {
    const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
    if (tl == 0) {
      return false;
    }
    if (optimistic) {
      return true;  // text leaf of a monotone program, first evaluation (kQOptimistic)
    }
    const uint8_t* term = bv.term_bytes + bv.term_boff[tid];
    const uint64_t b = iv.text_off[doc];
    const uint64_t e = iv.text_off[doc + 1];
    if (e - b < tl) {
      return false;
    }
    const uint8_t t0 = __ldg(term);
    for (uint64_t j = b; j + tl <= e; ++j) {
      if (__ldg(iv.text + j) != t0) {
        continue;
      }
      uint32_t i = 1;
      while (i < tl && __ldg(iv.text + j + i) == __ldg(term + i)) {
        ++i;
      }
      if (i == tl) {
        return true;
      }
    }
    return false;
  }
  if (bv.t_est[tid] == 0) {
    return false;  // one of its n-grams is not in the index
  }
  for (uint32_t kk = k0; kk < k1; ++kk) {
    if (!list_contains(make_list(iv, bv.key_list[kk], bv.key_len[kk]), doc)) {
      return false;
    }
  }
  return true;
}

// Bytes of the whitespace unit that starts at text[i] (0 = none): the delimiters of ContainsFuzzyMatch after its
// NormalizeUnicodeWhitespace pass (utils/edit_distance.cpp: " \t\r\n", U+3000 as E3 80 80 and U+00A0 as C2 A0, both
// matched on raw bytes with the same bounds checks).
__device__ __forceinline__ uint32_t fuzzy_ws_len(const uint8_t* text, uint64_t i, uint64_t e) {
  const uint32_t b = __ldg(text + i);
  if (b == ' ' || b == '\t' || b == '\r' || b == '\n') {
    return 1;
  }
  if (b == 0xE3u && i + 2 < e && __ldg(text + i + 1) == 0x80u && __ldg(text + i + 2) == 0x80u) {
    return 3;
  }
  if (b == 0xC2u && i + 1 < e && __ldg(text + i + 1) == 0xA0u) {
    return 2;
  }
  return 0;
}

// One row of the edit-distance recurrence: dp[j] = distance between the characters consumed so far and term[0..j).
__device__ __forceinline__ void fuzzy_row(uint16_t* dp, const uint32_t* tcp, uint32_t tl, uint32_t c, uint16_t first) {
  uint16_t prev = dp[0];
  dp[0] = first;
  for (uint32_t j = 0; j < tl; ++j) {
    const uint16_t cost = tcp[j] == c ? 0 : 1;
    const uint16_t v = min(min(static_cast<uint16_t>(dp[j + 1] + 1), static_cast<uint16_t>(dp[j] + 1)),
                           static_cast<uint16_t>(prev + cost));
    prev = dp[j + 1];
    dp[j + 1] = v;
  }
}

// ContainsFuzzyMatch(text, term, d), utils/edit_distance.cpp: the text is cut into whitespace-delimited words; an
// all-ASCII word matches when its Levenshtein distance to the term (in code points) is <= d; a word with non-ASCII
// characters matches when one of its code-point windows of length |term| - d .. |term| + d does. The window test is
// evaluated as approximate substring matching (first column held at 0): a window within distance d of the term has a
// length within d of the term's by itself, and for |term| <= d every one-character window already matches, so the
// two formulations accept the same words. Code points follow Utf8ToCodepoints (invalid bytes skipped one at a time).
__device__ __noinline__ bool fuzzy_text_holds(const IndexView& iv, const BatchView& bv, uint32_t tid, uint32_t d,
                                              uint32_t doc) {
  const uint32_t tb0 = bv.term_boff[tid];
  const uint32_t tbytes = bv.term_boff[tid + 1] - tb0;
  if (tbytes == 0) {
    return true;
  }
  const uint64_t b = iv.text_off[doc];
  const uint64_t e = iv.text_off[doc + 1];
  if (e == b) {
    return false;
  }
  uint32_t tcp[kFuzzyMaxTermCps];
  uint32_t tl = 0;
  {
    const uint8_t* term = bv.term_bytes + tb0;
    uint32_t i = 0;
    while (i < tbytes && tl < kFuzzyMaxTermCps) {
      const uint32_t avail = tbytes - i;
      uint32_t cp = 0;
      const int n = parse_utf8(__ldg(term + i), avail > 1 ? __ldg(term + i + 1) : 0, avail > 2 ? __ldg(term + i + 2) : 0,
                               avail > 3 ? __ldg(term + i + 3) : 0, avail, &cp);
      if (n > 0) {
        tcp[tl++] = cp;
        i += static_cast<uint32_t>(n);
      } else {
        ++i;
      }
    }
  }
  uint16_t dp[kFuzzyMaxTermCps + 1];
  const uint8_t* text = iv.text;
  uint64_t i = b;
  while (i < e) {
    uint32_t w = fuzzy_ws_len(text, i, e);
    if (w != 0) {
      i += w;
      continue;
    }
    const uint64_t ws = i;
    bool ascii = true;
    while (i < e && fuzzy_ws_len(text, i, e) == 0) {
      ascii = ascii && __ldg(text + i) < 0x80u;
      ++i;
    }
    const uint64_t we = i;
    for (uint32_t j = 0; j <= tl; ++j) {
      dp[j] = static_cast<uint16_t>(j);
    }
    if (ascii) {
      const uint64_t wl = we - ws;
      const uint64_t diff = wl > tl ? wl - tl : tl - wl;
      if (diff > d) {
        continue;
      }
      for (uint64_t p = ws; p < we; ++p) {
        fuzzy_row(dp, tcp, tl, __ldg(text + p), static_cast<uint16_t>(p - ws + 1));
      }
      if (dp[tl] <= d) {
        return true;
      }
    } else {
      uint64_t p = ws;
      while (p < we) {
        const uint64_t avail = we - p;
        uint32_t cp = 0;
        const int n = parse_utf8(__ldg(text + p), avail > 1 ? __ldg(text + p + 1) : 0, avail > 2 ? __ldg(text + p + 2) : 0,
                                 avail > 3 ? __ldg(text + p + 3) : 0, avail, &cp);
        if (n == 0) {
          ++p;
          continue;
        }
        p += static_cast<uint64_t>(n);
        if (tl == 0) {
          if (d >= 1) {
            return true;  // a one-character window against an empty term: distance 1
          }
          break;
        }
        fuzzy_row(dp, tcp, tl, cp, 0);
        if (dp[tl] <= d) {
          return true;
        }
      }
    }
  }
  return false;
}

// Postfix evaluation with a 64-bit stack (bit 0 = top). AND / OR of zero children and NOT without a child are
// false, as QueryNode::Evaluate returns an empty set for them (query_ast.cpp:96-140).
__device__ bool eval_program(const IndexView& iv, const BatchView& bv, uint32_t p0, uint32_t p1, uint32_t doc,
                             bool optimistic) {
  unsigned long long stack = 0;
  uint32_t sp = 0;
  for (uint32_t i = p0; i < p1; ++i) {
    const uint32_t op = bv.prog_op[i];
    const uint32_t arg = bv.prog_arg[i];
    if (op == kOpTerm) {
      stack = (stack << 1) | (program_term_holds(iv, bv, arg, doc, optimistic) ? 1ULL : 0ULL);
      ++sp;
    } else if (op == kOpFuzzyText) {
      stack = (stack << 1) |
              ((optimistic || fuzzy_text_holds(iv, bv, arg & 0xFFFFFFu, arg >> 24, doc)) ? 1ULL : 0ULL);
      ++sp;
    } else if (op == kOpAnd || op == kOpOr || op == kOpAtLeast) {
      const uint32_t n = op == kOpAtLeast ? (arg & 0xFFFFu) : arg;
      if (n > sp) {
        return false;
      }
      const unsigned long long mask = n >= 64 ? ~0ULL : ((1ULL << n) - 1ULL);
      const unsigned long long bits = stack & mask;
      bool v;
      if (op == kOpAtLeast) {
        const uint32_t need = arg >> 16;  // 0 or more than n children: nothing matches (index.cpp:489-501)
        v = need != 0 && static_cast<uint32_t>(__popcll(bits)) >= need;
      } else {
        v = n != 0 && (op == kOpAnd ? bits == mask : bits != 0);
      }
      stack = n >= 64 ? 0ULL : (stack >> n);
      sp -= n;
      stack = (stack << 1) | (v ? 1ULL : 0ULL);
      ++sp;
    } else {
      if (sp == 0) {
        return false;
      }
      stack ^= 1ULL;
    }
  }
  return sp > 0 && (stack & 1ULL) != 0;
}

// BM25 contribution of one term (bm25_scorer.cpp:74-85), evaluated operation by operation (no FMA contraction).
__device__ __forceinline__ double bm25_term(double idf, uint32_t tf_u, double length_norm, double k1) {
  const double tf = static_cast<double>(tf_u);
  const double numerator = __dmul_rn(tf, __dadd_rn(k1, 1.0));
  const double denominator = __dadd_rn(tf, __dmul_rn(k1, length_norm));
  return __ddiv_rn(__dmul_rn(idf, numerator), denominator);
}

// One CTA, one 1024-entry tile of query q's driver (tile_in_q-th tile of the query). tile_slot = the tile's index in
// tile_count / tile_total; its records go to rec_doc / rec_score [tile_slot * rec_slot, +written).
__device__ __forceinline__ void and_tile_body(const IndexView& iv, const BatchView& bv, const ScoreParams& sp,
                                              const uint32_t q, const uint64_t tile_in_q, const uint64_t tile_slot,
                                              const uint32_t rec_slot, uint32_t* __restrict__ tile_count,
                                              uint32_t* __restrict__ tile_total, uint32_t* __restrict__ rec_doc,
                                              double* __restrict__ rec_score, const uint32_t prune_k) {
  __shared__ uint32_t s_doc[kTile];     // local doc index of survivors (kNone = id unknown to this shard)
  __shared__ uint32_t s_gid[kTile];     // global doc id of survivors (explicit drivers only)
  __shared__ double s_score[kTile];
  __shared__ uint8_t s_keep[kTile];     // 0 drop, 1 keep, 2 = needs the warp-cooperative slow path
  // sub-list staging of the membership phase, text staging of the (rare) warp-per-document scan, key array of the
  // pruning step: three phases separated by block barriers, one buffer
  __shared__ __align__(16) uint32_t s_stage[kStageCap];
  static_assert((kTileThreads / 32) * kStageBuf <= sizeof(uint32_t) * kStageCap, "text staging must fit s_stage");
  uint8_t(*s_text)[kStageBuf] = reinterpret_cast<uint8_t(*)[kStageBuf]>(s_stage);
  __shared__ uint32_t s_range[2];
  __shared__ uint32_t s_warp[kTileThreads / 32];
  __shared__ uint32_t s_dmin;
  __shared__ uint32_t s_dmax;
  __shared__ uint32_t s_any_slow;
  __shared__ unsigned long long s_bytes;
  __shared__ ListRef s_lists[kMaxCachedLists];
  __shared__ uint32_t s_bounds[kPreSearchLists][2];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    s_bytes = 0;
    s_any_slow = 0;
  }
  const uint32_t flags = bv.q_flags[q];
  const uint32_t l0 = bv.q_loff[q];
  const uint32_t nl = bv.q_nlists[q];
  const uint32_t driver_len = bv.q_driver_len[q];
  const bool drv_all = (flags & kQDriverAll) != 0;
  const bool drv_explicit = (flags & kQDriverExplicit) != 0;
  const bool any_mode = (flags & kQAnyMode) != 0;
  const bool drv_list = !drv_all && !drv_explicit;
  if (threadIdx.x < nl && threadIdx.x < kMaxCachedLists) {
    s_lists[threadIdx.x] = make_list(iv, bv.q_list[l0 + threadIdx.x], bv.q_list_len[l0 + threadIdx.x]);
  }
  __syncthreads();
  const uint32_t n0 = bv.q_noff[q];
  const uint32_t n1 = bv.q_noff[q + 1];

  // ---- driver entries: each thread owns kTileItems CONSECUTIVE entries (keeps the order)
  const uint64_t e0 = tile_in_q * kTile;
  const uint32_t tile_n = static_cast<uint32_t>(umin64(kTile, driver_len - e0));
  uint32_t my_doc[kTileItems];
  uint32_t my_gid[kTileItems];
  uint32_t alive = 0;
#pragma unroll
  for (int k = 0; k < kTileItems; ++k) {
    const uint32_t i = threadIdx.x * kTileItems + k;
    my_doc[k] = kNone;
    my_gid[k] = 0;
    if (i < tile_n) {
      const uint64_t e = e0 + i;
      if (drv_list) {
        my_doc[k] = __ldg(s_lists[0].p + e);
      } else if (drv_all) {
        my_doc[k] = static_cast<uint32_t>(e);
      } else {
        my_gid[k] = __ldg(bv.explicit_ids + e);
        my_doc[k] = local_of(iv, my_gid[k]);
      }
      alive |= 1u << k;
    }
  }
  if (threadIdx.x == 0) {
    if (drv_list) {
      s_dmin = __ldg(s_lists[0].p + e0);
      s_dmax = __ldg(s_lists[0].p + e0 + tile_n - 1);
    } else {
      s_dmin = static_cast<uint32_t>(e0);
      s_dmax = static_cast<uint32_t>(e0 + tile_n - 1);
    }
  }
  __syncthreads();

  // ---- membership
  if ((flags & kQProgram) != 0) {
    const uint32_t p0 = bv.q_poff[q];
    const uint32_t p1 = bv.q_poff[q + 1];
    if ((flags & kQOptimistic) != 0) {  // lists first: most documents are ruled out without reading their text
#pragma unroll
      for (int k = 0; k < kTileItems; ++k) {
        if (((alive >> k) & 1u) && (my_doc[k] == kNone || !eval_program(iv, bv, p0, p1, my_doc[k], true))) {
          alive &= ~(1u << k);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
      if (((alive >> k) & 1u) && (my_doc[k] == kNone || !eval_program(iv, bv, p0, p1, my_doc[k], false))) {
        alive &= ~(1u << k);
      }
    }
  } else if (any_mode) {
    // Index::SearchOr / SearchByThreshold: keep the doc if at least `need` of the lists hold it
    const uint32_t need = bv.q_threshold[q];
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
      if ((alive >> k) & 1u) {
        uint32_t hit = 0;
        for (uint32_t j = 0; j < nl && hit < need && my_doc[k] != kNone; ++j) {
          const ListRef l = j < kMaxCachedLists ? s_lists[j] : make_list(iv, bv.q_list[l0 + j], bv.q_list_len[l0 + j]);
          hit += list_contains(l, my_doc[k]) ? 1u : 0u;
        }
        if (hit < need) {
          alive &= ~(1u << k);
        }
      }
    }
  } else {
    const bool narrow = !drv_explicit;  // caller-supplied candidates may be unsorted (FilterByNgrams keeps their order)
    const uint32_t dmin = s_dmin;
    const uint32_t dmax = s_dmax;
    const uint32_t j_first = drv_list ? 1u : 0u;
    if (narrow) {
      // sub-range bounds of the first kPreSearchLists sparse lists, all searched at once (a warp per bound): the
      // searches are chains of dependent loads, and one list after the other they were most of a tile's latency
      const uint32_t j = j_first + (warp >> 1);
      if ((warp >> 1) < kPreSearchLists && j < nl && j < kMaxCachedLists && s_lists[j].bm == nullptr) {
        const uint32_t r = warp_lower_bound(s_lists[j].p, s_lists[j].len, (warp & 1u) != 0 ? dmax + 1u : dmin);
        if (lane == 0) {
          s_bounds[warp >> 1][warp & 1u] = r;
        }
      }
      __syncthreads();
    }
    for (uint32_t j = j_first; j < nl; ++j) {
      const ListRef l = j < kMaxCachedLists ? s_lists[j] : make_list(iv, bv.q_list[l0 + j], bv.q_list_len[l0 + j]);
      const bool searched = narrow && j - j_first < kPreSearchLists && j < kMaxCachedLists;
      tile_filter_list(l, dmin, dmax, narrow, s_stage, s_range, my_doc, &alive, searched ? s_bounds[j - j_first] : nullptr);
      if (__syncthreads_or(alive != 0) == 0) {
        break;
      }
    }
  }
  // NOT terms (ApplyNotFilter :871-932): drop the doc if it is in ALL n-gram lists of a NOT term.
  if (n1 > n0) {
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
      if (((alive >> k) & 1u) == 0 || my_doc[k] == kNone) {
        continue;
      }
      bool keep = true;
      for (uint32_t i = n0; i < n1 && keep; ++i) {
        const uint32_t tid = bv.q_ntids[i];
        const uint32_t k0 = bv.term_koff[tid];
        const uint32_t k1 = bv.term_koff[tid + 1];
        if (k1 == k0 || bv.t_est[tid] == 0) {
          continue;  // no n-grams (the text stage decides) or an unknown n-gram (matches nothing)
        }
        bool in_all = true;
        for (uint32_t kk = k0; kk < k1 && in_all; ++kk) {
          in_all = list_contains(make_list(iv, bv.key_list[kk], bv.key_len[kk]), my_doc[k]);
        }
        keep = !in_all;
      }
      if (!keep) {
        alive &= ~(1u << k);
      }
    }
  }
  // column filters (Execute :849-852, after the NOT filter)
  {
    const uint32_t f0 = bv.q_foff[q];
    const uint32_t f1 = bv.q_foff[q + 1];
    for (uint32_t fi = f0; fi < f1 && alive != 0; ++fi) {
      const FilterPred f = bv.filters[fi];
#pragma unroll
      for (int k = 0; k < kTileItems; ++k) {
        if (((alive >> k) & 1u) && (my_doc[k] == kNone || !filter_pass(f, my_doc[k]))) {
          alive &= ~(1u << k);
        }
      }
    }
  }
  uint32_t total = 0;
  uint32_t off = block_offsets(__popc(alive), s_warp, &total);
#pragma unroll
  for (int k = 0; k < kTileItems; ++k) {
    if ((alive >> k) & 1u) {
      s_doc[off] = my_doc[k];
      s_gid[off] = my_gid[k];
      ++off;
    }
  }
  __syncthreads();

  // ---- fused epilogue: text scan -> tf -> BM25 (bm25_scorer.cpp:67-88), text constraints
  const uint32_t t0 = bv.q_toff[q];
  const uint32_t t1 = bv.q_toff[q + 1];
  // The terms of a scored boolean program are only what its results are SCORED with (search_handler.cpp:405-470):
  // membership was the program's, so they neither have to occur in the text nor drop a document without text.
  const bool terms_filter = (flags & kQProgram) == 0;
  bool need_text = sp.compute_score != 0 || (flags & kQVerify) != 0;
  for (uint32_t i = t0; i < t1 && !need_text; ++i) {
    const uint32_t tid = bv.q_tids[i];
    need_text = bv.term_koff[tid + 1] == bv.term_koff[tid];
  }
  for (uint32_t i = n0; i < n1 && !need_text; ++i) {
    const uint32_t tid = bv.q_ntids[i];
    need_text = bv.term_koff[tid + 1] == bv.term_koff[tid] && bv.term_boff[tid + 1] > bv.term_boff[tid];
  }
  if (need_text && total > 0) {
    // every term fits the register scanner? (one term per thread, then a block vote: the loads of all terms are in
    // flight together instead of one after the other in every thread)
    bool short_here = true;
    for (uint32_t i = t0 + threadIdx.x; i < t1; i += kTileThreads) {
      const uint32_t tid = bv.q_tids[i];
      short_here = short_here && (bv.term_boff[tid + 1] - bv.term_boff[tid]) <= kThreadScanMaxTerm;
    }
    for (uint32_t i = n0 + threadIdx.x; i < n1; i += kTileThreads) {
      const uint32_t tid = bv.q_ntids[i];
      short_here = short_here && (bv.term_boff[tid + 1] - bv.term_boff[tid]) <= kThreadScanMaxTerm;
    }
    const bool all_short = __syncthreads_and(short_here ? 1 : 0) != 0;
    unsigned long long text_bytes = 0;
    const uint32_t n_search = t1 - t0;
    // A query of ONE term that is exactly its single n-gram, driven by that n-gram's list with every entry of the
    // tile surviving (no NOT terms, no filters, no text verification): tf comes from the occurrences recorded with
    // the posting -- one occurrence, or two that do not overlap (CountTermOccurrences counts non-overlapping matches,
    // bm25_scorer.cpp:27-45) -- and the score needs only the document's length. No text, no text offsets: what a
    // C5-style batch of short single-term queries over long lists spends its time on. Documents with three or more
    // occurrences (or offsets beyond the recorded range) take the scanning path below.
    bool pay_only = false;
    if (sp.compute_score != 0 && terms_filter && drv_list && nl == 1 && n_search == 1 && n1 == n0 &&
        (flags & kQVerify) == 0 && total == tile_n && iv.post_pos != nullptr && iv.all_valid_utf8 != 0) {
      const uint32_t tid = bv.q_tids[t0];
      pay_only = (bv.term_flags[tid] & 8u) != 0 && bv.term_koff[tid + 1] - bv.term_koff[tid] == 1;
    }
    if (pay_only) {
      const uint32_t tid = bv.q_tids[t0];
      const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
      const double idf = bv.q_idf[t0];
      const uint64_t pbase = static_cast<uint64_t>(s_lists[0].p - iv.postings) + e0;
      for (uint32_t s = threadIdx.x; s < total; s += kTileThreads) {
        const uint32_t doc = s_doc[s];  // survivor s is entry s of the tile: every entry survived
        const uint32_t p1 = __ldg(iv.post_pos + pbase + s);
        const uint32_t dl_u = __ldg(iv.doc_len + doc);
        uint32_t tf_u = 1;
        bool known = true;
        if ((p1 & kPosMulti) != 0) {
          const uint32_t p2 = __ldg(iv.post_pos2 + pbase + s);
          known = (p1 & kPosUnknown) != kPosUnknown && (p2 & kPosUnknown) != kPosUnknown && (p2 & kPosMulti) == 0;
          tf_u = (p2 & kPosUnknown) - (p1 & kPosUnknown) >= tl ? 2u : 1u;
        }
        if (!known) {
          // three or more occurrences, or offsets beyond the recorded range: count them in the text (by this thread
          // when the document fits the register scanner, else by a warp through shared memory below)
          const uint64_t b = iv.text_off[doc];
          const uint32_t len = static_cast<uint32_t>(iv.text_off[doc + 1] - b);
          text_bytes += len + 4;
          if (!all_short || len > kThreadScanMaxDoc) {
            s_keep[s] = 2;
            s_score[s] = 0.0;
            s_any_slow = 1;
          } else {
            const uint8_t* term = bv.term_bytes + bv.term_boff[tid];
            tf_u = thread_count_term(iv.text, b, len, term, tl, load_term_regs(term, tl), false);
            const double dl = static_cast<double>(dl_u);
            const double length_norm = __dadd_rn(__dsub_rn(1.0, sp.b), __ddiv_rn(__dmul_rn(sp.b, dl), sp.avgdl_clamped));
            s_keep[s] = 1;
            s_score[s] = tf_u != 0 ? __dadd_rn(0.0, bm25_term(idf, tf_u, length_norm, sp.k1)) : 0.0;
          }
        } else {
          const double dl = static_cast<double>(dl_u);
          const double length_norm = __dadd_rn(__dsub_rn(1.0, sp.b), __ddiv_rn(__dmul_rn(sp.b, dl), sp.avgdl_clamped));
          s_keep[s] = 1;
          s_score[s] = __dadd_rn(0.0, bm25_term(idf, tf_u, length_norm, sp.k1));
          text_bytes += 4;  // B_score counts text bytes + 4 per scored document; the text was not read here
        }
      }
    } else if (all_short && n1 == n0 && n_search >= 2 && total * n_search <= kPairScanMaxItems) {
      // Very few survivors (the common tile ends with one): kGroupScanLanes lanes per (document, TERM) pair, so the
      // terms of a document are counted side by side instead of one after the other; the BM25 contributions are
      // added afterwards in term order, operation for operation as the per-document loops below do.
      constexpr int G = kGroupScanLanes;
      double* const s_contrib = reinterpret_cast<double*>(s_stage);
      uint32_t* const s_tf = s_stage + 2 * kPairScanMaxItems;
      static_assert(3 * kPairScanMaxItems <= kStageCap, "pair results live in the staging buffer");
      const bool leader = (threadIdx.x & (G - 1)) == 0;
      for (uint32_t item = threadIdx.x / G; item < total * n_search; item += kTileThreads / G) {
        const uint32_t s = item / n_search;
        const uint32_t i = t0 + (item - s * n_search);
        const uint32_t doc = s_doc[s];
        uint32_t tf_u = 0;
        double contrib = 0.0;
        if (doc != kNone) {
          const uint64_t b = iv.text_off[doc];
          const uint32_t len = static_cast<uint32_t>(iv.text_off[doc + 1] - b);
          const uint32_t dl_u = __ldg(iv.doc_len + doc);
          const uint32_t tid = bv.q_tids[i];
          const double idf = sp.compute_score != 0 ? bv.q_idf[i] : 0.0;
          if (len != 0 && len <= kThreadScanMaxDoc) {
            const uint8_t* term = bv.term_bytes + bv.term_boff[tid];
            const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
            tf_u = group_count_term<G>(iv.text, b, len, term, tl, load_term_regs(term, tl), sp.compute_score == 0);
            if (sp.compute_score != 0 && tf_u != 0) {
              const double dl = static_cast<double>(dl_u);
              const double length_norm =
                  __dadd_rn(__dsub_rn(1.0, sp.b), __ddiv_rn(__dmul_rn(sp.b, dl), sp.avgdl_clamped));
              contrib = bm25_term(idf, tf_u, length_norm, sp.k1);
            }
          }
        }
        if (leader) {
          s_tf[item] = tf_u;
          s_contrib[item] = contrib;
        }
      }
      __syncthreads();
      if (threadIdx.x < total) {
        const uint32_t s = threadIdx.x;
        const uint32_t doc = s_doc[s];
        uint8_t keep = 1;
        double score = 0.0;
        if (doc != kNone) {
          const uint64_t b = iv.text_off[doc];
          const uint32_t len = static_cast<uint32_t>(iv.text_off[doc + 1] - b);
          text_bytes += len + 4;
          if (len > kThreadScanMaxDoc) {
            keep = 2;
          } else {
            for (uint32_t i = t0; i < t1; ++i) {
              const uint32_t tid = bv.q_tids[i];
              const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
              const bool no_keys = bv.term_koff[tid + 1] == bv.term_koff[tid];
              if (len == 0) {
                // no stored text: a substring-only search term cannot match; verify keeps the doc
                if (no_keys && tl != 0 && terms_filter) {
                  keep = 0;
                }
                continue;
              }
              const bool must = terms_filter && ((flags & kQVerify) != 0 || no_keys);
              const uint32_t tf_u = s_tf[s * n_search + (i - t0)];
              if (must && tf_u == 0 && tl != 0) {
                keep = 0;
              }
              if (sp.compute_score != 0 && tf_u != 0) {
                score = __dadd_rn(score, s_contrib[s * n_search + (i - t0)]);
              }
            }
          }
        }
        s_keep[s] = keep;
        s_score[s] = score;
        if (keep == 2) {
          s_any_slow = 1;
        }
      }
    } else if (total <= kGroupScanMaxDocs) {
      // few survivors: kGroupScanLanes lanes per document (same decisions as the thread-per-document loop below)
      constexpr int G = kGroupScanLanes;
      const bool leader = (threadIdx.x & (G - 1)) == 0;
      for (uint32_t s = threadIdx.x / G; s < total; s += kTileThreads / G) {
        const uint32_t doc = s_doc[s];
        uint8_t keep = 1;
        double score = 0.0;
        if (doc != kNone) {
          const uint64_t b = iv.text_off[doc];
          const uint32_t len = static_cast<uint32_t>(iv.text_off[doc + 1] - b);
          const uint32_t dl_u = __ldg(iv.doc_len + doc);  // issued with the offsets, not after the branches below
          if (leader) {
            text_bytes += len + 4;
          }
          if (len == 0) {
            for (uint32_t i = t0; i < t1; ++i) {
              const uint32_t tid = bv.q_tids[i];
              if (terms_filter && bv.term_koff[tid + 1] == bv.term_koff[tid] && bv.term_boff[tid + 1] > bv.term_boff[tid]) {
                keep = 0;
              }
            }
          } else if (!all_short || len > kThreadScanMaxDoc) {
            keep = 2;
          } else {
            const double dl = static_cast<double>(dl_u);
            const double length_norm =
                __dadd_rn(__dsub_rn(1.0, sp.b), __ddiv_rn(__dmul_rn(sp.b, dl), sp.avgdl_clamped));
            for (uint32_t i = t0; i < t1; ++i) {
              const uint32_t tid = bv.q_tids[i];
              const uint8_t* term = bv.term_bytes + bv.term_boff[tid];
              const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
              const bool must = terms_filter && ((flags & kQVerify) != 0 || bv.term_koff[tid + 1] == bv.term_koff[tid]);
              const uint32_t tf_u =
                  group_count_term<G>(iv.text, b, len, term, tl, load_term_regs(term, tl), sp.compute_score == 0);
              if (must && tf_u == 0 && tl != 0) {
                keep = 0;
              }
              if (sp.compute_score != 0 && tf_u != 0) {
                score = __dadd_rn(score, bm25_term(bv.q_idf[i], tf_u, length_norm, sp.k1));
              }
            }
            for (uint32_t i = n0; i < n1 && keep != 0; ++i) {
              const uint32_t tid = bv.q_ntids[i];
              const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
              if (bv.term_koff[tid + 1] == bv.term_koff[tid] && tl != 0) {
                const uint8_t* term = bv.term_bytes + bv.term_boff[tid];
                if (group_count_term<G>(iv.text, b, len, term, tl, load_term_regs(term, tl), true) != 0) {
                  keep = 0;
                }
              }
            }
          }
        }
        if (leader) {
          s_keep[s] = keep;
          s_score[s] = score;
          if (keep == 2) {
            s_any_slow = 1;
          }
        }
      }
    } else
    // fast path: one thread per surviving document
    for (uint32_t s = threadIdx.x; s < total; s += kTileThreads) {
      const uint32_t doc = s_doc[s];
      uint8_t keep = 1;
      double score = 0.0;
      if (doc != kNone) {
        const uint64_t b = iv.text_off[doc];
        const uint32_t len = static_cast<uint32_t>(iv.text_off[doc + 1] - b);
        text_bytes += len + 4;
        if (len == 0) {
          // no stored text: a substring-only search term cannot match; verify keeps the doc
          for (uint32_t i = t0; i < t1; ++i) {
            const uint32_t tid = bv.q_tids[i];
            if (terms_filter && bv.term_koff[tid + 1] == bv.term_koff[tid] && bv.term_boff[tid + 1] > bv.term_boff[tid]) {
              keep = 0;
            }
          }
        } else if (!all_short || len > kThreadScanMaxDoc) {
          keep = 2;
        } else {
          const double dl = static_cast<double>(__ldg(iv.doc_len + doc));
          const double length_norm =
              __dadd_rn(__dsub_rn(1.0, sp.b), __ddiv_rn(__dmul_rn(sp.b, dl), sp.avgdl_clamped));
          for (uint32_t i = t0; i < t1; ++i) {
            const uint32_t tid = bv.q_tids[i];
            const uint8_t* term = bv.term_bytes + bv.term_boff[tid];
            const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
            const bool must = terms_filter && ((flags & kQVerify) != 0 || bv.term_koff[tid + 1] == bv.term_koff[tid]);
            const uint32_t tf_u =
                thread_count_term(iv.text, b, len, term, tl, load_term_regs(term, tl), sp.compute_score == 0);
            if (must && tf_u == 0 && tl != 0) {
              keep = 0;
            }
            if (sp.compute_score != 0 && tf_u != 0) {
              score = __dadd_rn(score, bm25_term(bv.q_idf[i], tf_u, length_norm, sp.k1));
            }
          }
          for (uint32_t i = n0; i < n1 && keep != 0; ++i) {
            const uint32_t tid = bv.q_ntids[i];
            const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
            if (bv.term_koff[tid + 1] == bv.term_koff[tid] && tl != 0) {
              const uint8_t* term = bv.term_bytes + bv.term_boff[tid];
              if (thread_count_term(iv.text, b, len, term, tl, load_term_regs(term, tl), true) != 0) {
                keep = 0;
              }
            }
          }
        }
      }
      s_keep[s] = keep;
      s_score[s] = score;
      if (keep == 2) {
        s_any_slow = 1;
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      text_bytes += __shfl_xor_sync(0xffffffffu, text_bytes, s);
    }
    if (lane == 0 && text_bytes != 0) {
      atomicAdd(&s_bytes, text_bytes);
    }
    __syncthreads();
    if (s_any_slow != 0) {
      // slow path (terms > 16 bytes or documents > 4 KB): one warp per document through shared memory
      for (uint32_t s = warp; s < total; s += kTileThreads / 32) {
        if (s_keep[s] != 2) {
          continue;
        }
        const uint32_t doc = s_doc[s];
        bool keep = true;
        double score = 0.0;
        DocText d = doc_open(iv, doc, s_text[warp]);
        const double dl = static_cast<double>(__ldg(iv.doc_len + doc));
        const double length_norm = __dadd_rn(__dsub_rn(1.0, sp.b), __ddiv_rn(__dmul_rn(sp.b, dl), sp.avgdl_clamped));
        for (uint32_t i = t0; i < t1; ++i) {
          const uint32_t tid = bv.q_tids[i];
          const uint8_t* term = bv.term_bytes + bv.term_boff[tid];
          const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
          const bool must = terms_filter && ((flags & kQVerify) != 0 || bv.term_koff[tid + 1] == bv.term_koff[tid]);
          const uint32_t tf_u = doc_count_term(d, term, tl, sp.compute_score == 0);
          if (must && tf_u == 0 && tl != 0) {
            keep = false;
          }
          if (sp.compute_score != 0 && tf_u != 0) {
            score = __dadd_rn(score, bm25_term(bv.q_idf[i], tf_u, length_norm, sp.k1));
          }
        }
        for (uint32_t i = n0; i < n1 && keep; ++i) {
          const uint32_t tid = bv.q_ntids[i];
          const uint32_t tl = bv.term_boff[tid + 1] - bv.term_boff[tid];
          if (bv.term_koff[tid + 1] == bv.term_koff[tid] && tl != 0) {
            keep = doc_count_term(d, bv.term_bytes + bv.term_boff[tid], tl, true) == 0;
          }
        }
        __syncwarp();
        if (lane == 0) {
          s_keep[s] = keep ? 1 : 0;
          s_score[s] = score;
        }
      }
      __syncthreads();
    }
  }

  // ---- ordered write of the tile's records
  const uint64_t out_base = tile_slot * rec_slot;
  uint32_t keep_mask = 0;
  const uint32_t s0 = threadIdx.x * kTileItems;
#pragma unroll
  for (int k = 0; k < kTileItems; ++k) {
    const uint32_t s = s0 + k;
    if (s < total && (!need_text || s_keep[s] != 0)) {
      keep_mask |= 1u << k;
    }
  }
  uint32_t kept_total = 0;
  uint32_t woff = block_offsets(__popc(keep_mask), s_warp, &kept_total);
  uint32_t written = kept_total;
  if (prune_k != 0 && sp.compute_score != 0) {
    // Block-level top-k inside the epilogue: a tile can contribute at most prune_k (= offset + limit) records to the
    // query's answer. Two filters, both exact. (1) The query's running threshold q_thr = the best "prune_k-th best
    // score of a tile" any of its tiles has published so far: prune_k records at or above it exist, so a record
    // strictly below it cannot reach the answer (which records get written depends on the order the tiles run in;
    // the answer does not). (2) Only if more than prune_k records pass, the tile sorts them (SortByScore order,
    // bitonic over the next power of two), keeps its best prune_k and publishes the score of the last one. Tiles of
    // a long list mostly end at (1): with t tiles done, about prune_k / t of a tile's records pass.
    // The full count still goes to tile_total. The staging buffer of the membership phase is the key array.
    const bool desc = sp.descending != 0;
    const unsigned long long thr = *reinterpret_cast<const volatile unsigned long long*>(bv.q_thr + q);
    uint32_t pass_mask = 0;
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
      if (keep_mask & (1u << k)) {
        const uint64_t o = ord_f64(s_score[s0 + k]);
        if ((desc ? o : ~o) >= thr) {
          pass_mask |= 1u << k;
        }
      }
    }
    uint32_t n_pass = 0;
    uint32_t poff = block_offsets(__popc(pass_mask), s_warp, &n_pass);
    if (n_pass > prune_k) {
      SortKey* keys = reinterpret_cast<SortKey*>(s_stage);
      static_assert(sizeof(SortKey) * kTile <= sizeof(uint32_t) * kStageCap, "key array must fit the staging buffer");
#pragma unroll
      for (int k = 0; k < kTileItems; ++k) {
        if (pass_mask & (1u << k)) {
          const uint32_t s = s0 + k;
          keys[poff++] = make_sort_key(s_score[s], drv_explicit ? s_gid[s] : gid_of(iv, s_doc[s]), desc);
        }
      }
      uint32_t n_pow2 = 256;
      while (n_pow2 < n_pass) {
        n_pow2 <<= 1;
      }
      for (uint32_t i = n_pass + threadIdx.x; i < n_pow2; i += kTileThreads) {
        keys[i].s = 0;
        keys[i].d = 0;
      }
      bitonic_sort_desc(keys, n_pow2);
      written = prune_k;
      for (uint32_t i = threadIdx.x; i < written; i += kTileThreads) {
        rec_doc[out_base + i] = desc ? keys[i].d : ~keys[i].d;
        rec_score[out_base + i] = unord_f64(desc ? keys[i].s : ~keys[i].s);
      }
      if (threadIdx.x == 0) {
        atomicMax(bv.q_thr + q, static_cast<unsigned long long>(keys[prune_k - 1].s));
      }
    } else {
#pragma unroll
      for (int k = 0; k < kTileItems; ++k) {
        if (pass_mask & (1u << k)) {
          const uint32_t s = s0 + k;
          rec_doc[out_base + poff] = drv_explicit ? s_gid[s] : gid_of(iv, s_doc[s]);
          rec_score[out_base + poff] = s_score[s];
          ++poff;
        }
      }
      written = n_pass;
    }
  } else {
    // Un-scored answers are the first limit + offset ids in driver order: a tile can contribute at most that many, so
    // only its first prune_k survivors are written (tile_total keeps the full count for the ranks and the total).
    const uint32_t cap = prune_k != 0 && sp.compute_score == 0 ? prune_k : static_cast<uint32_t>(kTile);
#pragma unroll
    for (int k = 0; k < kTileItems; ++k) {
      if (keep_mask & (1u << k)) {
        const uint32_t s = s0 + k;
        if (woff < cap) {
          rec_doc[out_base + woff] = drv_explicit ? s_gid[s] : gid_of(iv, s_doc[s]);
          if (sp.compute_score != 0) {
            rec_score[out_base + woff] = s_score[s];
          }
        }
        ++woff;
      }
    }
    written = min(kept_total, cap);
  }
  if (threadIdx.x == 0) {
    tile_count[tile_slot] = written;
    tile_total[tile_slot] = kept_total;
    if (kept_total != 0) {
      atomicAdd(bv.stats + kStatResultDocs * kStatStripes + (tile_slot & (kStatStripes - 1)),
                static_cast<unsigned long long>(kept_total));
    }
    if (s_bytes != 0) {
      atomicAdd(bv.stats + kStatScoreBytes * kStatStripes + (tile_slot & (kStatStripes - 1)), s_bytes);
    }
  }
}

// Order in which a launch visits its tiles. Consecutive tile numbers belong to one query; in a batch of long lists
// they would all run at the same moment, before any of them has published a pruning threshold (and_tile_body), and
// every one of them would sort. Large launches of scored, pruned batches therefore walk the tiles with a stride of
// about 0.618 n (coprime to n, so i -> i * stride mod n is a permutation): what runs together comes from far-apart
// queries, and the tiles of one query follow each other in time. Small launches keep the natural order (their tiles
// share staged sub-lists through L2).
constexpr uint32_t kInterleaveMinTiles = 1u << 15;
__host__ __device__ __forceinline__ uint32_t interleave_stride(uint32_t n, uint32_t prune_k, int compute_score) {
  if (n < kInterleaveMinTiles || prune_k == 0 || compute_score == 0) {
    return 1;
  }
  uint32_t p = static_cast<uint32_t>(static_cast<uint64_t>(n) * 618 / 1000) | 1u;
  for (;; p += 2) {
    uint32_t a = n, b = p;
    while (b != 0) {
      const uint32_t t = a % b;
      a = b;
      b = t;
    }
    if (a == 1) {
      return p;
    }
  }
}
__device__ __forceinline__ uint32_t interleaved_tile(uint32_t i, uint32_t n, uint32_t stride) {
  return stride == 1 ? i : static_cast<uint32_t>(static_cast<uint64_t>(i) * stride % n);
}

__global__ void __launch_bounds__(kTileThreads, MGX_AND_OCC)
and_tile_kernel(IndexView iv, BatchView bv, ScoreParams sp, uint64_t tile_base, uint32_t rec_slot,
                uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_total, uint32_t* __restrict__ rec_doc,
                double* __restrict__ rec_score, uint32_t prune_k, uint32_t tile_stride) {
  const uint32_t slot = interleaved_tile(blockIdx.x, gridDim.x, tile_stride);  // stride: interleave_stride, on the host
  const uint64_t tile_global = tile_base + slot;
  const uint32_t q = __ldg(bv.tile_query + tile_global);
  and_tile_body(iv, bv, sp, q, tile_global - bv.q_tile_off[q], slot, rec_slot, tile_count, tile_total, rec_doc,
                rec_score, prune_k);
}

// Streamed form: persistent CTAs pull tile indices from a device-side counter (the tile count never visits the
// host); tile -> query by a 32-ary search over the queries' tile prefix sums.
__global__ void __launch_bounds__(kTileThreads, MGX_AND_OCC)
and_tiles_kernel(IndexView iv, BatchView bv, ScoreParams sp, uint32_t rec_slot, uint32_t* __restrict__ tile_count,
                 uint32_t* __restrict__ tile_total, uint32_t* __restrict__ rec_doc, double* __restrict__ rec_score,
                 uint32_t prune_k) {
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_q;
  __shared__ uint32_t s_stride;
  const uint32_t n_tiles = bv.launch[kLaunchAndTiles];
  if (threadIdx.x == 0) {
    s_stride = interleave_stride(n_tiles, prune_k, sp.compute_score);  // once per persistent CTA
  }
  __syncthreads();
  const uint32_t stride = s_stride;
  for (;;) {
    if (threadIdx.x < 32) {
      uint32_t tile = 0;
      if (threadIdx.x == 0) {
        tile = atomicAdd(bv.launch + kLaunchAndNext, 1u);
      }
      tile = __shfl_sync(0xffffffffu, tile, 0);
      uint32_t q = 0;
      if (tile < n_tiles) {
        tile = interleaved_tile(tile, n_tiles, stride);
        q = warp_upper_bound_u64(bv.q_tile_off, bv.n_queries + 1, tile) - 1u;
      }
      if (threadIdx.x == 0) {
        s_tile = tile;
        s_q = q;
      }
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t q = s_q;
    if (tile >= n_tiles) {
      return;
    }
    and_tile_body(iv, bv, sp, q, tile - bv.q_tile_off[q], tile, rec_slot, tile_count, tile_total, rec_doc, rec_score,
                  prune_k);
    __syncthreads();  // the body's shared state (and s_tile / s_q) is rewritten by the next round
  }
}

// ------------------------------------------------------------------ top-k kernels
// One CTA per query. Records of tile t (index in the chunk) live at t * rec_slot .. + tile_count[t].
__global__ void __launch_bounds__(256)
topk_kernel(BatchView bv, uint32_t q_first, uint64_t tile_base, uint32_t rec_slot,
            const uint32_t* __restrict__ tile_count, const uint32_t* __restrict__ tile_total,
            const uint32_t* __restrict__ rec_doc, const double* __restrict__ rec_score, int compute_score,
            int descending, uint32_t limit, uint32_t offset,
            uint64_t stride, uint32_t* __restrict__ out_ids, double* __restrict__ out_scores,
            uint32_t* __restrict__ out_count, uint64_t* __restrict__ out_total) {
  __shared__ TopkShared sh;
  __shared__ uint64_t s_scan[8];
  __shared__ uint64_t s_scan_rec[8];
  __shared__ uint64_t s_carry;
  const uint32_t q = q_first + blockIdx.x;
  const uint64_t t0 = bv.q_tile_off[q] - tile_base;
  uint32_t ntiles = static_cast<uint32_t>(bv.q_tile_off[q + 1] - bv.q_tile_off[q]);
  if (bv.launch != nullptr && bv.launch[kLaunchOverflow] != 0) {
    ntiles = 0;  // the streamed batch did not fit its workspace: nothing ran, the host repeats it (see batch_search)
  }
  const uint64_t r0 = t0 * rec_slot;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  uint32_t* ids = out_ids + static_cast<uint64_t>(q) * stride;
  double* scs = out_scores != nullptr ? out_scores + static_cast<uint64_t>(q) * stride : nullptr;

  // total = sum of the tiles' full survivor counts; records = what the tiles actually wrote (pruned to top-k)
  uint64_t local = 0;
  uint64_t local_rec = 0;
  for (uint32_t t = threadIdx.x; t < ntiles; t += blockDim.x) {
    local += tile_total[t0 + t];
    local_rec += tile_count[t0 + t];
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    local += __shfl_xor_sync(0xffffffffu, local, s);
    local_rec += __shfl_xor_sync(0xffffffffu, local_rec, s);
  }
  if (lane == 0) {
    s_scan[warp] = local;
    s_scan_rec[warp] = local_rec;
  }
  if (threadIdx.x == 0) {
    s_carry = 0;
  }
  __syncthreads();
  uint64_t total = 0;
  uint64_t n_records = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    total += s_scan[w];
    n_records += s_scan_rec[w];
  }
  __syncthreads();
  const uint64_t want_end = limit == 0 ? total : umin64(total, static_cast<uint64_t>(offset) + limit);
  const uint64_t want_begin = umin64(offset, total);
  const uint32_t n_out = static_cast<uint32_t>(umin64(want_end - want_begin, stride));
  if (threadIdx.x == 0) {
    out_total[q] = total;
    out_count[q] = n_out;
  }
  if (n_out == 0) {
    return;
  }

  if (!compute_score) {
    // ascending ids: tiles are in driver order and each tile is compacted in order
    for (uint32_t tb = 0; tb < ntiles; tb += blockDim.x) {
      const uint32_t t = tb + threadIdx.x;
      const uint32_t c = t < ntiles ? tile_count[t0 + t] : 0;   // records the tile wrote (its first ones, in order)
      const uint32_t r = t < ntiles ? tile_total[t0 + t] : 0;   // survivors of the tile: what the ranks count
      uint64_t inc = r;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const uint64_t o = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= static_cast<unsigned>(s)) {
          inc += o;
        }
      }
      if (lane == 31) {
        s_scan[warp] = inc;
      }
      __syncthreads();
      uint64_t prefix = s_carry;
      uint64_t chunk_total = 0;
      for (unsigned w = 0; w < 8; ++w) {
        if (w < warp) {
          prefix += s_scan[w];
        }
        chunk_total += s_scan[w];
      }
      const uint64_t first_rank = prefix + inc - r;
      for (uint32_t i = 0; i < c; ++i) {
        const uint64_t rank = first_rank + i;
        if (rank >= want_begin && rank - want_begin < n_out) {
          ids[rank - want_begin] = rec_doc[r0 + static_cast<uint64_t>(t) * rec_slot + i];
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        s_carry += chunk_total;
      }
      __syncthreads();
      if (s_carry >= want_begin + n_out) {
        break;
      }
    }
    return;
  }

  const bool desc = descending != 0;
  auto scan = [&](auto f) {
    for (uint32_t t = warp; t < ntiles; t += 8) {
      const uint32_t c = tile_count[t0 + t];
      const uint64_t base = r0 + static_cast<uint64_t>(t) * rec_slot;
      for (uint32_t i = lane; i < c; i += 32) {
        f(make_sort_key(rec_score[base + i], rec_doc[base + i], desc));
      }
    }
  };
  block_topk_window(sh, scan, n_records, want_begin, n_out, [&](uint64_t i, const SortKey k) {
    ids[i] = desc ? k.d : ~k.d;
    if (scs != nullptr) {
      scs[i] = unord_f64(desc ? k.s : ~k.s);
    }
  });
}

// Pre-reduction for queries whose driver spans many tiles. One CTA takes a group of consecutive tiles of one query
// (at most kGroupKeyCap records), keeps the group's kk best records (SortByScore order) and writes them back into
// the first tile's record slot; the other tiles of the group are emptied. The final topk_kernel then reads
// groups x kk records instead of tiles x kk, which removes the single-CTA tail of a batch's largest queries.
// tile_total (the exact survivor counts) is not touched.
constexpr uint32_t kGroupKeyCap = 3072;
struct TopkGroup {
  uint32_t q;
  uint32_t t_begin;  // tile range inside the query
  uint32_t t_end;
};

__device__ __forceinline__ void topk_group_body(const BatchView& bv, const TopkGroup g, uint64_t tile_base,
                                                uint32_t rec_slot, uint32_t* __restrict__ tile_count,
                                                uint32_t* __restrict__ rec_doc, double* __restrict__ rec_score,
                                                int descending, uint32_t kk) {
  __shared__ uint64_t s_ks[kGroupKeyCap];
  __shared__ uint32_t s_kd[kGroupKeyCap];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_n;
  __shared__ uint32_t s_out;
  __shared__ SortKey s_prefix;
  __shared__ uint32_t s_need;
  const uint64_t t0 = bv.q_tile_off[g.q] - tile_base + g.t_begin;
  const uint64_t r0 = t0 * rec_slot;
  const uint32_t ntiles = g.t_end - g.t_begin;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const bool desc = descending != 0;
  if (threadIdx.x == 0) {
    s_n = 0;
    s_out = 0;
    s_prefix.s = 0;
    s_prefix.d = 0;
    s_need = kk;
  }
  __syncthreads();
  for (uint32_t t = warp; t < ntiles; t += 8) {
    const uint32_t c = tile_count[t0 + t];
    const uint64_t base = r0 + static_cast<uint64_t>(t) * rec_slot;
    uint32_t at = 0;
    if (lane == 0 && c != 0) {
      at = atomicAdd(&s_n, c);
    }
    at = __shfl_sync(0xffffffffu, at, 0);
    for (uint32_t i = lane; i < c; i += 32) {
      if (at + i < kGroupKeyCap) {
        const SortKey k = make_sort_key(rec_score[base + i], rec_doc[base + i], desc);
        s_ks[at + i] = k.s;
        s_kd[at + i] = k.d;
      }
    }
  }
  __syncthreads();
  const uint32_t n = s_n;
  if (n <= kk || n > kGroupKeyCap) {
    return;  // nothing to prune (or, never by construction, more records than the staging holds): leave the tiles as is
  }
  for (int pass = 0; pass < 12; ++pass) {
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const SortKey prefix = s_prefix;
    for (uint32_t i = threadIdx.x; i < n; i += 256) {
      SortKey k;
      k.s = s_ks[i];
      k.d = s_kd[i];
      if (key_has_prefix(k, prefix, pass)) {
        atomicAdd(&s_hist[key_digit(k, pass)], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t need = s_need;
      int digit = 255;
      for (; digit > 0; --digit) {
        if (s_hist[digit] >= need) {
          break;
        }
        need -= s_hist[digit];
      }
      s_need = need;
      if (pass < 8) {
        s_prefix.s |= static_cast<uint64_t>(digit) << (56 - 8 * pass);
      } else {
        s_prefix.d |= static_cast<uint32_t>(digit) << (24 - 8 * (pass - 8));
      }
    }
    __syncthreads();
  }
  const SortKey thr = s_prefix;  // the kk-th best key of the group (keys are unique)
  // all records of the group are staged: the first tile's slot can be overwritten now
  for (uint32_t i = threadIdx.x; i < n; i += 256) {
    SortKey k;
    k.s = s_ks[i];
    k.d = s_kd[i];
    if (!key_greater(thr, k)) {
      const uint32_t pos = atomicAdd(&s_out, 1u);
      rec_doc[r0 + pos] = desc ? k.d : ~k.d;
      rec_score[r0 + pos] = unord_f64(desc ? k.s : ~k.s);
    }
  }
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < ntiles; t += 256) {
    tile_count[t0 + t] = t == 0 ? s_out : 0u;
  }
}

__global__ void __launch_bounds__(256)
topk_group_kernel(BatchView bv, const TopkGroup* __restrict__ groups, uint64_t tile_base, uint32_t rec_slot,
                  uint32_t* __restrict__ tile_count, uint32_t* __restrict__ rec_doc, double* __restrict__ rec_score,
                  int descending, uint32_t kk) {
  topk_group_body(bv, groups[blockIdx.x], tile_base, rec_slot, tile_count, rec_doc, rec_score, descending, kk);
}

// Streamed form: the groups are never listed. Query q has q_group_off[q+1] - q_group_off[q] groups of group_tiles
// consecutive tiles (0 when it has fewer than 2 * group_tiles tiles; prefix sums by the planning tail); persistent
// CTAs pull group indices from a device-side counter.
__global__ void __launch_bounds__(256)
topk_groups_kernel(BatchView bv, uint32_t group_tiles, uint32_t rec_slot, uint32_t* __restrict__ tile_count,
                   uint32_t* __restrict__ rec_doc, double* __restrict__ rec_score, int descending, uint32_t kk) {
  __shared__ uint32_t s_group;
  __shared__ uint32_t s_gq;
  const uint32_t n_groups = bv.launch[kLaunchGroups];
  for (;;) {
    if (threadIdx.x < 32) {
      uint32_t gi = 0;
      if (threadIdx.x == 0) {
        gi = atomicAdd(bv.launch + kLaunchGroupNext, 1u);
      }
      gi = __shfl_sync(0xffffffffu, gi, 0);
      uint32_t q = 0;
      if (gi < n_groups) {
        q = warp_upper_bound_u64(bv.q_group_off, bv.n_queries + 1, gi) - 1u;
      }
      if (threadIdx.x == 0) {
        s_group = gi;
        s_gq = q;
      }
    }
    __syncthreads();
    const uint32_t gi = s_group;
    const uint32_t q = s_gq;
    if (gi >= n_groups) {
      return;
    }
    const uint32_t ntiles = static_cast<uint32_t>(bv.q_tile_off[q + 1] - bv.q_tile_off[q]);
    TopkGroup g;
    g.q = q;
    g.t_begin = (gi - static_cast<uint32_t>(bv.q_group_off[q])) * group_tiles;
    g.t_end = min(ntiles, g.t_begin + group_tiles);
    topk_group_body(bv, g, 0, rec_slot, tile_count, rec_doc, rec_score, descending, kk);
    __syncthreads();
  }
}

// Full ascending result sets (the Index::Search* style calls), two launches. set_tile_ranks_kernel: one CTA per query
// turns the tiles' record counts into the rank of each tile's first record inside the query's set (written over
// tile_total, which the counting pass no longer needs). gather_tiles_kernel: one CTA per TILE copies its records to
// set_off[q] + rank, coalesced -- a query with hundreds of thousands of results is gathered by thousands of CTAs, not
// by one thread per tile of a single CTA.
__global__ void __launch_bounds__(256)
set_tile_ranks_kernel(BatchView bv, uint32_t q_first, uint64_t tile_base, const uint32_t* __restrict__ tile_count,
                      uint32_t* __restrict__ tile_rank) {
  __shared__ uint64_t s_scan[8];
  __shared__ uint64_t s_carry;
  const uint32_t q = q_first + blockIdx.x;
  const uint64_t t0 = bv.q_tile_off[q] - tile_base;
  const uint32_t ntiles = static_cast<uint32_t>(bv.q_tile_off[q + 1] - bv.q_tile_off[q]);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    s_carry = 0;
  }
  __syncthreads();
  for (uint32_t tb = 0; tb < ntiles; tb += blockDim.x) {
    const uint32_t t = tb + threadIdx.x;
    const uint32_t c = t < ntiles ? tile_count[t0 + t] : 0;
    uint64_t inc = c;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const uint64_t o = __shfl_up_sync(0xffffffffu, inc, s);
      if (lane >= static_cast<unsigned>(s)) {
        inc += o;
      }
    }
    if (lane == 31) {
      s_scan[warp] = inc;
    }
    __syncthreads();
    uint64_t prefix = s_carry;
    uint64_t chunk_total = 0;
    for (unsigned w = 0; w < 8; ++w) {
      if (w < warp) {
        prefix += s_scan[w];
      }
      chunk_total += s_scan[w];
    }
    if (t < ntiles) {
      tile_rank[t0 + t] = static_cast<uint32_t>(prefix + inc - c);  // a set never exceeds the 2^32 documents of a shard
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      s_carry += chunk_total;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
gather_tiles_kernel(BatchView bv, uint64_t tile_base, uint32_t rec_slot, const uint32_t* __restrict__ tile_count,
                    const uint32_t* __restrict__ tile_rank, const uint32_t* __restrict__ rec_doc,
                    const uint64_t* __restrict__ set_off, uint32_t* __restrict__ out) {
  const uint64_t slot = blockIdx.x;  // tile slot inside the chunk
  const uint32_t c = tile_count[slot];
  if (c == 0) {
    return;
  }
  const uint32_t q = __ldg(bv.tile_query + tile_base + slot);
  uint32_t* const dst = out + set_off[q] + tile_rank[slot];
  const uint32_t* const src = rec_doc + slot * rec_slot;
  for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
    dst[i] = src[i];
  }
}

// The gathered per-shard runs: four arrays per shard, shard s at base + s * pitch (bytes). Dense [G][Q][S] arrays
// and the packed one-buffer-per-shard record of mgx_shard_record_bytes() are both expressed this way.
struct ShardRuns {
  const uint8_t* ids_base;
  const uint8_t* scores_base;
  const uint8_t* count_base;
  const uint8_t* total_base;
  uint64_t ids_pitch;
  uint64_t scores_pitch;
  uint64_t count_pitch;
  uint64_t total_pitch;
  __device__ __forceinline__ const uint32_t* ids(uint32_t s) const {
    return reinterpret_cast<const uint32_t*>(ids_base + s * ids_pitch);
  }
  __device__ __forceinline__ const double* scores(uint32_t s) const {
    return reinterpret_cast<const double*>(scores_base + s * scores_pitch);
  }
  __device__ __forceinline__ const uint32_t* count(uint32_t s) const {
    return reinterpret_cast<const uint32_t*>(count_base + s * count_pitch);
  }
  __device__ __forceinline__ const uint64_t* total(uint32_t s) const {
    return reinterpret_cast<const uint64_t*>(total_base + s * total_pitch);
  }
};

// ------------------------------------------------------------------ shard merge (rank based, no sort)
// Every shard's run is already in final order. A record's final position is the
// number of records (over all shards) that beat it, found by binary search in
// each other run; records with position in [0, K) are written there.
__global__ void __launch_bounds__(256)
merge_topk_kernel(uint32_t n_shards, uint64_t n_queries, uint64_t stride, int compute_score, int descending,
                  uint32_t limit, uint32_t offset, ShardRuns runs, uint32_t* __restrict__ ids_out,
                  double* __restrict__ scores_out, uint32_t* __restrict__ count_out, uint64_t* __restrict__ total_out) {
  const uint64_t q = blockIdx.x;
  const bool desc = descending != 0;
  uint64_t total = 0;
  uint64_t have = 0;
  for (uint32_t s = 0; s < n_shards; ++s) {
    total += runs.total(s)[q];
    have += runs.count(s)[q];
  }
  // shards return their best (offset + limit) records with offset 0; the offset is applied here
  const uint64_t skip = umin64(have, offset);
  const uint32_t k_out =
      static_cast<uint32_t>(umin64(umin64(have - skip, stride), limit == 0 ? have - skip : limit));
  if (threadIdx.x == 0) {
    total_out[q] = total;
    count_out[q] = k_out;
  }
  for (uint32_t s = 0; s < n_shards; ++s) {
    const uint32_t c = runs.count(s)[q];
    const uint32_t* my_ids = runs.ids(s) + q * stride;
    const double* my_scores = runs.scores(s) + q * stride;
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
      uint32_t rank = i;
      if (compute_score) {
        const SortKey me = make_sort_key(my_scores[i], my_ids[i], desc);
        for (uint32_t o = 0; o < n_shards; ++o) {
          if (o == s) {
            continue;
          }
          const uint32_t oc = runs.count(o)[q];
          const uint32_t* o_ids = runs.ids(o) + q * stride;
          const double* o_scores = runs.scores(o) + q * stride;
          uint32_t lo = 0;  // number of records in run o that beat `me`
          uint32_t hi = oc;
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const SortKey k = make_sort_key(o_scores[mid], o_ids[mid], desc);
            if (key_greater(k, me)) {
              lo = mid + 1;
            } else {
              hi = mid;
            }
          }
          rank += lo;
        }
      } else {
        // ascending ids over disjoint ascending shard ranges: concatenate in shard order
        for (uint32_t o = 0; o < s; ++o) {
          rank += runs.count(o)[q];
        }
      }
      if (rank >= skip && rank - skip < k_out) {
        ids_out[q * stride + (rank - skip)] = my_ids[i];
        if (compute_score && scores_out != nullptr) {
          scores_out[q * stride + (rank - skip)] = my_scores[i];
        }
      }
    }
  }
}

// ------------------------------------------------------------------ stand-alone scoring / sorting
// BM25Scorer::ScoreDocuments for explicit candidates: one warp per candidate.
__global__ void __launch_bounds__(256)
score_docs_kernel(IndexView iv, const uint32_t* __restrict__ cands, uint64_t n_cands,
                  const uint8_t* __restrict__ term_bytes, const uint32_t* __restrict__ term_boff,
                  const double* __restrict__ idf, uint32_t n_terms, ScoreParams sp, double* __restrict__ out) {
  __shared__ __align__(16) uint8_t s_text[8][kStageBuf];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const uint64_t c = static_cast<uint64_t>(blockIdx.x) * 8 + warp;
  if (c >= n_cands) {
    return;
  }
  const uint32_t pos = local_of(iv, cands[c]);
  double score = 0.0;
  if (pos != kNone) {
    DocText d = doc_open(iv, pos, s_text[warp]);
    if (d.len > 0) {
      const double dl = static_cast<double>(iv.doc_len[pos]);
      const double length_norm = __dadd_rn(__dsub_rn(1.0, sp.b), __ddiv_rn(__dmul_rn(sp.b, dl), sp.avgdl_clamped));
      for (uint32_t i = 0; i < n_terms; ++i) {
        const uint32_t tl = term_boff[i + 1] - term_boff[i];
        const uint32_t tf_u = doc_count_term(d, term_bytes + term_boff[i], tl, false);
        if (tf_u != 0) {
          const double tf = static_cast<double>(tf_u);
          const double numerator = __dmul_rn(tf, __dadd_rn(sp.k1, 1.0));
          const double denominator = __dadd_rn(tf, __dmul_rn(sp.k1, length_norm));
          score = __dadd_rn(score, __ddiv_rn(__dmul_rn(idf[i], numerator), denominator));
        }
      }
    }
  }
  if (lane == 0) {
    out[c] = score;
  }
}

__global__ void idf_plain_kernel(const uint64_t* __restrict__ dfs, uint32_t n, uint64_t total_docs,
                                 double* __restrict__ idf) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  double v = 0.0;
  if (total_docs != 0) {
    uint64_t df = dfs[i];
    if (df > total_docs) {
      df = total_docs;
    }
    const double nn = static_cast<double>(total_docs);
    const double dd = static_cast<double>(df);
    v = log(__dadd_rn(__ddiv_rn(__dadd_rn(__dsub_rn(nn, dd), 0.5), __dadd_rn(dd, 0.5)), 1.0));
  }
  idf[i] = v;
}

// SortByScore over explicit arrays: single CTA, the windowed top-k of topk_kernel. limit 0 = everything from offset.
__global__ void __launch_bounds__(256)
sort_by_score_kernel(const uint32_t* __restrict__ docs, const double* __restrict__ scores, uint64_t n, int descending,
                     uint32_t limit, uint32_t offset, uint32_t* __restrict__ out, uint32_t* __restrict__ out_count) {
  __shared__ TopkShared sh;
  const bool desc = descending != 0;
  const uint64_t want_begin = umin64(offset, n);
  const uint64_t want_end = limit == 0 ? n : umin64(n, static_cast<uint64_t>(offset) + limit);
  const uint64_t n_out = want_end - want_begin;
  if (threadIdx.x == 0) {
    *out_count = static_cast<uint32_t>(n_out);
  }
  if (n_out == 0) {
    return;
  }
  auto scan = [&](auto f) {
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
      f(make_sort_key(scores[i], docs[i], desc));
    }
  };
  block_topk_window(sh, scan, n, want_begin, n_out, [&](uint64_t i, const SortKey k) { out[i] = desc ? k.d : ~k.d; });
}

BatchView make_batch_view(Batch& b) {
  BatchView v{};
  v.term_bytes = b.d_term_bytes.p;
  v.term_boff = b.d_term_boff.p;
  v.term_koff = b.d_term_koff.p;
  v.key_list = b.d_key_list.p;
  v.key_len = b.d_key_len.p;
  v.key_toff = b.d_key_toff.p;
  v.t_est = b.d_t_est.p;
  v.t_df_tiles = b.d_t_df_tiles.p;
  v.t_df_tile_off = b.d_t_df_tile_off.p;
  v.t_df = b.d_t_df.p;
  v.n_terms = b.n_terms;
  v.q_toff = b.d_q_toff.p;
  v.q_tids = b.d_q_tids.p;
  v.q_noff = b.d_q_noff.p;
  v.q_ntids = b.d_q_ntids.p;
  v.q_loff = b.d_q_loff.p;
  v.q_list = b.d_q_list.p;
  v.q_list_len = b.d_q_list_len.p;
  v.q_nlists = b.d_q_nlists.p;
  v.q_flags = b.d_q_flags.p;
  v.q_host_flags = b.d_q_host_flags.p;
  v.q_threshold = b.d_q_threshold.p;
  v.q_poff = b.d_q_poff.p;
  v.prog_op = b.d_prog_op.p;
  v.prog_arg = b.d_prog_arg.p;
  v.q_coff = b.d_q_coff.p;
  v.q_conj = b.d_q_conj.p;
  v.q_foff = b.d_q_foff.p;
  v.filters = b.d_filters.p;
  v.q_driver_len = b.d_q_driver_len.p;
  v.q_ntiles = b.d_q_ntiles.p;
  v.q_tile_off = b.d_q_tile_off.p;
  v.q_group_off = b.d_q_group_off.p;
  v.q_idf = b.d_q_idf.p;
  v.n_queries = b.n_queries;
  v.explicit_ids = b.explicit_driver.d_ids;
  v.explicit_n = static_cast<uint32_t>(b.explicit_driver.n);
  v.stats = b.d_stats.p;
  v.q_thr = b.d_q_thr.p;
  v.df_tile_desc = b.d_df_tile_desc.p;
  v.key_ref = b.d_key_ref.p;
  v.tile_query = b.d_tile_query.p;
  v.stream_slots = b.d_stream_slots.p;
  v.stream_entries = b.d_stream_entries.p;
  v.stream_bloom = b.d_stream_bloom.p;
  v.stream_len8_mask = b.stream_len8_mask;
  v.stream_len12_mask = b.stream_len12_mask;
  v.stream_slot_mask = b.n_stream_slots > 0 ? b.n_stream_slots - 1 : 0;
  v.df_mode = b.d_df_mode.p;
  v.launch = b.streamed ? b.d_launch.p : nullptr;
  v.term_flags = b.d_term_flags.p;
  v.q_tids0 = b.d_q_tids0.p;
  v.key_glen = b.d_key_glen.p;
  return v;
}

template <typename T>
void upload(DevBuf<T>& dst, const std::vector<T>& src, cudaStream_t stream, uint64_t* bytes) {
  dst.alloc(src.size());
  if (!src.empty()) {
    MGX_CUDA(cudaMemcpyAsync(dst.p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
    *bytes += src.size() * sizeof(T);
  }
}

unsigned grid_for(uint64_t n, unsigned block) { return static_cast<unsigned>((n + block - 1) / block); }

}  // namespace

// ------------------------------------------------------------------ batch timing / accounting
void Batch::time_begin(int kind) {
  Timed t{nullptr, nullptr, kind};
  MGX_CUDA(cudaEventCreate(&t.a));
  MGX_CUDA(cudaEventCreate(&t.b));
  MGX_CUDA(cudaEventRecord(t.a, stream));
  if (ev_first == nullptr) {
    MGX_CUDA(cudaEventCreate(&ev_first));
    MGX_CUDA(cudaEventRecord(ev_first, stream));
  }
  timed.push_back(t);
}

void Batch::time_end() { MGX_CUDA(cudaEventRecord(timed.back().b, stream)); }

void Batch::mark_last() {
  if (ev_last == nullptr) {
    MGX_CUDA(cudaEventCreate(&ev_last));
  }
  MGX_CUDA(cudaEventRecord(ev_last, stream));
}

Batch::~Batch() {
  for (cudaEvent_t& e : ev_x) {
    if (e != nullptr) {
      cudaEventDestroy(e);
      e = nullptr;
    }
  }
  for (auto& t : timed) {
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  if (ev_first != nullptr) {
    cudaEventDestroy(ev_first);
  }
  if (ev_last != nullptr) {
    cudaEventDestroy(ev_last);
  }
}

void Batch::collect_stats(mgx_batch_stats_t* out) {
  MGX_CUDA(cudaStreamSynchronize(stream));
  mgx_batch_stats_t s{};
  for (const auto& t : timed) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t.a, t.b) != cudaSuccess) {
      (void)cudaGetLastError();
      continue;
    }
    if (t.kind == 0) {
      s.ms_plan += ms;
    } else if (t.kind == 1) {
      s.ms_df_kernel += ms;
    } else if (t.kind == 2) {
      s.ms_and_kernel += ms;
    } else if (t.kind == 4) {
      s.ms_df_stream_kernel += ms;
    } else {
      s.ms_topk_kernel += ms;
    }
  }
  if (ev_first != nullptr && ev_last != nullptr) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev_first, ev_last) == cudaSuccess) {
      s.ms_total = ms;
    } else {
      (void)cudaGetLastError();
    }
  }
  unsigned long long h[kStatCount] = {0};
  if (d_stats.p != nullptr) {
    std::vector<unsigned long long> raw(static_cast<size_t>(kStatCount) * kStatStripes, 0);
    MGX_CUDA(cudaMemcpy(raw.data(), d_stats.p, raw.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int slot = 0; slot < kStatCount; ++slot) {
      for (int i = 0; i < kStatStripes; ++i) {
        h[slot] += raw[static_cast<size_t>(slot) * kStatStripes + i];
      }
    }
  }
  const uint64_t k = params.limit + params.offset;
  s.n_df_tiles = n_df_tiles;
  s.n_and_tiles = n_and_tiles;
  s.algo_bytes_intersect = h[kStatIntersectLists] + 4ULL * h[kStatResultDocs];
  s.algo_bytes_score = params.compute_score != 0 ? h[kStatScoreBytes] + 12ULL * k * n_queries : 0;
  s.algo_bytes_df = h[kStatDfBytes];
  s.algo_bytes_df_lists = h[kStatDfLists];
  s.h2d_bytes = h2d_bytes;
  s.d2h_bytes = d2h_bytes;
  s.driver_entries = h[kStatDriverEntries];
  if (streamed && status_copied && h_launch.p != nullptr) {
    s.n_df_tiles = h_launch.p[kLaunchDfUnits];  // units of kDfUnit entries
    s.n_and_tiles = h_launch.p[kLaunchAndTiles];
    uint32_t mode = 0;
    MGX_CUDA(cudaMemcpy(&mode, d_df_mode.p, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    h_df_mode = static_cast<int>(mode);
  }
  s.result_docs = h[kStatResultDocs];
  s.df_candidates = h[kStatDfCandidates];
  s.unique_terms = n_terms;
  s.df_stream_terms = h_df_mode != 0 ? n_stream_terms : 0;
  s.df_stream_bytes = h_df_mode != 0 ? ix->text_bytes : 0;
  s.df_stream_hits = h[kStatStreamHits];
  s.df_scanned_docs = h[kStatDfScanned];
  *out = s;
}

// ------------------------------------------------------------------ host orchestration
void Batch::recycle() {
  for (auto& t : timed) {
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  timed.clear();
  if (ev_first != nullptr) {
    cudaEventDestroy(ev_first);
    ev_first = nullptr;
  }
  if (ev_last != nullptr) {
    cudaEventDestroy(ev_last);
    ev_last = nullptr;
  }
  h_slot_tid.clear();
  explicit_driver = ExplicitDriver{};
  h2d_bytes = d2h_bytes = 0;
  n_df_tiles = n_and_tiles = driver_entries = 0;
  planned = df_done = searched = false;
  streamed = allow_streamed = status_copied = false;
  sharded_enqueued = false;
  global_order = false;
  rec_slot = kTile;
  sc = nullptr;
  serial = 0;
  h_df_mode = 0;
  n_stream_slots = n_stream_terms = 0;
}

namespace {
// Literal of a condition in the class of the column it addresses: ParseFilterValue (search_pipeline.cpp:954-993)
// for the typed path, the type interpretations of BuildTypeUnionBitmap (:1021-1094) for the bitmap path.
FilterPred resolve_filter(const Index& ix, const HostFilter& hf, bool bitmap_mode) {
  FilterPred p{};
  p.op_flags = (hf.op & 7u) | (bitmap_mode ? kFilterBitmapMode : 0u);
  const FilterColumn* col = hf.col < ix.columns.size() ? ix.columns[hf.col] : nullptr;
  if (col == nullptr || col->n_docs != ix.n_docs) {
    p.cls = kFcNone;
    return p;
  }
  p.cls = col->cls;
  p.values = col->values.p;
  p.nulls = col->nulls.p;
  const std::string& v = hf.literal;
  const char* b = v.data();
  const char* e = v.data() + v.size();
  bool valid = false;
  switch (col->cls) {
    case kFcBool:
      if (bitmap_mode) {
        valid = v == "1" || v == "true" || v == "0" || v == "false";
        p.c = (v == "1" || v == "true") ? 1 : 0;
      } else {
        valid = true;
        p.c = (v == "1" || v == "true") ? 1 : 0;  // parsed_value.bool_val
      }
      break;
    case kFcSigned: {
      int64_t r = 0;
      auto [ptr, ec] = std::from_chars(b, e, r);
      valid = ec == std::errc() && ptr == e;
      p.c = static_cast<uint64_t>(r);
      break;
    }
    case kFcUnsigned: {
      uint64_t r = 0;
      auto [ptr, ec] = std::from_chars(b, e, r);
      valid = ec == std::errc() && ptr == e;
      p.c = r;
      break;
    }
    case kFcDouble: {
      double r = 0.0;
      auto [ptr, ec] = std::from_chars(b, e, r);
      valid = ec == std::errc() && ptr == e;
      std::memcpy(&p.c, &r, sizeof(r));
      break;
    }
    case kFcString: {
      const auto it = std::lower_bound(col->dict.begin(), col->dict.end(), v);
      p.c = static_cast<uint64_t>(it - col->dict.begin());
      const bool found = it != col->dict.end() && *it == v;
      valid = bitmap_mode ? found : true;
      if (found) {
        p.op_flags |= kFilterFound;
      }
      break;
    }
    default:
      break;
  }
  if (valid) {
    p.op_flags |= kFilterValid;
  }
  return p;
}
}  // namespace

void build_stream_table(std::vector<HostTerm>& terms, HostStreamTable* out) {
  out->slots.clear();
  out->entries.clear();
  out->bloom.clear();
  out->len8_mask = out->len12_mask = 0;
  size_t n = 0;
  for (const HostTerm& t : terms) {
    n += t.streamable ? 1 : 0;
  }
  if (n == 0) {
    return;
  }
  uint32_t n_slots = 64;
  while (n_slots < 2 * n) {
    n_slots <<= 1;
  }
  // scratch vectors live in the table object: a pooled batch re-uses them (no allocation, no fresh pages)
  std::vector<HostStreamTable::Item>& items = out->items;
  std::vector<uint32_t>& count = out->count;
  items.clear();
  items.reserve(n);
  count.assign(n_slots, 0);
  out->bloom.assign(kStreamBloomWords, 0);
  for (size_t t = 0; t < terms.size(); ++t) {
    HostTerm& ht = terms[t];
    if (!ht.streamable) {
      continue;
    }
    const uint32_t tl = static_cast<uint32_t>(ht.bytes.size());
    const uint32_t len12 = std::min(tl, kStreamKeyBytes);
    const uint32_t len8 = std::min(tl, kStreamFilterBytes);
    uint32_t w[3] = {0, 0, 0};
    for (uint32_t i = 0; i < len12; ++i) {
      w[i >> 2] |= static_cast<uint32_t>(static_cast<uint8_t>(ht.bytes[i])) << (8 * (i & 3));
    }
    const uint32_t slot = stream_hash(w[0], w[1], w[2], len12) & (n_slots - 1);
    if (count[slot] >= kStreamMaxBucket || items.size() >= kStreamMaxTerms || t >= (1u << 24)) {
      ht.streamable = false;  // stays on the candidate-tile path
      continue;
    }
    ++count[slot];
    const uint32_t f1 = len8 > 4 ? (w[1] & low_bytes_mask(len8 - 4)) : 0u;
    const uint32_t bit = stream_filter_bit(w[0] & low_bytes_mask(len8), f1, len8);
    out->bloom[bit >> 5] |= 1u << (bit & 31);
    out->len8_mask |= 1u << len8;
    out->len12_mask |= 1u << len12;
    items.push_back({slot, {w[0], w[1], w[2], len12 | (static_cast<uint32_t>(t) << 8)}});
  }
  out->slots.resize(n_slots);
  uint32_t first = 0;
  for (uint32_t i = 0; i < n_slots; ++i) {
    out->slots[i] = (first << 8) | count[i];
    const uint32_t c = count[i];
    count[i] = first;  // from here on: the bucket's fill cursor
    first += c;
  }
  out->entries.resize(items.size());
  for (const HostStreamTable::Item& it : items) {
    out->entries[count[it.slot]++] = it.e;
  }
}

void batch_upload(Batch& b, std::vector<HostTerm>& terms, const std::vector<HostQuery>& queries,
                  const std::vector<uint32_t>& slot_tid, const std::vector<uint32_t>* xoff) {
  cudaStream_t st = b.stream;
  HostStreamTable& stream_table = b.h_stream_table;
  build_stream_table(terms, &stream_table);
  const size_t T = terms.size();
  const size_t Q = queries.size();
  b.n_queries = static_cast<uint32_t>(Q);
  b.n_out_queries = b.n_queries;
  b.h_xoff.clear();
  if (xoff != nullptr && !xoff->empty()) {
    b.h_xoff = *xoff;
    b.n_out_queries = static_cast<uint32_t>(xoff->size() - 1);
  }
  b.n_terms = static_cast<uint32_t>(T);
  b.n_slots = static_cast<uint32_t>(slot_tid.size());
  b.h_slot_tid = slot_tid;

  // ---- sizes first, so that every array is written ONCE, straight into the pinned staging buffer
  size_t n_bytes = 0, n_keys = 0;
  for (const HostTerm& t : terms) {
    n_bytes += t.bytes.size();
    n_keys += t.keys.size();
  }
  n_bytes += 16;  // zero tail: the kernels read term bytes in words
  size_t n_tids = 0, n_ntids = 0, n_prog = 0, n_conj = 0, n_preds = 0;
  for (const HostQuery& q : queries) {
    n_tids += q.terms.size();
    n_ntids += q.not_terms.size();
    n_prog += q.prog_ops.size();
    n_conj += q.conjuncts.size();
    n_preds += q.filters.size();
  }
  b.n_keys = static_cast<uint32_t>(n_keys);
  b.n_qterms = static_cast<uint32_t>(n_tids);
  StageLayout& L = b.layout;
  L = StageLayout{};
  L.n_terms = T;
  L.n_queries = Q;
  L.n_out_queries = b.n_out_queries;
  L.n_slots = slot_tid.size();
  L.n_bytes = n_bytes;
  L.n_keys = n_keys;
  L.n_tids = n_tids;
  L.n_ntids = n_ntids;
  L.n_prog = n_prog;
  L.n_conj = n_conj;
  L.n_preds = n_preds;
  L.n_sslots = stream_table.slots.size();
  L.n_sentries = stream_table.entries.size();
  L.n_sbloom = stream_table.bloom.size();
  L.n_xoff = b.h_xoff.size();
  L.stream_len8_mask = stream_table.len8_mask;
  L.stream_len12_mask = stream_table.len12_mask;
  L.assumed_all_valid_utf8 = b.ix->all_valid_utf8 ? 1 : 0;
  L.wide_words = static_cast<uint32_t>(b.ix->wide_words);
  const StageOffsets O = stage_offsets(L);
  const size_t total = O.total;
  const size_t i_bytes = O.i_bytes, i_boff = O.i_boff, i_koff = O.i_koff, i_keys = O.i_keys, i_raw = O.i_raw;
  const size_t i_toff = O.i_toff, i_tids = O.i_tids, i_tids0 = O.i_tids0, i_noff = O.i_noff, i_ntids = O.i_ntids;
  const size_t i_loff = O.i_loff, i_hflags = O.i_hflags, i_slot = O.i_slot, i_ktoff = O.i_ktoff, i_thr = O.i_thr;
  const size_t i_poff = O.i_poff, i_pops = O.i_pops, i_pargs = O.i_pargs, i_coff = O.i_coff, i_conj = O.i_conj;
  const size_t i_foff = O.i_foff, i_preds = O.i_preds, i_sslots = O.i_sslots, i_sentries = O.i_sentries;
  const size_t i_sbloom = O.i_sbloom, i_xoff = O.i_xoff;
  b.staging.reserve(total + 256);
  uint8_t* const S = b.staging.p;

  {  // terms
    uint8_t* bytes = S + i_bytes;
    uint32_t* boff = reinterpret_cast<uint32_t*>(S + i_boff);
    uint32_t* koff = reinterpret_cast<uint32_t*>(S + i_koff);
    uint64_t* keys = reinterpret_cast<uint64_t*>(S + i_keys);
    uint32_t* key_toff = reinterpret_cast<uint32_t*>(S + i_ktoff);
    uint8_t* raw = S + i_raw;
    uint32_t nb = 0, nk = 0;
    boff[0] = 0;
    koff[0] = 0;
    for (size_t t = 0; t < T; ++t) {
      const HostTerm& ht = terms[t];
      std::memcpy(bytes + nb, ht.bytes.data(), ht.bytes.size());
      nb += static_cast<uint32_t>(ht.bytes.size());
      boff[t + 1] = nb;
      for (size_t k = 0; k < ht.keys.size(); ++k) {
        keys[nk + k] = ht.keys[k];
        if (L.wide_words > 0) {  // the words behind the handle travel with the batch
          uint64_t* dst = reinterpret_cast<uint64_t*>(S + O.i_wide) + static_cast<size_t>(nk + k) * L.wide_words;
          if (ht.keys[k] != kInvalidKey) {
            std::memcpy(dst, host_wide_words(ht.keys[k], static_cast<int>(L.wide_words)), L.wide_words * 8);
          } else {
            std::memset(dst, 0, L.wide_words * 8);
          }
        }
        key_toff[nk + k] = k < ht.key_toff.size() ? ht.key_toff[k] : static_cast<uint32_t>(kNoTermOffset);
      }
      nk += static_cast<uint32_t>(ht.keys.size());
      koff[t + 1] = nk;
      raw[t] = static_cast<uint8_t>((ht.raw ? 1 : 0) | (ht.exact_single ? 2 : 0) | (ht.streamable ? 4 : 0) |
                                    (ht.payload_tf ? 8 : 0));
    }
    raw[T] = 0;
    std::memset(bytes + nb, 0, 16);
  }
  size_t list_cap = 0;
  {  // queries
    uint32_t* toff = reinterpret_cast<uint32_t*>(S + i_toff);
    uint32_t* tids = reinterpret_cast<uint32_t*>(S + i_tids);
    uint32_t* tids0 = reinterpret_cast<uint32_t*>(S + i_tids0);
    uint32_t* noff = reinterpret_cast<uint32_t*>(S + i_noff);
    uint32_t* ntids = reinterpret_cast<uint32_t*>(S + i_ntids);
    uint32_t* loff = reinterpret_cast<uint32_t*>(S + i_loff);
    uint32_t* hflags = reinterpret_cast<uint32_t*>(S + i_hflags);
    uint32_t* thresholds = reinterpret_cast<uint32_t*>(S + i_thr);
    uint32_t* poff = reinterpret_cast<uint32_t*>(S + i_poff);
    uint8_t* prog_ops = S + i_pops;
    uint32_t* prog_args = reinterpret_cast<uint32_t*>(S + i_pargs);
    uint32_t* coff = reinterpret_cast<uint32_t*>(S + i_coff);
    uint32_t* conj = reinterpret_cast<uint32_t*>(S + i_conj);
    uint32_t* foff = reinterpret_cast<uint32_t*>(S + i_foff);
    FilterPred* preds = reinterpret_cast<FilterPred*>(S + i_preds);
    uint32_t nt = 0, nn = 0, np = 0, nc = 0, nf = 0, nl = 0;
    toff[0] = noff[0] = loff[0] = poff[0] = coff[0] = foff[0] = 0;
    for (size_t q = 0; q < Q; ++q) {
      const HostQuery& hq = queries[q];
      if (!hq.filters.empty()) {
        bool all_bitmap = true;  // AllFiltersHaveBitmapSupport, search_pipeline.cpp:995-1003
        for (const HostFilter& hf : hq.filters) {
          all_bitmap = all_bitmap && (hf.op == 0 || hf.op == 1);
        }
        for (const HostFilter& hf : hq.filters) {
          preds[nf++] = resolve_filter(*b.ix, hf, all_bitmap);
        }
      }
      foff[q + 1] = nf;
      for (size_t i = 0; i < hq.prog_ops.size(); ++i) {
        prog_ops[np + i] = hq.prog_ops[i];
        prog_args[np + i] = hq.prog_args[i];
      }
      np += static_cast<uint32_t>(hq.prog_ops.size());
      for (uint32_t c : hq.conjuncts) {
        conj[nc++] = c;
      }
      poff[q + 1] = np;
      coff[q + 1] = nc;
      thresholds[q] = hq.threshold;
      uint32_t cap = (hq.flags & kQProgram) != 0 ? 1u : 0u;  // a program's only list is its driver
      for (uint32_t tid : hq.terms) {
        tids0[nt] = tid;
        tids[nt++] = tid;
        cap += static_cast<uint32_t>(terms[tid].keys.size());
      }
      for (uint32_t tid : hq.not_terms) {
        ntids[nn++] = tid;
      }
      toff[q + 1] = nt;
      noff[q + 1] = nn;
      nl += cap;
      loff[q + 1] = nl;
      hflags[q] = hq.flags;
      if ((hq.flags & kQProgram) != 0 && std::getenv("MGX_NO_OPTIMISTIC") == nullptr) {
        // text leaves only under AND / OR / ATLEAST (never under a NOT)?
        bool any_text = false;
        bool monotone = true;
        std::vector<uint8_t> has_text;  // per stack entry: its subtree holds a text leaf
        for (size_t i = 0; i < hq.prog_ops.size() && monotone; ++i) {
          const uint8_t op = hq.prog_ops[i];
          const uint32_t arg = hq.prog_args[i];
          if (op == kOpTerm) {
            const bool text = arg < terms.size() && terms[arg].keys.empty() && !terms[arg].bytes.empty();
            has_text.push_back(text ? 1 : 0);
            any_text = any_text || text;
          } else if (op == kOpFuzzyText) {
            has_text.push_back(1);
            any_text = true;
          } else if (op == kOpAnd || op == kOpOr || op == kOpAtLeast) {
            const size_t n_kids = op == kOpAtLeast ? (arg & 0xFFFFu) : arg;
            if (n_kids > has_text.size()) {
              monotone = false;
              break;
            }
            uint8_t any = 0;
            for (size_t c = 0; c < n_kids; ++c) {
              any |= has_text[has_text.size() - 1 - c];
            }
            has_text.resize(has_text.size() - n_kids);
            has_text.push_back(any);
          } else {  // NOT
            if (has_text.empty() || has_text.back() != 0) {
              monotone = false;
            }
          }
        }
        if (any_text && monotone) {
          hflags[q] |= kQOptimistic;
        }
      }
    }
    hflags[Q] = 0;
    thresholds[Q] = 1;
    list_cap = nl;
    L.list_cap = nl;
  }
  if (!slot_tid.empty()) {
    std::memcpy(S + i_slot, slot_tid.data(), slot_tid.size() * 4);
  }
  if (!b.h_xoff.empty()) {
    std::memcpy(S + i_xoff, b.h_xoff.data(), b.h_xoff.size() * 4);
  }
  if (!stream_table.slots.empty()) {
    std::memcpy(S + i_sslots, stream_table.slots.data(), stream_table.slots.size() * 4);
    std::memcpy(S + i_sentries, stream_table.entries.data(), stream_table.entries.size() * sizeof(StreamEntry));
    std::memcpy(S + i_sbloom, stream_table.bloom.data(), stream_table.bloom.size() * 4);
  }
  (void)st;
  (void)list_cap;
  batch_bind(b);
}

StageOffsets stage_offsets(const StageLayout& L) {
  StageOffsets O{};
  size_t total = 0;
  auto add = [&](size_t nbytes) {
    const size_t at = total;
    total += DevArena::padded(nbytes == 0 ? 1 : nbytes);
    return at;
  };
  const size_t T = L.n_terms, Q = L.n_queries;
  O.i_bytes = add(L.n_bytes);
  O.i_boff = add((T + 1) * 4);
  O.i_koff = add((T + 1) * 4);
  O.i_keys = add(L.n_keys * 8);
  O.i_raw = add(T + 1);
  O.i_toff = add((Q + 1) * 4);
  O.i_tids = add(L.n_tids * 4);
  O.i_tids0 = add(L.n_tids * 4);
  O.i_noff = add((Q + 1) * 4);
  O.i_ntids = add(L.n_ntids * 4);
  O.i_loff = add((Q + 1) * 4);
  O.i_hflags = add((Q + 1) * 4);
  O.i_slot = add(L.n_slots * 4);
  O.i_ktoff = add(L.n_keys * 4);
  O.i_thr = add((Q + 1) * 4);
  O.i_poff = add((Q + 1) * 4);
  O.i_pops = add(L.n_prog);
  O.i_pargs = add(L.n_prog * 4);
  O.i_coff = add((Q + 1) * 4);
  O.i_conj = add(L.n_conj * 4);
  O.i_foff = add((Q + 1) * 4);
  O.i_preds = add(L.n_preds * sizeof(FilterPred));
  O.i_sslots = add(L.n_sslots * 4);
  O.i_sentries = add(L.n_sentries * sizeof(StreamEntry));
  O.i_sbloom = add(L.n_sbloom * 4);
  O.i_xoff = add(L.n_xoff * 4);
  O.i_wide = add(L.n_keys * 8 * L.wide_words);
  O.total = total;
  return O;
}

// The compiled batch lies in b.staging in the layout b.layout describes: copy it to the device in ONE transfer, point
// the batch's arrays into it and carve the device-only planning arrays. Shared by the local compile (batch_upload)
// and by a batch that was compiled elsewhere (batch_import).
void batch_bind(Batch& b) {
  cudaStream_t st = b.stream;
  const StageLayout& L = b.layout;
  const StageOffsets O = stage_offsets(L);
  const size_t T = L.n_terms, Q = L.n_queries, n_keys = L.n_keys, n_tids = L.n_tids, n_bytes = L.n_bytes;
  const size_t n_ntids = L.n_ntids, n_prog = L.n_prog, n_conj = L.n_conj, n_preds = L.n_preds, list_cap = L.list_cap;
  const size_t total = O.total;
  const size_t i_bytes = O.i_bytes, i_boff = O.i_boff, i_koff = O.i_koff, i_keys = O.i_keys, i_raw = O.i_raw;
  const size_t i_toff = O.i_toff, i_tids = O.i_tids, i_tids0 = O.i_tids0, i_noff = O.i_noff, i_ntids = O.i_ntids;
  const size_t i_loff = O.i_loff, i_hflags = O.i_hflags, i_slot = O.i_slot, i_ktoff = O.i_ktoff, i_thr = O.i_thr;
  const size_t i_poff = O.i_poff, i_pops = O.i_pops, i_pargs = O.i_pargs, i_coff = O.i_coff, i_conj = O.i_conj;
  const size_t i_foff = O.i_foff, i_preds = O.i_preds, i_sslots = O.i_sslots, i_sentries = O.i_sentries;
  const size_t i_sbloom = O.i_sbloom, i_xoff = O.i_xoff;
  b.n_queries = static_cast<uint32_t>(Q);
  b.n_out_queries = static_cast<uint32_t>(L.n_out_queries);
  b.n_terms = static_cast<uint32_t>(T);
  b.n_slots = static_cast<uint32_t>(L.n_slots);
  b.n_keys = static_cast<uint32_t>(n_keys);
  b.n_qterms = static_cast<uint32_t>(n_tids);
  b.n_stream_slots = static_cast<uint32_t>(L.n_sslots);
  b.n_stream_terms = static_cast<uint32_t>(L.n_sentries);
  b.stream_len8_mask = L.stream_len8_mask;
  b.stream_len12_mask = L.stream_len12_mask;
  b.in_arena.reserve(total + 256, true);
  uint8_t* base = b.in_arena.take<uint8_t>(total);
  MGX_CUDA(cudaMemcpyAsync(base, b.staging.p, total, cudaMemcpyHostToDevice, st));
  b.h2d_bytes = total;
  auto at = [&](size_t off) { return base + off; };
  b.d_term_bytes.borrow(at(i_bytes), n_bytes);
  b.d_term_boff.borrow(reinterpret_cast<uint32_t*>(at(i_boff)), T + 1);
  b.d_term_koff.borrow(reinterpret_cast<uint32_t*>(at(i_koff)), T + 1);
  b.d_keys.borrow(reinterpret_cast<uint64_t*>(at(i_keys)), n_keys);
  b.d_term_flags.borrow(at(i_raw), T + 1);
  b.d_q_toff.borrow(reinterpret_cast<uint32_t*>(at(i_toff)), Q + 1);
  b.d_q_tids.borrow(reinterpret_cast<uint32_t*>(at(i_tids)), n_tids);
  b.d_q_tids0.borrow(reinterpret_cast<uint32_t*>(at(i_tids0)), n_tids);
  b.d_q_noff.borrow(reinterpret_cast<uint32_t*>(at(i_noff)), Q + 1);
  b.d_q_ntids.borrow(reinterpret_cast<uint32_t*>(at(i_ntids)), n_ntids);
  b.d_q_loff.borrow(reinterpret_cast<uint32_t*>(at(i_loff)), Q + 1);
  b.d_q_host_flags.borrow(reinterpret_cast<uint32_t*>(at(i_hflags)), Q + 1);
  b.d_slot_tid.borrow(reinterpret_cast<uint32_t*>(at(i_slot)), L.n_slots);
  b.d_key_toff.borrow(reinterpret_cast<uint32_t*>(at(i_ktoff)), n_keys);
  b.d_q_threshold.borrow(reinterpret_cast<uint32_t*>(at(i_thr)), Q + 1);
  b.d_q_poff.borrow(reinterpret_cast<uint32_t*>(at(i_poff)), Q + 1);
  b.d_prog_op.borrow(at(i_pops), n_prog);
  b.d_prog_arg.borrow(reinterpret_cast<uint32_t*>(at(i_pargs)), n_prog);
  b.d_q_coff.borrow(reinterpret_cast<uint32_t*>(at(i_coff)), Q + 1);
  b.d_q_conj.borrow(reinterpret_cast<uint32_t*>(at(i_conj)), n_conj);
  b.d_q_foff.borrow(reinterpret_cast<uint32_t*>(at(i_foff)), Q + 1);
  b.d_filters.borrow(reinterpret_cast<FilterPred*>(at(i_preds)), n_preds);
  b.d_stream_slots.borrow(reinterpret_cast<uint32_t*>(at(i_sslots)), L.n_sslots);
  b.d_stream_entries.borrow(reinterpret_cast<StreamEntry*>(at(i_sentries)), L.n_sentries);
  b.d_stream_bloom.borrow(reinterpret_cast<uint32_t*>(at(i_sbloom)), L.n_sbloom);
  b.d_xoff.borrow(reinterpret_cast<uint32_t*>(at(i_xoff)), L.n_xoff);
  b.d_wide.borrow(L.wide_words > 0 ? reinterpret_cast<uint64_t*>(at(O.i_wide)) : nullptr, n_keys * L.wide_words);

  // ---- device-only planning arrays from the work arena
  const size_t K = n_keys;
  const size_t Lc = list_cap;
  const size_t scan_elems = scan_scratch_elems(std::max<uint64_t>(T, Q)) + 8;
  size_t work = 0;
  for (size_t nbytes : {K * sizeof(KeyRef), K * 4, K * 4, T * 8, T * 4, (T + 1) * 8, (T + K) * 8, Lc * 4, Lc * 4, Q * 4, Q * 4, Q * 4, Q * 4,
                        (Q + 1) * 8, (Q + 1) * 8, n_tids * 8, static_cast<size_t>(kStatCount) * kStatStripes * 8, (Q + 1) * 8,
                        scan_elems * 8, static_cast<size_t>(64), static_cast<size_t>(kLaunchCount) * 4}) {
    work += DevArena::padded(nbytes == 0 ? 1 : nbytes);
  }
  b.work_arena.reserve(work + 1024, true);
  b.d_key_ref.borrow(b.work_arena.take<KeyRef>(K), K);
  b.d_key_list.borrow(b.work_arena.take<uint32_t>(K), K);
  b.d_key_len.borrow(b.work_arena.take<uint32_t>(K), K);
  b.d_t_est.borrow(b.work_arena.take<uint64_t>(T), T);
  b.d_t_df_tiles.borrow(b.work_arena.take<uint32_t>(T), T);
  b.d_t_df_tile_off.borrow(b.work_arena.take<uint64_t>(T + 1), T + 1);
  // [per-term df | per-key posting sizes in upload order]: ONE array, so the sharded pipeline sums both in one exchange
  uint64_t* xchg = b.work_arena.take<uint64_t>(T + K);
  b.d_t_df.borrow(xchg, T);
  b.d_key_glen.borrow(xchg + T, K);
  b.d_q_list.borrow(b.work_arena.take<uint32_t>(Lc), Lc);
  b.d_q_list_len.borrow(b.work_arena.take<uint32_t>(Lc), Lc);
  b.d_q_nlists.borrow(b.work_arena.take<uint32_t>(Q), Q);
  b.d_q_flags.borrow(b.work_arena.take<uint32_t>(Q), Q);
  b.d_q_driver_len.borrow(b.work_arena.take<uint32_t>(Q), Q);
  b.d_q_ntiles.borrow(b.work_arena.take<uint32_t>(Q), Q);
  b.d_q_tile_off.borrow(b.work_arena.take<uint64_t>(Q + 1), Q + 1);
  b.d_q_group_off.borrow(b.work_arena.take<uint64_t>(Q + 1), Q + 1);
  b.d_q_idf.borrow(b.work_arena.take<double>(n_tids), n_tids);
  b.d_scan_scratch.borrow(b.work_arena.take<uint64_t>(scan_elems), scan_elems);
  // accounting counters, df choice and the launch block of streamed batches lie back to back: one memset clears them
  const size_t n_stats = static_cast<size_t>(kStatCount) * kStatStripes;
  b.d_stats.borrow(b.work_arena.take<unsigned long long>(n_stats), n_stats);
  b.d_df_mode.borrow(b.work_arena.take<uint32_t>(2), 2);
  b.d_q_thr.borrow(b.work_arena.take<unsigned long long>(Q + 1), Q + 1);  // inside the cleared range
  b.d_launch.borrow(b.work_arena.take<uint32_t>(kLaunchCount), kLaunchCount);
  b.in_base = base;
  b.in_total = total;
  batch_clear_counters(b);
}

void batch_import(Batch& b, const StageLayout& layout, const uint8_t* bytes) {
  if (layout.n_preds != 0) {
    set_last_error("a batch with column conditions cannot be shared (its predicates hold device addresses)");
    throw CudaFailure{MGX_ERR_UNSUPPORTED};
  }
  const StageOffsets O = stage_offsets(layout);
  b.layout = layout;
  b.staging.reserve(O.total + 256);
  std::memcpy(b.staging.p, bytes, O.total);
  if (layout.assumed_all_valid_utf8 != 0 && !b.ix->all_valid_utf8) {
    // the compiling shard held only valid UTF-8 and chose the streaming df pass for some terms; this shard cannot
    // (DESIGN: exactness of the two shortcuts), so those terms go back to the candidate-tile path here
    uint8_t* raw = b.staging.p + O.i_raw;
    for (uint64_t t = 0; t < layout.n_terms; ++t) {
      raw[t] &= static_cast<uint8_t>(~4u);
    }
  }
  b.h_slot_tid.clear();
  b.h_xoff.clear();
  if (layout.n_xoff != 0) {
    const uint32_t* x = reinterpret_cast<const uint32_t*>(b.staging.p + O.i_xoff);
    b.h_xoff.assign(x, x + layout.n_xoff);
  }
  batch_bind(b);
  if (layout.assumed_all_valid_utf8 != 0 && !b.ix->all_valid_utf8) {
    b.n_stream_slots = b.n_stream_terms = 0;
  }
}

void batch_clear_counters(Batch& b) {
  uint8_t* first = reinterpret_cast<uint8_t*>(b.d_stats.p);
  uint8_t* last = reinterpret_cast<uint8_t*>(b.d_launch.p + kLaunchCount);
  MGX_CUDA(cudaMemsetAsync(first, 0, static_cast<size_t>(last - first), b.stream));
}

namespace {
std::atomic<uint64_t> g_batch_serial{0};

// tile -> term / tile -> query maps of a planned batch, in the shared (index, stream) workspace. Another batch
// planned on the same stream in the meantime replaces them, so every stage re-checks the owner.
void ensure_tile_maps(Batch& b) {
  cudaStream_t st = b.stream;
  if (b.sc == nullptr) {
    b.sc = b.ix->scratch_for(st);
  }
  if (b.serial == 0) {
    b.serial = ++g_batch_serial;
  }
  SearchScratch& sc = *b.sc;
  const uint64_t n_and_tiles = b.n_queries > 0 ? b.h_q_tile_off[b.n_queries] : 0;
  if (sc.map_owner != b.serial) {
    // floors sized for batches of a few thousand queries, so that a steady stream of batches never reallocates
    // (cudaMalloc / cudaFree cost tens to hundreds of milliseconds and synchronise the device)
    sc.df_tile_desc.reserve(2 * std::max<uint64_t>(b.n_df_tiles, 1ULL << 19));
    sc.tile_query.reserve(std::max<uint64_t>(n_and_tiles, 1ULL << 16));
    if (b.n_df_tiles > 0) {
      fill_df_desc_kernel<<<grid_for(static_cast<uint64_t>(b.n_terms) * 32, 256), 256, 0, st>>>(
          b.d_t_df_tile_off.p, b.d_term_koff.p, b.d_key_ref.p, b.n_terms, reinterpret_cast<DfTileDesc*>(sc.df_tile_desc.p));
      MGX_LAUNCH_CHECK();
    }
    if (n_and_tiles > 0) {
      fill_tile_map_kernel<<<grid_for(static_cast<uint64_t>(b.n_queries) * 32, 256), 256, 0, st>>>(
          b.d_q_tile_off.p, b.n_queries, sc.tile_query.p);
      MGX_LAUNCH_CHECK();
    }
    sc.map_owner = b.serial;
  }
  b.d_df_tile_desc.borrow(reinterpret_cast<DfTileDesc*>(sc.df_tile_desc.p), sc.df_tile_desc.n / 2);
  b.d_tile_query.borrow(sc.tile_query.p, sc.tile_query.n);
}
}  // namespace

namespace {
int device_sm_count(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device] != 0) {
    return cached[device];
  }
  int n = 148;
  MGX_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
  if (device >= 0 && device < 64) {
    cached[device] = n;
  }
  return n;
}

// record slots a tile needs: a top-k answer keeps at most limit + offset records per tile (and_tile_body prunes)
uint32_t rec_slot_for(const Batch& b) {
  const uint64_t k = static_cast<uint64_t>(b.params.limit) + b.params.offset;
  if (b.params.limit != 0 && k < kTile) {
    return static_cast<uint32_t>((k + 31) & ~31ULL);
  }
  return kTile;
}
uint32_t group_tiles_for(const Batch& b) {
  const uint64_t k = static_cast<uint64_t>(b.params.limit) + b.params.offset;
  if (b.params.compute_score != 0 && b.params.limit != 0 && k * 2 <= kGroupKeyCap) {
    return static_cast<uint32_t>(kGroupKeyCap / k);
  }
  return 0;
}
int df_force_mode() {  // MGX_DF_MODE=tiles|stream pins the choice (tests exercise both paths); default: cost model
  if (const char* mode = std::getenv("MGX_DF_MODE")) {
    return std::strcmp(mode, "tiles") == 0 ? 1 : (std::strcmp(mode, "stream") == 0 ? 2 : 0);
  }
  return 0;
}

// Workspace of a streamed batch: sized before anything is launched, because nothing comes back to the host.
void bind_streamed_scratch(Batch& b) {
  if (b.sc == nullptr) {
    b.sc = b.ix->scratch_for(b.stream);
  }
  SearchScratch& sc = *b.sc;
  const uint64_t floor_tiles = 1ULL << 17;
  sc.tile_count.reserve(floor_tiles);
  sc.tile_total.reserve(floor_tiles);
  sc.rec_doc.reserve(1ULL << 24);
  if (b.params.compute_score != 0) {
    sc.rec_score.reserve(sc.rec_doc.n);
  }
  b.d_tile_count.borrow(sc.tile_count.p, sc.tile_count.n);
  b.d_tile_total.borrow(sc.tile_total.p, sc.tile_total.n);
  b.d_rec_doc.borrow(sc.rec_doc.p, sc.rec_doc.n);
  b.d_rec_score.borrow(sc.rec_score.p, sc.rec_score.n);
}

void batch_plan_streamed(Batch& b) {
  cudaStream_t st = b.stream;
  Index& ix = *b.ix;
  b.streamed = true;
  b.rec_slot = rec_slot_for(b);
  bind_streamed_scratch(b);
  uint64_t cap_tiles = std::min<uint64_t>(b.d_tile_count.n, b.d_tile_total.n);
  cap_tiles = std::min<uint64_t>(cap_tiles, b.d_rec_doc.n / b.rec_slot);
  if (b.params.compute_score != 0) {
    cap_tiles = std::min<uint64_t>(cap_tiles, b.d_rec_score.n / b.rec_slot);
  }
  cap_tiles = std::min<uint64_t>(cap_tiles, 0x7FFFFFFFULL);
  if (const char* e = std::getenv("MGX_STREAM_TILE_CAP")) {  // tests force the overflow path with a tiny workspace
    cap_tiles = std::min<uint64_t>(cap_tiles, std::strtoull(e, nullptr, 10));
  }
  b.time_begin(0);
  const IndexView iv = make_view(ix);
  const BatchView bv = make_batch_view(b);
  if (b.n_terms > 0) {
    plan_terms_kernel<<<grid_for(b.n_terms, 128), 128, 0, st>>>(iv, bv, b.d_keys.p, b.params.compute_score != 0 ? 1 : 0,
                                                                ix.all_valid_utf8 ? 1 : 0, b.d_term_flags.p,
                                                                (ix.n_docs + 7) / 8, b.d_wide.p);
    MGX_LAUNCH_CHECK();
  }
  const int df_choice = b.params.compute_score != 0 && b.n_stream_terms > 0 ? 1 : 0;
  const uint64_t n_plan = std::max<uint64_t>(std::max<uint64_t>(b.n_queries, df_choice != 0 ? b.n_terms : 0), 1);
  plan_queries_kernel<<<grid_for(n_plan, 256), 256, 0, st>>>(iv, bv, b.d_term_flags.p, df_choice, df_force_mode(),
                                                             ix.n_text_tiles > 0 ? 1 : 0,
                                                             static_cast<uint32_t>(cap_tiles), group_tiles_for(b));
  MGX_LAUNCH_CHECK();
  b.time_end();
  b.planned = true;
}

void batch_df_streamed(Batch& b) {
  cudaStream_t st = b.stream;
  Index& ix = *b.ix;
  const int sms = device_sm_count(ix.device);
  if (b.n_stream_terms > 0 && ix.n_text_tiles > 0) {
    static int per_sm_stream = 0;
    if (per_sm_stream == 0) {
      MGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_stream, df_stream_kernel, kStreamThreads, 0));
      per_sm_stream = std::max(per_sm_stream, 1);
    }
    const unsigned grid =
        static_cast<unsigned>(std::min<uint64_t>(ix.n_text_tiles, static_cast<uint64_t>(sms) * per_sm_stream));
    b.time_begin(4);
    df_stream_kernel<<<grid, kStreamThreads, 0, st>>>(make_view(ix), make_batch_view(b));
    MGX_LAUNCH_CHECK();
    b.time_end();
  }
  if (b.n_terms > 0) {
    static int per_sm_df = 0;
    if (per_sm_df == 0) {
      MGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_df, df_units_kernel, kTileThreads, 0));
      per_sm_df = std::max(per_sm_df, 1);
    }
    b.time_begin(1);
    df_units_kernel<<<static_cast<unsigned>(sms * per_sm_df), kTileThreads, 0, st>>>(make_view(ix), make_batch_view(b));
    MGX_LAUNCH_CHECK();
    b.time_end();
  }
  b.df_done = true;
}
}  // namespace

bool batch_overflowed(Batch& b) {
  return b.streamed && b.status_copied && b.h_launch.p != nullptr && b.h_launch.p[kLaunchOverflow] != 0;
}

// Back to the freshly uploaded state: the planning kernels permute key tables and term lists in place, so the compiled
// batch is copied from the pinned staging buffer again. The repeat takes the synchronous form, whose workspace grows
// with the batch, and tells the stream's workspace to stay in that form for a while when even that is chunked.
void batch_reset_for_repeat(Batch& b, bool after_overflow) {
  MGX_CUDA(cudaMemcpyAsync(b.in_base, b.staging.p, b.in_total, cudaMemcpyHostToDevice, b.stream));
  b.h2d_bytes += b.in_total;
  batch_clear_counters(b);
  if (after_overflow) {
    if (b.sc != nullptr) {
      b.sc->skip_streamed = 8;
    }
    b.allow_streamed = false;
  }
  b.streamed = false;
  b.planned = b.df_done = b.searched = b.status_copied = false;
  b.rec_slot = kTile;
  b.n_df_tiles = b.n_and_tiles = b.driver_entries = 0;
}

void batch_plan(Batch& b) {
  cudaStream_t st = b.stream;
  Index& ix = *b.ix;
  if (b.allow_streamed && b.explicit_driver.d_ids == nullptr) {
    const bool disabled = std::getenv("MGX_NO_STREAMED") != nullptr;  // read per batch: the tests flip it
    if (b.sc == nullptr) {
      b.sc = ix.scratch_for(st);
    }
    if (!disabled && b.sc->skip_streamed == 0) {
      batch_plan_streamed(b);
      return;
    }
    if (b.sc->skip_streamed > 0) {
      --b.sc->skip_streamed;
    }
  }
  b.streamed = false;
  b.rec_slot = kTile;
  b.time_begin(0);
  if (b.n_keys > 0) {
    lookup_kernel<<<grid_for(b.n_keys, 256), 256, 0, st>>>(ix.d_term_keys.p, ix.d_term_off.p, ix.n_terms, b.d_keys.p,
                                                          b.n_keys, b.d_key_list.p, b.d_key_len.p, b.d_key_glen.p,
                                                          ix.wide_words > 0 ? ix.d_wide_keys.p : nullptr, ix.wide_words,
                                                          b.d_wide.p);
    MGX_LAUNCH_CHECK();
  }
  BatchView bv = make_batch_view(b);
  if (b.n_terms > 0) {
    term_plan_kernel<<<grid_for(b.n_terms, 128), 128, 0, st>>>(bv, b.params.compute_score != 0 ? 1 : 0,
                                                               ix.all_valid_utf8 ? 1 : 0, b.d_term_flags.p,
                                                               (ix.n_docs + 7) / 8);
    MGX_LAUNCH_CHECK();
    if (b.n_keys > 0) {
      key_ref_kernel<<<grid_for(b.n_keys, 256), 256, 0, st>>>(make_view(ix), b.d_key_list.p, b.d_key_len.p, b.n_keys,
                                                              b.d_key_ref.p);
      MGX_LAUNCH_CHECK();
    }
    if (b.params.compute_score != 0 && b.n_stream_terms > 0) {
      df_mode_kernel<<<grid_for(b.n_terms, 128), 128, 0, st>>>(bv, b.d_term_flags.p, ix.text_bytes, df_force_mode(),
                                                               ix.n_text_tiles > 0 ? 1 : 0);
      MGX_LAUNCH_CHECK();
    }
  }
  const IndexView iv = make_view(ix);
  if (b.n_queries > 0) {
    query_plan_kernel<<<grid_for(b.n_queries, 128), 128, 0, st>>>(iv, bv);
    MGX_LAUNCH_CHECK();
  }
  if (b.n_terms <= kSmallScanMax && b.n_queries <= kSmallScanMax) {
    // tile offsets of the df stage and of the search stage: one launch
    SmallScanJobs jobs{{b.d_t_df_tiles.p, b.d_q_ntiles.p, nullptr},
                       {b.d_t_df_tile_off.p, b.d_q_tile_off.p, nullptr},
                       {b.n_terms, b.n_queries, 0}};
    exclusive_scans_small(jobs, 2, st);
  } else {
    exclusive_scan_u32_u64(b.d_t_df_tiles.p, b.d_t_df_tile_off.p, b.n_terms, b.d_scan_scratch.p, st);
    exclusive_scan_u32_u64(b.d_q_ntiles.p, b.d_q_tile_off.p, b.n_queries, b.d_scan_scratch.p, st);
  }
  b.h_q_tile_off.resize(b.n_queries + 1);
  MGX_CUDA(cudaMemcpyAsync(b.h_q_tile_off.data(), b.d_q_tile_off.p, (b.n_queries + 1) * sizeof(uint64_t),
                           cudaMemcpyDeviceToHost, st));
  MGX_CUDA(cudaMemcpyAsync(&b.n_df_tiles, b.d_t_df_tile_off.p + b.n_terms, sizeof(uint64_t), cudaMemcpyDeviceToHost,
                           st));
  uint32_t df_mode = 0;
  MGX_CUDA(cudaMemcpyAsync(&df_mode, b.d_df_mode.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  b.time_end();
  MGX_CUDA(cudaStreamSynchronize(st));
  b.h_df_mode = static_cast<int>(df_mode);
  b.d2h_bytes += (b.n_queries + 1) * sizeof(uint64_t) + sizeof(uint64_t);
  b.planned = true;
  b.time_begin(0);
  ensure_tile_maps(b);
  b.time_end();
}

void batch_df(Batch& b) {
  cudaStream_t st = b.stream;
  Index& ix = *b.ix;
  if (!b.planned) {
    batch_plan(b);
  }
  if (b.streamed) {
    batch_df_streamed(b);
    return;
  }
  ensure_tile_maps(b);
  const uint64_t total_tiles = b.n_df_tiles;
  if (b.h_df_mode != 0 && ix.n_text_tiles > 0) {
    // one pass over the text arena for every stream-eligible term; persistent CTAs stride over the tiles
    int sm_count = 148;
    MGX_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ix.device));
    int per_sm = 4;
    MGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, df_stream_kernel, kStreamThreads, 0));
    const unsigned grid = static_cast<unsigned>(
        std::min<uint64_t>(ix.n_text_tiles, static_cast<uint64_t>(sm_count) * std::max(per_sm, 1)));
    b.time_begin(4);
    df_stream_kernel<<<grid, kStreamThreads, 0, st>>>(make_view(ix), make_batch_view(b));
    MGX_LAUNCH_CHECK();
    b.time_end();
  }
  if (total_tiles > 0) {
    if (total_tiles > 0x7FFFFFFFULL) {
      set_last_error("df stage: too many tiles in one batch");
      throw CudaFailure{MGX_ERR_UNSUPPORTED};
    }
    b.time_begin(1);
    df_tile_kernel<<<static_cast<unsigned>(total_tiles), kTileThreads, 0, st>>>(make_view(ix), make_batch_view(b));
    MGX_LAUNCH_CHECK();
    b.time_end();
  }
  b.df_done = true;
}

void batch_df_to_slots(Batch& b, uint64_t* d_df_slots) {
  if (b.n_slots > 0) {
    df_to_slots_kernel<<<grid_for(b.n_slots, 256), 256, 0, b.stream>>>(b.d_t_df.p, b.d_slot_tid.p, b.n_slots,
                                                                       d_df_slots);
    MGX_LAUNCH_CHECK();
  }
}

namespace {

struct Chunk {
  uint32_t q0;
  uint32_t q1;
};

// Split the batch so that each chunk's record area fits the scratch budget.
std::vector<Chunk> make_chunks(const Batch& b, uint64_t max_records, uint32_t rec_slot) {
  std::vector<Chunk> chunks;
  uint32_t q0 = 0;
  while (q0 < b.n_queries) {
    uint32_t q1 = q0 + 1;
    while (q1 < b.n_queries && (b.h_q_tile_off[q1 + 1] - b.h_q_tile_off[q0]) * rec_slot <= max_records) {
      ++q1;
    }
    chunks.push_back({q0, q1});
    q0 = q1;
  }
  return chunks;
}

void prepare_scoring(Batch& b, const uint64_t* d_df_slots) {
  cudaStream_t st = b.stream;
  if (b.params.compute_score == 0) {
    return;
  }
  if (b.global_order) {
    // sharded pipeline: t_df and key_glen hold the sums over all shards
    const uint64_t docs = b.params.total_docs != 0 ? b.params.total_docs : b.ix->doc_count;
    if (b.n_queries > 0) {
      global_order_kernel<<<grid_for(b.n_queries, 128), 128, 0, st>>>(make_batch_view(b), docs, b.d_q_idf.p);
      MGX_LAUNCH_CHECK();
    }
    return;
  }
  if (d_df_slots != nullptr && b.n_slots > 0) {
    slots_to_df_kernel<<<grid_for(b.n_slots, 256), 256, 0, st>>>(d_df_slots, b.d_slot_tid.p, b.n_slots, b.d_t_df.p);
    MGX_LAUNCH_CHECK();
  }
  const uint64_t total_docs = b.params.total_docs != 0 ? b.params.total_docs : b.ix->doc_count;
  const uint32_t n = b.n_qterms;
  if (n > 0) {
    idf_kernel<<<grid_for(n, 256), 256, 0, st>>>(b.d_q_tids.p, n, b.d_t_df.p, total_docs, b.d_q_idf.p);
    MGX_LAUNCH_CHECK();
  }
}

ScoreParams score_params(const Batch& b) {
  ScoreParams sp;
  sp.k1 = b.params.k1;
  sp.b = b.params.b;
  const uint64_t total_docs = b.params.total_docs != 0 ? b.params.total_docs : b.ix->doc_count;
  const uint64_t total_len = b.params.total_docs != 0 ? b.params.total_doc_length : b.ix->total_doc_length;
  const double avgdl =
      total_docs > 0 ? static_cast<double>(total_len) / static_cast<double>(total_docs) : 0.0;  // server_types.h:182-187
  sp.avgdl_clamped = std::max(avgdl, 1.0);
  sp.compute_score = b.params.compute_score;
  sp.descending = b.params.descending;
  return sp;
}

uint64_t scratch_records(const Batch& b) {
  const uint64_t bytes = b.ix->cfg.scratch_bytes != 0 ? b.ix->cfg.scratch_bytes : (4ULL << 30);
  return std::max<uint64_t>(bytes / 12, 1ULL << 20);
}

// record slots per tile for a given pruning depth (0 = no pruning: a tile may keep all its kTile entries)
uint32_t rec_slot_for_prune(uint32_t prune_k) {
  return prune_k != 0 && prune_k < kTile ? ((prune_k + 31u) & ~31u) : static_cast<uint32_t>(kTile);
}

void run_tiles(Batch& b, const Chunk& c, const ScoreParams& sp, uint32_t prune_k) {
  cudaStream_t st = b.stream;
  const uint64_t tile_base = b.h_q_tile_off[c.q0];
  const uint64_t n_tiles = b.h_q_tile_off[c.q1] - tile_base;
  b.rec_slot = rec_slot_for_prune(prune_k);
  const uint64_t n_recs = n_tiles * b.rec_slot;
  ensure_tile_maps(b);
  SearchScratch& sc = *b.sc;
  sc.tile_count.reserve(std::max<uint64_t>(n_tiles, 1ULL << 16));
  sc.tile_total.reserve(std::max<uint64_t>(n_tiles, 1ULL << 16));
  sc.rec_doc.reserve(std::max<uint64_t>(n_recs + kTile, 1ULL << 24));
  if (sp.compute_score != 0) {
    sc.rec_score.reserve(std::max<uint64_t>(n_recs + kTile, 1ULL << 24));
  }
  b.d_tile_count.borrow(sc.tile_count.p, sc.tile_count.n);
  b.d_tile_total.borrow(sc.tile_total.p, sc.tile_total.n);
  b.d_rec_doc.borrow(sc.rec_doc.p, sc.rec_doc.n);
  b.d_rec_score.borrow(sc.rec_score.p, sc.rec_score.n);
  if (n_tiles > 0) {
    if (n_tiles > 0x7FFFFFFFULL) {
      set_last_error("search stage: too many tiles in one chunk");
      throw CudaFailure{MGX_ERR_UNSUPPORTED};
    }
    b.time_begin(2);
    and_tile_kernel<<<static_cast<unsigned>(n_tiles), kTileThreads, 0, st>>>(
        make_view(*b.ix), make_batch_view(b), sp, tile_base, b.rec_slot, b.d_tile_count.p, b.d_tile_total.p,
        b.d_rec_doc.p, b.d_rec_score.p, prune_k,
        interleave_stride(static_cast<uint32_t>(n_tiles), prune_k, sp.compute_score));
    MGX_LAUNCH_CHECK();
    b.time_end();
    b.n_and_tiles += n_tiles;
  }
}

}  // namespace

__global__ void fold_expanded_kernel(const uint32_t* __restrict__ xoff, uint32_t n_out, uint64_t x_stride,
                                     const uint32_t* __restrict__ x_ids, const uint32_t* __restrict__ x_count,
                                     const uint64_t* __restrict__ x_total, uint32_t limit, uint32_t offset,
                                     uint64_t stride, uint32_t* __restrict__ out_ids, uint32_t* __restrict__ out_count,
                                     uint64_t* __restrict__ out_total);

void batch_search(Batch& b, const uint64_t* d_df_slots, uint64_t stride, uint32_t* d_ids, double* d_scores,
                  uint32_t* d_count, uint64_t* d_total) {
  cudaStream_t st = b.stream;
  if (!b.planned) {
    batch_plan(b);
  }
  prepare_scoring(b, d_df_slots);
  const ScoreParams sp = score_params(b);
  // Expanded OR-rooted programs: the internal queries answer into the batch's own buffers with room for
  // offset + limit ids each, and fold_expanded_kernel (at the end) writes the caller's rows.
  const bool expanded = !b.h_xoff.empty();
  struct Fold {
    Batch& b;
    bool on;
    mgx_query_params_t user_params;
    uint64_t stride;
    uint32_t* d_ids;
    uint32_t* d_count;
    uint64_t* d_total;
    uint64_t x_stride;
    void run() {
      if (!on) {
        return;
      }
      fold_expanded_kernel<<<grid_for(static_cast<uint64_t>(b.n_out_queries) * 32, 256), 256, 0, b.stream>>>(
          b.d_xoff.p, b.n_out_queries, x_stride, b.x_ids.p, b.x_count.p, b.x_total.p, user_params.limit,
          user_params.offset, stride, d_ids, d_count, d_total);
      MGX_LAUNCH_CHECK();
      b.params = user_params;
    }
  } fold{b, expanded, b.params, stride, d_ids, d_count, d_total, 0};
  if (expanded) {
    if (sp.compute_score != 0) {
      set_last_error("internal: expanded programs are unscored");
      throw CudaFailure{MGX_ERR_INVALID_ARGUMENT};
    }
    fold.x_stride = (b.params.limit != 0 ? std::min<uint64_t>(b.params.limit, stride) : stride) + b.params.offset;
    b.x_ids.reserve(static_cast<uint64_t>(b.n_queries) * fold.x_stride);
    b.x_count.reserve(b.n_queries);
    b.x_total.reserve(b.n_queries);
    b.params.limit = static_cast<uint32_t>(std::min<uint64_t>(fold.x_stride, 0xFFFFFFFFULL));
    b.params.offset = 0;
    stride = fold.x_stride;
    d_ids = b.x_ids.p;
    d_count = b.x_count.p;
    d_total = b.x_total.p;
  }
  // per-tile top-k pruning is only valid when the answer is a top-k by score
  const uint32_t prune_k =
      b.params.limit != 0 && static_cast<uint64_t>(b.params.limit) + b.params.offset <= kTile
          ? b.params.limit + b.params.offset : 0u;
  if (b.streamed) {
    // one uninterrupted enqueue: persistent kernels size themselves from the device-side launch block
    if (rec_slot_for_prune(prune_k) > b.rec_slot) {
      set_last_error("internal: streamed batch searched with other limits than it was planned with");
      throw CudaFailure{MGX_ERR_INVALID_ARGUMENT};
    }
    const int sms = device_sm_count(b.ix->device);
    static int per_sm_and = 0;
    if (per_sm_and == 0) {
      MGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_and, and_tiles_kernel, kTileThreads, 0));
      per_sm_and = std::max(per_sm_and, 1);
    }
    const BatchView bv = make_batch_view(b);
    b.time_begin(2);
    and_tiles_kernel<<<static_cast<unsigned>(sms * per_sm_and), kTileThreads, 0, st>>>(
        make_view(*b.ix), bv, sp, b.rec_slot, b.d_tile_count.p, b.d_tile_total.p, b.d_rec_doc.p, b.d_rec_score.p, prune_k);
    MGX_LAUNCH_CHECK();
    b.time_end();
    b.time_begin(3);
    const uint32_t group_tiles = group_tiles_for(b);
    if (group_tiles != 0 && prune_k != 0) {
      topk_groups_kernel<<<static_cast<unsigned>(sms * 2), 256, 0, st>>>(bv, group_tiles, b.rec_slot, b.d_tile_count.p,
                                                                        b.d_rec_doc.p, b.d_rec_score.p,
                                                                        b.params.descending, prune_k);
      MGX_LAUNCH_CHECK();
    }
    if (b.n_queries > 0) {
      topk_kernel<<<b.n_queries, 256, 0, st>>>(bv, 0, 0, b.rec_slot, b.d_tile_count.p, b.d_tile_total.p, b.d_rec_doc.p,
                                               b.d_rec_score.p, b.params.compute_score, b.params.descending,
                                               b.params.limit, b.params.offset, stride, d_ids, d_scores, d_count, d_total);
      MGX_LAUNCH_CHECK();
    }
    b.time_end();
    fold.run();
    // sizes, counters and the overflow flag follow the results to the host
    b.h_launch.reserve(kLaunchCount);
    MGX_CUDA(cudaMemcpyAsync(b.h_launch.p, b.d_launch.p, kLaunchCount * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    b.status_copied = true;
    b.mark_last();
    b.searched = true;
    return;
  }
  const auto chunks = make_chunks(b, scratch_records(b), rec_slot_for_prune(prune_k));
  for (const Chunk& c : chunks) {
    if (chunks.size() > 1) {
      MGX_CUDA(cudaStreamSynchronize(st));  // scratch is reused by the next chunk
      if (b.sc != nullptr) {
        b.sc->skip_streamed = 8;  // batches of this size do not fit a streamed workspace either
      }
    }
    run_tiles(b, c, sp, prune_k);
    b.time_begin(3);
    if (sp.compute_score != 0 && prune_k != 0 && prune_k * 2 <= kGroupKeyCap) {
      // queries with many tiles: reduce groups of tiles to prune_k records each before the per-query top-k
      const uint32_t group_tiles = kGroupKeyCap / prune_k;
      std::vector<TopkGroup> groups;
      for (uint32_t q = c.q0; q < c.q1; ++q) {
        const uint32_t ntiles = static_cast<uint32_t>(b.h_q_tile_off[q + 1] - b.h_q_tile_off[q]);
        if (ntiles >= 2 * group_tiles) {
          for (uint32_t tb = 0; tb < ntiles; tb += group_tiles) {
            groups.push_back({q, tb, std::min(ntiles, tb + group_tiles)});
          }
        }
      }
      if (!groups.empty()) {
        SearchScratch& sc = *b.sc;
        sc.topk_groups.reserve(std::max<uint64_t>(groups.size() * 3, 1ULL << 14));
        MGX_CUDA(cudaMemcpyAsync(sc.topk_groups.p, groups.data(), groups.size() * sizeof(TopkGroup),
                                 cudaMemcpyHostToDevice, st));
        b.h2d_bytes += groups.size() * sizeof(TopkGroup);
        topk_group_kernel<<<static_cast<unsigned>(groups.size()), 256, 0, st>>>(
            make_batch_view(b), reinterpret_cast<const TopkGroup*>(sc.topk_groups.p), b.h_q_tile_off[c.q0],
            b.rec_slot, b.d_tile_count.p, b.d_rec_doc.p, b.d_rec_score.p, b.params.descending, prune_k);
        MGX_LAUNCH_CHECK();
      }
    }
    topk_kernel<<<c.q1 - c.q0, 256, 0, st>>>(make_batch_view(b), c.q0, b.h_q_tile_off[c.q0], b.rec_slot,
                                             b.d_tile_count.p, b.d_tile_total.p, b.d_rec_doc.p, b.d_rec_score.p,
                                             b.params.compute_score, b.params.descending, b.params.limit,
                                             b.params.offset, stride, d_ids, d_scores, d_count, d_total);
    MGX_LAUNCH_CHECK();
    b.time_end();
  }
  fold.run();
  b.mark_last();
  b.searched = true;
}

// Caller query q of a batch with expanded OR-rooted programs = internal queries [xoff[q], xoff[q+1]), whose answers
// are ascending and pairwise disjoint: total = sum of totals, ids = the merged runs cut to [offset, offset + limit).
// One warp per caller query; the merged rank of an id = its index in its own run + its rank in every other run.
__global__ void fold_expanded_kernel(const uint32_t* __restrict__ xoff, uint32_t n_out, uint64_t x_stride,
                                     const uint32_t* __restrict__ x_ids, const uint32_t* __restrict__ x_count,
                                     const uint64_t* __restrict__ x_total, uint32_t limit, uint32_t offset,
                                     uint64_t stride, uint32_t* __restrict__ out_ids, uint32_t* __restrict__ out_count,
                                     uint64_t* __restrict__ out_total) {
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (q >= n_out) {
    return;
  }
  const uint32_t s0 = xoff[q];
  const uint32_t s1 = xoff[q + 1];
  uint64_t total = 0;
  for (uint32_t s = s0; s < s1; ++s) {
    total += x_total[s];
  }
  const uint64_t want_end = limit == 0 ? total : umin64(total, static_cast<uint64_t>(offset) + limit);
  const uint64_t want_begin = umin64(offset, total);
  const uint64_t n_out_ids = umin64(want_end - want_begin, stride);
  if (lane == 0) {
    out_total[q] = total;
    out_count[q] = static_cast<uint32_t>(n_out_ids);
  }
  for (uint32_t s = s0; s < s1; ++s) {
    const uint32_t* run = x_ids + static_cast<uint64_t>(s) * x_stride;
    const uint32_t c = x_count[s];
    for (uint32_t i = lane; i < c; i += 32) {
      const uint32_t v = run[i];
      uint64_t rank = i;
      for (uint32_t o = s0; o < s1; ++o) {
        if (o != s) {
          rank += lower_bound_u32(x_ids + static_cast<uint64_t>(o) * x_stride, x_count[o], v);
        }
      }
      if (rank >= want_begin && rank - want_begin < n_out_ids) {
        out_ids[static_cast<uint64_t>(q) * stride + (rank - want_begin)] = v;
      }
    }
  }
}

// Union of ascending, pairwise DISJOINT runs laid back to back (the per-driver result sets of one expanded query):
// no sort is needed, the final position of an element is its index in its own run plus the number of smaller
// elements in every other run (a binary search each).
__global__ void merge_disjoint_runs_kernel(const uint32_t* __restrict__ in, const uint64_t* __restrict__ run_off,
                                           uint32_t n_runs, uint32_t* __restrict__ out) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= run_off[n_runs]) {
    return;
  }
  const uint32_t v = in[i];
  uint64_t pos = 0;
  for (uint32_t r = 0; r < n_runs; ++r) {
    const uint64_t r0 = run_off[r];
    const uint64_t r1 = run_off[r + 1];
    if (i >= r0 && i < r1) {
      pos += i - r0;
      continue;
    }
    uint64_t lo = r0;
    uint64_t hi = r1;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (in[mid] < v) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    pos += lo - r0;
  }
  out[pos] = v;
}

// Posting-list lengths of up to 64 packed keys (0 for a key the dictionary does not hold): one lookup launch and one
// small copy. Used by the single calls to choose between a list-driven plan and a pass over the shard.
void lookup_list_lengths(Batch& b, const uint64_t* h_keys, uint32_t n, uint32_t* h_lens) {
  Index& ix = *b.ix;
  cudaStream_t st = b.stream;
  if (n == 0) {
    return;
  }
  const int W = ix.wide_words;
  b.len_buf.reserve(3 * 64 * 8 + 64 * kMaxWideWords);  // keys | list ids | lengths | wide words
  uint64_t* d_keys = b.len_buf.p;
  uint32_t* d_list = reinterpret_cast<uint32_t*>(b.len_buf.p + 64);
  uint32_t* d_len = reinterpret_cast<uint32_t*>(b.len_buf.p + 128);
  uint64_t* d_wide = b.len_buf.p + 192;
  MGX_CUDA(cudaMemcpyAsync(d_keys, h_keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  uint64_t h_wide[64 * kMaxWideWords];
  if (W > 0) {
    for (uint32_t i = 0; i < n; ++i) {
      if (h_keys[i] != kInvalidKey) {
        std::memcpy(h_wide + static_cast<size_t>(i) * W, host_wide_words(h_keys[i], W), static_cast<size_t>(W) * 8);
      } else {
        std::memset(h_wide + static_cast<size_t>(i) * W, 0, static_cast<size_t>(W) * 8);
      }
    }
    MGX_CUDA(cudaMemcpyAsync(d_wide, h_wide, static_cast<size_t>(n) * W * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  }
  lookup_kernel<<<1, 64, 0, st>>>(ix.d_term_keys.p, ix.d_term_off.p, ix.n_terms, d_keys, n, d_list, d_len, nullptr,
                                  W > 0 ? ix.d_wide_keys.p : nullptr, W, d_wide);
  MGX_LAUNCH_CHECK();
  MGX_CUDA(cudaMemcpyAsync(h_lens, d_len, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  MGX_CUDA(cudaStreamSynchronize(st));  // also keeps h_wide alive until its copy is done
}

namespace {
// run_group[2 r], run_group[2 r + 1]: first and end run of the group run r belongs to. One thread per entry: its run
// by a bound search over the offsets, its place = start of the group + its rank in every run of the group.
__global__ void merge_grouped_runs_kernel(const uint32_t* __restrict__ in, const uint64_t* __restrict__ run_off,
                                          uint32_t n_runs, const uint32_t* __restrict__ run_group,
                                          uint32_t* __restrict__ out) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= run_off[n_runs]) {
    return;
  }
  uint32_t lo_r = 0;  // largest r with run_off[r] <= i (empty runs repeat an offset: the LAST such r holds i)
  uint32_t hi_r = n_runs;
  while (hi_r - lo_r > 1) {
    const uint32_t mid = (lo_r + hi_r) >> 1;
    if (run_off[mid] <= i) {
      lo_r = mid;
    } else {
      hi_r = mid;
    }
  }
  const uint32_t mine = lo_r;
  const uint32_t g0 = run_group[2 * mine];
  const uint32_t g1 = run_group[2 * mine + 1];
  const uint32_t v = in[i];
  uint64_t pos = run_off[g0];
  for (uint32_t r = g0; r < g1; ++r) {
    const uint64_t r0 = run_off[r];
    const uint64_t r1 = run_off[r + 1];
    if (r == mine) {
      pos += i - r0;
      continue;
    }
    uint64_t lo = r0;
    uint64_t hi = r1;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (in[mid] < v) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    pos += lo - r0;
  }
  out[pos] = v;
}
}  // namespace

void merge_grouped_runs(Batch& b, const uint32_t* d_in, const std::vector<uint64_t>& run_off,
                        const std::vector<uint32_t>& group_begin, DevBuf<uint32_t>* d_out) {
  cudaStream_t st = b.stream;
  const uint64_t total = run_off.back();
  const size_t n_runs = run_off.size() - 1;
  b.union_buf.reserve(std::max<uint64_t>(1, total));
  d_out->borrow(b.union_buf.p, total);
  if (total == 0) {
    return;
  }
  // offsets and the (first, end) run of every run's group, back to back in the batch object's grow-only buffer
  std::vector<uint64_t> host(run_off);
  std::vector<uint32_t> rg(2 * n_runs, 0);
  for (size_t g = 0; g + 1 < group_begin.size(); ++g) {
    for (uint32_t r = group_begin[g]; r < group_begin[g + 1]; ++r) {
      rg[2 * r] = group_begin[g];
      rg[2 * r + 1] = group_begin[g + 1];
    }
  }
  const size_t words = host.size() + (rg.size() + 1) / 2;
  host.resize(words, 0);
  std::memcpy(host.data() + run_off.size(), rg.data(), rg.size() * sizeof(uint32_t));
  b.run_off_buf.reserve(words);
  MGX_CUDA(cudaMemcpyAsync(b.run_off_buf.p, host.data(), words * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  MGX_CUDA(cudaStreamSynchronize(st));  // `host` is pageable and goes out of scope
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  merge_grouped_runs_kernel<<<blocks, 256, 0, st>>>(
      d_in, b.run_off_buf.p, static_cast<uint32_t>(n_runs),
      reinterpret_cast<const uint32_t*>(b.run_off_buf.p + run_off.size()), d_out->p);
  MGX_LAUNCH_CHECK();
  MGX_CUDA(cudaStreamSynchronize(st));
}

void merge_disjoint_runs(Batch& b, const uint32_t* d_in, const std::vector<uint64_t>& run_off, DevBuf<uint32_t>* d_out) {
  cudaStream_t st = b.stream;
  const uint64_t total = run_off.back();
  // the union and the run offsets live in the batch object's grow-only buffers: no cudaMalloc / cudaFree per call
  b.union_buf.reserve(std::max<uint64_t>(1, total));
  d_out->borrow(b.union_buf.p, total);
  if (total == 0) {
    return;
  }
  b.run_off_buf.reserve(run_off.size());
  MGX_CUDA(cudaMemcpyAsync(b.run_off_buf.p, run_off.data(), run_off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  merge_disjoint_runs_kernel<<<blocks, 256, 0, st>>>(d_in, b.run_off_buf.p, static_cast<uint32_t>(run_off.size() - 1),
                                                     d_out->p);
  MGX_LAUNCH_CHECK();
  MGX_CUDA(cudaStreamSynchronize(st));
}

void batch_search_sets(Batch& b, std::vector<uint64_t>* h_set_off, DevBuf<uint32_t>* d_sets) {
  cudaStream_t st = b.stream;
  if (!b.planned) {
    batch_plan(b);
  }
  ScoreParams sp = score_params(b);
  sp.compute_score = 0;
  const auto chunks = make_chunks(b, scratch_records(b), kTile);
  // small per-call arrays from the batch object's arena (grow-only): a single Index::Search* call allocates nothing
  const size_t Q = b.n_queries;
  b.sets_arena.reserve(DevArena::padded(Q * 4 + 4) + 2 * DevArena::padded((Q + 1) * 8) + DevArena::padded(4) + 256);
  uint32_t* d_count = b.sets_arena.take<uint32_t>(Q);
  uint64_t* d_total = b.sets_arena.take<uint64_t>(Q + 1);
  uint64_t* d_set_off = b.sets_arena.take<uint64_t>(Q + 1);
  uint32_t* d_dummy = b.sets_arena.take<uint32_t>(1);
  std::vector<uint64_t> totals(Q, 0);
  auto count_chunk = [&](const Chunk& c) {  // totals per query (topk_kernel with limit 0 / stride 0 only counts)
    topk_kernel<<<c.q1 - c.q0, 256, 0, st>>>(make_batch_view(b), c.q0, b.h_q_tile_off[c.q0], b.rec_slot,
                                             b.d_tile_count.p, b.d_tile_total.p, b.d_rec_doc.p, nullptr, 0, 0, 0, 0, 0,
                                             d_dummy, nullptr, d_count, d_total);
    MGX_LAUNCH_CHECK();
  };
  auto offsets_from_totals = [&]() {
    MGX_CUDA(cudaMemcpyAsync(totals.data(), d_total, Q * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    MGX_CUDA(cudaStreamSynchronize(st));
    h_set_off->assign(Q + 1, 0);
    for (uint32_t q = 0; q < Q; ++q) {
      (*h_set_off)[q + 1] = (*h_set_off)[q] + totals[q];
    }
    MGX_CUDA(cudaMemcpyAsync(d_set_off, h_set_off->data(), (Q + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    b.sets_buf.reserve(std::max<uint64_t>(1, h_set_off->back()));
    d_sets->borrow(b.sets_buf.p, h_set_off->back());
  };
  auto gather_chunk = [&](const Chunk& c) {  // the tile totals are no longer needed: their array takes the ranks
    const uint64_t tile_base = b.h_q_tile_off[c.q0];
    const uint64_t n_tiles = b.h_q_tile_off[c.q1] - tile_base;
    set_tile_ranks_kernel<<<c.q1 - c.q0, 256, 0, st>>>(make_batch_view(b), c.q0, tile_base, b.d_tile_count.p,
                                                      b.d_tile_total.p);
    MGX_LAUNCH_CHECK();
    if (n_tiles > 0) {
      gather_tiles_kernel<<<static_cast<unsigned>(n_tiles), 256, 0, st>>>(make_batch_view(b), tile_base, b.rec_slot,
                                                                        b.d_tile_count.p, b.d_tile_total.p,
                                                                        b.d_rec_doc.p, d_set_off, d_sets->p);
      MGX_LAUNCH_CHECK();
    }
  };
  if (chunks.size() == 1) {
    // single chunk: records stay valid between the counting and the gathering pass
    const Chunk& c = chunks[0];
    run_tiles(b, c, sp, 0);
    count_chunk(c);
    offsets_from_totals();
    gather_chunk(c);
    MGX_CUDA(cudaStreamSynchronize(st));
    return;
  }
  // several chunks: count everything first, then redo the tiles chunk by chunk and gather
  for (const Chunk& c : chunks) {
    MGX_CUDA(cudaStreamSynchronize(st));
    run_tiles(b, c, sp, 0);
    count_chunk(c);
  }
  offsets_from_totals();
  for (const Chunk& c : chunks) {
    MGX_CUDA(cudaStreamSynchronize(st));
    run_tiles(b, c, sp, 0);
    gather_chunk(c);
  }
  MGX_CUDA(cudaStreamSynchronize(st));
}

void launch_merge_topk(cudaStream_t stream, const mgx_query_params_t& params, uint32_t n_shards, uint64_t n_queries,
                       uint64_t stride, const uint32_t* d_ids_all, const double* d_scores_all,
                       const uint32_t* d_count_all, const uint64_t* d_total_all, uint64_t shard_pitch_bytes,
                       uint32_t* d_ids_out, double* d_scores_out, uint32_t* d_count_out, uint64_t* d_total_out) {
  if (n_queries == 0) {
    return;
  }
  ShardRuns runs;
  runs.ids_base = reinterpret_cast<const uint8_t*>(d_ids_all);
  runs.scores_base = reinterpret_cast<const uint8_t*>(d_scores_all);
  runs.count_base = reinterpret_cast<const uint8_t*>(d_count_all);
  runs.total_base = reinterpret_cast<const uint8_t*>(d_total_all);
  // pitch 0: four dense [n_shards][n_queries][stride] arrays; otherwise every array of shard s sits s * pitch further
  runs.ids_pitch = shard_pitch_bytes != 0 ? shard_pitch_bytes : n_queries * stride * sizeof(uint32_t);
  runs.scores_pitch = shard_pitch_bytes != 0 ? shard_pitch_bytes : n_queries * stride * sizeof(double);
  runs.count_pitch = shard_pitch_bytes != 0 ? shard_pitch_bytes : n_queries * sizeof(uint32_t);
  runs.total_pitch = shard_pitch_bytes != 0 ? shard_pitch_bytes : n_queries * sizeof(uint64_t);
  merge_topk_kernel<<<static_cast<unsigned>(n_queries), 256, 0, stream>>>(
      n_shards, n_queries, stride, params.compute_score, params.descending, params.limit, params.offset, runs,
      d_ids_out, d_scores_out, d_count_out, d_total_out);
  MGX_LAUNCH_CHECK();
}

namespace {
__global__ void or_status_kernel(const uint8_t* __restrict__ in, uint64_t pitch, uint32_t n_shards, uint8_t* __restrict__ out) {
  const uint32_t w = threadIdx.x;  // four status words
  if (w < 4) {
    uint32_t v = 0;
    for (uint32_t s = 0; s < n_shards; ++s) {
      v |= reinterpret_cast<const uint32_t*>(in + static_cast<uint64_t>(s) * pitch)[w];
    }
    reinterpret_cast<uint32_t*>(out)[w] = v;
  }
}
}  // namespace

void launch_or_status(cudaStream_t stream, const uint8_t* d_status_first, uint64_t shard_pitch_bytes, uint32_t n_shards,
                      uint8_t* d_status_out) {
  or_status_kernel<<<1, 32, 0, stream>>>(d_status_first, shard_pitch_bytes, n_shards, d_status_out);
  MGX_LAUNCH_CHECK();
}

void launch_score_documents(Index& ix, cudaStream_t stream, const uint32_t* d_cands, uint64_t n_cands,
                            const uint8_t* d_term_bytes, const uint32_t* d_term_boff, const uint64_t* d_dfs,
                            uint32_t n_terms, uint64_t total_docs, double avgdl, double k1, double b, double* d_scores) {
  DevBuf<double> d_idf;
  d_idf.alloc(n_terms);
  if (n_terms > 0) {
    idf_plain_kernel<<<grid_for(n_terms, 128), 128, 0, stream>>>(d_dfs, n_terms, total_docs, d_idf.p);
    MGX_LAUNCH_CHECK();
  }
  ScoreParams sp;
  sp.k1 = k1;
  sp.b = b;
  sp.avgdl_clamped = std::max(avgdl, 1.0);
  sp.compute_score = 1;
  sp.descending = 1;
  if (n_cands > 0) {
    score_docs_kernel<<<grid_for(n_cands, 8), 256, 0, stream>>>(make_view(ix), d_cands, n_cands, d_term_bytes,
                                                                d_term_boff, d_idf.p, n_terms, sp, d_scores);
    MGX_LAUNCH_CHECK();
  }
  MGX_CUDA(cudaStreamSynchronize(stream));
}

void launch_sort_by_score(cudaStream_t stream, const uint32_t* d_docs, const double* d_scores, uint64_t n,
                          bool descending, uint32_t limit, uint32_t offset, uint32_t* d_out, uint32_t* d_out_count) {
  sort_by_score_kernel<<<1, 256, 0, stream>>>(d_docs, d_scores, n, descending ? 1 : 0, limit, offset, d_out,
                                              d_out_count);
  MGX_LAUNCH_CHECK();
}

}  // namespace mgx
