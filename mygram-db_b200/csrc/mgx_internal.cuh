// mgx_internal.cuh — shared declarations of libmgx.so (host + device helpers).
//
// Data layout of one index shard in HBM (all arrays cudaMalloc'ed, 256-B aligned):
//   d_doc_ids   u32[N]        global DocIds, strictly ascending; postings store the
//                             LOCAL index (rank) so bitmaps are exactly N bits
//   d_text      u8[bytes]     normalised UTF-8 of all documents, back to back
//   d_text_off  u64[N+1]      byte range of document i
//   d_doc_len   u32[N]        CountCodePoints(text_i)   (string_utils.cpp:655-669)
//   d_term_keys u64[T]        packed n-grams, ascending == UTF-8 bytewise order
//   d_term_off  u64[T+1]      CSR offsets into d_postings
//   d_postings  u32[P]        local doc indices, ascending inside each term
//   d_term_bm   i32[T]        dense-bitmap slot of the term or -1
//   d_bitmaps   u32[D][W]     W = ceil(N/32) words per dense term
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <new>
#include <type_traits>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mgx.h"

namespace mgx {

// ---------------------------------------------------------------- errors
void set_last_error(const std::string& message);
void count_launch(uint64_t n = 1);

struct CudaFailure {
  int code;
};

#define MGX_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t mgx_err__ = (expr);                                                                 \
    if (mgx_err__ != cudaSuccess) {                                                                 \
      ::mgx::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(mgx_err__) + " at " +    \
                            __FILE__ + ":" + std::to_string(__LINE__));                             \
      throw ::mgx::CudaFailure{MGX_ERR_CUDA};                                                       \
    }                                                                                               \
  } while (0)

#define MGX_LAUNCH_CHECK()                \
  do {                                    \
    ::mgx::count_launch();                \
    MGX_CUDA(cudaGetLastError());         \
  } while (0)

// Owning device buffer (cudaMalloc / cudaFree).
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  bool owned = true;  // false: a view into a larger allocation (see DevArena)
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), owned(o.owned) {
    o.p = nullptr;
    o.n = 0;
  }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p;
      n = o.n;
      owned = o.owned;
      o.p = nullptr;
      o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p != nullptr && owned) {
      cudaFree(p);
    }
    p = nullptr;
    n = 0;
    owned = true;
  }
  void borrow(T* ptr, size_t count) {
    release();
    p = ptr;
    n = count;
    owned = false;
  }
  void alloc(size_t count) {
    release();
    n = count;
    if (count == 0) {
      count = 1;
    }
    MGX_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)));
  }
  // grow-only (keeps the allocation when it is already large enough). A buffer that has to grow takes 50 % more
  // than asked: cudaFree synchronises the device, so per-batch workspaces must stop reallocating after a few batches.
  void reserve(size_t count) {
    if (p == nullptr || count > n) {
      alloc(count + count / 2 + 64);
    }
  }
  size_t bytes() const { return n * sizeof(T); }
};

// One device allocation carved into 256-byte aligned pieces. cudaMalloc / cudaFree cost ~10 ms each on this
// platform (cudaFree synchronises the device), so a build makes a handful of them instead of dozens.
struct DevArena {
  DevBuf<uint8_t> blob;
  size_t used = 0;
  static size_t padded(size_t bytes) { return (bytes + 255) & ~static_cast<size_t>(255); }
  void reserve(size_t bytes, bool headroom = false) {  // grow-only: an arena that is big enough is reused as is
    if (blob.p == nullptr || blob.n < bytes + 256) {
      blob.alloc(bytes + 256 + (headroom ? bytes / 2 : 0));
    }
    used = 0;
  }
  template <typename T>
  T* take(size_t count) {
    const size_t bytes = padded((count == 0 ? 1 : count) * sizeof(T));
    if (used + bytes > blob.n) {
      set_last_error("internal: device arena exhausted");
      throw CudaFailure{MGX_ERR_CUDA};
    }
    T* out = reinterpret_cast<T*>(blob.p + used);
    used += bytes;
    return out;
  }
  void release() {
    blob.release();
    used = 0;
  }
};

// Pinned host buffer (cudaMallocHost), grow-only.
template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t n = 0;
  PinBuf() = default;
  PinBuf(const PinBuf&) = delete;
  PinBuf& operator=(const PinBuf&) = delete;
  ~PinBuf() {
    if (p != nullptr) {
      cudaFreeHost(p);
    }
  }
  void reserve(size_t count) {
    if (p != nullptr && count <= n) {
      return;
    }
    if (p != nullptr) {
      cudaFreeHost(p);
      p = nullptr;
    }
    n = count + count / 2 + 64;
    MGX_CUDA(cudaMallocHost(reinterpret_cast<void**>(&p), n * sizeof(T)));
  }
};

// MGX_BUILD_TRACE=1 prints the host-side wall time of build phases (stream synchronised) to stderr.
struct PhaseTrace {
  bool on;
  cudaStream_t stream;
  double t0;
  explicit PhaseTrace(cudaStream_t s);
  void mark(const char* name);
};

// ---------------------------------------------------------------- packed keys
// A key holds up to `width` (<= 3) code points as 21-bit fields storing cp+1,
// first code point in the most significant used field, absent trailing code
// points = 0. Integer order of keys == bytewise order of the UTF-8 n-gram
// strings (the reference sorts std::string keys, string_utils.h:192-196), and a
// shorter n-gram that is a prefix of a longer one sorts first, like strings do.
constexpr int kMaxKeyWidth = 3;
// Wider n-grams (sizes 4..10, config-schema.json:279-285) are "wide keys": kWords = ceil(width / 3) words of the same
// 21-bit packing (word 0 = code points 0..2, first code point in the most significant field), compared word by word,
// which is again the bytewise order of the UTF-8 strings. The index sorts them with one 64-bit radix sort per word
// (least significant word first, through a permutation) and then names every distinct n-gram by its RANK + 1: the
// rest of the build, the CSR and every query kernel see that rank as "the key" (d_term_keys[t] == t + 1), and only the
// dictionary lookup compares wide words (d_wide_keys). On the host a wide key is a handle into the calling thread's
// WidePool (host_make_key); batch staging copies the words next to the handles.
constexpr int kMaxNgramSize = 10;
constexpr int kMaxWideWords = 4;
__host__ __device__ inline int wide_words_for(int width) { return width <= kMaxKeyWidth ? 0 : (width + 2) / 3; }
constexpr uint32_t kTextTileBytes = 8192;  // arena tile of the streaming df pass (256 threads x 16 B x 2 rounds)
constexpr uint64_t kInvalidKey = ~0ULL;
constexpr int kPosBits = 22;               // spare low bits of a packed key of width <= 2 (2 * 21 + 22 = 64)
constexpr uint16_t kPosUnknown = 0x7FFF;   // post_pos value: first occurrence not recorded
constexpr uint16_t kPosMulti = 0x8000;     // post_pos flag: more than one occurrence in the document
inline int pos_bits_for_width(int width) { return 21 * width + kPosBits <= 64 ? kPosBits : 0; }

// Neighbour signatures (round 2). The kPosBits spare bits of a key are split per index into
//   [ prev signature : sig_prev_bits ][ next signature : sig_next_bits ][ byte offset : pos_field_bits ]
// where pos_field_bits is what the LONGEST document of the shard needs (8..15; 15 saturates at kPosUnknown as before).
// A signature is a few hash bits of the code point right before / right after the n-gram occurrence in the same
// document, 0 = "no such character". The verified-df kernels use them as an exact pre-filter: a term that contains
// the n-gram with a character after (before) it can only start at a recorded occurrence whose next (prev) signature
// equals that character's (a cached comparison of one character of text; false positives go on to the text check,
// false negatives cannot happen because a true occurrence of the term puts exactly that character there).
constexpr int kSigFieldBits = 7;  // widest signature (the posting payload keeps two 7-bit fields per occurrence)
__host__ __device__ inline uint32_t sig_hash7(uint32_t cp) { return (cp * 0x9E3779B1u) >> 25; }
// b-bit signature of a 7-bit hash, never 0 (0 = no neighbour)
__host__ __device__ inline uint32_t sig_of_hash(uint32_t h7, int bits) {
  const uint32_t s = h7 >> (kSigFieldBits - bits);
  return s != 0 ? s : 1u;
}
struct SigLayout {
  int pos_bits = 0;   // low bits holding the byte offset (0 = the keys carry no payload)
  int next_bits = 0;
  int prev_bits = 0;
};
inline SigLayout sig_layout_for(int payload_bits, uint64_t max_doc_bytes) {
  SigLayout l;
  if (payload_bits <= 0) {
    return l;
  }
  int need = 1;
  while (need < 15 && (1ULL << need) <= max_doc_bytes) {  // offsets 0 .. max_doc_bytes - 1 and never the saturated value
    ++need;
  }
  l.pos_bits = need < 8 ? 8 : need;
  const int spare = payload_bits - l.pos_bits;
  l.next_bits = (spare + 1) / 2 > kSigFieldBits ? kSigFieldBits : (spare + 1) / 2;
  l.prev_bits = spare - l.next_bits > kSigFieldBits ? kSigFieldBits : spare - l.next_bits;
  return l;
}
// Term side (key_toff, one uint32 per (term, n-gram)): low 16 bits as before (kNoTermOffset / offset | count << 12);
// bits 16..22 = hash7 of the character after the n-gram's first occurrence in the term, bit 23 = there is one,
// bits 24..30 / bit 31 = the same for the character before it.
constexpr uint32_t kToffNextShift = 16;
constexpr uint32_t kToffHasNext = 1u << 23;
constexpr uint32_t kToffPrevShift = 24;
constexpr uint32_t kToffHasPrev = 1u << 31;
__host__ __device__ inline uint32_t toff_neighbours(bool has_prev, uint32_t prev_cp, bool has_next, uint32_t next_cp) {
  return (has_next ? (kToffHasNext | (sig_hash7(next_cp) << kToffNextShift)) : 0u) |
         (has_prev ? (kToffHasPrev | (sig_hash7(prev_cp) << kToffPrevShift)) : 0u);
}

__host__ __device__ inline uint64_t pack_key(const uint32_t* cps, int n, int width) {
  uint64_t key = 0;
  for (int j = 0; j < width; ++j) {
    key = (key << 21) | (j < n ? static_cast<uint64_t>(cps[j]) + 1 : 0ULL);
  }
  return key;
}

// words[0..n_words) of the wide key of cps[0..n) (n <= 3 * n_words)
__host__ __device__ inline void pack_wide(const uint32_t* cps, int n, int n_words, uint64_t* words) {
  for (int w = 0; w < n_words; ++w) {
    uint64_t word = 0;
    for (int f = 0; f < 3; ++f) {
      const int j = 3 * w + f;
      word = (word << 21) | (j < n ? static_cast<uint64_t>(cps[j]) + 1 : 0ULL);
    }
    words[w] = word;
  }
}

// IsCJKIdeograph, string_utils.cpp:441-448 (ranges :176-187).
__host__ __device__ inline bool is_cjk_ideograph(uint32_t cp) {
  return (cp >= 0x4E00u && cp <= 0x9FFFu) || (cp >= 0x3400u && cp <= 0x4DBFu) || (cp >= 0x20000u && cp <= 0x2A6DFu) ||
         (cp >= 0x2A700u && cp <= 0x2B73Fu) || (cp >= 0x2B740u && cp <= 0x2B81Fu) || (cp >= 0xF900u && cp <= 0xFAFFu);
}

// TryParseUtf8Char (string_utils.cpp:92-164) on four bytes already in
// registers: b0 is the candidate lead byte, b1..b3 the following bytes (any
// value when beyond `available`). Returns the sequence length (1..4) or 0 when
// no valid character starts here.
__host__ __device__ inline int parse_utf8(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint64_t available,
                                          uint32_t* cp) {
  if (available == 0) {
    return 0;
  }
  if (b0 < 0x80u) {
    *cp = b0;
    return 1;
  }
  if ((b0 & 0xE0u) == 0xC0u) {
    if (b0 < 0xC2u || available < 2 || (b1 & 0xC0u) != 0x80u) {
      return 0;
    }
    *cp = ((b0 & 0x1Fu) << 6) | (b1 & 0x3Fu);
    return 2;
  }
  if ((b0 & 0xF0u) == 0xE0u) {
    if (available < 3 || (b1 & 0xC0u) != 0x80u || (b2 & 0xC0u) != 0x80u) {
      return 0;
    }
    const uint32_t c = ((b0 & 0x0Fu) << 12) | ((b1 & 0x3Fu) << 6) | (b2 & 0x3Fu);
    if (c < 0x800u || (c >= 0xD800u && c <= 0xDFFFu)) {
      return 0;
    }
    *cp = c;
    return 3;
  }
  if ((b0 & 0xF8u) == 0xF0u) {
    if (b0 > 0xF4u || available < 4 || (b1 & 0xC0u) != 0x80u || (b2 & 0xC0u) != 0x80u || (b3 & 0xC0u) != 0x80u) {
      return 0;
    }
    const uint32_t c = ((b0 & 0x07u) << 18) | ((b1 & 0x3Fu) << 12) | ((b2 & 0x3Fu) << 6) | (b3 & 0x3Fu);
    if (c < 0x10000u || c > 0x10FFFFu) {
      return 0;
    }
    *cp = c;
    return 4;
  }
  return 0;
}

// ---------------------------------------------------------------- host tokenizer (queries)
// Utf8ToCodepoints, string_utils.cpp:200-219.
std::vector<uint32_t> host_utf8_to_codepoints(const uint8_t* text, uint64_t len);
// Vector of trivially copyable elements with room for N of them inside the object: the query compiler builds ~10^4
// terms and queries per batch, nearly all with <= 4 n-grams / terms, and a heap allocation for each of them was most
// of its run time. Only the std::vector operations the compiler uses are provided.
template <class T, unsigned N>
class SmallVec {
 public:
  SmallVec() = default;
  SmallVec(const SmallVec& o) { assign(o.begin(), o.end()); }
  SmallVec(SmallVec&& o) noexcept { steal(o); }
  SmallVec(std::initializer_list<T> il) { assign(il.begin(), il.end()); }
  ~SmallVec() { release(); }
  SmallVec& operator=(const SmallVec& o) {
    if (this != &o) {
      assign(o.begin(), o.end());
    }
    return *this;
  }
  SmallVec& operator=(SmallVec&& o) noexcept {
    if (this != &o) {
      release();
      steal(o);
    }
    return *this;
  }
  SmallVec& operator=(const std::vector<T>& v) {
    assign(v.data(), v.data() + v.size());
    return *this;
  }
  SmallVec& operator=(std::initializer_list<T> il) {
    assign(il.begin(), il.end());
    return *this;
  }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  T* data() { return p_; }
  const T* data() const { return p_; }
  T* begin() { return p_; }
  T* end() { return p_ + n_; }
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }
  T& operator[](size_t i) { return p_[i]; }
  const T& operator[](size_t i) const { return p_[i]; }
  T& back() { return p_[n_ - 1]; }
  void clear() { n_ = 0; }
  void reserve(size_t want) {
    if (want > cap_) {
      grow(want);
    }
  }
  void push_back(const T& v) {
    if (n_ == cap_) {
      grow(static_cast<size_t>(cap_) * 2);
    }
    p_[n_++] = v;
  }
  T* erase(T* first, T* last) {  // std::vector::erase(first, last)
    if (first != last) {
      std::memmove(first, last, static_cast<size_t>(end() - last) * sizeof(T));
      n_ -= static_cast<uint32_t>(last - first);
    }
    return first;
  }

 private:
  static_assert(std::is_trivially_copyable<T>::value, "SmallVec holds plain data only");
  void assign(const T* first, const T* last) {
    const size_t n = static_cast<size_t>(last - first);
    n_ = 0;
    reserve(n);
    if (n > 0) {
      std::memcpy(p_, first, n * sizeof(T));
    }
    n_ = static_cast<uint32_t>(n);
  }
  void grow(size_t want) {
    T* fresh = static_cast<T*>(std::malloc(want * sizeof(T)));
    if (fresh == nullptr) {
      throw std::bad_alloc();
    }
    if (n_ > 0) {
      std::memcpy(fresh, p_, n_ * sizeof(T));
    }
    release();
    p_ = fresh;
    cap_ = static_cast<uint32_t>(want);
  }
  void release() {
    if (p_ != inline_) {
      std::free(p_);
      p_ = inline_;
      cap_ = N;
    }
  }
  void steal(SmallVec& o) {
    if (o.p_ != o.inline_) {
      p_ = o.p_;
      cap_ = o.cap_;
      o.p_ = o.inline_;
      o.cap_ = N;
    } else {
      p_ = inline_;
      cap_ = N;
      if (o.n_ > 0) {
        std::memcpy(inline_, o.inline_, o.n_ * sizeof(T));
      }
    }
    n_ = o.n_;
    o.n_ = 0;
  }
  T* p_ = inline_;
  uint32_t n_ = 0;
  uint32_t cap_ = N;
  T inline_[N];
};
using KeyVec = SmallVec<uint64_t, 4>;
using TermOffsetVec = SmallVec<uint32_t, 4>;
using TermIdVec = SmallVec<uint32_t, 4>;

// GenerateQueryNgrams (string_utils.cpp:639-653) + DeduplicateSorted as packed
// keys. Returns false if a window is wider than kMaxKeyWidth.
// key_toff (optional): per returned key, (byte offset of the n-gram's FIRST occurrence inside the term) |
// (min(number of occurrences in the term, 3) << kTermCountShift); kNoTermOffset when the term is not valid UTF-8
// (its windows skip bytes, so they are no contiguous byte runs).
constexpr uint16_t kNoTermOffset = 0xFFFF;  // low 16 bits of a key_toff entry
constexpr uint32_t kTermOffsetMask = 0x0FFF;
constexpr uint32_t kTermCountShift = 12;
bool host_query_keys(const uint8_t* term, uint64_t len, int ngram_size, int kanji_ngram_size, bool cross_boundary,
                     int key_width, KeyVec* keys, TermOffsetVec* key_toff = nullptr, bool* all_valid = nullptr);
// all_valid (optional): the term decoded without skipping a byte (it is valid UTF-8).
// One n-gram string -> packed key (for Index::SearchAnd style calls). False if
// it is not valid UTF-8 of 1..width code points (such a term cannot be in the index).
bool host_ngram_to_key(const uint8_t* term, uint64_t len, int key_width, uint64_t* key);
// Key of the n-gram cps[0..n) for an index of key width `width`: the packed key for width <= 3, otherwise a handle
// (>= 1) into the calling thread's pool of wide keys -- equal n-grams get equal handles for as long as the pool lives,
// i.e. until the next top-level API call on this thread starts (WideScope).
uint64_t host_make_key(const uint32_t* cps, int n, int width);
// the n_words words behind a handle of the calling thread's pool
const uint64_t* host_wide_words(uint64_t handle, int n_words);
// UTF-8 bytes of a key (packed or handle), returns their number; out: 4 * width bytes
int host_key_to_utf8(uint64_t key, int width, uint8_t* out);
int wide_words_to_utf8(const uint64_t* words, int n_words, uint8_t* out);
// Pool of the wide keys one API call makes (thread-local). Scopes nest: only the outermost one clears the pool.
struct WideScope {
  WideScope();
  ~WideScope();
  WideScope(const WideScope&) = delete;
  WideScope& operator=(const WideScope&) = delete;
};
// Handles of a pool that worker threads filled: append another thread's words to this thread's pool and return the
// offset to add to its handles.
struct WidePoolSnapshot {
  std::vector<uint64_t> words;  // 4 words per key
};
WidePoolSnapshot wide_pool_snapshot();
uint64_t wide_pool_adopt(const WidePoolSnapshot& other);

// ---------------------------------------------------------------- device-wide primitives (primitives.cu)
// out[i] = sum_{j<i} in[j]  (u32 -> u64), out has n+1 entries (out[n] = total).
// d_scratch: scan_scratch_elems(n) uint64 of caller-provided device memory (no allocation inside).
size_t scan_scratch_elems(uint64_t n);
// Up to three short, independent scans of the same kind in one launch (one CTA per array; meant for n <~ 10^5).
struct SmallScanJobs {
  const uint32_t* in[3];
  uint64_t* out[3];  // n + 1 entries each
  uint64_t n[3];
};
constexpr uint64_t kSmallScanMax = 1ULL << 17;
void exclusive_scans_small(const SmallScanJobs& jobs, int n_jobs, cudaStream_t stream);
void exclusive_scan_u32_u64(const uint32_t* d_in, uint64_t* d_out, uint64_t n, uint64_t* d_scratch,
                            cudaStream_t stream);
// Stable LSD radix sort of (u64 key, u32 value) pairs on key bits [0, key_bits).
// Buffers ping-pong; returns which pair of pointers holds the sorted data.
struct SortResult {
  uint64_t* keys;
  uint32_t* vals;
};
// d_scratch: radix_sort_scratch_bytes(n) bytes of caller-provided device memory (no allocation inside).
size_t radix_sort_scratch_bytes(uint64_t n);
// Only key bits [bit_base, bit_base + key_bits) take part; lower bits ride along (stable order is kept).
SortResult radix_sort_pairs(uint64_t* d_keys_a, uint32_t* d_vals_a, uint64_t* d_keys_b, uint32_t* d_vals_b, uint64_t n,
                            int key_bits, int bit_base, uint8_t* d_scratch, cudaStream_t stream);

// Search workspace shared by all batches that run on one stream of one index (batches on a stream execute one after
// the other, so they can share it; growing it happens during the first batches only). Grow-only.
struct SearchScratch {
  cudaStream_t stream = nullptr;
  DevBuf<uint32_t> tile_count;
  DevBuf<uint32_t> tile_total;
  DevBuf<uint32_t> rec_doc;
  DevBuf<double> rec_score;
  DevBuf<uint4> df_tile_desc;  // DfTileDesc records (two uint4 each), see query.cuh
  DevBuf<uint32_t> tile_query;
  DevBuf<uint32_t> topk_groups;  // (query, first tile, end tile) triples of the top-k pre-reduction
  uint64_t map_owner = 0;  // serial of the batch whose tile maps are in df_tile_term / tile_query
  int skip_streamed = 0;   // batches to run in the synchronous form after one that cannot fit any workspace
};

// One DocumentStore filter column mirrored on the device (see mgx_index_set_filter_column).
constexpr uint32_t kMaxFilterColumns = 64;
enum FilterClass : uint32_t { kFcNone = 0, kFcBool = 1, kFcSigned = 2, kFcUnsigned = 3, kFcString = 4, kFcDouble = 5 };
struct FilterColumn {
  FilterClass cls = kFcNone;
  uint64_t n_docs = 0;
  DevBuf<uint64_t> values;           // integer / bool value, double bits, or the rank of the string in `dict`
  DevBuf<uint8_t> nulls;
  std::vector<std::string> dict;     // sorted distinct strings (string columns)
};

// ---------------------------------------------------------------- index object
struct Index {
  mgx_index_config_t cfg{};
  int ngram = 2;
  int kanji = 2;  // effective
  bool cross = true;
  int width = 2;
  int device = 0;

  uint64_t n_docs = 0;
  bool sequential_ids = true;
  uint32_t first_id = 1;
  DevBuf<uint32_t> d_doc_ids;
  DevBuf<uint8_t> d_text;
  DevBuf<uint64_t> d_text_off;
  DevBuf<uint32_t> d_doc_len;
  uint64_t text_bytes = 0;
  // document holding the first byte of every kTextTileBytes tile of the arena (streaming df pass)
  DevBuf<uint32_t> d_tile_first_doc;
  uint64_t n_text_tiles = 0;

  uint64_t n_terms = 0;
  uint64_t n_postings = 0;
  DevBuf<uint64_t> d_term_keys;
  int wide_words = 0;              // 0: packed keys; otherwise words per wide key (width > 3)
  DevBuf<uint64_t> d_wide_keys;    // [n_terms * wide_words], term-major, ascending; d_term_keys[t] == t + 1 then
  DevBuf<uint64_t> d_term_off;
  DevBuf<uint32_t> d_postings;
  // Payload of every posting (only when the packed key leaves room to carry it through the sort, i.e. key width
  // <= 2). Low half = first-occurrence position: bits 0..14 = byte offset of the n-gram's first occurrence in the
  // document's text (0x7FFF = unknown / beyond 32 KB), bit 15 = the n-gram occurs more than once in the document.
  // High half = neighbour signatures of that occurrence: bits 16..22 next, bits 24..30 prev (sig layout above).
  DevBuf<uint32_t> d_post_pos;
  // Second occurrence, same encoding: bits 0..14 = byte offset (0x7FFF = none / unknown), bit 15 = a third exists.
  DevBuf<uint32_t> d_post_pos2;
  bool has_positions = false;
  SigLayout sig;  // how the build split the payload bits (sig.next_bits == sig.prev_bits == 0: no signatures)
  bool text_less = false;  // loaded from an MGIX stream: posting lists only, no document text (load_index_device)
  DevBuf<int32_t> d_term_bm;
  DevBuf<uint32_t> d_bitmaps;
  DevArena resident_a;  // doc ids, text, text offsets, doc lengths
  DevArena resident_b;  // dictionary, CSR offsets, postings, bitmap slots
  DevArena build_arena0;  // build workspaces (tokenizer scratch; pair arrays + sort scratch), kept between builds:
  DevArena build_arena;   // see build_index_device; released by mgx_index_trim
  uint64_t n_dense = 0;
  uint64_t bm_words = 0;
  uint64_t dense_min_len = 0;

  uint64_t total_doc_length = 0;
  uint64_t doc_count = 0;
  bool all_valid_utf8 = true;
  uint64_t n_pair_slots = 0;
  double last_build_ms = 0.0;

  std::vector<FilterColumn*> columns = std::vector<FilterColumn*>(kMaxFilterColumns, nullptr);
  void drop_filter_columns();
  uint64_t optimized_total_docs = 0;  // argument of the last Index::Optimize call, 0 = never optimised
  mgx_batch_stats_t last_stats{};
  cudaStream_t stream = nullptr;  // owned, non-blocking; used by the non-staged calls
  // recycled batch workspaces (device arenas + pinned staging), so a steady stream of batches allocates nothing
  std::vector<void*> batch_pool;
  std::mutex pool_mu;
  std::vector<SearchScratch*> scratches;  // one per stream that has run a batch; guarded by pool_mu
  SearchScratch* scratch_for(cudaStream_t stream);
  ~Index();

  uint64_t device_bytes() const;
};

// Device-side view passed to kernels by value.
struct IndexView {
  const uint32_t* doc_ids;
  const uint8_t* text;
  const uint64_t* text_off;
  const uint32_t* doc_len;
  const uint32_t* tile_first_doc;
  uint64_t text_bytes;
  uint64_t n_text_tiles;
  const uint64_t* term_keys;
  const uint64_t* wide_keys;  // nullptr unless the index has wide keys
  int wide_words;
  const uint64_t* term_off;
  const uint32_t* postings;
  const uint32_t* post_pos;  // nullptr when the index carries no positions; low 16 bits position, high 16 signatures
  const uint32_t* post_pos2;
  int sig_next_bits;         // signature widths of the payload (0: none)
  int sig_prev_bits;
  const int32_t* term_bm;
  const uint32_t* bitmaps;
  uint64_t n_docs;
  uint64_t n_terms;
  uint64_t bm_words;
  uint32_t first_id;    // DocId of local index 0
  int sequential_ids;   // doc_ids[i] == first_id + i (DocumentStore assigns ids sequentially, document_store.h:520)
  int all_valid_utf8;   // no document of the shard contains an invalid byte
};

// local index -> global DocId
__device__ __forceinline__ uint32_t gid_of(const IndexView& iv, uint32_t local) {
  return iv.sequential_ids ? iv.first_id + local : __ldg(iv.doc_ids + local);
}
// global DocId -> local index, or 0xFFFFFFFF if the shard does not hold it
__device__ __forceinline__ uint32_t local_of(const IndexView& iv, uint32_t gid) {
  if (iv.sequential_ids) {
    const uint32_t d = gid - iv.first_id;
    return (gid >= iv.first_id && d < iv.n_docs) ? d : 0xFFFFFFFFu;
  }
  uint32_t lo = 0;
  uint32_t hi = static_cast<uint32_t>(iv.n_docs);
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(iv.doc_ids + mid) < gid) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return (lo < iv.n_docs && __ldg(iv.doc_ids + lo) == gid) ? lo : 0xFFFFFFFFu;
}

inline IndexView make_view(const Index& ix) {
  IndexView v;
  v.doc_ids = ix.d_doc_ids.p;
  v.text = ix.d_text.p;
  v.text_off = ix.d_text_off.p;
  v.doc_len = ix.d_doc_len.p;
  v.tile_first_doc = ix.d_tile_first_doc.p;
  v.text_bytes = ix.text_bytes;
  v.n_text_tiles = ix.n_text_tiles;
  v.term_keys = ix.d_term_keys.p;
  v.wide_keys = ix.wide_words > 0 ? ix.d_wide_keys.p : nullptr;
  v.wide_words = ix.wide_words;
  v.term_off = ix.d_term_off.p;
  v.postings = ix.d_postings.p;
  v.post_pos = ix.has_positions ? ix.d_post_pos.p : nullptr;
  v.post_pos2 = ix.has_positions ? ix.d_post_pos2.p : nullptr;
  // MGX_DF_NO_SIG (read per view, the tests and the bench flip it): the signature pre-filter of the df kernels off
  const bool use_sig = ix.has_positions && std::getenv("MGX_DF_NO_SIG") == nullptr;
  v.sig_next_bits = use_sig ? ix.sig.next_bits : 0;
  v.sig_prev_bits = use_sig ? ix.sig.prev_bits : 0;
  v.term_bm = ix.d_term_bm.p;
  v.bitmaps = ix.d_bitmaps.p;
  v.n_docs = ix.n_docs;
  v.n_terms = ix.n_terms;
  v.bm_words = ix.bm_words;
  v.first_id = ix.first_id;
  v.sequential_ids = ix.sequential_ids ? 1 : 0;
  v.all_valid_utf8 = ix.all_valid_utf8 ? 1 : 0;
  return v;
}

// build.cu
// The three inputs may be host or device pointers (cudaMemcpyDefault); the index keeps its own copies.
void build_index_device(Index& ix, const uint32_t* doc_ids, const uint8_t* text, const uint64_t* text_off,
                        uint64_t n_docs, uint64_t text_bytes, cudaStream_t stream);
// Index::LoadFromStream's device side: the index becomes the given CSR (ascending packed keys, GLOBAL ascending doc ids
// per list, host arrays). No document text comes with a stream, so the shard holds none (text_less): the set calls
// answer from the lists, text-dependent paths see documents without stored text.
void load_index_device(Index& ix, const uint64_t* h_keys, const uint64_t* h_term_off, const uint32_t* h_postings,
                       uint64_t n_terms, uint64_t n_postings, cudaStream_t stream);
// Folds a journal of document mutations (host arrays sorted by id; removed[j] != 0 deletes, otherwise the text
// replaces / adds the document) into the resident corpus of `ix` on the device and builds the resulting shard INTO
// `next` (another Index object with the same configuration; `ix` is only read, so calls that read it may run beside
// this one). The filter columns of `ix` follow their documents into `next`. swap_generation then makes it current.
void apply_journal_device(const Index& ix, Index& next, const uint32_t* h_ids, const uint8_t* h_removed,
                          const uint8_t* h_text, const uint64_t* h_off, uint64_t n_j, cudaStream_t stream);
// Exchanges everything a (re)build produces -- corpus mirror, dictionary, lists, payload, bitmaps, counters, filter
// columns, resident arenas -- between two Index objects; configuration, streams, workspaces and pools stay.
void swap_generation(Index& a, Index& b);
// configuration of `from` into `to` (what a build reads)
void copy_index_config(Index& to, const Index& from);
// Number of posting lists the reference would hold as Roaring bitmaps (see mgx_index_get_statistics).
uint64_t count_roaring_lists(const Index& ix, double roaring_threshold, uint64_t optimized_total_docs,
                             cudaStream_t stream);
// Flat tile-based tokenizer (build.cu). Stage 1: per-document code-point counts (d_doc_len), per-tile n-gram counts
// scanned into d_tile_off (tokenize_tile_count() + 1 entries; the last is the total). counters_out: [0] non-empty
// docs, [1] docs with invalid bytes, [2] total code points. d_scratch: tokenize_scratch_elems() uint64 of device
// memory, to be passed again (untouched) to stage 2.
size_t tokenize_tile_count(uint64_t n_docs, uint64_t text_bytes);
size_t tokenize_scratch_elems(uint64_t n_docs, uint64_t text_bytes);
void tokenize_count(int ngram, int kanji, bool cross, int width, const uint8_t* d_text, const uint64_t* d_text_off,
                    uint64_t n_docs, uint64_t text_bytes, uint32_t* d_doc_len, uint64_t* d_tile_off,
                    uint64_t* d_scratch, uint64_t* n_slots, uint64_t* counters_out, cudaStream_t stream);
// Stage 2: (packed key, doc) pairs, exactly d_tile_off[n_tiles] of them, in text order. pos_bits > 0: every key is
// shifted left by pos_bits and carries the byte offset of the n-gram in its document (saturated) in the low bits.
void tokenize_emit(int ngram, int kanji, bool cross, int width, const uint8_t* d_text, const uint64_t* d_text_off,
                   uint64_t n_docs, uint64_t text_bytes, const uint64_t* d_tile_off, uint64_t* d_scratch,
                   uint64_t* d_keys, uint32_t* d_docs, int pos_bits, cudaStream_t stream, uint64_t wide_stride = 0);
// wide_stride (key width > 3): d_keys holds wide_words_for(width) arrays of wide_stride words, word w of slot s at
// d_keys[w * wide_stride + s].

}  // namespace mgx
