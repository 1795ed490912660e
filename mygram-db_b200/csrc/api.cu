// api.cu — the C ABI of libmgx.so (include/mgx.h) and the host-side query compiler.
//
// Host work is limited to what the reference also does per request on the CPU
// before touching the index: decoding the (short) query terms and cutting them
// into n-grams (GenerateQueryNgrams, utils/string_utils.cpp:639-653). Everything
// that touches postings, document text or scores runs in the kernels of
// build.cu / query.cu. There is no CPU fallback: without a CUDA device every
// entry point fails with MGX_ERR_NO_DEVICE.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <nccl.h>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "query.cuh"

namespace mgx {

namespace {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};
}  // namespace

void set_last_error(const std::string& message) { g_last_error = message; }
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ------------------------------------------------------------------ host tokenizer for query terms
std::vector<uint32_t> host_utf8_to_codepoints(const uint8_t* text, uint64_t len) {
  std::vector<uint32_t> cps;
  cps.reserve(len / 2 + 1);
  uint64_t i = 0;
  while (i < len) {
    uint32_t cp = 0;
    const uint64_t avail = len - i;
    const int n = parse_utf8(text[i], avail > 1 ? text[i + 1] : 0, avail > 2 ? text[i + 2] : 0,
                             avail > 3 ? text[i + 3] : 0, avail, &cp);
    if (n > 0) {
      cps.push_back(cp);
      i += static_cast<uint64_t>(n);
    } else {
      ++i;  // string_utils.cpp:212-215: skip one byte and retry
    }
  }
  return cps;
}

namespace {

// GenerateHybridNgrams (string_utils.cpp:452-509) as packed keys; windows wider
// than the index's key width cannot exist in the dictionary => kInvalidKey.
void hybrid_keys(const std::vector<uint32_t>& cps, int ascii_n, int kanji_n, bool cross, int width,
                 KeyVec* keys, std::vector<uint32_t>* starts = nullptr) {
  if (ascii_n <= 0 || kanji_n <= 0) {
    return;
  }
  for (size_t i = 0; i < cps.size(); ++i) {
    const bool cjk = is_cjk_ideograph(cps[i]);
    const int size = cjk ? kanji_n : ascii_n;
    if (i + static_cast<size_t>(size) > cps.size()) {
      continue;
    }
    if (!cross) {
      bool crossed = false;
      for (int j = 1; j < size; ++j) {
        if (is_cjk_ideograph(cps[i + static_cast<size_t>(j)]) != cjk) {
          crossed = true;
          break;
        }
      }
      if (crossed) {
        continue;
      }
    }
    keys->push_back(size <= width ? host_make_key(cps.data() + i, size, width) : kInvalidKey);
    if (starts != nullptr) {
      starts->push_back(static_cast<uint32_t>(i));  // index of the window's first code point
    }
  }
}

// the pipeline's private, slightly wider CJK classifier (search_pipeline.cpp:73-78)
bool pipeline_is_cjk(uint32_t cp) {
  return is_cjk_ideograph(cp) || (cp >= 0x2B820u && cp <= 0x2CEAFu);
}

// HasUncoveredHybridFragment, search_pipeline.cpp:80-136
bool has_uncovered_hybrid_fragment(const uint8_t* term, uint64_t len, int ngram_size, int kanji_ngram_size,
                                   bool cross) {
  if (len == 0 || kanji_ngram_size <= 0) {
    return false;
  }
  const int ascii_n = ngram_size > 0 ? ngram_size : 2;
  const auto cps = host_utf8_to_codepoints(term, len);
  if (cps.size() < 2) {
    return false;
  }
  bool has_cjk = false;
  bool has_non = false;
  for (uint32_t cp : cps) {
    (pipeline_is_cjk(cp) ? has_cjk : has_non) = true;
  }
  if (!has_cjk || !has_non) {
    return false;
  }
  std::vector<bool> covered(cps.size(), false);
  for (size_t i = 0; i < cps.size(); ++i) {
    const bool start_cjk = pipeline_is_cjk(cps[i]);
    const int size = start_cjk ? kanji_ngram_size : ascii_n;
    if (size <= 0 || i + static_cast<size_t>(size) > cps.size()) {
      continue;
    }
    if (!cross) {
      bool crossed = false;
      for (int j = 1; j < size; ++j) {
        if (pipeline_is_cjk(cps[i + static_cast<size_t>(j)]) != start_cjk) {
          crossed = true;
          break;
        }
      }
      if (crossed) {
        continue;
      }
    }
    for (int j = 0; j < size; ++j) {
      covered[i + static_cast<size_t>(j)] = true;
    }
  }
  return std::any_of(covered.begin(), covered.end(), [](bool c) { return !c; });
}

}  // namespace

namespace {
// host_query_keys for the common case -- fixed-size n-grams (GenerateNgrams, string_utils.cpp:382-423) of a term of
// at most kFastTermBytes bytes: same result as the general path below, without its scratch vectors.
constexpr uint64_t kFastTermBytes = 96;
void fixed_ngram_keys_small(const uint8_t* term, uint64_t len, int ngram_size, int key_width, KeyVec* keys,
                            TermOffsetVec* key_toff, bool* all_valid) {
  uint32_t cps[kFastTermBytes];
  uint16_t cp_byte[kFastTermBytes];
  uint32_t n_cp = 0;
  bool valid = true;
  for (uint64_t i = 0; i < len;) {
    uint32_t cp = 0;
    const uint64_t avail = len - i;
    const int n = parse_utf8(term[i], avail > 1 ? term[i + 1] : 0, avail > 2 ? term[i + 2] : 0,
                             avail > 3 ? term[i + 3] : 0, avail, &cp);
    if (n > 0) {
      cps[n_cp] = cp;
      cp_byte[n_cp++] = static_cast<uint16_t>(i);
      i += static_cast<uint64_t>(n);
    } else {
      valid = false;
      ++i;  // string_utils.cpp:212-215: skip one byte and retry
    }
  }
  if (all_valid != nullptr) {
    *all_valid = valid;
  }
  keys->clear();
  if (key_toff != nullptr) {
    key_toff->clear();
  }
  const uint32_t n = static_cast<uint32_t>(ngram_size);
  if (n_cp < n) {
    return;
  }
  const uint32_t n_win = n_cp - n + 1;
  uint64_t wkey[kFastTermBytes];
  uint16_t wstart[kFastTermBytes];
  for (uint32_t i = 0; i < n_win; ++i) {  // insertion sort by (key, start): windows arrive in start order
    const uint64_t k = ngram_size <= key_width ? host_make_key(cps + i, ngram_size, key_width) : kInvalidKey;
    uint32_t j = i;
    while (j > 0 && wkey[j - 1] > k) {
      wkey[j] = wkey[j - 1];
      wstart[j] = wstart[j - 1];
      --j;
    }
    wkey[j] = k;
    wstart[j] = static_cast<uint16_t>(i);
  }
  for (uint32_t i = 0; i < n_win;) {  // sorted + unique == DeduplicateSorted, string_utils.h:192-196
    uint32_t j = i;
    while (j < n_win && wkey[j] == wkey[i]) {
      ++j;
    }
    keys->push_back(wkey[i]);
    if (key_toff != nullptr) {
      const uint32_t ws = wstart[i];  // equal keys keep their start order: wstart[i] is the first
      const uint32_t off = cp_byte[ws];
      const uint32_t cnt = std::min<uint32_t>(j - i, 3);
      key_toff->push_back(valid && off <= kTermOffsetMask
                              ? (off | (cnt << kTermCountShift) |
                                 toff_neighbours(ws > 0, ws > 0 ? cps[ws - 1] : 0, ws + n < n_cp, ws + n < n_cp ? cps[ws + n] : 0))
                              : static_cast<uint32_t>(kNoTermOffset));
    }
    i = j;
  }
}
}  // namespace

bool host_query_keys(const uint8_t* term, uint64_t len, int ngram_size, int kanji_ngram_size, bool cross_boundary,
                     int key_width, KeyVec* keys, TermOffsetVec* key_toff, bool* all_valid) {
  if (kanji_ngram_size <= 0 && ngram_size > 0 && len <= kFastTermBytes) {
    fixed_ngram_keys_small(term, len, ngram_size, key_width, keys, key_toff, all_valid);
    return true;
  }
  // scratch reused across calls: a batch compiles ~10^4 terms and must not allocate per term
  static thread_local std::vector<uint32_t> cps;
  static thread_local std::vector<uint32_t> cp_byte;
  static thread_local std::vector<uint32_t> starts;
  static thread_local std::vector<std::pair<uint64_t, uint32_t>> occ;
  keys->clear();
  cps.clear();
  cp_byte.clear();
  starts.clear();
  bool valid = true;  // no byte was skipped: windows are contiguous byte runs of the term
  {
    uint64_t i = 0;
    while (i < len) {
      uint32_t cp = 0;
      const uint64_t avail = len - i;
      const int n = parse_utf8(term[i], avail > 1 ? term[i + 1] : 0, avail > 2 ? term[i + 2] : 0,
                               avail > 3 ? term[i + 3] : 0, avail, &cp);
      if (n > 0) {
        cps.push_back(cp);
        cp_byte.push_back(static_cast<uint32_t>(i));
        i += static_cast<uint64_t>(n);
      } else {
        valid = false;
        ++i;  // string_utils.cpp:212-215: skip one byte and retry
      }
    }
  }
  if (all_valid != nullptr) {
    *all_valid = valid;
  }
  if (kanji_ngram_size > 0) {
    hybrid_keys(cps, ngram_size > 0 ? ngram_size : 2, kanji_ngram_size, cross_boundary, key_width, keys, &starts);
  } else if (ngram_size == 0) {
    hybrid_keys(cps, 2, 1, true, key_width, keys, &starts);  // GenerateHybridNgrams defaults, string_utils.h
  } else if (ngram_size > 0 && !cps.empty()) {
    // GenerateNgrams, string_utils.cpp:382-423
    const size_t n = static_cast<size_t>(ngram_size);
    if (cps.size() >= n) {
      for (size_t i = 0; i + n <= cps.size(); ++i) {
        keys->push_back(ngram_size <= key_width ? host_make_key(cps.data() + i, ngram_size, key_width) : kInvalidKey);
        starts.push_back(static_cast<uint32_t>(i));
      }
    }
  }
  if (key_toff != nullptr) {
    // Per n-gram: byte offset of its first occurrence inside the term and how often it occurs there (1, 2, 3+);
    // kNoTermOffset for every n-gram of a term that is not valid UTF-8 (its windows are no contiguous byte runs).
    occ.resize(keys->size());
    for (size_t i = 0; i < keys->size(); ++i) {
      occ[i] = {(*keys)[i], starts[i]};
    }
    std::sort(occ.begin(), occ.end());
    key_toff->clear();
    keys->clear();
    for (size_t i = 0; i < occ.size();) {
      size_t j = i;
      while (j < occ.size() && occ[j].first == occ[i].first) {
        ++j;
      }
      const uint32_t ws = occ[i].second;  // window start (code point index) of the first occurrence
      const uint32_t off = cp_byte[ws];
      keys->push_back(occ[i].first);  // sorted + unique == DeduplicateSorted, string_utils.h:192-196
      const uint32_t cnt = static_cast<uint32_t>(std::min<size_t>(j - i, 3));
      // the window's size follows the tokenizer's rule (fixed, or chosen by the class of the start code point)
      uint32_t wn = static_cast<uint32_t>(ngram_size);
      if (kanji_ngram_size > 0) {  // same rule as hybrid_keys above
        wn = static_cast<uint32_t>(is_cjk_ideograph(cps[ws]) ? kanji_ngram_size : (ngram_size > 0 ? ngram_size : 2));
      } else if (ngram_size == 0) {
        wn = is_cjk_ideograph(cps[ws]) ? 1u : 2u;
      }
      const bool has_next = ws + wn < cps.size();
      key_toff->push_back(valid && off <= kTermOffsetMask
                              ? (off | (cnt << kTermCountShift) |
                                 toff_neighbours(ws > 0, ws > 0 ? cps[ws - 1] : 0, has_next, has_next ? cps[ws + wn] : 0))
                              : static_cast<uint32_t>(kNoTermOffset));
      i = j;
    }
    return true;
  }
  std::sort(keys->begin(), keys->end());  // DeduplicateSorted, string_utils.h:192-196
  keys->erase(std::unique(keys->begin(), keys->end()), keys->end());
  return true;
}

bool host_ngram_to_key(const uint8_t* term, uint64_t len, int key_width, uint64_t* key) {
  // a dictionary key is the UTF-8 re-encoding of 1..width decoded code points;
  // a byte string with invalid sequences can therefore never equal one
  uint32_t cps[kMaxNgramSize];
  int n = 0;
  uint64_t i = 0;
  while (i < len) {
    uint32_t cp = 0;
    const uint64_t avail = len - i;
    const int l = parse_utf8(term[i], avail > 1 ? term[i + 1] : 0, avail > 2 ? term[i + 2] : 0,
                             avail > 3 ? term[i + 3] : 0, avail, &cp);
    if (l <= 0 || n >= key_width) {
      return false;
    }
    cps[n++] = cp;
    i += static_cast<uint64_t>(l);
  }
  if (n == 0) {
    return false;
  }
  *key = host_make_key(cps, n, key_width);
  return true;
}

// ---------------------------------------------------------------- wide keys on the host
namespace {
struct WidePool {
  std::vector<uint64_t> words;  // kMaxWideWords per key, unused words 0
  std::vector<uint32_t> table;  // open addressing: handle, 0 = empty
  int depth = 0;
  void clear() {
    words.clear();
    if (!table.empty()) {
      std::fill(table.begin(), table.end(), 0u);
    }
  }
  static uint64_t hash(const uint64_t* w) {
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    for (int i = 0; i < kMaxWideWords; ++i) {
      h = (h ^ w[i]) * 0xff51afd7ed558ccdULL;
      h ^= h >> 32;
    }
    return h;
  }
  void rehash(size_t cap) {
    table.assign(cap, 0u);
    const size_t n = words.size() / kMaxWideWords;
    for (size_t k = 0; k < n; ++k) {
      size_t slot = static_cast<size_t>(hash(words.data() + k * kMaxWideWords)) & (cap - 1);
      while (table[slot] != 0) {
        slot = (slot + 1) & (cap - 1);
      }
      table[slot] = static_cast<uint32_t>(k + 1);
    }
  }
  uint64_t intern(const uint64_t* w) {
    const size_t n = words.size() / kMaxWideWords;
    if (table.size() < 2 * (n + 1) + 16) {
      size_t cap = 64;
      while (cap < 4 * (n + 1) + 16) {
        cap <<= 1;
      }
      rehash(cap);
    }
    const size_t cap = table.size();
    size_t slot = static_cast<size_t>(hash(w)) & (cap - 1);
    while (table[slot] != 0) {
      if (std::memcmp(words.data() + static_cast<size_t>(table[slot] - 1) * kMaxWideWords, w,
                      kMaxWideWords * sizeof(uint64_t)) == 0) {
        return table[slot];
      }
      slot = (slot + 1) & (cap - 1);
    }
    words.insert(words.end(), w, w + kMaxWideWords);
    table[slot] = static_cast<uint32_t>(n + 1);
    return n + 1;
  }
};
thread_local WidePool tl_wide_pool;
}  // namespace

// The pool is emptied when the OUTERMOST scope ends: handles made before the call entered its guarded section (argument
// checks that tokenize) are still good inside it.
WideScope::WideScope() { ++tl_wide_pool.depth; }
WideScope::~WideScope() {
  if (--tl_wide_pool.depth == 0 && !tl_wide_pool.words.empty()) {
    tl_wide_pool.clear();
  }
}

uint64_t host_make_key(const uint32_t* cps, int n, int width) {
  if (width <= kMaxKeyWidth) {
    return pack_key(cps, n, width);
  }
  uint64_t w[kMaxWideWords];
  pack_wide(cps, n, kMaxWideWords, w);
  return tl_wide_pool.intern(w);
}

const uint64_t* host_wide_words(uint64_t handle, int n_words) {
  (void)n_words;
  static const uint64_t kNone[kMaxWideWords] = {0, 0, 0, 0};
  if (handle == 0 || handle > tl_wide_pool.words.size() / kMaxWideWords) {
    return kNone;  // a handle of another call: matches nothing (no n-gram has word 0 == 0)
  }
  return tl_wide_pool.words.data() + static_cast<size_t>(handle - 1) * kMaxWideWords;
}

WidePoolSnapshot wide_pool_snapshot() {
  WidePoolSnapshot snap;
  snap.words = tl_wide_pool.words;
  return snap;
}

uint64_t wide_pool_adopt(const WidePoolSnapshot& other) {
  WidePool& pool = tl_wide_pool;
  const uint64_t base = pool.words.size() / kMaxWideWords;
  pool.words.insert(pool.words.end(), other.words.begin(), other.words.end());
  if (!other.words.empty()) {
    pool.table.clear();  // rebuilt by the next intern
  }
  return base;
}

namespace {
int encode_utf8(uint32_t cp, uint8_t* out) {
  int n = 0;
  if (cp <= 0x7F) {
    out[n++] = static_cast<uint8_t>(cp);
  } else if (cp <= 0x7FF) {
    out[n++] = static_cast<uint8_t>(0xC0 | (cp >> 6));
    out[n++] = static_cast<uint8_t>(0x80 | (cp & 0x3F));
  } else if (cp <= 0xFFFF) {
    out[n++] = static_cast<uint8_t>(0xE0 | (cp >> 12));
    out[n++] = static_cast<uint8_t>(0x80 | ((cp >> 6) & 0x3F));
    out[n++] = static_cast<uint8_t>(0x80 | (cp & 0x3F));
  } else {
    out[n++] = static_cast<uint8_t>(0xF0 | (cp >> 18));
    out[n++] = static_cast<uint8_t>(0x80 | ((cp >> 12) & 0x3F));
    out[n++] = static_cast<uint8_t>(0x80 | ((cp >> 6) & 0x3F));
    out[n++] = static_cast<uint8_t>(0x80 | (cp & 0x3F));
  }
  return n;
}
}  // namespace

int wide_words_to_utf8(const uint64_t* words, int n_words, uint8_t* out) {
  int n = 0;
  for (int w = 0; w < n_words; ++w) {
    for (int f = 2; f >= 0; --f) {
      const uint64_t field = (words[w] >> (21 * f)) & 0x1FFFFFULL;
      if (field != 0) {
        n += encode_utf8(static_cast<uint32_t>(field - 1), out + n);
      }
    }
  }
  return n;
}

int host_key_to_utf8(uint64_t key, int width, uint8_t* out) {
  if (width <= kMaxKeyWidth) {
    return mgx_key_to_utf8(key, width, out);
  }
  return wide_words_to_utf8(host_wide_words(key, kMaxWideWords), kMaxWideWords, out);
}

namespace {

int require_device() {
  int count = 0;
  const cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count <= 0) {
    (void)cudaGetLastError();
    set_last_error("no CUDA device is visible; libmgx has no CPU fallback");
    return MGX_ERR_NO_DEVICE;
  }
  return MGX_OK;
}

template <typename F>
int guarded(F&& fn) {
  WideScope wide_scope;  // wide-key handles made by this call live until the next one starts
  try {
    return fn();
  } catch (const CudaFailure& f) {
    return f.code;
  } catch (const std::bad_alloc&) {
    set_last_error("host allocation failed");
    return MGX_ERR_CUDA;
  } catch (const std::exception& e) {
    set_last_error(std::string("unexpected: ") + e.what());
    return MGX_ERR_CUDA;
  }
}

int invalid(const char* what) {
  set_last_error(what);
  return MGX_ERR_INVALID_ARGUMENT;
}

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) {
      cudaSetDevice(dev);
    }
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

// Compile the caller's flat query description into unique terms + queries.
// Validates a postfix program and collects the terms every result must satisfy (the operands of the root when it is
// a TERM, or the TERM operands of a root AND, recursively through nested ANDs).
// positive_terms (optional): the TERM operands that are not below a NOT, left to right, duplicates kept -- the terms
// a boolean query is SCORED with (CollectAstScoringTerms, search_pipeline.cpp:232-254).
int analyse_program(const int32_t* ops, const int32_t* args, uint64_t n_ops, uint64_t n_terms,
                    std::vector<uint32_t>* conjuncts, std::vector<uint32_t>* positive_terms = nullptr) {
  struct Node {
    int op;
    int32_t arg;
    std::vector<size_t> kids;
  };
  std::vector<Node> nodes;
  std::vector<size_t> stack;
  size_t depth_max = 0;
  for (uint64_t i = 0; i < n_ops; ++i) {
    Node nd{ops[i], args[i], {}};
    if (ops[i] == kOpTerm) {
      if (args[i] < 0 || static_cast<uint64_t>(args[i]) >= n_terms) {
        return invalid("boolean program: TERM index out of range");
      }
    } else if (ops[i] == kOpFuzzyText) {
      if (args[i] < 0 || static_cast<uint64_t>(args[i] & 0xFFFFFF) >= n_terms) {
        return invalid("boolean program: FUZZYTEXT term index out of range");
      }
    } else if (ops[i] == kOpAnd || ops[i] == kOpOr || ops[i] == kOpAtLeast) {
      const int32_t n_kids = ops[i] == kOpAtLeast ? (args[i] & 0xFFFF) : args[i];
      if (args[i] < 0 || static_cast<size_t>(n_kids) > stack.size()) {
        return invalid("boolean program: operator has more children than the stack holds");
      }
      nd.kids.assign(stack.end() - n_kids, stack.end());
      stack.resize(stack.size() - static_cast<size_t>(n_kids));
    } else if (ops[i] == kOpNot) {
      if (stack.empty()) {
        return invalid("boolean program: NOT without an operand");
      }
      nd.kids.push_back(stack.back());
      stack.pop_back();
    } else {
      return invalid("boolean program: unknown op");
    }
    nodes.push_back(std::move(nd));
    stack.push_back(nodes.size() - 1);
    depth_max = std::max(depth_max, stack.size());
  }
  if (depth_max > kMaxProgramDepth) {
    set_last_error("boolean program deeper than 64 operands");
    return MGX_ERR_UNSUPPORTED;
  }
  if (stack.empty()) {
    return MGX_OK;
  }
  if (positive_terms != nullptr) {
    // pre-order, children left to right: the order the reference walks its AST in
    std::vector<std::pair<size_t, bool>> walk{{stack.back(), false}};
    while (!walk.empty()) {
      const auto [at, under_not] = walk.back();
      walk.pop_back();
      const Node& nd = nodes[at];
      if (nd.op == kOpTerm) {
        if (!under_not) {
          positive_terms->push_back(static_cast<uint32_t>(nd.arg));
        }
        continue;
      }
      for (size_t k = nd.kids.size(); k > 0; --k) {
        walk.push_back({nd.kids[k - 1], under_not || nd.op == kOpNot});
      }
    }
  }
  std::vector<size_t> todo{stack.back()};  // the value of the program is the top of the stack
  while (!todo.empty()) {
    const Node& nd = nodes[todo.back()];
    todo.pop_back();
    if (nd.op == kOpTerm) {
      conjuncts->push_back(static_cast<uint32_t>(nd.arg));
    } else if (nd.op == kOpAnd ||
               (nd.op == kOpAtLeast && !nd.kids.empty() && static_cast<size_t>(nd.arg >> 16) == nd.kids.size())) {
      todo.insert(todo.end(), nd.kids.begin(), nd.kids.end());  // "all n of n" is an AND
    }
  }
  return MGX_OK;
}

// Compiles queries [q_first, q_last) of the caller's flat description: unique terms of THAT range into `terms` (ids
// local to the range), the queries and the term id of every search-term slot into the caller-sized `queries` /
// `slot_tid`. Ranges are independent, so a large batch is compiled by several threads (compile_batch below).
int compile_range(const Index& ix, const mgx_query_params_t& p, uint64_t q_first, uint64_t q_last,
                  const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                  const uint8_t* not_bytes, const uint64_t* not_offsets, const uint64_t* q_not_begin,
                  const mgx_query_ext_t* ext, std::vector<HostTerm>* terms, std::vector<HostQuery>* queries,
                  std::vector<uint32_t>* slot_tid) {
  // open-addressing table of term ids keyed by the term bytes (no string is built for a lookup)
  const uint64_t n_search_slots = q_last > q_first ? q_term_begin[q_last] - q_term_begin[q_first] : 0;
  const uint64_t n_not_slots =
      (q_last > q_first && q_not_begin != nullptr) ? q_not_begin[q_last] - q_not_begin[q_first] : 0;
  size_t table_cap = 64;
  while (table_cap < 2 * (n_search_slots + n_not_slots) + 16) {
    table_cap <<= 1;
  }
  std::vector<uint32_t> table(table_cap, 0);  // term id + 1, 0 = empty
  terms->reserve(n_search_slots + n_not_slots + 16);
  // The streaming df pass (df_stream_kernel) counts "documents whose text contains the term". That equals the
  // reference's df (documents of SearchAnd(term n-grams) whose text contains the term) when every document is valid
  // UTF-8 and the query-side windows of a term are windows the index stores for any text containing it, i.e. both
  // sides cut with the same sizes and the index does not reject boundary windows the query side keeps.
  const int q_ascii = p.kanji_ngram_size > 0 ? (p.ngram_size > 0 ? p.ngram_size : 2) : (p.ngram_size == 0 ? 2 : p.ngram_size);
  const int q_kanji = p.kanji_ngram_size > 0 ? p.kanji_ngram_size : (p.ngram_size == 0 ? 1 : p.ngram_size);
  const bool q_cross = p.kanji_ngram_size > 0 ? p.cross_boundary != 0 : true;
  // tok_agree: every window the query side cuts from a term is a window the index stores for any text that contains
  // the term. Needed by both text-free shortcuts of the df stage (streaming pass, first-occurrence positions).
  const bool tok_agree = q_ascii == ix.ngram && q_kanji == ix.kanji && (ix.cross || !q_cross);
  const bool stream_ok = p.compute_score != 0 && ix.all_valid_utf8 && tok_agree;
  auto intern = [&](const uint8_t* bytes, uint64_t b, uint64_t e, uint32_t* out) -> int {
    if (e - b > kMaxTermBytes) {
      set_last_error("query term longer than 256 bytes is not supported");
      return MGX_ERR_UNSUPPORTED;
    }
    uint64_t h = 0xcbf29ce484222325ULL;  // FNV-1a
    for (uint64_t i = b; i < e; ++i) {
      h = (h ^ bytes[i]) * 0x100000001b3ULL;
    }
    h ^= h >> 29;
    size_t slot = static_cast<size_t>(h) & (table_cap - 1);
    while (table[slot] != 0) {
      const HostTerm& known = (*terms)[table[slot] - 1];
      if (known.bytes.size() == e - b && std::memcmp(known.bytes.data(), bytes + b, e - b) == 0) {
        *out = table[slot] - 1;
        return MGX_OK;
      }
      slot = (slot + 1) & (table_cap - 1);
    }
    const uint32_t id = static_cast<uint32_t>(terms->size());
    terms->emplace_back();
    HostTerm& t = terms->back();
    t.hash = h;
    t.bytes.assign(reinterpret_cast<const char*>(bytes) + b, e - b);
    bool valid_utf8 = false;
    host_query_keys(bytes + b, e - b, p.ngram_size, p.kanji_ngram_size, p.cross_boundary != 0, ix.width, &t.keys,
                    tok_agree ? &t.key_toff : nullptr, &valid_utf8);
    if (t.keys.size() == 1 && t.keys[0] != kInvalidKey) {
      uint8_t enc[4 * kMaxNgramSize];
      const int n = host_key_to_utf8(t.keys[0], ix.width, enc);
      t.exact_single = static_cast<uint64_t>(n) == e - b && std::memcmp(enc, bytes + b, e - b) == 0;
    }
    t.payload_tf = t.exact_single && valid_utf8 && tok_agree;
    // streaming df pass: valid UTF-8 of at least kStreamMinTermBytes bytes that needs a text check
    t.streamable = stream_ok && valid_utf8 && e - b >= kStreamMinTermBytes && !t.keys.empty() &&
                   (t.keys.size() > 1 || !t.exact_single) &&
                   std::find(t.keys.begin(), t.keys.end(), kInvalidKey) == t.keys.end();
    table[slot] = id + 1;
    *out = id;
    return MGX_OK;
  };
  for (uint64_t q = q_first; q < q_last; ++q) {
    HostQuery& hq = (*queries)[q];
    bool all_ascii = true;
    bool hybrid_exact = false;
    if (q_term_begin[q + 1] < q_term_begin[q]) {
      return invalid("q_term_begin must be non-decreasing");
    }
    for (uint64_t s = q_term_begin[q]; s < q_term_begin[q + 1]; ++s) {
      uint32_t tid = 0;
      const int rc = intern(term_bytes, term_offsets[s], term_offsets[s + 1], &tid);
      if (rc != MGX_OK) {
        return rc;
      }
      (*slot_tid)[s] = tid;
      hq.terms.push_back(tid);
      for (uint64_t i = term_offsets[s]; i < term_offsets[s + 1]; ++i) {
        all_ascii &= term_bytes[i] < 0x80;
      }
      hybrid_exact |= has_uncovered_hybrid_fragment(term_bytes + term_offsets[s], term_offsets[s + 1] - term_offsets[s],
                                                    p.ngram_size, p.kanji_ngram_size, p.cross_boundary != 0);
    }
    if (hq.terms.size() > 64) {
      set_last_error("more than 64 search terms in one query (reference limit, query_ast.h:184-185)");
      return MGX_ERR_UNSUPPORTED;
    }
    if (q_not_begin != nullptr) {
      for (uint64_t s = q_not_begin[q]; s < q_not_begin[q + 1]; ++s) {
        uint32_t tid = 0;
        const int rc = intern(not_bytes, not_offsets[s], not_offsets[s + 1], &tid);
        if (rc != MGX_OK) {
          return rc;
        }
        hq.not_terms.push_back(tid);
      }
    }
    // ShouldApplyVerifyText (search_pipeline.cpp:48-66) and the hybrid-fragment rule (:858-866)
    const bool verify = p.verify_text == 1 || (p.verify_text == 2 && all_ascii) || hybrid_exact;
    hq.flags = verify ? kQVerify : 0u;
    if (ext != nullptr && ext->q_prog_begin != nullptr && ext->q_prog_begin[q + 1] > ext->q_prog_begin[q]) {
      // boolean program over the query's own terms (QueryNode::Evaluate): TERM args are local term indices
      const uint64_t p0 = ext->q_prog_begin[q];
      const uint64_t pn = ext->q_prog_begin[q + 1] - p0;
      std::vector<uint32_t> local_conj;
      std::vector<uint32_t> local_scored;
      if (int rc = analyse_program(ext->prog_ops + p0, ext->prog_args + p0, pn, hq.terms.size(), &local_conj,
                                   p.compute_score != 0 ? &local_scored : nullptr);
          rc != MGX_OK) {
        return rc;
      }
      if (local_scored.size() > 64) {
        set_last_error("more than 64 scored terms in one boolean query");
        return MGX_ERR_UNSUPPORTED;
      }
      for (uint32_t c : local_conj) {
        hq.conjuncts.push_back(hq.terms[c]);
      }
      for (uint64_t i = 0; i < pn; ++i) {
        const int32_t op = ext->prog_ops[p0 + i];
        if (op == kOpFuzzyText) {
          return invalid("boolean program: FUZZYTEXT nodes are built by mgx_search_fuzzy only");
        }
        hq.prog_ops.push_back(static_cast<uint8_t>(op));
        hq.prog_args.push_back(op == kOpTerm ? hq.terms[static_cast<size_t>(ext->prog_args[p0 + i])]
                                             : static_cast<uint32_t>(ext->prog_args[p0 + i]));
      }
      // operands of the program, not AND-ed search terms. A scored program keeps the terms its results are scored
      // with (search_handler.cpp:405-470 scores every result shape with all_search_terms): membership is the
      // program's alone, the epilogue only counts these terms in the text of the survivors.
      TermIdVec scored;
      for (uint32_t c : local_scored) {
        scored.push_back(hq.terms[c]);
      }
      hq.terms = std::move(scored);
      hq.flags = kQProgram;
    }
    if (ext != nullptr && ext->q_filter_begin != nullptr) {
      for (uint64_t f = ext->q_filter_begin[q]; f < ext->q_filter_begin[q + 1]; ++f) {
        HostFilter hf;
        hf.col = ext->filter_col[f];
        hf.op = ext->filter_op[f];
        if (hf.op > 5) {
          return invalid("filter op must be 0..5 (EQ, NE, GT, GTE, LT, LTE)");
        }
        hf.literal.assign(reinterpret_cast<const char*>(ext->filter_bytes) + ext->filter_offsets[f],
                          ext->filter_offsets[f + 1] - ext->filter_offsets[f]);
        hq.filters.push_back(std::move(hf));
      }
    }
  }
  return MGX_OK;
}

unsigned compile_threads(uint64_t n_queries) {
  if (n_queries < 2048) {
    return 1;
  }
  // one thread compiles a 4096-query batch in about a millisecond; starting workers and merging their term tables
  // only pays for much larger batches (MGX_COMPILE_THREADS overrides)
  unsigned t = n_queries < 16384 ? 1u : std::max(1u, std::min(4u, std::thread::hardware_concurrency() / 8));
  if (const char* env = std::getenv("MGX_COMPILE_THREADS")) {
    t = static_cast<unsigned>(std::max(1, std::atoi(env)));
  }
  return std::min<unsigned>(t, 16);
}

// Compile the caller's flat query description into unique terms + queries. Large batches are cut into ranges of
// queries compiled concurrently (tokenising the terms is the expensive part); the ranges' term tables are then
// merged into one.
int compile_batch(const Index& ix, const mgx_query_params_t& p, uint64_t n_queries, const uint8_t* term_bytes,
                  const uint64_t* term_offsets, const uint64_t* q_term_begin, const uint8_t* not_bytes,
                  const uint64_t* not_offsets, const uint64_t* q_not_begin, const mgx_query_ext_t* ext,
                  std::vector<HostTerm>* terms, std::vector<HostQuery>* queries, std::vector<uint32_t>* slot_tid) {
  const uint64_t n_slots = n_queries > 0 ? q_term_begin[n_queries] : 0;
  slot_tid->assign(n_slots, 0);
  queries->clear();
  queries->resize(n_queries);
  const unsigned T = compile_threads(n_queries);
  if (T <= 1) {
    return compile_range(ix, p, 0, n_queries, term_bytes, term_offsets, q_term_begin, not_bytes, not_offsets,
                         q_not_begin, ext, terms, queries, slot_tid);
  }
  std::vector<std::vector<HostTerm>> part_terms(T);
  std::vector<WidePoolSnapshot> part_wide(T);  // wide keys: every worker interns into its own thread's pool
  std::vector<int> rcs(T, MGX_OK);
  std::vector<std::string> errs(T);
  std::vector<std::thread> workers;
  auto range_of = [&](unsigned t) { return std::make_pair(n_queries * t / T, n_queries * (t + 1) / T); };
  for (unsigned t = 0; t < T; ++t) {
    workers.emplace_back([&, t]() {
      const auto [q0, q1] = range_of(t);
      rcs[t] = compile_range(ix, p, q0, q1, term_bytes, term_offsets, q_term_begin, not_bytes, not_offsets, q_not_begin,
                             ext, &part_terms[t], queries, slot_tid);
      if (rcs[t] != MGX_OK) {
        errs[t] = mgx_last_error();  // thread-local in the worker
      }
      if (ix.wide_words > 0) {
        part_wide[t] = wide_pool_snapshot();
      }
    });
  }
  for (auto& w : workers) {
    w.join();
  }
  for (unsigned t = 0; t < T; ++t) {
    if (rcs[t] != MGX_OK) {
      set_last_error(errs[t]);
      return rcs[t];
    }
  }
  // merge: one id per distinct term over all ranges (the expensive terms are the frequent ones, which every range
  // meets: leaving them duplicated would repeat exactly the heaviest df work)
  size_t total_terms = 0;
  for (const auto& pt : part_terms) {
    total_terms += pt.size();
  }
  size_t cap = 64;
  while (cap < 2 * total_terms + 16) {
    cap <<= 1;
  }
  std::vector<uint32_t> table(cap, 0);  // batch term id + 1
  terms->clear();
  terms->reserve(total_terms);
  std::vector<uint32_t> remap;
  for (unsigned t = 0; t < T; ++t) {
    remap.assign(part_terms[t].size(), 0);
    const uint64_t wide_base = ix.wide_words > 0 ? wide_pool_adopt(part_wide[t]) : 0;
    for (size_t i = 0; i < part_terms[t].size(); ++i) {
      HostTerm& ht = part_terms[t][i];
      if (wide_base != 0) {  // handles of the worker's pool -> handles of this thread's pool
        for (uint64_t& k : ht.keys) {
          k = k == kInvalidKey ? k : k + wide_base;
        }
      }
      size_t slot = static_cast<size_t>(ht.hash) & (cap - 1);
      uint32_t id = 0;
      for (;;) {
        if (table[slot] == 0) {
          id = static_cast<uint32_t>(terms->size());
          table[slot] = id + 1;
          terms->push_back(std::move(ht));
          break;
        }
        const HostTerm& known = (*terms)[table[slot] - 1];
        if (known.hash == ht.hash && known.bytes == ht.bytes) {
          id = table[slot] - 1;
          break;
        }
        slot = (slot + 1) & (cap - 1);
      }
      remap[i] = id;
    }
    const auto [q0, q1] = range_of(t);
    for (uint64_t q = q0; q < q1; ++q) {  // range-local term ids -> batch ids
      HostQuery& hq = (*queries)[q];
      for (uint32_t& id : hq.terms) id = remap[id];
      for (uint32_t& id : hq.not_terms) id = remap[id];
      for (uint32_t& id : hq.conjuncts) id = remap[id];
      for (size_t i = 0; i < hq.prog_ops.size(); ++i) {
        if (hq.prog_ops[i] == kOpTerm) {
          hq.prog_args[i] = remap[hq.prog_args[i]];
        }
      }
    }
    for (uint64_t s2 = q_term_begin[q0]; s2 < q_term_begin[q1]; ++s2) {
      (*slot_tid)[s2] = remap[(*slot_tid)[s2]];
    }
  }
  return MGX_OK;
}

// Driver expansion of OR-rooted boolean programs. A program whose root is an OR has no term every result must
// satisfy, so the tile kernel would evaluate it on EVERY document of the shard (~6 ms per query at 10M documents).
// When every child of the root can drive (a TERM with n-grams, or an AND that holds one), the query becomes one
// internal query per child: child_i AND NOT child_1 ... AND NOT child_{i-1} -- each driven by child_i's shortest
// list, pairwise disjoint by construction, folded back on the device (fold_expanded_kernel). `xoff` stays empty when
// no query of the batch qualifies.
bool expand_or_roots(const std::vector<HostTerm>& terms, const std::vector<HostQuery>& queries,
                     std::vector<HostQuery>* expanded, std::vector<uint32_t>* xoff) {
  constexpr size_t kMaxChildren = 16;
  struct Plan {
    std::vector<std::pair<size_t, size_t>> child;  // op ranges [begin, end) of the root's children
    std::vector<std::vector<uint32_t>> conj;       // terms every result of a child satisfies
  };
  auto analyse = [&](const HostQuery& hq, Plan* plan) {
    const size_t n = hq.prog_ops.size();
    if ((hq.flags & kQProgram) == 0 || n < 3 || !hq.conjuncts.empty() || hq.prog_ops[n - 1] != kOpOr) {
      return false;
    }
    // subtree of op i = ops [start[i], i]; children of an n-ary op are the subtrees that end right before it
    std::vector<size_t> start(n, 0);
    std::vector<size_t> stack;
    size_t depth_max = 0;
    for (size_t i = 0; i < n; ++i) {
      const uint8_t op = hq.prog_ops[i];
      size_t kids = 0;
      if (op == kOpAnd || op == kOpOr) {
        kids = hq.prog_args[i];
      } else if (op == kOpAtLeast) {
        kids = hq.prog_args[i] & 0xFFFFu;
      } else if (op == kOpNot) {
        kids = 1;
      }
      if (kids > stack.size()) {
        return false;
      }
      start[i] = kids == 0 ? i : start[stack[stack.size() - kids]];
      stack.resize(stack.size() - kids);
      stack.push_back(i);
      depth_max = std::max(depth_max, stack.size());
    }
    if (stack.size() != 1) {
      return false;
    }
    const size_t k = hq.prog_args[n - 1];
    if (k < 2 || k > kMaxChildren || depth_max + k + 1 > kMaxProgramDepth) {
      return false;
    }
    size_t end = n - 1;  // children, right to left
    std::vector<std::pair<size_t, size_t>> rev;
    for (size_t c = 0; c < k; ++c) {
      if (end == 0) {
        return false;
      }
      const size_t root = end - 1;
      rev.emplace_back(start[root], root + 1);
      end = start[root];
    }
    plan->child.assign(rev.rbegin(), rev.rend());
    for (const auto& [b0, e0] : plan->child) {
      // conjunct terms of the child: itself when it is a TERM, the TERM operands of (nested) ANDs otherwise
      std::vector<uint32_t> conj;
      std::vector<size_t> todo{e0 - 1};
      while (!todo.empty()) {
        const size_t i = todo.back();
        todo.pop_back();
        const uint8_t op = hq.prog_ops[i];
        if (op == kOpTerm) {
          conj.push_back(hq.prog_args[i]);
        } else if (op == kOpAnd) {
          size_t e = i;
          for (uint32_t c = 0; c < hq.prog_args[i]; ++c) {
            todo.push_back(e - 1);
            e = start[e - 1];
          }
        }
      }
      bool drives = false;
      for (uint32_t tid : conj) {
        drives = drives || !terms[tid].keys.empty();
      }
      if (!drives) {
        return false;  // a child that is a text scan (or an OR / NOT itself) needs every document anyway
      }
      plan->conj.push_back(std::move(conj));
    }
    return true;
  };
  std::vector<Plan> plans(queries.size());
  std::vector<uint8_t> take(queries.size(), 0);
  bool any = false;
  for (size_t q = 0; q < queries.size(); ++q) {
    take[q] = analyse(queries[q], &plans[q]) ? 1 : 0;
    any = any || take[q] != 0;
  }
  if (!any) {
    return false;
  }
  expanded->clear();
  xoff->assign(1, 0);
  for (size_t q = 0; q < queries.size(); ++q) {
    const HostQuery& hq = queries[q];
    if (take[q] == 0) {
      expanded->push_back(hq);
    } else {
      const Plan& plan = plans[q];
      for (size_t i = 0; i < plan.child.size(); ++i) {
        HostQuery sub;
        sub.flags = hq.flags;
        sub.threshold = hq.threshold;
        sub.filters = hq.filters;
        sub.conjuncts = plan.conj[i];
        auto append = [&](const std::pair<size_t, size_t>& r) {
          sub.prog_ops.insert(sub.prog_ops.end(), hq.prog_ops.begin() + static_cast<ptrdiff_t>(r.first),
                              hq.prog_ops.begin() + static_cast<ptrdiff_t>(r.second));
          sub.prog_args.insert(sub.prog_args.end(), hq.prog_args.begin() + static_cast<ptrdiff_t>(r.first),
                               hq.prog_args.begin() + static_cast<ptrdiff_t>(r.second));
        };
        append(plan.child[i]);
        for (size_t j = 0; j < i; ++j) {
          append(plan.child[j]);
          sub.prog_ops.push_back(kOpNot);
          sub.prog_args.push_back(0);
        }
        if (i > 0) {
          sub.prog_ops.push_back(kOpAnd);
          sub.prog_args.push_back(static_cast<uint32_t>(i + 1));
        }
        expanded->push_back(std::move(sub));
      }
    }
    xoff->push_back(static_cast<uint32_t>(expanded->size()));
  }
  return true;
}

int check_params(const mgx_query_params_t& p) {
  // SortByScore takes any offset and limit (limit 0 = everything from offset, result_sorter.cpp:689-710); the only
  // bound here is that offset + limit is computed in 32 bits by the shard / merge stages
  if (static_cast<uint64_t>(p.limit) + p.offset > 0xFFFFFFFFULL) {
    return invalid("limit + offset must fit 32 bits");
  }
  return MGX_OK;
}

}  // namespace
}  // namespace mgx

using namespace mgx;

// Readers / writer gate of one index handle (the reference's table generation lock, server_types.h:245: shared for a
// whole request, exclusive for a rebuild). Not std::shared_mutex: a staged batch is prepared by one thread and
// destroyed by another, which a std::shared_mutex does not allow. Writers have preference.
struct RwGate {
  std::mutex m;
  std::condition_variable cv;
  int readers = 0;
  int writers_waiting = 0;
  bool writer = false;
  void lock_shared() {
    std::unique_lock<std::mutex> l(m);
    cv.wait(l, [&] { return !writer && writers_waiting == 0; });
    ++readers;
  }
  void unlock_shared() {
    std::unique_lock<std::mutex> l(m);
    if (--readers == 0) {
      cv.notify_all();
    }
  }
  void lock() {
    std::unique_lock<std::mutex> l(m);
    ++writers_waiting;
    cv.wait(l, [&] { return !writer && readers == 0; });
    --writers_waiting;
    writer = true;
  }
  void unlock() {
    std::unique_lock<std::mutex> l(m);
    writer = false;
    cv.notify_all();
  }
};

constexpr int kMaxLanes = 8;

struct mgx_index {
  Index ix;
  RwGate gate;  // build / column upload / the generation swap of a commit exclusive; every reading call shared
  // Commits of journaled mutations build the next generation of the shard BESIDE the current one (into `next`, while
  // reading calls go on under shared access) and make it current in an exclusive section that only exchanges the
  // two (swap_generation). writer_mu keeps the writers (build, commit, column upload, trim, clear, stream load) from
  // overlapping each other for their whole duration; commit_mu is held by the one commit in progress.
  Index next;
  std::mutex writer_mu;
  std::mutex commit_mu;
  cudaStream_t commit_stream = nullptr;
  std::atomic<int> commit_mode{0};  // mgx_index_set_commit_mode: 0 = a read waits for a commit in progress, 1 = it does not
  std::atomic<bool> committing{false};  // a commit has taken its journal and not finished yet
  // Execution lanes of the single-call readers: a stream with its own search workspace each, so that concurrent
  // Index::Search* style calls from the server's worker threads (thread_pool.cpp:33) run side by side on the device
  // instead of queueing behind one mutex. Lane 0 is ix.stream.
  std::mutex lane_mu;
  std::condition_variable lane_cv;
  cudaStream_t lane_stream[kMaxLanes] = {nullptr};
  bool lane_busy[kMaxLanes] = {false};
  int n_lanes = 0;
  std::mutex stats_mu;  // ix.last_stats
  // Journal of Index::AddDocument / AddDocumentBatch / UpdateDocument / RemoveDocument calls not folded in yet: an
  // append-only log in arrival order (ids, removed flags, text ranges in one arena -- a 1000-document loader batch is
  // three memcpys, not a thousand map nodes). The last call for an id wins; the next read applies the whole journal
  // (apply_journal_device).
  struct Journal {
    std::vector<uint32_t> ids;
    std::vector<uint8_t> removed;
    std::vector<uint64_t> off = std::vector<uint64_t>(1, 0);
    std::vector<uint8_t> text;
    bool ascending = true;  // ids strictly ascending in arrival order: neither a sort nor a de-duplication is needed
    bool empty() const { return ids.empty(); }
    void put(uint32_t id, bool rem, const uint8_t* t, uint64_t len) {
      if (!ids.empty() && id <= ids.back()) {
        ascending = false;
      }
      ids.push_back(id);
      removed.push_back(rem ? 1 : 0);
      if (!rem && len > 0) {
        text.insert(text.end(), t, t + len);
      }
      off.push_back(text.size());
    }
    void clear() {
      ids.clear();
      removed.clear();
      off.assign(1, 0);
      text.clear();
      ascending = true;
    }
  } journal;
  std::mutex journal_mu;
  std::atomic<bool> dirty{false};
};

namespace {
struct WriteGuard {
  mgx_index* h;
  explicit WriteGuard(mgx_index* index) : h(index) {
    h->writer_mu.lock();
    h->gate.lock();
  }
  ~WriteGuard() {
    h->gate.unlock();
    h->writer_mu.unlock();
  }
  WriteGuard(const WriteGuard&) = delete;
  WriteGuard& operator=(const WriteGuard&) = delete;
};

// Shared access to the index plus the exclusive use of one execution lane for the duration of a reading call.
struct Reader {
  mgx_index* h;
  int lane = -1;
  explicit Reader(mgx_index* index) : h(index) {
    h->gate.lock_shared();
    std::unique_lock<std::mutex> l(h->lane_mu);
    for (;;) {
      for (int i = 0; i < h->n_lanes; ++i) {
        if (!h->lane_busy[i]) {
          lane = i;
          break;
        }
      }
      if (lane >= 0) {
        break;
      }
      if (h->n_lanes < kMaxLanes) {
        lane = h->n_lanes;
        if (lane == 0) {
          h->lane_stream[0] = h->ix.stream;
        } else {
          DeviceGuard guard(h->ix.device);
          if (cudaStreamCreateWithFlags(&h->lane_stream[lane], cudaStreamNonBlocking) != cudaSuccess) {
            (void)cudaGetLastError();
            lane = -1;
            if (h->n_lanes == 0) {
              h->lane_stream[0] = h->ix.stream;
            }
            h->lane_cv.wait(l);  // no further stream to be had: wait for a lane like everybody else
            continue;
          }
        }
        ++h->n_lanes;
        break;
      }
      h->lane_cv.wait(l);
    }
    h->lane_busy[lane] = true;
  }
  ~Reader() {
    {
      std::unique_lock<std::mutex> l(h->lane_mu);
      h->lane_busy[lane] = false;
    }
    h->lane_cv.notify_one();
    h->gate.unlock_shared();
  }
  Reader(const Reader&) = delete;
  Reader& operator=(const Reader&) = delete;
  cudaStream_t stream() const { return h->lane_stream[lane]; }
};

// Applies the pending mutations; called at the top of every entry point that reads the index, BEFORE the call
// takes its shared access (never under it: the commit needs the gate exclusively).
int commit_pending(mgx_index_t* index, bool explicit_commit = false) {
  if (index == nullptr ||
      (!index->dirty.load(std::memory_order_acquire) && !index->committing.load(std::memory_order_acquire))) {
    return MGX_OK;
  }
  // Overlapped mode: reading calls neither commit nor wait for a commit -- they answer from the current generation,
  // and mutations become visible when the mgx_index_commit of the mutating thread returns. (A reading thread may hold
  // staged batches; a commit of its own would wait for them at the exchange.)
  if (!explicit_commit && index->commit_mode.load(std::memory_order_relaxed) != 0) {
    return MGX_OK;
  }
  return guarded([&]() {
    // One commit at a time; whoever finds one in progress waits for it (it may hold mutations that returned before
    // this call began).
    std::unique_lock<std::mutex> cl(index->commit_mu);
    // The journal as it stands now is this commit's; calls that arrive from here on start the next one. `committing`
    // is raised before `dirty` falls, so a read that arrives in between still waits for this commit.
    struct Committing {
      mgx_index* h;
      explicit Committing(mgx_index* i) : h(i) { h->committing.store(true, std::memory_order_release); }
      ~Committing() { h->committing.store(false, std::memory_order_release); }
    } committing{index};
    mgx_index::Journal j;
    {
      std::lock_guard<std::mutex> jl(index->journal_mu);
      if (index->journal.empty()) {
        index->dirty.store(false);
        return MGX_OK;
      }
      std::swap(j, index->journal);
      index->journal.clear();
      index->dirty.store(false, std::memory_order_release);
    }
    // A failed commit leaves the CURRENT generation as it was (the next one is built beside it) and the journal is
    // discarded: retrying the same journal on every later read would fail the same way. The caller learns it from this
    // call's status.
    j.text.push_back(0);
    std::vector<uint32_t> ids;
    std::vector<uint8_t> removed;
    std::vector<uint8_t> text;
    std::vector<uint64_t> off(1, 0);
    if (!j.ascending) {
      // arrival order -> ascending ids, the LAST entry of an id wins (stable sort keeps arrival order among equals)
      const size_t n = j.ids.size();
      std::vector<uint32_t> order(n);
      for (size_t i = 0; i < n; ++i) {
        order[i] = static_cast<uint32_t>(i);
      }
      std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return j.ids[a] < j.ids[b]; });
      ids.reserve(n);
      removed.reserve(n);
      off.reserve(n + 1);
      text.reserve(j.text.size());
      for (size_t k = 0; k < n; ++k) {
        if (k + 1 < n && j.ids[order[k + 1]] == j.ids[order[k]]) {
          continue;  // superseded by a later call for the same id
        }
        const uint32_t e = order[k];
        ids.push_back(j.ids[e]);
        removed.push_back(j.removed[e]);
        text.insert(text.end(), j.text.begin() + static_cast<ptrdiff_t>(j.off[e]),
                    j.text.begin() + static_cast<ptrdiff_t>(j.off[e + 1]));
        off.push_back(text.size());
      }
      text.push_back(0);
    }
    const uint32_t* p_ids = j.ascending ? j.ids.data() : ids.data();
    const uint8_t* p_removed = j.ascending ? j.removed.data() : removed.data();
    const uint8_t* p_text = j.ascending ? j.text.data() : text.data();
    const uint64_t* p_off = j.ascending ? j.off.data() : off.data();
    const uint64_t n_j = j.ascending ? j.ids.size() : ids.size();

    std::lock_guard<std::mutex> wl(index->writer_mu);  // no build / column upload / trim while the next generation forms
    Index& ix = index->ix;
    Index& next = index->next;
    DeviceGuard guard(ix.device);
    if (index->commit_stream == nullptr) {
      MGX_CUDA(cudaStreamCreateWithFlags(&index->commit_stream, cudaStreamNonBlocking));
    }
    // the build workspaces (pairs, sort scratch) belong to whichever generation is being built
    struct LendArenas {
      Index& a;
      Index& b;
      LendArenas(Index& from, Index& to) : a(from), b(to) {
        std::swap(a.build_arena0, b.build_arena0);
        std::swap(a.build_arena, b.build_arena);
      }
      ~LendArenas() {
        std::swap(a.build_arena0, b.build_arena0);
        std::swap(a.build_arena, b.build_arena);
      }
    };
    {
      index->gate.lock_shared();  // the current generation is read (its corpus, its columns), never written
      struct Unshare {
        mgx_index* h;
        ~Unshare() { h->gate.unlock_shared(); }
      } unshare{index};
      LendArenas lend(ix, next);
      copy_index_config(next, ix);
      apply_journal_device(ix, next, p_ids, p_removed, p_text, p_off, n_j, index->commit_stream);
      MGX_CUDA(cudaStreamSynchronize(index->commit_stream));
    }
    {
      index->gate.lock();  // exclusive only for the exchange
      swap_generation(ix, next);
      index->gate.unlock();
    }
    // `next` holds the previous generation now. Its arrays are released unless MGX_COMMIT_KEEP_SPARE asks to keep
    // them for the next commit (no cudaMalloc / cudaFree per commit, at twice the resident memory).
    if (std::getenv("MGX_COMMIT_KEEP_SPARE") == nullptr) {
      next.drop_filter_columns();
      next.d_bitmaps.release();
      next.resident_a.release();
      next.resident_b.release();
      next.n_docs = next.n_terms = next.n_postings = 0;
    }
    return MGX_OK;
  });
}

int journal_put(mgx_index_t* index, uint32_t doc_id, bool removed, const uint8_t* text, uint64_t len) {
  if (index == nullptr || (len > 0 && text == nullptr)) {
    return invalid("null argument");
  }
  if (int rc = require_device(); rc != MGX_OK) {
    return rc;
  }
  if (index->ix.text_less) {
    set_last_error("the index was loaded from an MGIX stream and holds no document text: single-document mutations "
                   "need the documents (mgx_index_build)");
    return MGX_ERR_UNSUPPORTED;
  }
  std::lock_guard<std::mutex> jl(index->journal_mu);
  index->journal.put(doc_id, removed, text, len);
  index->dirty.store(true, std::memory_order_release);
  return MGX_OK;
}
}  // namespace

struct mgx_batch {
  Batch b;
  mgx_index* reader_of = nullptr;  // staged batches keep shared access to their index until mgx_batch_destroy
};

extern "C" {

const char* mgx_last_error(void) { return g_last_error.c_str(); }
const char* mgx_version(void) { return "mgx 0.1 (sm_100a, CUDA kernels only, no CPU fallback)"; }
uint64_t mgx_kernel_launch_count(void) { return g_launches.load(); }

int mgx_index_create(const mgx_index_config_t* config, mgx_index_t** out) {
  if (config == nullptr || out == nullptr) {
    return invalid("config/out is null");
  }
  *out = nullptr;
  if (int rc = require_device(); rc != MGX_OK) {
    return rc;
  }
  const int kanji = config->kanji_ngram_size > 0 ? config->kanji_ngram_size : config->ngram_size;  // index.cpp:32
  if (config->ngram_size < 1 || config->ngram_size > kMaxNgramSize || kanji < 1 || kanji > kMaxNgramSize) {
    set_last_error("n-gram sizes must be 1..10 (config-schema.json:279-285)");
    return MGX_ERR_INVALID_ARGUMENT;
  }
  return guarded([&]() {
    int count = 0;
    MGX_CUDA(cudaGetDeviceCount(&count));
    if (config->device < 0 || config->device >= count) {
      return invalid("device ordinal out of range");
    }
    auto h = std::make_unique<mgx_index>();
    h->ix.cfg = *config;
    h->ix.ngram = config->ngram_size;
    h->ix.kanji = kanji;
    h->ix.cross = config->cross_boundary_ngrams != 0;
    h->ix.width = std::max(h->ix.ngram, h->ix.kanji);
    h->ix.wide_words = wide_words_for(h->ix.width);
    h->ix.device = config->device;
    DeviceGuard guard(config->device);
    MGX_CUDA(cudaStreamCreateWithFlags(&h->ix.stream, cudaStreamNonBlocking));
    *out = h.release();
    return MGX_OK;
  });
}

void mgx_index_destroy(mgx_index_t* index) {
  if (index == nullptr) {
    return;
  }
  {
    DeviceGuard guard(index->ix.device);
    for (int i = 1; i < index->n_lanes; ++i) {
      if (index->lane_stream[i] != nullptr) {
        cudaStreamSynchronize(index->lane_stream[i]);
        cudaStreamDestroy(index->lane_stream[i]);
      }
    }
    if (index->commit_stream != nullptr) {
      cudaStreamSynchronize(index->commit_stream);
      cudaStreamDestroy(index->commit_stream);
    }
    if (index->ix.stream != nullptr) {
      cudaStreamSynchronize(index->ix.stream);
      cudaStreamDestroy(index->ix.stream);
    }
    for (void* pooled : index->ix.batch_pool) {
      delete static_cast<mgx_batch*>(pooled);
    }
    index->ix.batch_pool.clear();
    delete index;
  }
}

static int build_common(mgx_index_t* index, const uint32_t* doc_ids, const uint8_t* text, const uint64_t* text_offsets,
                        uint64_t n_docs, bool device_inputs) {
  if (index == nullptr || (n_docs > 0 && (doc_ids == nullptr || text_offsets == nullptr))) {
    return invalid("null argument");
  }
  if (n_docs >= (1ULL << 32) - 1) {
    return invalid("too many documents in one shard");
  }
  {
    std::lock_guard<std::mutex> jl(index->journal_mu);  // a bulk build replaces everything, pending mutations included
    index->journal.clear();
    index->dirty.store(false);
  }
  return guarded([&]() {
    WriteGuard lock(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    uint64_t text_bytes = 0;
    uint64_t first_off = 0;
    static const uint64_t kZeroOff[1] = {0};
    const uint64_t* offs = n_docs > 0 ? text_offsets : kZeroOff;
    if (device_inputs && n_docs > 0) {
      MGX_CUDA(cudaMemcpy(&first_off, text_offsets, sizeof(uint64_t), cudaMemcpyDeviceToHost));
      MGX_CUDA(cudaMemcpy(&text_bytes, text_offsets + n_docs, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    } else if (n_docs > 0) {
      first_off = text_offsets[0];
      text_bytes = text_offsets[n_docs];
      for (uint64_t i = 1; i < n_docs; ++i) {
        if (doc_ids[i] <= doc_ids[i - 1]) {
          return invalid("doc_ids must be strictly ascending");
        }
      }
      for (uint64_t i = 0; i < n_docs; ++i) {
        if (text_offsets[i + 1] < text_offsets[i]) {
          return invalid("text_offsets must be non-decreasing");
        }
      }
    }
    if (first_off != 0) {
      return invalid("text_offsets[0] must be 0");
    }
    static const uint8_t kNoText[1] = {0};
    build_index_device(ix, doc_ids, text != nullptr ? text : kNoText, offs, n_docs, text_bytes, ix.stream);
    return MGX_OK;
  });
}

int mgx_index_build(mgx_index_t* index, const uint32_t* doc_ids, const uint8_t* text, const uint64_t* text_offsets,
                    uint64_t n_docs) {
  return build_common(index, doc_ids, text, text_offsets, n_docs, false);
}

int mgx_index_build_device(mgx_index_t* index, const uint32_t* d_doc_ids, const uint8_t* d_text,
                           const uint64_t* d_text_offsets, uint64_t n_docs) {
  return build_common(index, d_doc_ids, d_text, d_text_offsets, n_docs, true);
}

int mgx_index_add_document(mgx_index_t* index, uint32_t doc_id, const uint8_t* text, uint64_t text_len,
                           int32_t* out_indexed) {
  if (out_indexed != nullptr && index != nullptr) {
    // Index::AddDocument returns false when the text yields no n-gram (index.cpp:49-57)
    KeyVec keys;
    host_query_keys(text, text_len, index->ix.ngram, index->ix.kanji, index->ix.cross, index->ix.width, &keys);
    *out_indexed = keys.empty() ? 0 : 1;
  }
  return journal_put(index, doc_id, false, text, text_len);
}

int mgx_index_add_document_batch(mgx_index_t* index, const uint32_t* doc_ids, const uint8_t* text,
                                 const uint64_t* text_offsets, uint64_t n_docs, uint64_t* out_indexed) {
  if (index == nullptr || (n_docs > 0 && (doc_ids == nullptr || text_offsets == nullptr))) {
    return invalid("null argument");
  }
  if (int rc = require_device(); rc != MGX_OK) {
    return rc;
  }
  for (uint64_t i = 0; i < n_docs; ++i) {
    if (text_offsets[i + 1] < text_offsets[i] || (text_offsets[i + 1] > text_offsets[i] && text == nullptr)) {
      return invalid("text_offsets must be non-decreasing (and text non-null)");
    }
  }
  if (index->ix.text_less && n_docs > 0) {
    set_last_error("the index was loaded from an MGIX stream and holds no document text: mutations need the "
                   "documents (mgx_index_build)");
    return MGX_ERR_UNSUPPORTED;
  }
  uint64_t indexed = 0;
  if (out_indexed != nullptr) {
    // Index::AddDocumentBatch silently skips a document whose text yields no n-gram (index.cpp:93-101); the count of
    // the others is what a caller can compare with the reference's log line
    for (uint64_t i = 0; i < n_docs; ++i) {
      KeyVec keys;
      host_query_keys(text + text_offsets[i], text_offsets[i + 1] - text_offsets[i], index->ix.ngram, index->ix.kanji,
                      index->ix.cross, index->ix.width, &keys);
      indexed += keys.empty() ? 0 : 1;
    }
    *out_indexed = indexed;
  }
  std::lock_guard<std::mutex> jl(index->journal_mu);
  mgx_index::Journal& j = index->journal;
  j.ids.reserve(j.ids.size() + n_docs);
  for (uint64_t i = 0; i < n_docs; ++i) {
    j.put(doc_ids[i], false, text + text_offsets[i], text_offsets[i + 1] - text_offsets[i]);
  }
  if (n_docs > 0) {
    index->dirty.store(true, std::memory_order_release);
  }
  return MGX_OK;
}

int mgx_index_update_document(mgx_index_t* index, uint32_t doc_id, const uint8_t* old_text, uint64_t old_len,
                              const uint8_t* new_text, uint64_t new_len) {
  (void)old_text;  // the resident text of doc_id is what gets replaced (consistent use: it equals old_text)
  (void)old_len;
  return journal_put(index, doc_id, false, new_text, new_len);
}

int mgx_index_remove_document(mgx_index_t* index, uint32_t doc_id, const uint8_t* text, uint64_t text_len) {
  (void)text;
  (void)text_len;
  return journal_put(index, doc_id, true, nullptr, 0);
}

int mgx_index_set_commit_mode(mgx_index_t* index, int overlapped) {
  if (index == nullptr) {
    return invalid("null argument");
  }
  index->commit_mode.store(overlapped != 0 ? 1 : 0, std::memory_order_relaxed);
  return MGX_OK;
}

int mgx_index_commit(mgx_index_t* index) {
  if (index == nullptr) {
    return invalid("null argument");
  }
  return commit_pending(index, true);
}

int mgx_index_get_stats(const mgx_index_t* index, mgx_index_stats_t* out) {
  if (index == nullptr || out == nullptr) {
    return invalid("null argument");
  }
  if (int rc = commit_pending(const_cast<mgx_index_t*>(index)); rc != MGX_OK) {
    return rc;
  }
  Reader rd(const_cast<mgx_index_t*>(index));
  const Index& ix = index->ix;
  out->n_docs = ix.n_docs;
  out->n_terms = ix.n_terms;
  out->n_postings = ix.n_postings;
  out->n_dense_terms = ix.n_dense;
  out->text_bytes = ix.text_bytes;
  out->total_doc_length = ix.total_doc_length;
  out->doc_count = ix.doc_count;
  out->device_bytes = ix.device_bytes();
  out->n_pair_slots = ix.n_pair_slots;
  out->all_valid_utf8 = ix.all_valid_utf8 ? 1 : 0;
  out->key_width = ix.width;
  out->last_build_ms = ix.last_build_ms;
  return MGX_OK;
}

int mgx_index_set_filter_column(mgx_index_t* index, uint32_t column, int32_t type, const uint64_t* values,
                                const uint8_t* nulls, uint64_t n_docs, const uint8_t* str_bytes,
                                const uint64_t* str_offsets, uint64_t n_strings) {
  if (index == nullptr || (n_docs > 0 && values == nullptr)) {
    return invalid("null argument");
  }
  if (column >= kMaxFilterColumns) {
    return invalid("filter column id must be < 64");
  }
  FilterClass cls = kFcNone;
  switch (type) {
    case 1: cls = kFcBool; break;
    case 2: case 4: case 6: case 8: case 10: cls = kFcSigned; break;   // int8..int64, TIME seconds
    case 3: case 5: case 7: case 9: cls = kFcUnsigned; break;
    case 11: cls = kFcString; break;
    case 12: cls = kFcDouble; break;
    default: return invalid("filter column type must be a FilterValue variant index 1..12");
  }
  if (cls == kFcString && n_docs > 0 && (str_bytes == nullptr || str_offsets == nullptr)) {
    return invalid("string column without a string table");
  }
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    WriteGuard lock(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    if (n_docs != ix.n_docs) {
      return invalid("a filter column needs one value per document of the index");
    }
    auto col = std::make_unique<FilterColumn>();
    col->cls = cls;
    col->n_docs = n_docs;
    std::vector<uint64_t> ranks;
    const uint64_t* src = values;
    if (cls == kFcString) {
      // dictionary-encode: the rank among the distinct strings keeps std::string's order (bytewise)
      std::vector<std::string> strs(n_strings);
      for (uint64_t i = 0; i < n_strings; ++i) {
        strs[i].assign(reinterpret_cast<const char*>(str_bytes) + str_offsets[i], str_offsets[i + 1] - str_offsets[i]);
      }
      col->dict = strs;
      std::sort(col->dict.begin(), col->dict.end());
      col->dict.erase(std::unique(col->dict.begin(), col->dict.end()), col->dict.end());
      std::vector<uint64_t> rank_of(n_strings);
      for (uint64_t i = 0; i < n_strings; ++i) {
        rank_of[i] = static_cast<uint64_t>(std::lower_bound(col->dict.begin(), col->dict.end(), strs[i]) - col->dict.begin());
      }
      ranks.resize(n_docs);
      for (uint64_t i = 0; i < n_docs; ++i) {
        const bool is_null = nulls != nullptr && nulls[i] != 0;
        if (!is_null && values[i] >= n_strings) {
          return invalid("string column: value index outside the string table");
        }
        ranks[i] = is_null ? 0 : rank_of[values[i]];
      }
      src = ranks.data();
    }
    col->values.alloc(n_docs);
    col->nulls.alloc(n_docs);
    if (n_docs > 0) {
      MGX_CUDA(cudaMemcpyAsync(col->values.p, src, n_docs * sizeof(uint64_t), cudaMemcpyHostToDevice, ix.stream));
      if (nulls != nullptr) {
        MGX_CUDA(cudaMemcpyAsync(col->nulls.p, nulls, n_docs, cudaMemcpyHostToDevice, ix.stream));
      } else {
        MGX_CUDA(cudaMemsetAsync(col->nulls.p, 0, n_docs, ix.stream));
      }
      MGX_CUDA(cudaStreamSynchronize(ix.stream));
    }
    delete ix.columns[column];
    ix.columns[column] = col.release();
    return MGX_OK;
  });
}

int mgx_index_get_statistics(const mgx_index_t* index_c, mgx_index_statistics_t* out) {
  mgx_index_t* index = const_cast<mgx_index_t*>(index_c);
  if (index == nullptr || out == nullptr) {
    return invalid("null argument");
  }
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    const double thr = ix.cfg.roaring_threshold > 0.0 ? ix.cfg.roaring_threshold : 0.18;
    const uint64_t roaring = count_roaring_lists(ix, thr, ix.optimized_total_docs, rd.stream());
    out->total_terms = ix.n_terms;
    out->total_postings = ix.n_postings;
    out->roaring_bitmap_lists = roaring;
    out->delta_encoded_lists = ix.n_terms - roaring;
    out->memory_usage_bytes = ix.device_bytes();
    return MGX_OK;
  });
}

int mgx_index_trim(mgx_index_t* index) {
  if (index == nullptr) {
    return invalid("null argument");
  }
  return guarded([&]() {
    WriteGuard lock(index);
    DeviceGuard guard(index->ix.device);
    index->ix.build_arena.release();
    index->ix.build_arena0.release();
    return MGX_OK;
  });
}

int mgx_index_optimize(mgx_index_t* index, uint64_t total_docs) {
  if (index == nullptr) {
    return invalid("null argument");
  }
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  // Index::Optimize re-encodes each list by density (posting_list.cpp:799-834). The device representation (sorted
  // uint32 + bitmaps chosen at build time for probing speed) does not change; the call records total_docs so that
  // the representation counters follow the reference's rule. total_docs == 0 is a no-op there too (:801-803).
  if (total_docs != 0) {
    WriteGuard lock(index);
    index->ix.optimized_total_docs = total_docs;
  }
  return MGX_OK;
}

int mgx_index_clear(mgx_index_t* index) {
  if (index == nullptr) {
    return invalid("null argument");
  }
  static const uint64_t kZeroOff[1] = {0};
  static const uint32_t kNoIds[1] = {0};
  static const uint8_t kNoText[1] = {0};
  const int rc = mgx_index_build(index, kNoIds, kNoText, kZeroOff, 0);  // also drops pending mutations
  if (rc == MGX_OK) {
    WriteGuard lock(index);
    index->ix.optimized_total_docs = 0;
  }
  return rc;
}

int mgx_key_to_utf8(uint64_t key, int32_t width, uint8_t* out) {
  if (out == nullptr || width < 1 || width > kMaxKeyWidth) {
    return invalid("bad width/out");
  }
  int n = 0;
  for (int j = width - 1; j >= 0; --j) {
    const uint64_t f = (key >> (21 * j)) & 0x1FFFFFULL;
    if (f == 0) {
      continue;
    }
    const uint32_t cp = static_cast<uint32_t>(f - 1);
    if (cp <= 0x7F) {
      out[n++] = static_cast<uint8_t>(cp);
    } else if (cp <= 0x7FF) {
      out[n++] = static_cast<uint8_t>(0xC0 | (cp >> 6));
      out[n++] = static_cast<uint8_t>(0x80 | (cp & 0x3F));
    } else if (cp <= 0xFFFF) {
      out[n++] = static_cast<uint8_t>(0xE0 | (cp >> 12));
      out[n++] = static_cast<uint8_t>(0x80 | ((cp >> 6) & 0x3F));
      out[n++] = static_cast<uint8_t>(0x80 | (cp & 0x3F));
    } else {
      out[n++] = static_cast<uint8_t>(0xF0 | (cp >> 18));
      out[n++] = static_cast<uint8_t>(0x80 | ((cp >> 12) & 0x3F));
      out[n++] = static_cast<uint8_t>(0x80 | ((cp >> 6) & 0x3F));
      out[n++] = static_cast<uint8_t>(0x80 | (cp & 0x3F));
    }
  }
  return n;
}

int mgx_key_words(int32_t width) {
  if (width < 1 || width > kMaxNgramSize) {
    return invalid("key width must be 1..10");
  }
  return std::max(wide_words_for(width), 1);
}

int mgx_wide_key_to_utf8(const uint64_t* words, int32_t width, uint8_t* out) {
  if (words == nullptr || out == nullptr || width < 1 || width > kMaxNgramSize) {
    return invalid("bad width/words/out");
  }
  const int W = wide_words_for(width);
  return W > 0 ? wide_words_to_utf8(words, W, out) : mgx_key_to_utf8(words[0], width, out);
}

int mgx_tokenize_batch(const mgx_index_config_t* config, const uint8_t* text, const uint64_t* text_offsets,
                       uint64_t n_docs, uint64_t* out_keys, uint32_t* out_doc, uint64_t cap, uint64_t* out_count) {
  if (config == nullptr || out_count == nullptr || (n_docs > 0 && text_offsets == nullptr)) {
    return invalid("null argument");
  }
  if (int rc = require_device(); rc != MGX_OK) {
    return rc;
  }
  const int kanji = config->kanji_ngram_size > 0 ? config->kanji_ngram_size : config->ngram_size;
  if (config->ngram_size < 1 || config->ngram_size > kMaxNgramSize || kanji < 1 || kanji > kMaxNgramSize) {
    set_last_error("n-gram sizes must be 1..10 (config-schema.json:279-285)");
    return MGX_ERR_INVALID_ARGUMENT;
  }
  return guarded([&]() {
    DeviceGuard guard(config->device);
    *out_count = 0;
    if (n_docs == 0) {
      return MGX_OK;
    }
    cudaStream_t st = nullptr;
    MGX_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const uint64_t bytes = text_offsets[n_docs];
    DevBuf<uint8_t> d_text;
    DevBuf<uint64_t> d_off;
    d_text.alloc(bytes + 64);
    d_off.alloc(n_docs + 1);
    MGX_CUDA(cudaMemsetAsync(d_text.p, 0, bytes + 64, st));
    if (bytes > 0) {
      MGX_CUDA(cudaMemcpyAsync(d_text.p, text, bytes, cudaMemcpyHostToDevice, st));
    }
    MGX_CUDA(cudaMemcpyAsync(d_off.p, text_offsets, (n_docs + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const int width = std::max(config->ngram_size, kanji);
    const int W = wide_words_for(width);
    const uint64_t nw = static_cast<uint64_t>(std::max(W, 1));
    DevBuf<uint32_t> d_len;
    DevBuf<uint64_t> d_tile_off;
    DevBuf<uint64_t> d_scratch;
    DevBuf<uint64_t> d_keys;
    DevBuf<uint32_t> d_docs;
    d_len.alloc(n_docs);
    d_tile_off.alloc(tokenize_tile_count(n_docs, bytes) + 1);
    d_scratch.alloc(tokenize_scratch_elems(n_docs, bytes));
    uint64_t n_slots = 0;
    uint64_t counters[3];
    tokenize_count(config->ngram_size, kanji, config->cross_boundary_ngrams != 0, width, d_text.p, d_off.p, n_docs,
                   bytes, d_len.p, d_tile_off.p, d_scratch.p, &n_slots, counters, st);
    d_keys.alloc(n_slots * nw);
    d_docs.alloc(n_slots);
    if (n_slots > 0) {
      tokenize_emit(config->ngram_size, kanji, config->cross_boundary_ngrams != 0, width, d_text.p, d_off.p, n_docs,
                    bytes, d_tile_off.p, d_scratch.p, d_keys.p, d_docs.p, 0, st, n_slots);
    }
    std::vector<uint64_t> keys(n_slots * nw);  // wide keys: word w of slot s at keys[w * n_slots + s]
    std::vector<uint32_t> docs(n_slots);
    if (n_slots > 0) {
      MGX_CUDA(cudaMemcpyAsync(keys.data(), d_keys.p, n_slots * nw * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
      MGX_CUDA(cudaMemcpyAsync(docs.data(), d_docs.p, n_slots * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    MGX_CUDA(cudaStreamSynchronize(st));
    cudaStreamDestroy(st);
    uint64_t n = 0;
    for (uint64_t i = 0; i < n_slots; ++i) {
      if (keys[i] != kInvalidKey) {
        if (n < cap && out_keys != nullptr && out_doc != nullptr) {
          for (uint64_t w = 0; w < nw; ++w) {
            out_keys[n * nw + w] = keys[w * n_slots + i];
          }
          out_doc[n] = docs[i];
        }
        ++n;
      }
    }
    *out_count = n;
    if (n > cap) {
      set_last_error("mgx_tokenize_batch: output capacity too small");
      return MGX_ERR_CAPACITY;
    }
    return MGX_OK;
  });
}

// ---------------------------------------------------------------- set-algebra single calls
namespace {
bool run_at_least_expanded(Index& ix, cudaStream_t stream, const KeyVec& keys, size_t need, uint32_t* out,
                           uint64_t cap, uint64_t* out_count, int* rc_out);  // with the expanded paths, below

enum class SetOp { kAnd, kOr, kNot, kFilter };

// One query whose full ascending result set is wanted (the Index::Search* style calls): upload, plan, run the
// tiles, gather; then the limit / reverse rules of index.cpp:356-366.
// A batch object from the index's pool for the duration of one single-query call. A fresh Batch allocates its pinned
// staging buffer and two device arenas and frees them again at the end of the call — most of the ~5 ms a single
// Index::Search* style call cost at 10M documents (profiles/r01_h_expanded_paths_10M.md); the pooled objects keep them.
struct PooledBatch {
  Index& ix;
  std::unique_ptr<mgx_batch> h;
  int exceptions_at_entry = std::uncaught_exceptions();
  explicit PooledBatch(Index& index) : ix(index) {
    {
      std::lock_guard<std::mutex> pool_lock(ix.pool_mu);
      if (!ix.batch_pool.empty()) {
        h.reset(static_cast<mgx_batch*>(ix.batch_pool.back()));
        ix.batch_pool.pop_back();
      }
    }
    if (!h) {
      h = std::make_unique<mgx_batch>();
    }
  }
  ~PooledBatch() {
    if (std::uncaught_exceptions() > exceptions_at_entry) {
      return;  // a failed call does not hand its workspace on
    }
    cudaStreamSynchronize(h->b.stream);
    h->b.recycle();
    std::lock_guard<std::mutex> pool_lock(ix.pool_mu);
    if (ix.batch_pool.size() < 16) {
      ix.batch_pool.push_back(h.release());
    }
  }
};

int run_single_set_query(Index& ix, cudaStream_t lane_stream, std::vector<HostTerm>& terms,
                         const std::vector<HostQuery>& queries, const uint32_t* driver_ids, uint64_t n_driver, uint64_t limit, bool reverse, uint32_t* out,
                         uint64_t cap, uint64_t* out_count) {
  PooledBatch pooled(ix);
  Batch& b = pooled.h->b;
  b.ix = &ix;
  b.stream = lane_stream;
  b.params = mgx_query_params_t{};
  b.params.compute_score = 0;
  b.launches_at_start = g_launches.load();
  if (driver_ids != nullptr) {
    if (n_driver >= (1ULL << 32)) {
      return invalid("too many candidate ids");
    }
    b.driver_buf.reserve(std::max<uint64_t>(1, n_driver));  // grow-only with the pooled batch object
    MGX_CUDA(cudaMemcpyAsync(b.driver_buf.p, driver_ids, n_driver * sizeof(uint32_t), cudaMemcpyHostToDevice, b.stream));
    b.explicit_driver.d_ids = b.driver_buf.p;
    b.explicit_driver.n = n_driver;
  }
  batch_upload(b, terms, queries, {});
  batch_plan(b);
  std::vector<uint64_t> set_off;
  DevBuf<uint32_t> d_sets;
  batch_search_sets(b, &set_off, &d_sets);
  if (queries.size() > 1) {
    // one logical query expanded into several with different drivers (finish_expanded): their result sets are
    // disjoint by construction and their union is the answer
    DevBuf<uint32_t> d_union;
    merge_disjoint_runs(b, d_sets.p, set_off, &d_union);
    d_sets.borrow(d_union.p, d_union.n);  // both are views into the batch object's buffers
    set_off = {0, set_off.back()};
  }
  const uint64_t total = set_off[1];
  uint64_t first = 0;
  uint64_t n = total;
  if (limit > 0 && total > limit) {  // index.cpp:356-366
    n = limit;
    first = reverse ? total - limit : 0;
  }
  *out_count = n;
  if (n > cap) {
    set_last_error("output capacity too small");
    return MGX_ERR_CAPACITY;
  }
  if (n > 0) {
    MGX_CUDA(cudaMemcpyAsync(out, d_sets.p + first, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, b.stream));
    MGX_CUDA(cudaStreamSynchronize(b.stream));
    if (reverse) {
      std::reverse(out, out + n);
    }
  }
  return MGX_OK;
}


// Many logical queries in ONE batch (the batch forms of the fuzzy / synonym calls): kernel queries
// [group_begin[g], group_begin[g + 1]) belong to logical query g (several when finish_expanded split it by driver,
// none when the reference answers the empty set without touching a list); their disjoint result sets are united per
// group on the device. out: the groups' ascending doc ids back to back, out_offsets[g .. g + 1) = group g.
int run_grouped_set_queries(Index& ix, cudaStream_t lane_stream, std::vector<HostTerm>& terms,
                            const std::vector<HostQuery>& queries, const std::vector<uint32_t>& group_begin,
                            uint32_t* out, uint64_t cap, uint64_t* out_offsets) {
  const size_t n_groups = group_begin.size() - 1;
  for (size_t g = 0; g <= n_groups; ++g) {
    out_offsets[g] = 0;
  }
  if (queries.empty()) {
    return MGX_OK;
  }
  PooledBatch pooled(ix);
  Batch& b = pooled.h->b;
  b.ix = &ix;
  b.stream = lane_stream;
  b.params = mgx_query_params_t{};
  b.params.compute_score = 0;
  b.launches_at_start = g_launches.load();
  batch_upload(b, terms, queries, {});
  batch_plan(b);
  std::vector<uint64_t> set_off;
  DevBuf<uint32_t> d_sets;
  batch_search_sets(b, &set_off, &d_sets);
  bool any_split = false;
  for (size_t g = 0; g < n_groups; ++g) {
    any_split = any_split || group_begin[g + 1] - group_begin[g] > 1;
  }
  if (any_split) {
    DevBuf<uint32_t> d_union;
    merge_grouped_runs(b, d_sets.p, set_off, group_begin, &d_union);
    d_sets.borrow(d_union.p, d_union.n);  // both are views into the batch object's buffers
  }
  for (size_t g = 0; g <= n_groups; ++g) {
    out_offsets[g] = set_off[group_begin[g]];
  }
  const uint64_t total = set_off.back();
  if (total > cap) {
    set_last_error("output capacity too small (out_offsets holds the sizes)");
    return MGX_ERR_CAPACITY;
  }
  if (total > 0) {
    MGX_CUDA(cudaMemcpyAsync(out, d_sets.p, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, b.stream));
    MGX_CUDA(cudaStreamSynchronize(b.stream));
  }
  return MGX_OK;
}

// term ids of a query built against its own ProgramBuilder -> ids in the batch's common term table
void rebase_query_terms(HostQuery* q, uint32_t base) {
  for (size_t i = 0; i < q->prog_ops.size(); ++i) {
    if (q->prog_ops[i] == kOpTerm || q->prog_ops[i] == kOpFuzzyText) {
      q->prog_args[i] += base;  // (the edit distance of a FUZZYTEXT node sits in the top byte)
    }
  }
  for (uint32_t& c : q->conjuncts) {
    c += base;
  }
  for (size_t i = 0; i < q->terms.size(); ++i) {
    q->terms[i] += base;
  }
  for (size_t i = 0; i < q->not_terms.size(); ++i) {
    q->not_terms[i] += base;
  }
}

int run_set_op(mgx_index_t* index, SetOp op, const uint32_t* driver_ids, uint64_t n_driver, const uint8_t* term_bytes,
               const uint64_t* term_offsets, uint64_t n_terms, uint64_t limit, bool reverse, uint32_t* out,
               uint64_t cap, uint64_t* out_count) {
  if (index == nullptr || out_count == nullptr || (n_terms > 0 && (term_bytes == nullptr || term_offsets == nullptr))) {
    return invalid("null argument");
  }
  *out_count = 0;
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    auto copy_ids = [&](const uint32_t* src, uint64_t n) {
      *out_count = n;
      if (n > cap) {
        set_last_error("output capacity too small");
        return MGX_ERR_CAPACITY;
      }
      if (n > 0) {
        std::memcpy(out, src, n * sizeof(uint32_t));
      }
      return MGX_OK;
    };
    // trivial cases decided exactly as the reference does before touching any list
    if (op == SetOp::kAnd && n_terms == 0) {
      return MGX_OK;  // index.cpp:203-205
    }
    if (op == SetOp::kOr && n_terms == 0) {
      return MGX_OK;  // index.cpp:420-422
    }
    if (op == SetOp::kNot && n_terms == 0) {
      return copy_ids(driver_ids, n_driver);  // index.cpp:451-453
    }
    if (op == SetOp::kFilter && n_driver == 0) {
      return MGX_OK;  // index.cpp:372-374
    }
    if (op == SetOp::kFilter && n_terms == 0) {
      return copy_ids(driver_ids, n_driver);  // index.cpp:378-380
    }
    if ((op == SetOp::kNot) && n_driver == 0) {
      return MGX_OK;
    }

    std::vector<uint64_t> keys;
    for (uint64_t i = 0; i < n_terms; ++i) {
      uint64_t key = kInvalidKey;
      if (!host_ngram_to_key(term_bytes + term_offsets[i], term_offsets[i + 1] - term_offsets[i], ix.width, &key)) {
        key = kInvalidKey;
      }
      keys.push_back(key);
    }
    std::vector<HostTerm> terms;
    std::vector<HostQuery> queries(1);
    if (op == SetOp::kNot) {
      for (uint64_t key : keys) {  // one NOT group per n-gram: excluded if in ANY of the lists
        HostTerm t;
        t.raw = true;
        t.keys = {key};
        queries[0].not_terms.push_back(static_cast<uint32_t>(terms.size()));
        terms.push_back(std::move(t));
      }
      queries[0].flags = kQDriverExplicit;
    } else {
      HostTerm t;
      t.raw = true;
      t.keys = keys;
      std::sort(t.keys.begin(), t.keys.end());
      t.keys.erase(std::unique(t.keys.begin(), t.keys.end()), t.keys.end());
      bool any_mode = op == SetOp::kOr;
      if (op == SetOp::kOr) {
        // unknown terms are ignored (index.cpp:437-445)
        t.keys.erase(std::remove(t.keys.begin(), t.keys.end(), kInvalidKey), t.keys.end());
        // the union is driven by the lists themselves, not by a pass over every document of the shard
        int rc = MGX_OK;
        if (run_at_least_expanded(ix, rd.stream(), t.keys, 1, out, cap, out_count, &rc)) {
          return rc;
        }
        any_mode = t.keys.size() != 1;  // one list: the list itself (an AND of one n-gram)
      }
      terms.push_back(std::move(t));
      queries[0].terms.push_back(0);
      queries[0].flags = any_mode ? kQAnyMode : (op == SetOp::kFilter ? kQDriverExplicit : 0u);
    }

    return run_single_set_query(ix, rd.stream(), terms, queries,
                                (op == SetOp::kNot || op == SetOp::kFilter) ? driver_ids : nullptr,
                                n_driver, limit, reverse, out, cap, out_count);
  });
}

}  // namespace

int mgx_search_and(const mgx_index_t* index, const uint8_t* term_bytes, const uint64_t* term_offsets,
                   uint64_t n_terms, uint64_t limit, int32_t reverse, uint32_t* out, uint64_t cap,
                   uint64_t* out_count) {
  return run_set_op(const_cast<mgx_index_t*>(index), SetOp::kAnd, nullptr, 0, term_bytes, term_offsets, n_terms, limit,
                    reverse != 0, out, cap, out_count);
}

int mgx_search_or(const mgx_index_t* index, const uint8_t* term_bytes, const uint64_t* term_offsets,
                  uint64_t n_terms, uint32_t* out, uint64_t cap, uint64_t* out_count) {
  return run_set_op(const_cast<mgx_index_t*>(index), SetOp::kOr, nullptr, 0, term_bytes, term_offsets, n_terms, 0,
                    false, out, cap, out_count);
}

int mgx_search_not(const mgx_index_t* index, const uint32_t* all_docs, uint64_t n_all, const uint8_t* term_bytes,
                   const uint64_t* term_offsets, uint64_t n_terms, uint32_t* out, uint64_t cap, uint64_t* out_count) {
  return run_set_op(const_cast<mgx_index_t*>(index), SetOp::kNot, all_docs, n_all, term_bytes, term_offsets, n_terms,
                    0, false, out, cap, out_count);
}

int mgx_filter_by_ngrams(const mgx_index_t* index, const uint32_t* candidates, uint64_t n_candidates,
                         const uint8_t* term_bytes, const uint64_t* term_offsets, uint64_t n_terms, uint32_t* out,
                         uint64_t cap, uint64_t* out_count) {
  return run_set_op(const_cast<mgx_index_t*>(index), SetOp::kFilter, candidates, n_candidates, term_bytes,
                    term_offsets, n_terms, 0, false, out, cap, out_count);
}

int mgx_search_by_threshold(const mgx_index_t* index_c, const uint8_t* term_bytes, const uint64_t* term_offsets,
                            uint64_t n_terms, uint64_t threshold, uint32_t* out, uint64_t cap, uint64_t* out_count) {
  mgx_index_t* index = const_cast<mgx_index_t*>(index_c);
  if (index == nullptr || out_count == nullptr || (n_terms > 0 && (term_bytes == nullptr || term_offsets == nullptr))) {
    return invalid("null argument");
  }
  *out_count = 0;
  if (n_terms == 0 || threshold == 0) {
    return MGX_OK;  // index.cpp:489-491
  }
  // index.cpp:496-497: repeated terms count once
  std::vector<std::string> uniq;
  for (uint64_t i = 0; i < n_terms; ++i) {
    uniq.emplace_back(reinterpret_cast<const char*>(term_bytes) + term_offsets[i], term_offsets[i + 1] - term_offsets[i]);
  }
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  if (threshold > uniq.size()) {
    return MGX_OK;  // :499-501
  }
  std::vector<uint8_t> flat;
  std::vector<uint64_t> offs(1, 0);
  for (const std::string& t : uniq) {
    flat.insert(flat.end(), t.begin(), t.end());
    offs.push_back(flat.size());
  }
  flat.push_back(0);
  if (threshold == uniq.size()) {
    return mgx_search_and(index_c, flat.data(), offs.data(), uniq.size(), 0, 0, out, cap, out_count);  // :504-506
  }
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    HostTerm t;
    t.raw = true;
    for (size_t i = 0; i < uniq.size(); ++i) {
      uint64_t key = kInvalidKey;
      if (host_ngram_to_key(flat.data() + offs[i], offs[i + 1] - offs[i], ix.width, &key)) {
        t.keys.push_back(key);  // a term that cannot be an n-gram of this index has no posting list: it never counts
      }
    }
    std::sort(t.keys.begin(), t.keys.end());
    t.keys.erase(std::unique(t.keys.begin(), t.keys.end()), t.keys.end());
    if (threshold > t.keys.size()) {
      return MGX_OK;  // fewer lists than the threshold: nothing can qualify
    }
    int rc = MGX_OK;
    if (run_at_least_expanded(ix, rd.stream(), t.keys, threshold, out, cap, out_count, &rc)) {
      return rc;
    }
    std::vector<HostTerm> terms;
    terms.push_back(std::move(t));
    std::vector<HostQuery> queries(1);
    queries[0].terms.push_back(0);
    queries[0].flags = kQAnyMode;
    queries[0].threshold = static_cast<uint32_t>(threshold);
    return run_single_set_query(ix, rd.stream(), terms, queries, nullptr, 0, 0, false, out, cap, out_count);
  });
}

int mgx_eval_boolean(const mgx_index_t* index_c, const int32_t* ops, const int32_t* args, uint64_t n_ops,
                     const uint8_t* term_bytes, const uint64_t* term_offsets, uint64_t n_terms, uint32_t* out,
                     uint64_t cap, uint64_t* out_count) {
  mgx_index_t* index = const_cast<mgx_index_t*>(index_c);
  if (index == nullptr || out_count == nullptr || (n_ops > 0 && (ops == nullptr || args == nullptr)) ||
      (n_terms > 0 && (term_bytes == nullptr || term_offsets == nullptr))) {
    return invalid("null argument");
  }
  *out_count = 0;
  if (n_ops == 0) {
    return MGX_OK;
  }
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    std::vector<HostQuery> queries(1);
    HostQuery& hq = queries[0];
    if (int rc = analyse_program(ops, args, n_ops, n_terms, &hq.conjuncts); rc != MGX_OK) {
      return rc;
    }
    // TERM nodes tokenise with the index's own configuration (query_ast.cpp:80-84: GetNgramSize / the EFFECTIVE
    // GetKanjiNgramSize / GetCrossBoundaryNgrams)
    std::vector<HostTerm> terms(n_terms);
    for (uint64_t t = 0; t < n_terms; ++t) {
      const uint64_t b = term_offsets[t];
      const uint64_t e = term_offsets[t + 1];
      if (e - b > kMaxTermBytes) {
        set_last_error("query term longer than 256 bytes is not supported");
        return MGX_ERR_UNSUPPORTED;
      }
      terms[t].bytes.assign(reinterpret_cast<const char*>(term_bytes) + b, e - b);
      host_query_keys(term_bytes + b, e - b, ix.ngram, ix.kanji, ix.cross, ix.width, &terms[t].keys);
    }
    hq.flags = kQProgram;
    for (uint64_t i = 0; i < n_ops; ++i) {
      hq.prog_ops.push_back(static_cast<uint8_t>(ops[i]));
      hq.prog_args.push_back(static_cast<uint32_t>(args[i]));
    }
    return run_single_set_query(ix, rd.stream(), terms, queries, nullptr, 0, 0, false, out, cap, out_count);
  });
}

// ------------------------------------------------ fuzzy / synonym execution paths
namespace {

// A postfix program under construction over a private term table.
struct ProgramBuilder {
  std::vector<HostTerm> terms;
  std::vector<int32_t> ops;
  std::vector<int32_t> args;

  void leaf(HostTerm&& t) {
    ops.push_back(kOpTerm);
    args.push_back(static_cast<int32_t>(terms.size()));
    terms.push_back(std::move(t));
  }
  void node(uint8_t op, int32_t arg) {
    ops.push_back(op);
    args.push_back(arg);
  }
};

// ShouldApplyVerifyText, search_pipeline.cpp:48-66 (0 "off", 1 "all", 2 "ascii": only when every term is ASCII)
bool should_verify(int32_t mode, const uint8_t* bytes, const uint64_t* offsets, uint64_t n) {
  if (mode == 1) {
    return true;
  }
  if (mode != 2) {
    return false;
  }
  for (uint64_t i = offsets[0]; i < offsets[n]; ++i) {
    if (bytes[i] >= 0x80) {
      return false;
    }
  }
  return true;
}

int check_expanded(const mgx_expanded_query_t* eq) {
  if (eq == nullptr) {
    return invalid("null query description");
  }
  if (eq->n_not > 0 && (eq->not_bytes == nullptr || eq->not_offsets == nullptr)) {
    return invalid("NOT terms without bytes / offsets");
  }
  if (eq->n_filters > 0 && (eq->filter_col == nullptr || eq->filter_op == nullptr || eq->filter_bytes == nullptr ||
                            eq->filter_offsets == nullptr)) {
    return invalid("filters without columns / ops / literals");
  }
  if (eq->verify_text < 0 || eq->verify_text > 2) {
    return invalid("verify_text must be 0 (off), 1 (all) or 2 (ascii)");
  }
  return MGX_OK;
}

// A search term as the pipeline sees it (GenerateTermInfos, search_pipeline.cpp:569-603): its unique sorted
// n-grams cut with the RAW table configuration; without n-grams the leaf falls back to the substring test
// (SearchTermDocuments, :454-462).
int term_leaf(const Index& ix, const mgx_expanded_query_t& eq, const uint8_t* bytes, uint64_t len, HostTerm* t) {
  if (len > kMaxTermBytes) {
    set_last_error("query term longer than 256 bytes is not supported");
    return MGX_ERR_UNSUPPORTED;
  }
  t->bytes.assign(reinterpret_cast<const char*>(bytes), len);
  host_query_keys(bytes, len, eq.ngram_size, eq.kanji_ngram_size, eq.cross_boundary != 0, ix.width, &t->keys);
  return MGX_OK;
}

// ApplyNotAndFilters (search_pipeline.cpp:470-485): a document of ANY NOT term (ApplyNotFilter, :871-932) is
// dropped, then the column conditions; closes the program with the AND over `children` operands.
// `drivers`: leaves (term ids) of ONE child of the root AND such that every result lies in the posting lists of at
// least one of them (the variants of a synonym group; any n - t + 1 leaves of an "at least t of n" node). Without a
// conjunct term the engine would have to evaluate the program on every document of the shard; instead the query is
// expanded into one query per driver leaf i, AND(leaf_i, NOT leaf_1, ..., NOT leaf_{i-1}, program): each is driven by
// leaf_i's shortest list, the result sets are pairwise disjoint, and their union (merge_disjoint_runs) is the answer.
int finish_expanded(const Index& ix, const mgx_expanded_query_t& eq, ProgramBuilder* pb, int32_t children,
                    const std::vector<int32_t>& drivers, std::vector<HostQuery>* queries) {
  queries->assign(std::max<size_t>(1, drivers.size()), HostQuery{});
  HostQuery* hq = &(*queries)[0];
  if (eq.n_not > 0) {
    for (uint64_t i = 0; i < eq.n_not; ++i) {
      HostTerm t;
      if (int rc = term_leaf(ix, eq, eq.not_bytes + eq.not_offsets[i], eq.not_offsets[i + 1] - eq.not_offsets[i], &t);
          rc != MGX_OK) {
        return rc;
      }
      pb->leaf(std::move(t));
      if (i > 0) {
        pb->node(kOpOr, 2);  // folded pairwise: the evaluation stack stays shallow
      }
    }
    pb->node(kOpNot, 0);
    ++children;
  }
  pb->node(kOpAnd, children);
  auto install = [&](HostQuery* q, const std::vector<int32_t>& ops, const std::vector<int32_t>& args) {
    q->conjuncts.clear();
    if (int rc = analyse_program(ops.data(), args.data(), ops.size(), pb->terms.size(), &q->conjuncts); rc != MGX_OK) {
      return rc;
    }
    q->flags = kQProgram;
    q->prog_ops.assign(ops.begin(), ops.end());
    q->prog_args.assign(args.begin(), args.end());
    return MGX_OK;
  };
  bool expanded = !drivers.empty();
  if (expanded) {
    for (size_t i = 0; i < drivers.size() && expanded; ++i) {
      std::vector<int32_t> ops{kOpTerm};
      std::vector<int32_t> args{drivers[i]};
      for (size_t j = 0; j < i; ++j) {
        ops.insert(ops.end(), {kOpTerm, kOpNot, kOpAnd});
        args.insert(args.end(), {drivers[j], 0, 2});
      }
      ops.insert(ops.end(), pb->ops.begin(), pb->ops.end());
      args.insert(args.end(), pb->args.begin(), pb->args.end());
      ops.push_back(kOpAnd);
      args.push_back(2);
      expanded = install(&(*queries)[i], ops, args) == MGX_OK;  // too deep for the evaluation stack: one plain query
    }
  }
  if (!expanded) {
    queries->assign(1, HostQuery{});
    hq = &(*queries)[0];
    if (int rc = install(hq, pb->ops, pb->args); rc != MGX_OK) {
      return rc;
    }
  }
  for (uint64_t f = 0; f < eq.n_filters; ++f) {
    HostFilter hf;
    hf.col = eq.filter_col[f];
    hf.op = eq.filter_op[f];
    if (hf.op > 5) {
      return invalid("filter op must be 0..5 (EQ, NE, GT, GTE, LT, LTE)");
    }
    hf.literal.assign(reinterpret_cast<const char*>(eq.filter_bytes) + eq.filter_offsets[f],
                      eq.filter_offsets[f + 1] - eq.filter_offsets[f]);
    for (HostQuery& q : *queries) {
      q.filters.push_back(hf);
    }
  }
  return MGX_OK;
}

constexpr size_t kMaxDriverLeaves = 16;  // beyond this one pass over every document is the cheaper plan

// "Documents in at least `need` of these lists" (Index::SearchOr: need = 1; Index::SearchByThreshold, index.cpp:488-578)
// driven by lists instead of by every document of the shard: a document in >= need of n lists is in one of ANY
// n - need + 1 of them, so the query is expanded into one query per such list (finish_expanded) and the disjoint
// answers are united. Returns false when the expansion does not apply (too many driver lists): the caller then takes
// the single pass over all documents.
bool run_at_least_expanded(Index& ix, cudaStream_t stream, const KeyVec& keys, size_t need, uint32_t* out,
                           uint64_t cap, uint64_t* out_count, int* rc_out) {
  const size_t n = keys.size();
  if (n < 2 || need < 1 || need >= n || n - need + 1 > kMaxDriverLeaves || n > 0xFFFF ||
      std::getenv("MGX_NO_OR_EXPANSION") != nullptr) {
    return false;
  }
  // ANY n - need + 1 lists will do for exactness; the SHORTEST ones are the cheapest drivers. Their lengths cost one
  // lookup launch; when they add up to a sizeable part of the shard the single pass over all documents (bit probes
  // into the dense lists, 1024 documents per tile) is the cheaper plan.
  std::vector<uint64_t> sorted_keys(keys.begin(), keys.end());
  if (n <= 64) {
    uint32_t lens[64];
    {
      PooledBatch pooled(ix);
      Batch& b = pooled.h->b;
      b.ix = &ix;
      b.stream = stream;
      lookup_list_lengths(b, sorted_keys.data(), static_cast<uint32_t>(n), lens);
    }
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; ++i) {
      order[i] = static_cast<uint32_t>(i);
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b2) { return lens[a] < lens[b2]; });
    uint64_t driven = 0;
    for (size_t i = 0; i < n - need + 1; ++i) {
      driven += lens[order[i]];
    }
    if (driven > ix.n_docs / 4) {
      return false;
    }
    std::vector<uint64_t> by_len(n);
    for (size_t i = 0; i < n; ++i) {
      by_len[i] = sorted_keys[order[i]];
    }
    sorted_keys.swap(by_len);
  }
  ProgramBuilder pb;
  std::vector<int32_t> drivers;
  for (size_t i = 0; i < n; ++i) {
    HostTerm leaf;
    leaf.raw = true;
    leaf.keys.push_back(sorted_keys[i]);
    if (i < n - need + 1) {
      drivers.push_back(static_cast<int32_t>(pb.terms.size()));
    }
    pb.leaf(std::move(leaf));
  }
  pb.node(kOpAtLeast, static_cast<int32_t>(n | (need << 16)));
  mgx_expanded_query_t eq{};
  std::vector<HostQuery> queries;
  int rc = finish_expanded(ix, eq, &pb, 1, drivers, &queries);
  if (rc == MGX_OK) {
    rc = run_single_set_query(ix, stream, pb.terms, queries, nullptr, 0, 0, false, out, cap, out_count);
  }
  *rc_out = rc;
  return true;
}

}  // namespace

namespace {
// The program(s) of ONE fuzzy query (ExecuteWithFuzzy, search_pipeline.cpp:1659-1755) appended to pb / written to
// queries; *empty: the reference answers the empty set before touching a list. Shared by the single and batch calls.
int build_fuzzy(Index& ix, const mgx_expanded_query_t* eq, const uint8_t* term_bytes, const uint64_t* term_offsets,
                uint64_t n_terms, uint32_t max_distance, ProgramBuilder* pb, std::vector<HostQuery>* queries,
                bool* empty) {
  *empty = false;
  queries->clear();
  if (n_terms == 0) {
    *empty = true;  // search_pipeline.cpp:1667-1670
    return MGX_OK;
  }
  const bool fuzzy_verify = should_verify(eq->verify_text, term_bytes, term_offsets, n_terms);
  if (fuzzy_verify && max_distance > 127) {
    return invalid("max_distance above 127");
  }
  int32_t children = 0;
  bool hybrid_exact = false;
  bool has_conjunct = false;      // a term whose n-grams are ALL required drives by itself
  std::vector<int32_t> drivers;   // else: the smallest sufficient leaf set over the terms
  for (uint64_t t = 0; t < n_terms; ++t) {
    const uint8_t* tb = term_bytes + term_offsets[t];
    const uint64_t tl = term_offsets[t + 1] - term_offsets[t];
    HostTerm whole;
    if (int rc = term_leaf(ix, *eq, tb, tl, &whole); rc != MGX_OK) {
      return rc;
    }
    const size_t n = whole.keys.size();
    if (n == 0) {
      *empty = true;  // too short for an n-gram: no candidates, empty_term_detected (:1674-1680)
      return MGX_OK;
    }
    if (n > kMaxProgramDepth - 2) {
      set_last_error("fuzzy term with more than 62 distinct n-grams is not supported");
      return MGX_ERR_UNSUPPORTED;
    }
    // effective n-gram size of the term (:1682-1695): the kanji size when most of its n-grams are <= 3 bytes
    int eff = eq->ngram_size > 0 ? eq->ngram_size : 2;
    if (eq->kanji_ngram_size > 0) {
      size_t short_count = 0;
      for (uint64_t key : whole.keys) {
        uint8_t enc[4 * kMaxNgramSize];
        if (key == kInvalidKey) {
          set_last_error("fuzzy term with n-grams wider than the index key");
          return MGX_ERR_UNSUPPORTED;
        }
        short_count += host_key_to_utf8(key, ix.width, enc) <= 3 ? 1 : 0;
      }
      if (short_count > n / 2) {
        eff = eq->kanji_ngram_size;
      }
    }
    const size_t drop = static_cast<size_t>(max_distance) * static_cast<size_t>(eff);
    const size_t need = n > drop ? n - drop : 1;  // :1697-1700
    const int32_t first_leaf = static_cast<int32_t>(pb->terms.size());
    for (uint64_t key : whole.keys) {  // Index::SearchByThreshold(ngrams, need), index.cpp:488-578
      HostTerm leaf;
      leaf.raw = true;
      leaf.keys.push_back(key);
      pb->leaf(std::move(leaf));
    }
    pb->node(kOpAtLeast, static_cast<int32_t>(n | (need << 16)));
    ++children;
    // a document in at least `need` of the n lists is in one of ANY n - need + 1 of them
    const size_t sufficient = n - need + 1;
    has_conjunct |= need == n;
    if (need < n && sufficient <= kMaxDriverLeaves && (drivers.empty() || sufficient < drivers.size())) {
      drivers.clear();
      for (size_t i = 0; i < sufficient; ++i) {
        drivers.push_back(first_leaf + static_cast<int32_t>(i));
      }
    }
    hybrid_exact |= has_uncovered_hybrid_fragment(tb, tl, eq->ngram_size, eq->kanji_ngram_size,
                                                  eq->cross_boundary != 0);
  }
  if (fuzzy_verify) {
    // PostFilterByFuzzyText (:1742-1752): every term occurs in the text exactly, or a word of the text is within
    // max_distance edits of it (ContainsFuzzyMatch, utils/edit_distance.cpp)
    for (uint64_t t = 0; t < n_terms; ++t) {
      const uint8_t* tb = term_bytes + term_offsets[t];
      const uint64_t tl = term_offsets[t + 1] - term_offsets[t];
      if (host_utf8_to_codepoints(tb, tl).size() > kFuzzyMaxTermCps) {
        set_last_error("fuzzy verification of a term longer than 64 code points is not supported");
        return MGX_ERR_UNSUPPORTED;
      }
      HostTerm exact;
      exact.bytes.assign(reinterpret_cast<const char*>(tb), tl);
      const int32_t tid = static_cast<int32_t>(pb->terms.size());
      pb->leaf(std::move(exact));
      pb->node(kOpFuzzyText, tid | static_cast<int32_t>(max_distance << 24));  // same term bytes, fuzzy test
      pb->node(kOpOr, 2);
      ++children;
    }
  }
  if (hybrid_exact) {  // RequiresExactTextForHybridFragments -> PostFilterByText (:1728-1737): every term, exactly
    for (uint64_t t = 0; t < n_terms; ++t) {
      HostTerm text_only;
      text_only.bytes.assign(reinterpret_cast<const char*>(term_bytes) + term_offsets[t],
                             term_offsets[t + 1] - term_offsets[t]);
      pb->leaf(std::move(text_only));
      ++children;
    }
  }
  if (has_conjunct) {
    drivers.clear();
  }
  return finish_expanded(ix, *eq, pb, children, drivers, queries);
}
}  // namespace

int mgx_search_fuzzy(const mgx_index_t* index_c, const mgx_expanded_query_t* eq, const uint8_t* term_bytes,
                     const uint64_t* term_offsets, uint64_t n_terms, uint32_t max_distance, uint32_t* out,
                     uint64_t cap, uint64_t* out_count) {
  mgx_index_t* index = const_cast<mgx_index_t*>(index_c);
  if (index == nullptr || out_count == nullptr || (n_terms > 0 && (term_bytes == nullptr || term_offsets == nullptr))) {
    return invalid("null argument");
  }
  *out_count = 0;
  if (int rc = check_expanded(eq); rc != MGX_OK) {
    return rc;
  }
  if (n_terms == 0) {
    return MGX_OK;  // search_pipeline.cpp:1667-1670
  }
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    ProgramBuilder pb;
    std::vector<HostQuery> queries;
    bool empty = false;
    if (int rc = build_fuzzy(ix, eq, term_bytes, term_offsets, n_terms, max_distance, &pb, &queries, &empty);
        rc != MGX_OK || empty) {
      return rc;
    }
    return run_single_set_query(ix, rd.stream(), pb.terms, queries, nullptr, 0, 0, false, out, cap, out_count);
  });
}

namespace {
// The program(s) of ONE synonym query (ExecuteWithSynonyms, search_pipeline.cpp:1580-1657); group_begin[0..n_groups]
// index the variants of this query. Shared by the single and batch calls.
int build_synonyms(Index& ix, const mgx_expanded_query_t* eq, const uint8_t* variant_bytes,
                   const uint64_t* variant_offsets, const uint64_t* group_begin, uint64_t n_groups, ProgramBuilder* pb,
                   std::vector<HostQuery>* queries, bool* empty) {
  *empty = false;
  queries->clear();
  if (n_groups == 0) {
    *empty = true;  // no group was processed: empty_term_detected (search_pipeline.cpp:1618-1621)
    return MGX_OK;
  }
  for (uint64_t g = 0; g < n_groups; ++g) {
    if (group_begin[g + 1] < group_begin[g]) {
      return invalid("group_begin must be non-decreasing");
    }
  }
  int32_t children = 0;
  const uint64_t n_variants = group_begin[n_groups];
  // ShouldApplyVerifyTextSynonyms (:154-172): decided over the variants of all groups
  const bool verify = n_variants > 0 && should_verify(eq->verify_text, variant_bytes, variant_offsets, n_variants);
  bool has_conjunct = false;     // a group with a single variant drives by itself
  std::vector<int32_t> drivers;  // else: the variants of the group with the fewest of them
  for (int pass = 0; pass < (verify ? 2 : 1); ++pass) {
    // pass 0: OR within a group of SearchTermDocuments(variant) (:1589-1608);
    // pass 1: PostFilterByTextWithSynonyms (:1633-1657): some variant of every group occurs in the text
    for (uint64_t g = 0; g < n_groups; ++g) {
      bool trivially_true = false;
      if (pass == 1) {
        for (uint64_t v = group_begin[g]; v < group_begin[g + 1]; ++v) {
          trivially_true |= variant_offsets[v + 1] == variant_offsets[v];  // text.find("") always succeeds
        }
      }
      if (trivially_true) {
        continue;
      }
      std::vector<int32_t> group_leaves;
      bool can_drive = pass == 0;
      for (uint64_t v = group_begin[g]; v < group_begin[g + 1]; ++v) {
        const uint8_t* vb = variant_bytes + variant_offsets[v];
        const uint64_t vl = variant_offsets[v + 1] - variant_offsets[v];
        HostTerm t;
        if (pass == 0) {
          if (int rc = term_leaf(ix, *eq, vb, vl, &t); rc != MGX_OK) {
            return rc;
          }
          if (!t.keys.empty()) {
            group_leaves.push_back(static_cast<int32_t>(pb->terms.size()));
          } else if (vl > 0) {
            can_drive = false;  // a variant shorter than an n-gram is a text scan: no list to drive with
          }
        } else {
          if (vl > kMaxTermBytes) {
            set_last_error("query term longer than 256 bytes is not supported");
            return MGX_ERR_UNSUPPORTED;
          }
          t.bytes.assign(reinterpret_cast<const char*>(vb), vl);  // text-only leaf
        }
        pb->leaf(std::move(t));
        if (v > group_begin[g]) {
          pb->node(kOpOr, 2);
        }
      }
      if (group_begin[g + 1] == group_begin[g]) {
        pb->node(kOpOr, 0);  // a group without variants matches nothing
      }
      ++children;
      if (can_drive && !group_leaves.empty()) {
        has_conjunct |= group_begin[g + 1] - group_begin[g] == 1;
        if (group_leaves.size() <= kMaxDriverLeaves && (drivers.empty() || group_leaves.size() < drivers.size())) {
          drivers = group_leaves;
        }
      }
    }
  }
  if (has_conjunct) {
    drivers.clear();
  }
  return finish_expanded(ix, *eq, pb, children, drivers, queries);
}
}  // namespace

int mgx_search_synonyms(const mgx_index_t* index_c, const mgx_expanded_query_t* eq, const uint8_t* variant_bytes,
                        const uint64_t* variant_offsets, const uint64_t* group_begin, uint64_t n_groups,
                        uint32_t* out, uint64_t cap, uint64_t* out_count) {
  mgx_index_t* index = const_cast<mgx_index_t*>(index_c);
  if (index == nullptr || out_count == nullptr ||
      (n_groups > 0 && (variant_bytes == nullptr || variant_offsets == nullptr || group_begin == nullptr))) {
    return invalid("null argument");
  }
  *out_count = 0;
  if (int rc = check_expanded(eq); rc != MGX_OK) {
    return rc;
  }
  if (n_groups == 0) {
    return MGX_OK;  // no group was processed: empty_term_detected (search_pipeline.cpp:1618-1621)
  }
  for (uint64_t g = 0; g < n_groups; ++g) {
    if (group_begin[g + 1] < group_begin[g]) {
      return invalid("group_begin must be non-decreasing");
    }
  }
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    ProgramBuilder pb;
    std::vector<HostQuery> queries;
    bool empty = false;
    if (int rc = build_synonyms(ix, eq, variant_bytes, variant_offsets, group_begin, n_groups, &pb, &queries, &empty);
        rc != MGX_OK || empty) {
      return rc;
    }
    return run_single_set_query(ix, rd.stream(), pb.terms, queries, nullptr, 0, 0, false, out, cap, out_count);
  });
}

namespace {
// appends one logical query's programs to the batch under construction
void append_group(ProgramBuilder* one, std::vector<HostQuery>* its_queries, bool empty, std::vector<HostTerm>* terms,
                  std::vector<HostQuery>* queries, std::vector<uint32_t>* group_begin) {
  if (!empty) {
    const uint32_t base = static_cast<uint32_t>(terms->size());
    for (HostTerm& t : one->terms) {
      terms->push_back(std::move(t));
    }
    for (HostQuery& q : *its_queries) {
      rebase_query_terms(&q, base);
      queries->push_back(std::move(q));
    }
  }
  group_begin->push_back(static_cast<uint32_t>(queries->size()));
}
}  // namespace

int mgx_search_fuzzy_batch(const mgx_index_t* index_c, const mgx_expanded_query_t* eq, uint64_t n_queries,
                           const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                           uint32_t max_distance, uint32_t* out, uint64_t cap, uint64_t* out_offsets) {
  mgx_index_t* index = const_cast<mgx_index_t*>(index_c);
  if (index == nullptr || out_offsets == nullptr || (n_queries > 0 && q_term_begin == nullptr)) {
    return invalid("null argument");
  }
  for (uint64_t q = 0; q <= n_queries; ++q) {
    out_offsets[q] = 0;
  }
  if (int rc = check_expanded(eq); rc != MGX_OK) {
    return rc;
  }
  for (uint64_t q = 0; q < n_queries; ++q) {
    if (q_term_begin[q + 1] < q_term_begin[q]) {
      return invalid("q_term_begin must be non-decreasing");
    }
  }
  if (n_queries == 0) {
    return MGX_OK;
  }
  if (q_term_begin[n_queries] > 0 && (term_bytes == nullptr || term_offsets == nullptr)) {
    return invalid("terms without bytes / offsets");
  }
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    std::vector<HostTerm> terms;
    std::vector<HostQuery> queries;
    std::vector<uint32_t> group_begin(1, 0);
    for (uint64_t q = 0; q < n_queries; ++q) {
      ProgramBuilder pb;
      std::vector<HostQuery> its;
      bool empty = false;
      const uint64_t t0 = q_term_begin[q];
      const uint64_t nt = q_term_begin[q + 1] - t0;
      if (int rc = build_fuzzy(ix, eq, term_bytes, nt > 0 ? term_offsets + t0 : nullptr, nt, max_distance, &pb, &its, &empty);
          rc != MGX_OK) {
        return rc;
      }
      append_group(&pb, &its, empty, &terms, &queries, &group_begin);
    }
    return run_grouped_set_queries(ix, rd.stream(), terms, queries, group_begin, out, cap, out_offsets);
  });
}

int mgx_search_synonyms_batch(const mgx_index_t* index_c, const mgx_expanded_query_t* eq, uint64_t n_queries,
                              const uint8_t* variant_bytes, const uint64_t* variant_offsets, const uint64_t* group_begin_in,
                              const uint64_t* q_group_begin, uint32_t* out, uint64_t cap, uint64_t* out_offsets) {
  mgx_index_t* index = const_cast<mgx_index_t*>(index_c);
  if (index == nullptr || out_offsets == nullptr || (n_queries > 0 && q_group_begin == nullptr)) {
    return invalid("null argument");
  }
  for (uint64_t q = 0; q <= n_queries; ++q) {
    out_offsets[q] = 0;
  }
  if (int rc = check_expanded(eq); rc != MGX_OK) {
    return rc;
  }
  for (uint64_t q = 0; q < n_queries; ++q) {
    if (q_group_begin[q + 1] < q_group_begin[q]) {
      return invalid("q_group_begin must be non-decreasing");
    }
  }
  if (n_queries == 0) {
    return MGX_OK;
  }
  if (q_group_begin[n_queries] > 0 && (variant_bytes == nullptr || variant_offsets == nullptr || group_begin_in == nullptr)) {
    return invalid("groups without variants");
  }
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    std::vector<HostTerm> terms;
    std::vector<HostQuery> queries;
    std::vector<uint32_t> group_begin(1, 0);
    std::vector<uint64_t> local;
    for (uint64_t q = 0; q < n_queries; ++q) {
      ProgramBuilder pb;
      std::vector<HostQuery> its;
      bool empty = false;
      const uint64_t g0 = q_group_begin[q];
      const uint64_t ng = q_group_begin[q + 1] - g0;
      // the query's own view: variants renumbered from 0 (the verification rule looks at the query's variants only)
      local.assign(ng + 1, 0);
      const uint64_t v0 = ng > 0 ? group_begin_in[g0] : 0;
      for (uint64_t g = 0; g <= ng && ng > 0; ++g) {
        if (group_begin_in[g0 + g] < v0) {
          return invalid("group_begin must be non-decreasing");
        }
        local[g] = group_begin_in[g0 + g] - v0;
      }
      if (int rc = build_synonyms(ix, eq, variant_bytes, ng > 0 ? variant_offsets + v0 : nullptr, local.data(), ng, &pb,
                                  &its, &empty);
          rc != MGX_OK) {
        return rc;
      }
      append_group(&pb, &its, empty, &terms, &queries, &group_begin);
    }
    return run_grouped_set_queries(ix, rd.stream(), terms, queries, group_begin, out, cap, out_offsets);
  });
}

int mgx_index_get_postings(const mgx_index_t* index, const uint8_t* term, uint64_t term_len, uint32_t* out,
                           uint64_t cap, uint64_t* out_count) {
  const uint64_t offs[2] = {0, term_len};
  static const uint8_t kEmpty[1] = {0};
  return mgx_search_and(index, term != nullptr ? term : kEmpty, offs, 1, 0, 0, out, cap, out_count);
}

namespace {
// CSR range [b, e) of one n-gram's list (binary search over the device dictionary, one key per probe); false if the
// index does not hold it. The caller holds a Reader and the device.
bool find_term_range(Index& ix, const uint8_t* term, uint64_t term_len, uint64_t* b, uint64_t* e) {
  uint64_t key = 0;
  if (term == nullptr || !host_ngram_to_key(term, term_len, ix.width, &key) || ix.n_terms == 0) {
    return false;
  }
  const int W = ix.wide_words;
  uint64_t want[kMaxWideWords] = {key, 0, 0, 0};
  if (W > 0) {
    std::memcpy(want, host_wide_words(key, W), static_cast<size_t>(W) * sizeof(uint64_t));
  }
  const int nw = W > 0 ? W : 1;
  const uint64_t* dict = W > 0 ? ix.d_wide_keys.p : ix.d_term_keys.p;
  auto probe = [&](uint64_t at) {  // <0: dict[at] < want
    uint64_t v[kMaxWideWords] = {0, 0, 0, 0};
    MGX_CUDA(cudaMemcpy(v, dict + at * nw, static_cast<size_t>(nw) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    for (int w = 0; w < nw; ++w) {
      if (v[w] != want[w]) {
        return v[w] < want[w] ? -1 : 1;
      }
    }
    return 0;
  };
  uint64_t lo = 0;
  uint64_t hi = ix.n_terms;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) / 2;
    if (probe(mid) < 0) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  if (lo >= ix.n_terms || probe(lo) != 0) {
    return false;
  }
  uint64_t off[2];
  MGX_CUDA(cudaMemcpy(off, ix.d_term_off.p + lo, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  *b = off[0];
  *e = off[1];
  return true;
}
}  // namespace

int mgx_index_posting_size(const mgx_index_t* index, const uint8_t* term, uint64_t term_len, uint64_t* out) {
  if (index == nullptr || out == nullptr) {
    return invalid("null argument");
  }
  *out = 0;
  mgx_index_t* h = const_cast<mgx_index_t*>(index);
  if (int rc = commit_pending(h); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(h);
    Index& ix = h->ix;
    DeviceGuard guard(ix.device);
    uint64_t b = 0, e = 0;
    if (find_term_range(ix, term, term_len, &b, &e)) {
      *out = e - b;
    }
    return MGX_OK;
  });
}

int mgx_index_get_posting_payload(const mgx_index_t* index, const uint8_t* term, uint64_t term_len, uint32_t* docs,
                                  uint32_t* first, uint32_t* second, uint64_t cap, uint64_t* out_count,
                                  int* layout) {
  if (index == nullptr || out_count == nullptr) {
    return invalid("null argument");
  }
  *out_count = 0;
  mgx_index_t* h = const_cast<mgx_index_t*>(index);
  if (int rc = commit_pending(h); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(h);
    Index& ix = h->ix;
    DeviceGuard guard(ix.device);
    if (layout != nullptr) {
      layout[0] = ix.has_positions ? ix.sig.pos_bits : 0;
      layout[1] = ix.has_positions ? ix.sig.next_bits : 0;
      layout[2] = ix.has_positions ? ix.sig.prev_bits : 0;
    }
    uint64_t b = 0, e = 0;
    if (!ix.has_positions || !find_term_range(ix, term, term_len, &b, &e)) {
      return MGX_OK;
    }
    *out_count = e - b;
    const uint64_t n = std::min<uint64_t>(e - b, cap);
    if (n != 0 && docs != nullptr) {  // LOCAL document indices (position in the shard's ascending id list)
      MGX_CUDA(cudaMemcpy(docs, ix.d_postings.p + b, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    if (n != 0 && first != nullptr) {
      MGX_CUDA(cudaMemcpy(first, ix.d_post_pos.p + b, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    if (n != 0 && second != nullptr) {
      MGX_CUDA(cudaMemcpy(second, ix.d_post_pos2.p + b, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    return MGX_OK;
  });
}

namespace {
// CSR of the index with global doc ids; the caller holds shared access.
// keys: n_terms packed keys, or (wide keys) n_terms * wide_words words
int export_unlocked(const Index& ix, uint64_t* keys, uint64_t* offsets, uint32_t* postings, bool wide = false) {
  DeviceGuard guard(ix.device);
  if (ix.n_terms > 0 && wide && ix.wide_words > 0) {
    MGX_CUDA(cudaMemcpy(keys, ix.d_wide_keys.p, ix.n_terms * ix.wide_words * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  } else if (ix.n_terms > 0) {
    MGX_CUDA(cudaMemcpy(keys, ix.d_term_keys.p, ix.n_terms * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  }
  if (ix.d_term_off.p != nullptr && ix.d_term_off.n >= ix.n_terms + 1) {
    MGX_CUDA(cudaMemcpy(offsets, ix.d_term_off.p, (ix.n_terms + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  } else {
    offsets[0] = 0;
  }
  if (ix.n_postings > 0) {
    MGX_CUDA(cudaMemcpy(postings, ix.d_postings.p, ix.n_postings * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    // local index -> global DocId (a pure relabelling done while copying out)
    std::vector<uint32_t> ids(ix.n_docs);
    MGX_CUDA(cudaMemcpy(ids.data(), ix.d_doc_ids.p, ix.n_docs * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < ix.n_postings; ++i) {
      postings[i] = ids[postings[i]];
    }
  }
  return MGX_OK;
}
}  // namespace

int mgx_index_export(const mgx_index_t* index, uint64_t* keys, uint64_t* offsets, uint32_t* postings) {
  if (index == nullptr || keys == nullptr || offsets == nullptr || postings == nullptr) {
    return invalid("null argument");
  }
  mgx_index_t* h = const_cast<mgx_index_t*>(index);
  if (int rc = commit_pending(h); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(h);
    return export_unlocked(h->ix, keys, offsets, postings);
  });
}

int mgx_index_export_terms(const mgx_index_t* index, uint8_t* term_bytes, uint64_t cap_bytes, uint64_t* term_offsets,
                           uint64_t* out_bytes) {
  if (index == nullptr || term_offsets == nullptr || out_bytes == nullptr) {
    return invalid("null argument");
  }
  *out_bytes = 0;
  mgx_index_t* h = const_cast<mgx_index_t*>(index);
  if (int rc = commit_pending(h); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(h);
    const Index& ix = h->ix;
    DeviceGuard guard(ix.device);
    const int W = ix.wide_words;
    const int nw = std::max(W, 1);
    std::vector<uint64_t> keys(ix.n_terms * nw + 1);
    if (ix.n_terms > 0) {
      MGX_CUDA(cudaMemcpy(keys.data(), W > 0 ? ix.d_wide_keys.p : ix.d_term_keys.p, ix.n_terms * nw * sizeof(uint64_t),
                          cudaMemcpyDeviceToHost));
    }
    uint64_t tb = 0;
    uint8_t enc[4 * kMaxNgramSize];
    for (uint64_t t = 0; t < ix.n_terms; ++t) {
      term_offsets[t] = tb;
      const int n = W > 0 ? wide_words_to_utf8(keys.data() + t * W, W, enc) : mgx_key_to_utf8(keys[t], ix.width, enc);
      if (term_bytes != nullptr && tb + static_cast<uint64_t>(n) <= cap_bytes) {
        std::memcpy(term_bytes + tb, enc, static_cast<size_t>(n));
      }
      tb += static_cast<uint64_t>(n);
    }
    term_offsets[ix.n_terms] = tb;
    *out_bytes = tb;
    if (term_bytes != nullptr && tb > cap_bytes) {
      set_last_error("mgx_index_export_terms: output capacity too small");
      return MGX_ERR_CAPACITY;
    }
    return MGX_OK;
  });
}

int mgx_index_save_mgix(const mgx_index_t* index, int32_t normalize_nfkc, const char* normalize_width,
                        int32_t normalize_lower, uint8_t* out, uint64_t cap, uint64_t* out_len) {
  if (index == nullptr || out_len == nullptr) {
    return invalid("null argument");
  }
  *out_len = 0;
  const char* width = normalize_width != nullptr ? normalize_width : "keep";
  mgx_mgix_info_t info{};
  if (std::strlen(width) >= sizeof(info.normalize_width)) {
    set_last_error("normalize_width longer than 31 bytes is not supported");
    return MGX_ERR_UNSUPPORTED;
  }
  mgx_index_t* h = const_cast<mgx_index_t*>(index);
  if (int rc = commit_pending(h); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    // sizes, CSR and header are read under ONE shared access, so a mutation committed by another thread cannot
    // grow the index between the sizing and the copy. The CSR comes back once (global doc ids, ascending term order).
    std::vector<uint64_t> keys;
    std::vector<uint64_t> offsets;
    std::vector<uint32_t> postings;
    int width_cp = 0;
    double roaring_min_len = 0.0;
    uint64_t n_terms = 0;
    {
      Reader rd(h);
      const Index& ix = h->ix;
      n_terms = ix.n_terms;
      keys.resize(ix.n_terms * std::max(ix.wide_words, 1) + 1);
      offsets.resize(ix.n_terms + 1);
      postings.resize(ix.n_postings + 1);
      if (int rc = export_unlocked(ix, keys.data(), offsets.data(), postings.data(), true); rc != MGX_OK) {
        return rc;
      }
      width_cp = ix.width;
      info.ngram_size = ix.ngram;
      info.kanji_ngram_size = ix.kanji;  // effective: kanji > 0 ? kanji : ngram, as index.cpp:29-37 stores it
      info.cross_boundary = ix.cross ? 1 : 0;
      const double thr = ix.cfg.roaring_threshold > 0.0 ? ix.cfg.roaring_threshold : 0.18;
      roaring_min_len = ix.optimized_total_docs > 0 ? thr * static_cast<double>(ix.optimized_total_docs) : 0.0;
    }
    info.normalize_nfkc = normalize_nfkc;
    info.normalize_lower = normalize_lower;
    std::strcpy(info.normalize_width, width);
    info.n_terms = n_terms;
    std::vector<uint8_t> term_bytes(n_terms * 4 * static_cast<size_t>(width_cp) + 1);
    std::vector<uint64_t> term_offsets(n_terms + 1, 0);
    uint64_t tb = 0;
    const int W = wide_words_for(width_cp);
    for (uint64_t t = 0; t < n_terms; ++t) {
      term_offsets[t] = tb;
      tb += static_cast<uint64_t>(W > 0 ? wide_words_to_utf8(keys.data() + t * W, W, term_bytes.data() + tb)
                                        : mgx_key_to_utf8(keys[t], width_cp, term_bytes.data() + tb));
    }
    term_offsets[n_terms] = tb;
    return mgx_mgix_encode(&info, term_bytes.data(), term_offsets.data(), offsets.data(), postings.data(),
                           roaring_min_len, out, cap, out_len);
  });
}

int mgx_index_load_mgix(mgx_index_t* index, const uint8_t* data, uint64_t len) {
  if (index == nullptr || data == nullptr) {
    return invalid("null argument");
  }
  if (int rc = require_device(); rc != MGX_OK) {
    return rc;
  }
  // host side: the reference's validation and decoding (mgx_mgix_decode), the configuration check of LoadFromData
  // (index_serialization.cpp:371-447), n-grams -> packed keys
  mgx_mgix_info_t info{};
  if (int rc = mgx_mgix_decode(data, len, &info, nullptr, nullptr, nullptr, nullptr); rc != MGX_OK) {
    return rc;
  }
  const Index& cfg = index->ix;
  if (info.ngram_size != cfg.ngram || info.kanji_ngram_size != cfg.kanji || (info.cross_boundary != 0) != cfg.cross) {
    set_last_error("MGIX stream was written with another n-gram configuration (kStorageConfigMismatch)");
    return MGX_ERR_FORMAT;
  }
  std::vector<uint8_t> term_bytes(info.term_bytes + 1);
  std::vector<uint64_t> term_offsets(info.n_terms + 1, 0);
  std::vector<uint64_t> posting_offsets(info.n_terms + 1, 0);
  std::vector<uint32_t> postings(info.n_postings + 1);
  if (int rc = mgx_mgix_decode(data, len, &info, term_bytes.data(), term_offsets.data(), posting_offsets.data(),
                               postings.data());
      rc != MGX_OK) {
    return rc;
  }
  const int W = wide_words_for(cfg.width);
  std::vector<uint64_t> keys(info.n_terms * std::max(W, 1) + 1);
  for (uint64_t t = 0; t < info.n_terms; ++t) {
    bool ok = true;
    if (W == 0) {
      ok = host_ngram_to_key(term_bytes.data() + term_offsets[t], term_offsets[t + 1] - term_offsets[t], cfg.width,
                             &keys[t]) &&
           (t == 0 || keys[t] > keys[t - 1]);
    } else {
      // strict UTF-8 of 1..width code points -> the words of the wide key, ascending word by word
      uint32_t cps[kMaxNgramSize];
      int n = 0;
      const uint8_t* tp = term_bytes.data() + term_offsets[t];
      const uint64_t tl = term_offsets[t + 1] - term_offsets[t];
      for (uint64_t i = 0; i < tl && ok;) {
        uint32_t cp = 0;
        const uint64_t avail = tl - i;
        const int l = parse_utf8(tp[i], avail > 1 ? tp[i + 1] : 0, avail > 2 ? tp[i + 2] : 0, avail > 3 ? tp[i + 3] : 0,
                                 avail, &cp);
        if (l <= 0 || n >= cfg.width) {
          ok = false;
        } else {
          cps[n++] = cp;
          i += static_cast<uint64_t>(l);
        }
      }
      ok = ok && n > 0;
      if (ok) {
        pack_wide(cps, n, W, keys.data() + t * W);
        ok = t == 0 || std::lexicographical_compare(keys.data() + (t - 1) * W, keys.data() + t * W, keys.data() + t * W,
                                                    keys.data() + (t + 1) * W);
      }
    }
    if (!ok) {
      set_last_error("MGIX stream holds a term that is not an n-gram of this configuration (or terms out of order)");
      return MGX_ERR_FORMAT;
    }
  }
  {
    std::lock_guard<std::mutex> jl(index->journal_mu);  // the stream replaces everything, pending mutations included
    index->journal.clear();
    index->dirty.store(false);
  }
  return guarded([&]() {
    WriteGuard lock(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    load_index_device(ix, keys.data(), posting_offsets.data(), postings.data(), info.n_terms, info.n_postings, ix.stream);
    return MGX_OK;
  });
}

int mgx_index_doc_lengths(const mgx_index_t* index, uint32_t* out) {
  if (index == nullptr || out == nullptr) {
    return invalid("null argument");
  }
  mgx_index_t* h = const_cast<mgx_index_t*>(index);
  if (int rc = commit_pending(h); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(h);
    DeviceGuard guard(h->ix.device);
    if (h->ix.n_docs > 0) {
      MGX_CUDA(cudaMemcpy(out, h->ix.d_doc_len.p, h->ix.n_docs * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    return MGX_OK;
  });
}

// ---------------------------------------------------------------- batched pipeline
int mgx_batch_prepare(mgx_index_t* index, const mgx_query_params_t* params, uint64_t n_queries,
                      const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                      const uint8_t* not_bytes, const uint64_t* not_offsets, const uint64_t* q_not_begin,
                      void* stream, mgx_batch_t** out) {
  return mgx_batch_prepare_ex(index, params, n_queries, term_bytes, term_offsets, q_term_begin, not_bytes,
                              not_offsets, q_not_begin, nullptr, stream, out);
}

namespace {
// Host compile + upload of one batch on `stream`. The caller has committed pending mutations and holds shared access
// to the index (its own Reader, or the one a staged batch keeps until mgx_batch_destroy).
std::unique_ptr<mgx_batch> take_pooled_batch(Index& ix) {
  std::unique_ptr<mgx_batch> h;
  {
    std::lock_guard<std::mutex> pool_lock(ix.pool_mu);
    if (!ix.batch_pool.empty()) {
      h.reset(static_cast<mgx_batch*>(ix.batch_pool.back()));
      ix.batch_pool.pop_back();
    }
  }
  if (!h) {
    h = std::make_unique<mgx_batch>();
  }
  return h;
}

int prepare_unlocked(mgx_index_t* index, const mgx_query_params_t* params, uint64_t n_queries,
                     const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                     const uint8_t* not_bytes, const uint64_t* not_offsets, const uint64_t* q_not_begin,
                     const mgx_query_ext_t* ext, cudaStream_t stream, mgx_batch_t** out) {
  Index& ix = index->ix;
  DeviceGuard guard(ix.device);
  WideScope wide_scope;
  std::unique_ptr<mgx_batch> h = take_pooled_batch(ix);
  Batch& b = h->b;
  b.ix = &ix;
  b.params = *params;
  b.stream = stream;  // NULL = the legacy default stream, as documented
  b.launches_at_start = g_launches.load();
  // the compile workspace stays with the pooled batch object (capacity is kept from batch to batch)
  std::vector<HostTerm>& terms = b.h_terms;
  std::vector<HostQuery>& queries = b.h_queries;
  std::vector<uint32_t>& slot_tid = b.h_slot_scratch;
  terms.clear();
  static const uint8_t kEmpty[1] = {0};
  static const bool trace = std::getenv("MGX_BATCH_TRACE") != nullptr;  // host wall time of the two host stages
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = compile_batch(ix, *params, n_queries, term_bytes != nullptr ? term_bytes : kEmpty, term_offsets,
                               q_term_begin, not_bytes != nullptr ? not_bytes : kEmpty, not_offsets, q_not_begin, ext,
                               &terms, &queries, &slot_tid);
  if (rc != MGX_OK) {
    std::lock_guard<std::mutex> pool_lock(ix.pool_mu);
    if (ix.batch_pool.size() < 16) {
      ix.batch_pool.push_back(h.release());  // nothing was enqueued: the workspace goes straight back
    }
    return rc;
  }
  const auto t1 = std::chrono::steady_clock::now();
  std::vector<HostQuery> expanded;
  std::vector<uint32_t> xoff;
  const bool no_expand = std::getenv("MGX_NO_OR_EXPANSION") != nullptr;  // read per batch: the tests flip it
  if (ext != nullptr && ext->q_prog_begin != nullptr && params->compute_score == 0 && !no_expand &&
      expand_or_roots(terms, queries, &expanded, &xoff)) {
    batch_upload(b, terms, expanded, slot_tid, &xoff);
  } else {
    batch_upload(b, terms, queries, slot_tid);
  }
  if (trace) {
    const auto t2 = std::chrono::steady_clock::now();
    fprintf(stderr, "[mgx batch] compile %.3f ms, stage+upload %.3f ms (%zu unique terms)\n",
            std::chrono::duration<double, std::milli>(t1 - t0).count(),
            std::chrono::duration<double, std::milli>(t2 - t1).count(), terms.size());
  }
  *out = h.release();
  return MGX_OK;
}

void release_batch(mgx_batch_t* batch) {
  Index& ix = *batch->b.ix;
  mgx_index* reader_of = batch->reader_of;
  batch->reader_of = nullptr;
  batch->b.recycle();
  {
    std::lock_guard<std::mutex> pool_lock(ix.pool_mu);
    if (ix.batch_pool.size() < 16) {
      ix.batch_pool.push_back(batch);  // keep the workspace for the next batch
      batch = nullptr;
    }
  }
  delete batch;
  if (reader_of != nullptr) {
    reader_of->gate.unlock_shared();
  }
}
}  // namespace

int mgx_batch_prepare_ex(mgx_index_t* index, const mgx_query_params_t* params, uint64_t n_queries,
                         const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                         const uint8_t* not_bytes, const uint64_t* not_offsets, const uint64_t* q_not_begin,
                         const mgx_query_ext_t* ext, void* stream, mgx_batch_t** out) {
  if (index == nullptr || params == nullptr || out == nullptr ||
      (n_queries > 0 && (term_offsets == nullptr || q_term_begin == nullptr))) {
    return invalid("null argument");
  }
  *out = nullptr;
  if (n_queries >= (1ULL << 31)) {
    return invalid("too many queries in one batch");
  }
  if (int rc = check_params(*params); rc != MGX_OK) {
    return rc;
  }
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  // A staged batch reads the index from here until mgx_batch_destroy: it keeps shared access for that long, so a
  // commit of journaled mutations (or a rebuild) waits for it instead of releasing the arrays under its kernels.
  index->gate.lock_shared();
  const int rc = guarded([&]() {
    return prepare_unlocked(index, params, n_queries, term_bytes, term_offsets, q_term_begin, not_bytes, not_offsets,
                            q_not_begin, ext, static_cast<cudaStream_t>(stream), out);
  });
  if (rc != MGX_OK || *out == nullptr) {
    index->gate.unlock_shared();
    return rc;
  }
  (*out)->reader_of = index;
  return MGX_OK;
}

// ---------------------------------------------------------------- compiled batches shared between shard processes
//
// Every shard answers every query, so without this every rank of a node compiles every batch (hash the terms, cut the
// n-grams, lay the batch out): at 8 GPUs that host work, not the device, bounded the end-to-end rate. The channel is a
// POSIX shared-memory ring: the rank whose turn it is compiles the batch once and publishes the staging bytes, the
// others copy them into their own pinned staging buffer and bind them (one H2D copy, as after a local compile).
namespace {
constexpr uint64_t kShareMagic = 0x317261685378676dULL;  // "mgxShar1"
struct ShareHeader {
  std::atomic<uint64_t> magic;
  uint32_t n_ranks, n_slots;
  uint64_t slot_bytes;   // payload capacity of a slot
  uint64_t slot_stride;  // header + payload, 4 KB aligned
  std::atomic<uint32_t> attached;
  uint32_t pad[7];
};
struct ShareSlot {
  std::atomic<uint64_t> published;     // sequence number + 1 of the batch the slot holds, 0 = never used
  std::atomic<uint32_t> readers_left;  // importers that have not copied it yet
  int32_t status;                      // MGX_OK, or why the publisher could not share the batch (importers compile locally)
  uint64_t config_sig;                 // n-gram configuration of the compiling index
  uint64_t bytes;
  StageLayout layout;
};
static_assert(std::atomic<uint64_t>::is_always_lock_free, "the ring relies on lock-free 64-bit atomics in shared memory");

uint64_t index_config_sig(const Index& ix) {
  return (static_cast<uint64_t>(ix.ngram) << 32) | (static_cast<uint64_t>(ix.kanji) << 16) |
         (static_cast<uint64_t>(ix.width) << 8) | (ix.cross ? 1u : 0u);
}
}  // namespace

struct mgx_share {
  std::string name;
  int n_ranks = 0, rank = 0;
  uint8_t* base = nullptr;
  size_t map_bytes = 0;
  ShareHeader* hdr = nullptr;
  bool owner = false;
  ShareSlot* slot(uint64_t seq) const {
    return reinterpret_cast<ShareSlot*>(base + 4096 + (seq % hdr->n_slots) * hdr->slot_stride);
  }
  uint8_t* payload(uint64_t seq) const { return reinterpret_cast<uint8_t*>(slot(seq)) + 512; }
};
static_assert(sizeof(ShareSlot) <= 512, "slot header");

namespace {
extern "C++" template <class Pred>
bool spin_until(Pred done, int32_t timeout_ms) {
  const auto t0 = std::chrono::steady_clock::now();
  for (uint32_t spins = 0; !done(); ++spins) {
    if (spins < 2000) {
      continue;
    }
    std::this_thread::sleep_for(std::chrono::microseconds(20));
    if (timeout_ms >= 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(timeout_ms)) {
      return false;
    }
  }
  return true;
}
}  // namespace

int mgx_share_open(const char* name, int32_t n_ranks, int32_t rank, int32_t n_slots, uint64_t slot_bytes,
                   mgx_share_t** out) {
  if (name == nullptr || out == nullptr || n_ranks < 1 || rank < 0 || rank >= n_ranks || n_slots < 1 || n_slots > 64 ||
      slot_bytes == 0) {
    return invalid("mgx_share_open: bad argument");
  }
  *out = nullptr;
  return guarded([&]() {
    auto sh = std::make_unique<mgx_share>();
    sh->name = name;
    sh->n_ranks = n_ranks;
    sh->rank = rank;
    sh->owner = rank == 0;
    const uint64_t stride = (512 + slot_bytes + 4095) / 4096 * 4096;
    const size_t total = 4096 + static_cast<size_t>(n_slots) * stride;
    int fd = -1;
    if (sh->owner) {
      shm_unlink(name);
      fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
      if (fd < 0 || ftruncate(fd, static_cast<off_t>(total)) != 0) {
        if (fd >= 0) {
          close(fd);
          shm_unlink(name);
        }
        set_last_error("mgx_share_open: cannot create the shared-memory segment");
        return MGX_ERR_UNSUPPORTED;
      }
    } else {
      const auto t0 = std::chrono::steady_clock::now();
      struct stat sb {};
      // rank 0 sizes the segment before it writes the header, and writes the magic last
      while ((fd = shm_open(name, O_RDWR, 0600)) < 0 || fstat(fd, &sb) != 0 || sb.st_size < 4096) {
        if (fd >= 0) {
          close(fd);
          fd = -1;
        }
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) {
          set_last_error("mgx_share_open: rank 0 did not create the segment within 60 s");
          return MGX_ERR_TIMEOUT;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
      }
      void* hm = mmap(nullptr, 4096, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (hm == MAP_FAILED) {
        close(fd);
        set_last_error("mgx_share_open: mmap failed");
        return MGX_ERR_UNSUPPORTED;
      }
      const ShareHeader* hd = static_cast<const ShareHeader*>(hm);
      const bool up = spin_until([&]() { return hd->magic.load(std::memory_order_acquire) == kShareMagic; }, 60000);
      const bool same = up && hd->n_ranks == static_cast<uint32_t>(n_ranks) &&
                        hd->n_slots == static_cast<uint32_t>(n_slots) && hd->slot_bytes == slot_bytes;
      munmap(hm, 4096);
      if (!same) {
        close(fd);
        set_last_error(up ? "mgx_share_open: the segment was created with other parameters"
                          : "mgx_share_open: rank 0 did not initialise the segment within 60 s");
        return up ? MGX_ERR_INVALID_ARGUMENT : MGX_ERR_TIMEOUT;
      }
    }
    void* m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) {
      set_last_error("mgx_share_open: mmap failed");
      return MGX_ERR_UNSUPPORTED;
    }
    sh->base = static_cast<uint8_t*>(m);
    sh->map_bytes = total;
    sh->hdr = reinterpret_cast<ShareHeader*>(m);
    if (sh->owner) {  // a fresh segment is zero-filled: every slot starts unpublished with no readers pending
      sh->hdr->n_ranks = static_cast<uint32_t>(n_ranks);
      sh->hdr->n_slots = static_cast<uint32_t>(n_slots);
      sh->hdr->slot_bytes = slot_bytes;
      sh->hdr->slot_stride = stride;
      sh->hdr->magic.store(kShareMagic, std::memory_order_release);
    }
    sh->hdr->attached.fetch_add(1);
    *out = sh.release();
    return MGX_OK;
  });
}

void mgx_share_close(mgx_share_t* share) {
  if (share == nullptr) {
    return;
  }
  if (share->base != nullptr) {
    munmap(share->base, share->map_bytes);
  }
  if (share->owner) {
    shm_unlink(share->name.c_str());
  }
  delete share;
}

int mgx_share_publish(mgx_share_t* share, uint64_t seq, const mgx_batch_t* batch, int32_t timeout_ms) {
  if (share == nullptr || batch == nullptr) {
    return invalid("null argument");
  }
  return guarded([&]() {
    ShareSlot* sl = share->slot(seq);
    // the slot's previous batch (seq - n_slots) must have been copied by every importer
    if (!spin_until([&]() { return sl->readers_left.load(std::memory_order_acquire) == 0; }, timeout_ms)) {
      set_last_error("mgx_share_publish: the other ranks did not take the slot's previous batch in time");
      return MGX_ERR_TIMEOUT;
    }
    const Batch& b = batch->b;
    const StageOffsets O = stage_offsets(b.layout);
    int status = MGX_OK;
    if (b.layout.n_preds != 0 || b.explicit_driver.d_ids != nullptr) {
      status = MGX_ERR_UNSUPPORTED;  // predicates / driver sets hold device addresses of this shard
    } else if (O.total > share->hdr->slot_bytes) {
      status = MGX_ERR_CAPACITY;
    }
    sl->status = status;
    sl->layout = b.layout;
    sl->config_sig = index_config_sig(*b.ix);
    sl->bytes = status == MGX_OK ? O.total : 0;
    if (status == MGX_OK) {
      std::memcpy(share->payload(seq), b.staging.p, O.total);
    }
    sl->readers_left.store(static_cast<uint32_t>(share->n_ranks - 1), std::memory_order_relaxed);
    sl->published.store(seq + 1, std::memory_order_release);
    if (status != MGX_OK) {
      set_last_error(status == MGX_ERR_CAPACITY ? "mgx_share_publish: the compiled batch exceeds the slot size"
                                                 : "mgx_share_publish: batches with column conditions or an explicit "
                                                   "driver set cannot be shared");
    }
    return status;
  });
}

int mgx_share_import(mgx_share_t* share, uint64_t seq, mgx_index_t* index, const mgx_query_params_t* params,
                     void* stream, int32_t timeout_ms, mgx_batch_t** out) {
  if (share == nullptr || index == nullptr || params == nullptr || out == nullptr) {
    return invalid("null argument");
  }
  *out = nullptr;
  if (int rc = check_params(*params); rc != MGX_OK) {
    return rc;
  }
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  ShareSlot* sl = share->slot(seq);
  if (!spin_until([&]() { return sl->published.load(std::memory_order_acquire) == seq + 1; }, timeout_ms)) {
    set_last_error("mgx_share_import: the batch was not published in time");
    return MGX_ERR_TIMEOUT;
  }
  index->gate.lock_shared();
  const int rc = guarded([&]() {
    Index& ix = index->ix;
    if (sl->status != MGX_OK) {
      set_last_error("mgx_share_import: the publisher could not share this batch; compile it locally");
      return static_cast<int>(sl->status);
    }
    if (sl->config_sig != index_config_sig(ix)) {
      return invalid("mgx_share_import: the batch was compiled for another n-gram configuration");
    }
    DeviceGuard guard(ix.device);
    std::unique_ptr<mgx_batch> h = take_pooled_batch(ix);
    Batch& b = h->b;
    b.ix = &ix;
    b.params = *params;
    b.stream = static_cast<cudaStream_t>(stream);
    b.launches_at_start = g_launches.load();
    batch_import(b, sl->layout, share->payload(seq));
    *out = h.release();
    return MGX_OK;
  });
  sl->readers_left.fetch_sub(1, std::memory_order_acq_rel);  // taken (or refused): the slot is free for this rank
  if (rc != MGX_OK || *out == nullptr) {
    index->gate.unlock_shared();
    return rc;
  }
  (*out)->reader_of = index;
  return MGX_OK;
}

int mgx_batch_set_streamed(mgx_batch_t* batch, int32_t enable) {
  if (batch == nullptr) {
    return invalid("null batch");
  }
  if (batch->b.planned) {
    return invalid("mgx_batch_set_streamed must precede the planning stage");
  }
  batch->b.allow_streamed = enable != 0;
  return MGX_OK;
}

int mgx_batch_plan_device(mgx_batch_t* batch) {
  if (batch == nullptr) {
    return invalid("null batch");
  }
  return guarded([&]() {
    DeviceGuard guard(batch->b.ix->device);
    if (!batch->b.planned) {
      batch_plan(batch->b);
    }
    return MGX_OK;
  });
}

int mgx_batch_get_stats(mgx_batch_t* batch, mgx_batch_stats_t* out) {
  if (batch == nullptr || out == nullptr) {
    return invalid("null argument");
  }
  return guarded([&]() {
    DeviceGuard guard(batch->b.ix->device);
    batch->b.collect_stats(out);
    out->launches = g_launches.load() - batch->b.launches_at_start;
    return MGX_OK;
  });
}

uint64_t mgx_batch_term_slots(const mgx_batch_t* batch) { return batch != nullptr ? batch->b.n_slots : 0; }

int mgx_batch_df_device(mgx_batch_t* batch, uint64_t* d_df) {
  if (batch == nullptr) {
    return invalid("null batch");
  }
  return guarded([&]() {
    Batch& b = batch->b;
    DeviceGuard guard(b.ix->device);
    if (!b.planned) {
      batch_plan(b);
    }
    if (b.params.compute_score != 0) {
      batch_df(b);
    }
    if (d_df != nullptr) {
      batch_df_to_slots(b, d_df);
    }
    return MGX_OK;
  });
}

namespace {
// a shard returns its best (offset + limit) records un-offset; the merge applies the offset
struct ShardParams {
  Batch& b;
  mgx_query_params_t saved;
  explicit ShardParams(Batch& batch) : b(batch), saved(batch.params) {
    if (b.params.limit != 0) {
      b.params.limit = saved.limit + saved.offset;
    }
    b.params.offset = 0;
  }
  ~ShardParams() { b.params = saved; }
};

int check_stride(const mgx_query_params_t& p, uint64_t stride) {
  if (p.limit == 0 && p.compute_score != 0) {
    set_last_error("sharded _score batches need an explicit limit: a shard returns its best limit + offset records, "
                   "and 'everything' (limit 0) has no bound to size the exchange with");
    return MGX_ERR_UNSUPPORTED;
  }
  if (p.limit != 0 && stride < static_cast<uint64_t>(p.limit) + p.offset) {
    set_last_error("stride must hold limit + offset records per query: a shard returns its best limit + offset records "
                   "un-offset and the merge skips the offset");
    return MGX_ERR_INVALID_ARGUMENT;
  }
  return MGX_OK;
}
}  // namespace

int mgx_batch_search_device(mgx_batch_t* batch, const uint64_t* d_df, uint64_t stride, uint32_t* d_ids,
                            double* d_scores, uint32_t* d_count, uint64_t* d_total) {
  if (batch == nullptr || d_ids == nullptr || d_count == nullptr || d_total == nullptr) {
    return invalid("null argument");
  }
  if (int rc = check_stride(batch->b.params, stride); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Batch& b = batch->b;
    DeviceGuard guard(b.ix->device);
    if (b.params.compute_score != 0 && !b.df_done) {
      batch_df(b);
    }
    ShardParams shard(b);
    batch_search(b, d_df, stride, d_ids, d_scores, d_count, d_total);
    return MGX_OK;
  });
}

int mgx_batch_overflowed(mgx_batch_t* batch, int32_t* out_overflowed) {
  if (batch == nullptr || out_overflowed == nullptr) {
    return invalid("null argument");
  }
  return guarded([&]() {
    Batch& b = batch->b;
    DeviceGuard guard(b.ix->device);
    *out_overflowed = 0;
    if (b.streamed && b.searched) {
      if (b.ev_last != nullptr) {
        MGX_CUDA(cudaEventSynchronize(b.ev_last));
      } else {
        MGX_CUDA(cudaStreamSynchronize(b.stream));
      }
      *out_overflowed = batch_overflowed(b) ? 1 : 0;
    }
    return MGX_OK;
  });
}

int mgx_batch_reset(mgx_batch_t* batch) {
  if (batch == nullptr) {
    return invalid("null batch");
  }
  return guarded([&]() {
    Batch& b = batch->b;
    DeviceGuard guard(b.ix->device);
    // after an overflow the repeat takes the synchronous form; otherwise this is a plain re-arm of the batch
    batch_reset_for_repeat(b, batch_overflowed(b));
    b.sharded_enqueued = false;
    return MGX_OK;
  });
}

void mgx_batch_destroy(mgx_batch_t* batch) {
  if (batch == nullptr) {
    return;
  }
  Index& ix = *batch->b.ix;
  DeviceGuard guard(ix.device);
  // wait for THIS batch only (its last kernel is followed by ev_last), not for whatever the caller has queued
  // behind it on the same stream: a pipelined caller releases batch i while batch i+1 is already running
  if (batch->b.ev_last != nullptr && batch->b.searched) {
    cudaEventSynchronize(batch->b.ev_last);
  } else {
    cudaStreamSynchronize(batch->b.stream);
  }
  release_batch(batch);
}

int mgx_merge_topk_device(int32_t device, void* stream, const mgx_query_params_t* params, uint32_t n_shards,
                          uint64_t n_queries, uint64_t stride, const uint32_t* d_ids_all, const double* d_scores_all,
                          const uint32_t* d_count_all, const uint64_t* d_total_all, uint32_t* d_ids_out,
                          double* d_scores_out, uint32_t* d_count_out, uint64_t* d_total_out) {
  if (params == nullptr || d_ids_all == nullptr || d_count_all == nullptr || d_total_all == nullptr ||
      d_ids_out == nullptr || d_count_out == nullptr || d_total_out == nullptr) {
    return invalid("null argument");
  }
  if (int rc = require_device(); rc != MGX_OK) {
    return rc;
  }
  if (int rc = check_stride(*params, stride); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    DeviceGuard guard(device);
    launch_merge_topk(static_cast<cudaStream_t>(stream), *params, n_shards, n_queries, stride, d_ids_all, d_scores_all,
                      d_count_all, d_total_all, 0, d_ids_out, d_scores_out, d_count_out, d_total_out);
    return MGX_OK;
  });
}

int mgx_shard_record_layout(uint64_t n_queries, uint64_t stride, mgx_shard_record_layout_t* out) {
  if (out == nullptr) {
    return invalid("null argument");
  }
  // 8-byte parts first, so every part is aligned for its element type; the record ends with a 16-byte status block
  out->scores_offset = 0;
  out->total_offset = n_queries * stride * sizeof(double);
  out->ids_offset = out->total_offset + n_queries * sizeof(uint64_t);
  out->count_offset = out->ids_offset + n_queries * stride * sizeof(uint32_t);
  out->status_offset = (out->count_offset + n_queries * sizeof(uint32_t) + 15) & ~static_cast<uint64_t>(15);
  out->bytes = out->status_offset + 16;
  return MGX_OK;
}

int mgx_batch_search_packed_device(mgx_batch_t* batch, const uint64_t* d_df, uint64_t stride, void* d_record) {
  if (batch == nullptr || d_record == nullptr) {
    return invalid("null argument");
  }
  mgx_shard_record_layout_t lay;
  mgx_shard_record_layout(batch->b.n_out_queries, stride, &lay);
  uint8_t* base = static_cast<uint8_t*>(d_record);
  const int rc = mgx_batch_search_device(batch, d_df, stride, reinterpret_cast<uint32_t*>(base + lay.ids_offset),
                                         reinterpret_cast<double*>(base + lay.scores_offset),
                                         reinterpret_cast<uint32_t*>(base + lay.count_offset),
                                         reinterpret_cast<uint64_t*>(base + lay.total_offset));
  if (rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Batch& b = batch->b;
    DeviceGuard guard(b.ix->device);
    // status block: [0] = 1 when this shard's streamed batch did not fit its workspace (nothing was computed)
    if (b.streamed) {
      MGX_CUDA(cudaMemcpyAsync(base + lay.status_offset, b.d_launch.p + kLaunchOverflow, sizeof(uint32_t),
                               cudaMemcpyDeviceToDevice, b.stream));
    } else {
      MGX_CUDA(cudaMemsetAsync(base + lay.status_offset, 0, 16, b.stream));
    }
    b.mark_last();
    return MGX_OK;
  });
}

int mgx_merge_topk_packed_device(int32_t device, void* stream, const mgx_query_params_t* params, uint32_t n_shards,
                                 uint64_t n_queries, uint64_t stride, const void* d_records, void* d_record_out) {
  if (params == nullptr || d_records == nullptr || d_record_out == nullptr) {
    return invalid("null argument");
  }
  if (int rc = require_device(); rc != MGX_OK) {
    return rc;
  }
  if (int rc = check_stride(*params, stride); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    DeviceGuard guard(device);
    mgx_shard_record_layout_t lay;
    mgx_shard_record_layout(n_queries, stride, &lay);
    const uint8_t* in = static_cast<const uint8_t*>(d_records);
    uint8_t* out = static_cast<uint8_t*>(d_record_out);
    launch_merge_topk(static_cast<cudaStream_t>(stream), *params, n_shards, n_queries, stride,
                      reinterpret_cast<const uint32_t*>(in + lay.ids_offset),
                      reinterpret_cast<const double*>(in + lay.scores_offset),
                      reinterpret_cast<const uint32_t*>(in + lay.count_offset),
                      reinterpret_cast<const uint64_t*>(in + lay.total_offset), lay.bytes,
                      reinterpret_cast<uint32_t*>(out + lay.ids_offset), reinterpret_cast<double*>(out + lay.scores_offset),
                      reinterpret_cast<uint32_t*>(out + lay.count_offset),
                      reinterpret_cast<uint64_t*>(out + lay.total_offset));
    // status of the merged record: any shard that overflowed
    launch_or_status(static_cast<cudaStream_t>(stream), in + lay.status_offset, lay.bytes, n_shards,
                     out + lay.status_offset);
    return MGX_OK;
  });
}

int mgx_query_batch(mgx_index_t* index, const mgx_query_params_t* params, uint64_t n_queries,
                    const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                    const uint8_t* not_bytes, const uint64_t* not_offsets, const uint64_t* q_not_begin,
                    uint64_t stride, uint32_t* out_ids, double* out_scores, uint32_t* out_count, uint64_t* out_total,
                    uint64_t* out_df) {
  return mgx_query_batch_ex(index, params, n_queries, term_bytes, term_offsets, q_term_begin, not_bytes, not_offsets,
                            q_not_begin, nullptr, stride, out_ids, out_scores, out_count, out_total, out_df);
}

int mgx_query_batch_ex(mgx_index_t* index, const mgx_query_params_t* params, uint64_t n_queries,
                       const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                       const uint8_t* not_bytes, const uint64_t* not_offsets, const uint64_t* q_not_begin,
                       const mgx_query_ext_t* ext, uint64_t stride, uint32_t* out_ids, double* out_scores,
                       uint32_t* out_count, uint64_t* out_total, uint64_t* out_df) {
  if (index == nullptr || params == nullptr || out_ids == nullptr || out_count == nullptr || out_total == nullptr ||
      (n_queries > 0 && (term_offsets == nullptr || q_term_begin == nullptr))) {
    return invalid("null argument");
  }
  if (params->compute_score != 0 && out_scores == nullptr) {
    return invalid("out_scores is required for SORT _score");
  }
  if (n_queries == 0) {
    return MGX_OK;
  }
  if (n_queries >= (1ULL << 31)) {
    return invalid("too many queries in one batch");
  }
  if (int rc = check_params(*params); rc != MGX_OK) {
    return rc;
  }
  // pending mutations first and OUTSIDE the shared access (the commit needs the index exclusively); a mutation
  // journaled by another thread after this point is seen by the next call
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(index);
    mgx_batch_t* batch = nullptr;
    if (int rc = prepare_unlocked(index, params, n_queries, term_bytes, term_offsets, q_term_begin, not_bytes, not_offsets,
                                  q_not_begin, ext, rd.stream(), &batch);
        rc != MGX_OK) {
      return rc;
    }
    struct Release {
      mgx_batch_t* b;
      ~Release() {
        cudaStreamSynchronize(b->b.stream);
        release_batch(b);
      }
    } release{batch};
    Batch& b = batch->b;
    Index& ix = *b.ix;
    DeviceGuard guard(ix.device);
    cudaStream_t st = b.stream;
    DevBuf<uint32_t>& d_ids = b.o_ids;
    DevBuf<double>& d_scores = b.o_scores;
    DevBuf<uint32_t>& d_count = b.o_count;
    DevBuf<uint64_t>& d_total = b.o_total;
    DevBuf<uint64_t>& d_df = b.o_df;
    d_ids.reserve(n_queries * stride);
    d_scores.reserve(params->compute_score != 0 ? n_queries * stride : 1);
    d_count.reserve(n_queries);
    d_total.reserve(n_queries);
    d_df.reserve(b.n_slots);
    b.allow_streamed = true;  // planned without a host read-back; repeated below if it does not fit the workspace
    for (int attempt = 0; attempt < 2; ++attempt) {
      batch_plan(b);
      if (params->compute_score != 0) {
        batch_df(b);
      }
      batch_df_to_slots(b, d_df.p);
      batch_search(b, nullptr, stride, d_ids.p, params->compute_score != 0 ? d_scores.p : nullptr, d_count.p, d_total.p);
      uint64_t d2h = 0;
      MGX_CUDA(cudaMemcpyAsync(out_ids, d_ids.p, n_queries * stride * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      d2h += n_queries * stride * sizeof(uint32_t);
      if (params->compute_score != 0) {
        MGX_CUDA(cudaMemcpyAsync(out_scores, d_scores.p, n_queries * stride * sizeof(double), cudaMemcpyDeviceToHost, st));
        d2h += n_queries * stride * sizeof(double);
      }
      MGX_CUDA(cudaMemcpyAsync(out_count, d_count.p, n_queries * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      MGX_CUDA(cudaMemcpyAsync(out_total, d_total.p, n_queries * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
      d2h += n_queries * 12;
      if (out_df != nullptr && b.n_slots > 0) {
        MGX_CUDA(cudaMemcpyAsync(out_df, d_df.p, b.n_slots * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        d2h += b.n_slots * sizeof(uint64_t);
      }
      b.mark_last();
      MGX_CUDA(cudaStreamSynchronize(st));
      b.d2h_bytes += d2h;
      if (!batch_overflowed(b)) {
        break;
      }
      batch_reset_for_repeat(b, true);
    }
    mgx_batch_stats_t stats{};
    b.collect_stats(&stats);
    stats.launches = g_launches.load() - b.launches_at_start;
    {
      std::lock_guard<std::mutex> sl(index->stats_mu);
      ix.last_stats = stats;
    }
    return MGX_OK;
  });
}

// ---------------------------------------------------------------- sharded pipeline behind the C ABI
// One process per GPU, one doc-id range each. NCCL is looked up at run time (dlopen of libnccl.so.2: the copy a
// hosting PyTorch process has already loaded, or the system library for a plain C++ host), so libmgx.so has no link
// dependency on it and single-GPU hosts never touch it.
namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    const char* override_path = std::getenv("MGX_NCCL_LIB");
    const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    for (const char* name : names) {
      if (name == nullptr) {
        continue;
      }
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib != nullptr) {
        break;
      }
    }
    if (api.lib == nullptr) {
      api.error = std::string("libnccl.so.2 not found: ") + (dlerror() != nullptr ? dlerror() : "");
      return;
    }
    auto sym = [&](const char* name) {
      void* s = dlsym(api.lib, name);
      if (s == nullptr) {
        api.error = std::string("NCCL symbol missing: ") + name;
      }
      return s;
    };
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &api;
}

int nccl_ready(NcclApi** out) {
  NcclApi* api = nccl_api();
  if (api->lib == nullptr || !api->error.empty()) {
    set_last_error("NCCL is not available: " + api->error);
    return MGX_ERR_UNSUPPORTED;
  }
  *out = api;
  return MGX_OK;
}

#define MGX_NCCL(api, expr)                                                                       \
  do {                                                                                            \
    ncclResult_t mgx_nccl_rc__ = (expr);                                                          \
    if (mgx_nccl_rc__ != ncclSuccess) {                                                           \
      ::mgx::set_last_error(std::string(#expr) + ": " + (api)->GetErrorString(mgx_nccl_rc__));    \
      throw ::mgx::CudaFailure{MGX_ERR_CUDA};                                                     \
    }                                                                                             \
  } while (0)
}  // namespace

struct mgx_shard_comm {
  int n_ranks = 1;
  int rank = 0;
  int device = 0;
  int n_lanes = 0;
  ncclComm_t comm[MGX_COMM_MAX_LANES] = {nullptr};
  cudaStream_t stream[MGX_COMM_MAX_LANES] = {nullptr};  // highest priority: the exchanges are latency-sized and must
                                                        // not queue behind another batch's grid-filling kernels
};

int mgx_comm_unique_id(uint8_t* out_id) {
  if (out_id == nullptr) {
    return invalid("null argument");
  }
  NcclApi* api = nullptr;
  if (int rc = nccl_ready(&api); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == MGX_COMM_ID_BYTES, "unique id size");
    MGX_NCCL(api, api->GetUniqueId(&id));
    std::memcpy(out_id, &id, sizeof(id));
    return MGX_OK;
  });
}

int mgx_comm_create(const uint8_t* ids, int32_t n_lanes, int32_t n_ranks, int32_t rank, int32_t device,
                    mgx_shard_comm_t** out) {
  if (ids == nullptr || out == nullptr) {
    return invalid("null argument");
  }
  *out = nullptr;
  if (n_lanes < 1 || n_lanes > MGX_COMM_MAX_LANES || n_ranks < 1 || rank < 0 || rank >= n_ranks) {
    return invalid("mgx_comm_create: bad lane / rank arguments");
  }
  if (int rc = require_device(); rc != MGX_OK) {
    return rc;
  }
  NcclApi* api = nullptr;
  if (int rc = nccl_ready(&api); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    DeviceGuard guard(device);
    auto c = std::make_unique<mgx_shard_comm>();
    c->n_ranks = n_ranks;
    c->rank = rank;
    c->device = device;
    int least = 0;
    int greatest = 0;
    MGX_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    for (int l = 0; l < n_lanes; ++l) {
      ncclUniqueId id;
      std::memcpy(&id, ids + static_cast<size_t>(l) * MGX_COMM_ID_BYTES, sizeof(id));
      MGX_NCCL(api, api->CommInitRank(&c->comm[l], n_ranks, id, rank));
      MGX_CUDA(cudaStreamCreateWithPriority(&c->stream[l], cudaStreamNonBlocking, greatest));
      c->n_lanes = l + 1;
    }
    *out = c.release();
    return MGX_OK;
  });
}

void mgx_comm_destroy(mgx_shard_comm_t* comm) {
  if (comm == nullptr) {
    return;
  }
  NcclApi* api = nccl_api();
  DeviceGuard guard(comm->device);
  for (int l = 0; l < comm->n_lanes; ++l) {
    if (comm->stream[l] != nullptr) {
      cudaStreamSynchronize(comm->stream[l]);
    }
    if (comm->comm[l] != nullptr && api->CommDestroy != nullptr) {
      api->CommDestroy(comm->comm[l]);
    }
    if (comm->stream[l] != nullptr) {
      cudaStreamDestroy(comm->stream[l]);
    }
  }
  delete comm;
}

int mgx_comm_info(const mgx_shard_comm_t* comm, int32_t* n_ranks, int32_t* rank, int32_t* n_lanes,
                  int32_t* nccl_version) {
  if (n_ranks != nullptr) {
    *n_ranks = comm != nullptr ? comm->n_ranks : 1;
  }
  if (rank != nullptr) {
    *rank = comm != nullptr ? comm->rank : 0;
  }
  if (n_lanes != nullptr) {
    *n_lanes = comm != nullptr ? comm->n_lanes : 0;
  }
  if (nccl_version != nullptr) {
    *nccl_version = 0;
    NcclApi* api = nccl_api();
    if (api->GetVersion != nullptr) {
      int v = 0;
      if (api->GetVersion(&v) == ncclSuccess) {
        *nccl_version = v;
      }
    }
  }
  return MGX_OK;
}

namespace {
cudaEvent_t batch_event(Batch& b, int i) {
  if (b.ev_x[i] == nullptr) {
    MGX_CUDA(cudaEventCreateWithFlags(&b.ev_x[i], cudaEventDisableTiming));
  }
  return b.ev_x[i];
}

// Everything of one batch, enqueued without a host synchronisation when the batch is streamed:
//   plan -> df -> [all-reduce of the per-term df] -> search -> [all-gather of the packed records] -> merge -> D2H
void sharded_enqueue(mgx_shard_comm_t* comm, int lane, Batch& b, uint64_t stride, void* h_record_out) {
  const int G = comm != nullptr ? comm->n_ranks : 1;
  const int rank = comm != nullptr ? comm->rank : 0;
  cudaStream_t st = b.stream;
  mgx_shard_record_layout_t lay;
  mgx_shard_record_layout(b.n_out_queries, stride, &lay);
  b.o_merged.reserve(lay.bytes);
  if (G > 1) {
    b.o_gather.reserve(lay.bytes * static_cast<uint64_t>(G));
  }
  batch_plan(b);
  if (b.params.compute_score != 0) {
    batch_df(b);
  }
  NcclApi* api = nullptr;
  if (G > 1) {
    if (int rc = nccl_ready(&api); rc != MGX_OK) {
      throw CudaFailure{rc};
    }
    if (b.params.compute_score != 0 && b.n_terms > 0) {
      // exchange 1: every shard scores with the GLOBAL document frequencies. The batch was compiled from the same
      // input on every rank, so the per-term array itself is reduced, in place.
      cudaStream_t cs = comm->stream[lane];
      MGX_CUDA(cudaEventRecord(batch_event(b, 0), st));
      MGX_CUDA(cudaStreamWaitEvent(cs, batch_event(b, 0), 0));
      // Behind it in the same array: every key's posting size, which gives each shard the GLOBAL estimated sizes
      // the reference orders (and therefore sums) a query's terms by.
      MGX_NCCL(api, api->AllReduce(b.d_t_df.p, b.d_t_df.p, static_cast<size_t>(b.n_terms) + b.n_keys, ncclUint64,
                                   ncclSum, comm->comm[lane], cs));
      b.global_order = true;
      MGX_CUDA(cudaEventRecord(batch_event(b, 1), cs));
      MGX_CUDA(cudaStreamWaitEvent(st, batch_event(b, 1), 0));
    }
  }
  uint8_t* rec = G > 1 ? b.o_gather.p + lay.bytes * static_cast<uint64_t>(rank) : b.o_merged.p;
  {
    // a shard returns its best (offset + limit) records un-offset and the merge applies the offset; a single shard
    // answers with the caller's own window
    mgx_query_params_t saved = b.params;
    if (G > 1) {
      if (b.params.limit != 0) {
        b.params.limit = saved.limit + saved.offset;
      }
      b.params.offset = 0;
    }
    try {
      batch_search(b, nullptr, stride, reinterpret_cast<uint32_t*>(rec + lay.ids_offset),
                   reinterpret_cast<double*>(rec + lay.scores_offset), reinterpret_cast<uint32_t*>(rec + lay.count_offset),
                   reinterpret_cast<uint64_t*>(rec + lay.total_offset));
    } catch (...) {
      b.params = saved;
      throw;
    }
    b.params = saved;
  }
  if (b.streamed) {
    MGX_CUDA(cudaMemsetAsync(rec + lay.status_offset, 0, 16, st));
    MGX_CUDA(cudaMemcpyAsync(rec + lay.status_offset, b.d_launch.p + kLaunchOverflow, sizeof(uint32_t),
                             cudaMemcpyDeviceToDevice, st));
  } else {
    MGX_CUDA(cudaMemsetAsync(rec + lay.status_offset, 0, 16, st));
  }
  if (G > 1) {
    // exchange 2: ONE all-gather of the fixed-size packed records, in place (every rank wrote its own slot)
    cudaStream_t cs = comm->stream[lane];
    MGX_CUDA(cudaEventRecord(batch_event(b, 2), st));
    MGX_CUDA(cudaStreamWaitEvent(cs, batch_event(b, 2), 0));
    MGX_NCCL(api, api->AllGather(rec, b.o_gather.p, lay.bytes, ncclUint8, comm->comm[lane], cs));
    MGX_CUDA(cudaEventRecord(batch_event(b, 3), cs));
    MGX_CUDA(cudaStreamWaitEvent(st, batch_event(b, 3), 0));
    const uint8_t* in = b.o_gather.p;
    uint8_t* out = b.o_merged.p;
    launch_merge_topk(st, b.params, static_cast<uint32_t>(G), b.n_out_queries, stride,
                      reinterpret_cast<const uint32_t*>(in + lay.ids_offset),
                      reinterpret_cast<const double*>(in + lay.scores_offset),
                      reinterpret_cast<const uint32_t*>(in + lay.count_offset),
                      reinterpret_cast<const uint64_t*>(in + lay.total_offset), lay.bytes,
                      reinterpret_cast<uint32_t*>(out + lay.ids_offset), reinterpret_cast<double*>(out + lay.scores_offset),
                      reinterpret_cast<uint32_t*>(out + lay.count_offset), reinterpret_cast<uint64_t*>(out + lay.total_offset));
    launch_or_status(st, in + lay.status_offset, lay.bytes, static_cast<uint32_t>(G), out + lay.status_offset);
  }
  b.h_status.reserve(4);
  MGX_CUDA(cudaMemcpyAsync(b.h_status.p, b.o_merged.p + lay.status_offset, 16, cudaMemcpyDeviceToHost, st));
  if (h_record_out != nullptr) {
    MGX_CUDA(cudaMemcpyAsync(h_record_out, b.o_merged.p, lay.bytes, cudaMemcpyDeviceToHost, st));
    b.d2h_bytes += lay.bytes;
  }
  b.mark_last();
  b.searched = true;
  b.sharded_enqueued = true;
}
}  // namespace

int mgx_sharded_batch_enqueue(mgx_shard_comm_t* comm, int32_t lane, mgx_batch_t* batch, uint64_t stride,
                              void* h_record_out) {
  if (batch == nullptr) {
    return invalid("null batch");
  }
  if (comm != nullptr && (lane < 0 || lane >= comm->n_lanes)) {
    return invalid("communicator lane out of range");
  }
  if (batch->b.planned) {
    return invalid("mgx_sharded_batch_enqueue needs a freshly prepared batch");
  }
  if (int rc = check_stride(batch->b.params, stride); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Batch& b = batch->b;
    DeviceGuard guard(b.ix->device);
    b.allow_streamed = true;
    sharded_enqueue(comm, lane, b, stride, h_record_out);
    return MGX_OK;
  });
}

int mgx_sharded_batch_finish(mgx_shard_comm_t* comm, int32_t lane, mgx_batch_t* batch, uint64_t stride,
                             void* h_record_out, const void** d_record_out, int32_t* out_repeated) {
  if (batch == nullptr) {
    return invalid("null batch");
  }
  if (!batch->b.sharded_enqueued) {
    return invalid("mgx_sharded_batch_finish without mgx_sharded_batch_enqueue");
  }
  return guarded([&]() {
    Batch& b = batch->b;
    DeviceGuard guard(b.ix->device);
    if (out_repeated != nullptr) {
      *out_repeated = 0;
    }
    MGX_CUDA(cudaEventSynchronize(b.ev_last));
    // the status block of the MERGED record: set when any shard's streamed batch did not fit its workspace. Every
    // rank reads the same value, so all of them repeat the batch together (in the synchronous form, which sizes its
    // workspace from the batch) and the exchanges stay matched.
    if (b.h_status.p != nullptr && b.h_status.p[0] != 0) {
      batch_reset_for_repeat(b, true);
      sharded_enqueue(comm, lane, b, stride, h_record_out);
      MGX_CUDA(cudaEventSynchronize(b.ev_last));
      if (out_repeated != nullptr) {
        *out_repeated = 1;
      }
    }
    if (d_record_out != nullptr) {
      *d_record_out = b.o_merged.p;
    }
    return MGX_OK;
  });
}

int mgx_index_last_batch_stats(const mgx_index_t* index, mgx_batch_stats_t* out) {
  if (index == nullptr || out == nullptr) {
    return invalid("null argument");
  }
  std::lock_guard<std::mutex> sl(const_cast<mgx_index_t*>(index)->stats_mu);
  *out = index->ix.last_stats;
  return MGX_OK;
}

// ---------------------------------------------------------------- scoring
int mgx_score_documents(mgx_index_t* index, const uint32_t* candidates, uint64_t n_candidates,
                        const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* term_doc_freqs,
                        uint64_t n_terms, uint64_t total_docs, double avg_doc_length, double k1, double b,
                        double* out_scores) {
  if (index == nullptr || (n_candidates > 0 && (candidates == nullptr || out_scores == nullptr)) ||
      (n_terms > 0 && (term_bytes == nullptr || term_offsets == nullptr || term_doc_freqs == nullptr))) {
    return invalid("null argument");
  }
  if (n_candidates == 0) {
    return MGX_OK;
  }
  if (int rc = commit_pending(index); rc != MGX_OK) {
    return rc;
  }
  return guarded([&]() {
    Reader rd(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    cudaStream_t st = rd.stream();
    std::vector<uint32_t> boff(n_terms + 1, 0);
    for (uint64_t i = 0; i < n_terms; ++i) {
      if (term_offsets[i + 1] - term_offsets[i] > kMaxTermBytes) {
        set_last_error("term longer than 256 bytes is not supported");
        return MGX_ERR_UNSUPPORTED;
      }
      boff[i + 1] = static_cast<uint32_t>(term_offsets[i + 1] - term_offsets[0]);
    }
    const uint64_t nbytes = n_terms > 0 ? term_offsets[n_terms] - term_offsets[0] : 0;
    DevBuf<uint8_t> d_bytes;
    DevBuf<uint32_t> d_boff;
    DevBuf<uint64_t> d_dfs;
    DevBuf<uint32_t> d_cands;
    DevBuf<double> d_scores;
    d_bytes.alloc(nbytes + 16);
    d_boff.alloc(n_terms + 1);
    d_dfs.alloc(n_terms);
    d_cands.alloc(n_candidates);
    d_scores.alloc(n_candidates);
    if (nbytes > 0) {
      MGX_CUDA(cudaMemcpyAsync(d_bytes.p, term_bytes + term_offsets[0], nbytes, cudaMemcpyHostToDevice, st));
    }
    MGX_CUDA(cudaMemcpyAsync(d_boff.p, boff.data(), (n_terms + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if (n_terms > 0) {
      MGX_CUDA(cudaMemcpyAsync(d_dfs.p, term_doc_freqs, n_terms * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    }
    MGX_CUDA(cudaMemcpyAsync(d_cands.p, candidates, n_candidates * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    launch_score_documents(ix, st, d_cands.p, n_candidates, d_bytes.p, d_boff.p, d_dfs.p,
                           static_cast<uint32_t>(n_terms), total_docs, avg_doc_length, k1, b, d_scores.p);
    MGX_CUDA(cudaMemcpyAsync(out_scores, d_scores.p, n_candidates * sizeof(double), cudaMemcpyDeviceToHost, st));
    MGX_CUDA(cudaStreamSynchronize(st));
    return MGX_OK;
  });
}

int mgx_sort_by_score(mgx_index_t* index, const uint32_t* results, const double* scores, uint64_t n,
                      int32_t descending, uint32_t limit, uint32_t offset, uint32_t* out, uint64_t* out_count) {
  if (index == nullptr || out_count == nullptr || (n > 0 && (results == nullptr || scores == nullptr || out == nullptr))) {
    return invalid("null argument");
  }
  *out_count = 0;
  if (n == 0) {
    return MGX_OK;  // result_sorter.cpp:663-665
  }
  return guarded([&]() {
    Reader rd(index);
    Index& ix = index->ix;
    DeviceGuard guard(ix.device);
    cudaStream_t st = rd.stream();
    DevBuf<uint32_t> d_docs;
    DevBuf<double> d_scores;
    DevBuf<uint32_t> d_out;
    DevBuf<uint32_t> d_cnt;
    d_docs.alloc(n);
    d_scores.alloc(n);
    d_out.alloc(limit == 0 ? n : std::min<uint64_t>(limit, n));  // the window never exceeds min(limit, n)
    d_cnt.alloc(1);
    MGX_CUDA(cudaMemcpyAsync(d_docs.p, results, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    MGX_CUDA(cudaMemcpyAsync(d_scores.p, scores, n * sizeof(double), cudaMemcpyHostToDevice, st));
    launch_sort_by_score(st, d_docs.p, d_scores.p, n, descending != 0, limit, offset, d_out.p, d_cnt.p);
    uint32_t cnt = 0;
    MGX_CUDA(cudaMemcpyAsync(&cnt, d_cnt.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    MGX_CUDA(cudaStreamSynchronize(st));
    if (cnt > 0) {
      MGX_CUDA(cudaMemcpy(out, d_out.p, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    *out_count = cnt;
    return MGX_OK;
  });
}

}  // extern "C"
