// oracle/shim/roaring/roaring.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Functional stand-in for the subset of CRoaring v4.6.1 (pinned by the
// reference at third_party/CMakeLists.txt:104-112, fetched at configure time,
// not vendored, and not present in this offline image) that the reference's
// hot-path translation units call. It lets `oracle/Makefile` compile the
// reference's own, unmodified sources (src/index/posting_list.cpp,
// src/storage/filter_index.cpp, ...) into `oracle/_ref/`.
//
// Roaring is a pure *set ADT* on this path: every call site only observes set
// membership, cardinality, order of iteration and set algebra. This shim
// restates the published Roaring layout (Chambi/Lemire et al.): a sorted
// vector of 64Ki-wide chunks keyed by the high 16 bits, each chunk either a
// sorted uint16 array (<= 4096 values) or a 65536-bit bitset. Run containers
// are not implemented (`run_optimize` is a no-op); the portable
// serialisation follows the published Roaring interchange format (written
// without run containers, read with them), so MGIX streams of this build and
// of a CRoaring build are mutually readable. Set results are identical by construction;
// CPU timings taken with it are labelled "reference sources + Roaring shim".
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace roaring_shim {

constexpr uint32_t kArrayMax = 4096;  // array container holds at most this many values
constexpr uint32_t kWords = 1024;     // 65536 bits

struct Chunk {
  uint16_t key = 0;
  uint32_t card = 0;
  std::vector<uint16_t> arr;   // sorted, used when bits is empty
  std::vector<uint64_t> bits;  // kWords words when in bitset form

  bool is_bitset() const { return !bits.empty(); }

  bool contains(uint16_t low) const {
    if (is_bitset()) {
      return (bits[low >> 6] >> (low & 63)) & 1ULL;
    }
    return std::binary_search(arr.begin(), arr.end(), low);
  }

  void to_bitset() {
    bits.assign(kWords, 0);
    for (uint16_t v : arr) {
      bits[v >> 6] |= 1ULL << (v & 63);
    }
    arr.clear();
    arr.shrink_to_fit();
  }

  void to_array() {
    arr.clear();
    arr.reserve(card);
    for (uint32_t w = 0; w < kWords; ++w) {
      uint64_t word = bits[w];
      while (word != 0) {
        arr.push_back(static_cast<uint16_t>((w << 6) | static_cast<uint32_t>(__builtin_ctzll(word))));
        word &= word - 1;
      }
    }
    bits.clear();
    bits.shrink_to_fit();
  }

  void normalize() {
    if (is_bitset() && card <= kArrayMax) {
      to_array();
    } else if (!is_bitset() && card > kArrayMax) {
      to_bitset();
    }
  }

  bool add(uint16_t low) {
    if (is_bitset()) {
      uint64_t& word = bits[low >> 6];
      const uint64_t mask = 1ULL << (low & 63);
      if ((word & mask) != 0) {
        return false;
      }
      word |= mask;
      ++card;
      return true;
    }
    if (arr.empty() || arr.back() < low) {
      arr.push_back(low);
    } else {
      auto pos = std::lower_bound(arr.begin(), arr.end(), low);
      if (pos != arr.end() && *pos == low) {
        return false;
      }
      arr.insert(pos, low);
    }
    ++card;
    if (card > kArrayMax) {
      to_bitset();
    }
    return true;
  }

  bool remove(uint16_t low) {
    if (is_bitset()) {
      uint64_t& word = bits[low >> 6];
      const uint64_t mask = 1ULL << (low & 63);
      if ((word & mask) == 0) {
        return false;
      }
      word &= ~mask;
      --card;
      if (card <= kArrayMax) {
        to_array();
      }
      return true;
    }
    auto pos = std::lower_bound(arr.begin(), arr.end(), low);
    if (pos == arr.end() || *pos != low) {
      return false;
    }
    arr.erase(pos);
    --card;
    return true;
  }

  template <typename F>
  void for_each(F&& fn) const {
    const uint32_t base = static_cast<uint32_t>(key) << 16;
    if (is_bitset()) {
      for (uint32_t w = 0; w < kWords; ++w) {
        uint64_t word = bits[w];
        while (word != 0) {
          fn(base | (w << 6) | static_cast<uint32_t>(__builtin_ctzll(word)));
          word &= word - 1;
        }
      }
    } else {
      for (uint16_t v : arr) {
        fn(base | v);
      }
    }
  }

  uint16_t nth(uint32_t rank) const {  // rank < card
    if (!is_bitset()) {
      return arr[rank];
    }
    for (uint32_t w = 0; w < kWords; ++w) {
      const uint32_t pc = static_cast<uint32_t>(__builtin_popcountll(bits[w]));
      if (rank < pc) {
        uint64_t word = bits[w];
        for (uint32_t i = 0; i < rank; ++i) {
          word &= word - 1;
        }
        return static_cast<uint16_t>((w << 6) | static_cast<uint32_t>(__builtin_ctzll(word)));
      }
      rank -= pc;
    }
    return 0;
  }
};

inline Chunk chunk_and(const Chunk& a, const Chunk& b) {
  Chunk out;
  out.key = a.key;
  if (a.is_bitset() && b.is_bitset()) {
    out.bits.resize(kWords);
    uint32_t card = 0;
    for (uint32_t w = 0; w < kWords; ++w) {
      out.bits[w] = a.bits[w] & b.bits[w];
      card += static_cast<uint32_t>(__builtin_popcountll(out.bits[w]));
    }
    out.card = card;
    out.normalize();
    return out;
  }
  if (!a.is_bitset() && !b.is_bitset()) {
    std::set_intersection(a.arr.begin(), a.arr.end(), b.arr.begin(), b.arr.end(), std::back_inserter(out.arr));
  } else {
    const Chunk& arr = a.is_bitset() ? b : a;
    const Chunk& bs = a.is_bitset() ? a : b;
    for (uint16_t v : arr.arr) {
      if ((bs.bits[v >> 6] >> (v & 63)) & 1ULL) {
        out.arr.push_back(v);
      }
    }
  }
  out.card = static_cast<uint32_t>(out.arr.size());
  return out;
}

inline Chunk chunk_or(const Chunk& a, const Chunk& b) {
  Chunk out;
  out.key = a.key;
  if (!a.is_bitset() && !b.is_bitset()) {
    std::set_union(a.arr.begin(), a.arr.end(), b.arr.begin(), b.arr.end(), std::back_inserter(out.arr));
    out.card = static_cast<uint32_t>(out.arr.size());
    out.normalize();
    return out;
  }
  const Chunk& bs = a.is_bitset() ? a : b;
  const Chunk& other = a.is_bitset() ? b : a;
  out.bits = bs.bits;
  if (other.is_bitset()) {
    for (uint32_t w = 0; w < kWords; ++w) {
      out.bits[w] |= other.bits[w];
    }
  } else {
    for (uint16_t v : other.arr) {
      out.bits[v >> 6] |= 1ULL << (v & 63);
    }
  }
  uint32_t card = 0;
  for (uint32_t w = 0; w < kWords; ++w) {
    card += static_cast<uint32_t>(__builtin_popcountll(out.bits[w]));
  }
  out.card = card;
  return out;
}

inline Chunk chunk_andnot(const Chunk& a, const Chunk& b) {
  Chunk out;
  out.key = a.key;
  if (a.is_bitset()) {
    out.bits = a.bits;
    if (b.is_bitset()) {
      for (uint32_t w = 0; w < kWords; ++w) {
        out.bits[w] &= ~b.bits[w];
      }
    } else {
      for (uint16_t v : b.arr) {
        out.bits[v >> 6] &= ~(1ULL << (v & 63));
      }
    }
    uint32_t card = 0;
    for (uint32_t w = 0; w < kWords; ++w) {
      card += static_cast<uint32_t>(__builtin_popcountll(out.bits[w]));
    }
    out.card = card;
    out.normalize();
    return out;
  }
  for (uint16_t v : a.arr) {
    if (!b.contains(v)) {
      out.arr.push_back(v);
    }
  }
  out.card = static_cast<uint32_t>(out.arr.size());
  return out;
}

}  // namespace roaring_shim

struct roaring_bitmap_t {
  std::vector<roaring_shim::Chunk> chunks;  // sorted by key, no empty chunks

  size_t find_chunk(uint16_t key) const {  // index of first chunk with key >= `key`
    size_t lo = 0;
    size_t hi = chunks.size();
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (chunks[mid].key < key) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    return lo;
  }
};

struct roaring_bulk_context_t {
  const void* container = nullptr;
  int idx = 0;
  uint16_t key = 0;
  uint8_t typecode = 0;
};

struct roaring_uint32_iterator_t {
  const roaring_bitmap_t* parent = nullptr;
  int64_t chunk_index = 0;
  int64_t rank_in_chunk = 0;
  uint32_t current_value = 0;
  bool has_value = false;
};

inline roaring_bitmap_t* roaring_bitmap_create() { return new roaring_bitmap_t(); }
inline void roaring_bitmap_free(const roaring_bitmap_t* r) { delete r; }
inline roaring_bitmap_t* roaring_bitmap_copy(const roaring_bitmap_t* r) { return new roaring_bitmap_t(*r); }

inline bool roaring_bitmap_add_checked(roaring_bitmap_t* r, uint32_t x) {
  const uint16_t key = static_cast<uint16_t>(x >> 16);
  size_t pos;
  if (!r->chunks.empty() && r->chunks.back().key == key) {
    pos = r->chunks.size() - 1;
  } else {
    pos = r->find_chunk(key);
    if (pos == r->chunks.size() || r->chunks[pos].key != key) {
      roaring_shim::Chunk chunk;
      chunk.key = key;
      r->chunks.insert(r->chunks.begin() + static_cast<std::ptrdiff_t>(pos), std::move(chunk));
    }
  }
  return r->chunks[pos].add(static_cast<uint16_t>(x & 0xFFFF));
}
inline void roaring_bitmap_add(roaring_bitmap_t* r, uint32_t x) { (void)roaring_bitmap_add_checked(r, x); }
inline void roaring_bitmap_add_many(roaring_bitmap_t* r, size_t n, const uint32_t* vals) {
  for (size_t i = 0; i < n; ++i) {
    (void)roaring_bitmap_add_checked(r, vals[i]);
  }
}
inline bool roaring_bitmap_remove_checked(roaring_bitmap_t* r, uint32_t x) {
  const uint16_t key = static_cast<uint16_t>(x >> 16);
  const size_t pos = r->find_chunk(key);
  if (pos == r->chunks.size() || r->chunks[pos].key != key) {
    return false;
  }
  const bool removed = r->chunks[pos].remove(static_cast<uint16_t>(x & 0xFFFF));
  if (removed && r->chunks[pos].card == 0) {
    r->chunks.erase(r->chunks.begin() + static_cast<std::ptrdiff_t>(pos));
  }
  return removed;
}
inline void roaring_bitmap_remove(roaring_bitmap_t* r, uint32_t x) { (void)roaring_bitmap_remove_checked(r, x); }

inline bool roaring_bitmap_contains(const roaring_bitmap_t* r, uint32_t x) {
  const uint16_t key = static_cast<uint16_t>(x >> 16);
  const size_t pos = r->find_chunk(key);
  return pos != r->chunks.size() && r->chunks[pos].key == key &&
         r->chunks[pos].contains(static_cast<uint16_t>(x & 0xFFFF));
}
inline bool roaring_bitmap_contains_bulk(const roaring_bitmap_t* r, roaring_bulk_context_t* ctx, uint32_t x) {
  const uint16_t key = static_cast<uint16_t>(x >> 16);
  if (ctx->container != nullptr && ctx->key == key && static_cast<size_t>(ctx->idx) < r->chunks.size() &&
      &r->chunks[static_cast<size_t>(ctx->idx)] == ctx->container) {
    return r->chunks[static_cast<size_t>(ctx->idx)].contains(static_cast<uint16_t>(x & 0xFFFF));
  }
  const size_t pos = r->find_chunk(key);
  if (pos == r->chunks.size() || r->chunks[pos].key != key) {
    ctx->container = nullptr;
    return false;
  }
  ctx->container = &r->chunks[pos];
  ctx->idx = static_cast<int>(pos);
  ctx->key = key;
  return r->chunks[pos].contains(static_cast<uint16_t>(x & 0xFFFF));
}

inline uint64_t roaring_bitmap_get_cardinality(const roaring_bitmap_t* r) {
  uint64_t total = 0;
  for (const auto& chunk : r->chunks) {
    total += chunk.card;
  }
  return total;
}
inline bool roaring_bitmap_is_empty(const roaring_bitmap_t* r) { return r->chunks.empty(); }
inline void roaring_bitmap_to_uint32_array(const roaring_bitmap_t* r, uint32_t* out) {
  for (const auto& chunk : r->chunks) {
    chunk.for_each([&](uint32_t v) { *out++ = v; });
  }
}
inline uint32_t roaring_bitmap_maximum(const roaring_bitmap_t* r) {
  if (r->chunks.empty()) {
    return 0;
  }
  const auto& chunk = r->chunks.back();
  return (static_cast<uint32_t>(chunk.key) << 16) | chunk.nth(chunk.card - 1);
}

inline roaring_bitmap_t* roaring_bitmap_and(const roaring_bitmap_t* a, const roaring_bitmap_t* b) {
  auto* out = new roaring_bitmap_t();
  size_t i = 0;
  size_t j = 0;
  while (i < a->chunks.size() && j < b->chunks.size()) {
    if (a->chunks[i].key < b->chunks[j].key) {
      ++i;
    } else if (b->chunks[j].key < a->chunks[i].key) {
      ++j;
    } else {
      auto chunk = roaring_shim::chunk_and(a->chunks[i], b->chunks[j]);
      if (chunk.card != 0) {
        out->chunks.push_back(std::move(chunk));
      }
      ++i;
      ++j;
    }
  }
  return out;
}
inline uint64_t roaring_bitmap_and_cardinality(const roaring_bitmap_t* a, const roaring_bitmap_t* b) {
  std::unique_ptr<roaring_bitmap_t> tmp(roaring_bitmap_and(a, b));
  return roaring_bitmap_get_cardinality(tmp.get());
}
inline roaring_bitmap_t* roaring_bitmap_or(const roaring_bitmap_t* a, const roaring_bitmap_t* b) {
  auto* out = new roaring_bitmap_t();
  size_t i = 0;
  size_t j = 0;
  while (i < a->chunks.size() || j < b->chunks.size()) {
    if (j == b->chunks.size() || (i < a->chunks.size() && a->chunks[i].key < b->chunks[j].key)) {
      out->chunks.push_back(a->chunks[i++]);
    } else if (i == a->chunks.size() || b->chunks[j].key < a->chunks[i].key) {
      out->chunks.push_back(b->chunks[j++]);
    } else {
      out->chunks.push_back(roaring_shim::chunk_or(a->chunks[i++], b->chunks[j++]));
    }
  }
  return out;
}
inline void roaring_bitmap_and_inplace(roaring_bitmap_t* a, const roaring_bitmap_t* b) {
  std::unique_ptr<roaring_bitmap_t> tmp(roaring_bitmap_and(a, b));
  a->chunks = std::move(tmp->chunks);
}
inline void roaring_bitmap_or_inplace(roaring_bitmap_t* a, const roaring_bitmap_t* b) {
  std::unique_ptr<roaring_bitmap_t> tmp(roaring_bitmap_or(a, b));
  a->chunks = std::move(tmp->chunks);
}
inline void roaring_bitmap_andnot_inplace(roaring_bitmap_t* a, const roaring_bitmap_t* b) {
  std::vector<roaring_shim::Chunk> out;
  size_t j = 0;
  for (auto& chunk : a->chunks) {
    while (j < b->chunks.size() && b->chunks[j].key < chunk.key) {
      ++j;
    }
    if (j < b->chunks.size() && b->chunks[j].key == chunk.key) {
      auto diff = roaring_shim::chunk_andnot(chunk, b->chunks[j]);
      if (diff.card != 0) {
        out.push_back(std::move(diff));
      }
    } else {
      out.push_back(std::move(chunk));
    }
  }
  a->chunks = std::move(out);
}

inline bool roaring_bitmap_run_optimize(roaring_bitmap_t* /*r*/) { return false; }

// Private (non-CRoaring) byte layout: u32 n_chunks, then per chunk
// u16 key, u32 card, u16 values[card]. Only used for size accounting and
// same-process round trips.
// Portable serialisation in the published Roaring interchange format (RoaringFormatSpec): this shim only ever
// WRITES the layout without run containers
//   u32 cookie 12346 | u32 n | n x (u16 key, u16 cardinality-1) | n x u32 byte offset | containers
//   (cardinality <= 4096: sorted u16 values; else 1024 x u64 bitset words)
// and READS that one plus the run-container layout CRoaring emits after run_optimize
//   u32 (12347 | (n-1) << 16) | ceil(n/8) run-flag bytes | n x (u16 key, u16 cardinality-1) |
//   [n >= 4: n x u32 offset] | containers (run: u16 n_runs, n_runs x (u16 start, u16 length-1)).
inline size_t roaring_bitmap_portable_size_in_bytes(const roaring_bitmap_t* r) {
  size_t total = 8 + r->chunks.size() * 8;
  for (const auto& chunk : r->chunks) {
    total += chunk.is_bitset() ? static_cast<size_t>(roaring_shim::kWords) * 8 : static_cast<size_t>(chunk.card) * 2;
  }
  return total;
}
inline size_t roaring_bitmap_portable_serialize(const roaring_bitmap_t* r, char* buf) {
  char* out = buf;
  auto put16 = [&](uint16_t v) {
    out[0] = static_cast<char>(v & 0xFF);
    out[1] = static_cast<char>(v >> 8);
    out += 2;
  };
  auto put32 = [&](uint32_t v) {
    put16(static_cast<uint16_t>(v & 0xFFFF));
    put16(static_cast<uint16_t>(v >> 16));
  };
  const uint32_t n = static_cast<uint32_t>(r->chunks.size());
  put32(12346u);
  put32(n);
  for (const auto& chunk : r->chunks) {
    put16(chunk.key);
    put16(static_cast<uint16_t>(chunk.card - 1));
  }
  uint32_t offset = 8 + n * 8;
  for (const auto& chunk : r->chunks) {
    put32(offset);
    offset += chunk.is_bitset() ? roaring_shim::kWords * 8 : chunk.card * 2;
  }
  for (const auto& chunk : r->chunks) {
    if (chunk.is_bitset()) {
      for (uint64_t w : chunk.bits) {
        put32(static_cast<uint32_t>(w & 0xFFFFFFFFu));
        put32(static_cast<uint32_t>(w >> 32));
      }
    } else {
      for (uint16_t v : chunk.arr) {
        put16(v);
      }
    }
  }
  return static_cast<size_t>(out - buf);
}
inline roaring_bitmap_t* roaring_bitmap_portable_deserialize_safe(const char* buf, size_t maxbytes) {
  const unsigned char* in = reinterpret_cast<const unsigned char*>(buf);
  const unsigned char* end = in + maxbytes;
  auto need = [&](size_t k) { return static_cast<size_t>(end - in) >= k; };
  auto get16 = [&]() {
    const uint16_t v = static_cast<uint16_t>(in[0] | (in[1] << 8));
    in += 2;
    return v;
  };
  auto get32 = [&]() {
    const uint32_t lo = get16();
    const uint32_t hi = get16();
    return lo | (hi << 16);
  };
  if (!need(4)) {
    return nullptr;
  }
  const uint32_t cookie = get32();
  uint32_t n = 0;
  std::vector<uint8_t> run_flags;
  bool has_runs = false;
  if ((cookie & 0xFFFF) == 12347u) {
    has_runs = true;
    n = (cookie >> 16) + 1;
    const size_t flag_bytes = (n + 7) / 8;
    if (!need(flag_bytes)) {
      return nullptr;
    }
    run_flags.assign(in, in + flag_bytes);
    in += flag_bytes;
  } else if (cookie == 12346u) {
    if (!need(4)) {
      return nullptr;
    }
    n = get32();
    if (n > 65536) {
      return nullptr;
    }
  } else {
    return nullptr;
  }
  if (!need(static_cast<size_t>(n) * 4)) {
    return nullptr;
  }
  auto out = std::make_unique<roaring_bitmap_t>();
  out->chunks.resize(n);
  for (uint32_t c = 0; c < n; ++c) {
    out->chunks[c].key = get16();
    out->chunks[c].card = static_cast<uint32_t>(get16()) + 1;
  }
  if (!has_runs || n >= 4) {
    if (!need(static_cast<size_t>(n) * 4)) {
      return nullptr;
    }
    in += static_cast<size_t>(n) * 4;  // the offset header only serves random access
  }
  for (uint32_t c = 0; c < n; ++c) {
    roaring_shim::Chunk& chunk = out->chunks[c];
    const bool is_run = has_runs && ((run_flags[c / 8] >> (c % 8)) & 1) != 0;
    if (is_run) {
      if (!need(2)) {
        return nullptr;
      }
      const uint32_t n_runs = get16();
      if (!need(static_cast<size_t>(n_runs) * 4)) {
        return nullptr;
      }
      std::vector<uint16_t> values;
      uint32_t next_free = 0;
      for (uint32_t i = 0; i < n_runs; ++i) {
        const uint32_t start = get16();
        const uint32_t len = get16();
        if (start < next_free || start + len > 65535) {
          return nullptr;
        }
        for (uint32_t v = start; v <= start + len; ++v) {
          values.push_back(static_cast<uint16_t>(v));
        }
        next_free = start + len + 1;
      }
      if (values.size() != chunk.card) {
        return nullptr;
      }
      chunk.arr = std::move(values);
      if (chunk.card > roaring_shim::kArrayMax) {
        chunk.to_bitset();
      }
    } else if (chunk.card > roaring_shim::kArrayMax) {
      const size_t bytes = static_cast<size_t>(roaring_shim::kWords) * 8;
      if (!need(bytes)) {
        return nullptr;
      }
      chunk.bits.resize(roaring_shim::kWords);
      uint32_t pop = 0;
      for (uint32_t w = 0; w < roaring_shim::kWords; ++w) {
        const uint64_t lo = get32();
        const uint64_t hi = get32();
        chunk.bits[w] = lo | (hi << 32);
        pop += static_cast<uint32_t>(__builtin_popcountll(chunk.bits[w]));
      }
      if (pop != chunk.card) {
        return nullptr;
      }
    } else {
      if (!need(static_cast<size_t>(chunk.card) * 2)) {
        return nullptr;
      }
      chunk.arr.resize(chunk.card);
      for (uint32_t i = 0; i < chunk.card; ++i) {
        chunk.arr[i] = get16();
      }
    }
  }
  return out.release();
}
inline roaring_bitmap_t* roaring_bitmap_portable_deserialize(const char* buf) {
  return roaring_bitmap_portable_deserialize_safe(buf, static_cast<size_t>(-1) >> 1);
}
inline bool roaring_bitmap_internal_validate(const roaring_bitmap_t* r, const char** reason) {
  for (size_t i = 0; i < r->chunks.size(); ++i) {
    const auto& chunk = r->chunks[i];
    const bool sorted_keys = i == 0 || r->chunks[i - 1].key < chunk.key;
    const bool arr_ok = chunk.is_bitset() || (chunk.arr.size() == chunk.card &&
                                              std::adjacent_find(chunk.arr.begin(), chunk.arr.end(),
                                                                 std::greater_equal<uint16_t>()) == chunk.arr.end());
    if (!sorted_keys || chunk.card == 0 || !arr_ok) {
      if (reason != nullptr) {
        *reason = "roaring shim: malformed container";
      }
      return false;
    }
  }
  return true;
}

namespace roaring_shim {
inline void iterator_load(roaring_uint32_iterator_t* it) {
  const auto* r = it->parent;
  if (it->chunk_index < 0 || it->chunk_index >= static_cast<int64_t>(r->chunks.size())) {
    it->has_value = false;
    return;
  }
  const auto& chunk = r->chunks[static_cast<size_t>(it->chunk_index)];
  it->current_value =
      (static_cast<uint32_t>(chunk.key) << 16) | chunk.nth(static_cast<uint32_t>(it->rank_in_chunk));
  it->has_value = true;
}
}  // namespace roaring_shim

inline void roaring_iterator_init(const roaring_bitmap_t* r, roaring_uint32_iterator_t* it) {
  it->parent = r;
  it->chunk_index = 0;
  it->rank_in_chunk = 0;
  roaring_shim::iterator_load(it);
}
inline void roaring_iterator_init_last(const roaring_bitmap_t* r, roaring_uint32_iterator_t* it) {
  it->parent = r;
  it->chunk_index = static_cast<int64_t>(r->chunks.size()) - 1;
  it->rank_in_chunk = r->chunks.empty() ? 0 : static_cast<int64_t>(r->chunks.back().card) - 1;
  roaring_shim::iterator_load(it);
}
inline bool roaring_uint32_iterator_advance(roaring_uint32_iterator_t* it) {
  if (!it->has_value) {
    return false;
  }
  const auto& chunk = it->parent->chunks[static_cast<size_t>(it->chunk_index)];
  if (it->rank_in_chunk + 1 < static_cast<int64_t>(chunk.card)) {
    ++it->rank_in_chunk;
  } else {
    ++it->chunk_index;
    it->rank_in_chunk = 0;
  }
  roaring_shim::iterator_load(it);
  return it->has_value;
}
inline bool roaring_uint32_iterator_previous(roaring_uint32_iterator_t* it) {
  if (!it->has_value) {
    return false;
  }
  if (it->rank_in_chunk > 0) {
    --it->rank_in_chunk;
  } else {
    --it->chunk_index;
    if (it->chunk_index >= 0) {
      it->rank_in_chunk = static_cast<int64_t>(it->parent->chunks[static_cast<size_t>(it->chunk_index)].card) - 1;
    }
  }
  roaring_shim::iterator_load(it);
  return it->has_value;
}
