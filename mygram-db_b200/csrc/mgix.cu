// mgix.cu — the MGIX index stream: the on-disk form either side of the device index.
//
// Host-side codec of what Index::SaveToStream / Index::LoadFromStream read and write
// (src/index/index_serialization.cpp:111-224, 279-613) and of the posting-list bodies inside it
// (PostingList::Serialize / Deserialize, src/index/posting_list.cpp:973-1102):
//
//   "MGIX" | u32 version (4) | u32 ngram_size | u32 kanji_ngram_size | u8 cross_boundary | u8 normalize_nfkc |
//   u32 len + normalize_width | u8 normalize_lower | u64 term_count |
//   term_count x ( u32 len + term | u64 body_len | body ) | u32 CRC32 of everything before it
//   body = u8 strategy | u32 size | data
//     strategy 0 (kFixedWidthDelta): size x u32 = first doc id, then the gaps (fixed width, not varint)
//     strategy 1 (kRoaringBitmap):   size bytes of the Roaring portable interchange format
//
// All integers little-endian. The Roaring bytes come from CRoaring v4.6.1 in the reference (a pinned dependency that
// is not in its tree, third_party/CMakeLists.txt:104-112); this file restates the published interchange format
// (cookie 12346 / 12347, descriptive header, offset header, array / bitset / run containers). It is a data-format
// codec, not a compute path: nothing here touches the device, and the search core never calls it.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "../../include/mgx.h"
#include "mgx_internal.cuh"

namespace mgx {
namespace {

// ---- CRC-32 (the zlib polynomial the reference uses through utils/crc32.h), slicing-by-8
struct Crc32Tables {
  uint32_t t[8][256];
  Crc32Tables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) {
        c = (c & 1u) != 0 ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      }
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i) {
      for (int s = 1; s < 8; ++s) {
        t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
      }
    }
  }
};

uint32_t crc32(const uint8_t* p, uint64_t n) {
  static const Crc32Tables tables;
  const auto& t = tables.t;
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {
    uint32_t lo;
    uint32_t hi;
    std::memcpy(&lo, p, 4);
    std::memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = t[7][lo & 0xFF] ^ t[6][(lo >> 8) & 0xFF] ^ t[5][(lo >> 16) & 0xFF] ^ t[4][lo >> 24] ^ t[3][hi & 0xFF] ^
        t[2][(hi >> 8) & 0xFF] ^ t[1][(hi >> 16) & 0xFF] ^ t[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n-- > 0) {
    c = t[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  }
  return c ^ 0xFFFFFFFFu;
}

// CRC of the concatenation of two byte strings from their CRCs (GF(2) matrix method, as zlib's crc32_combine):
// lets the checksum of a multi-gigabyte stream be computed in independent pieces.
uint32_t gf2_times(const uint32_t* mat, uint32_t vec) {
  uint32_t sum = 0;
  while (vec != 0) {
    if ((vec & 1u) != 0) {
      sum ^= *mat;
    }
    vec >>= 1;
    ++mat;
  }
  return sum;
}

void gf2_square(uint32_t* square, const uint32_t* mat) {
  for (int n = 0; n < 32; ++n) {
    square[n] = gf2_times(mat, mat[n]);
  }
}

uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
  if (len2 == 0) {
    return crc1;
  }
  uint32_t even[32];
  uint32_t odd[32];
  odd[0] = 0xEDB88320u;  // the operator for one zero bit
  uint32_t row = 1;
  for (int n = 1; n < 32; ++n) {
    odd[n] = row;
    row <<= 1;
  }
  gf2_square(even, odd);  // two zero bits
  gf2_square(odd, even);  // four
  do {
    gf2_square(even, odd);
    if ((len2 & 1u) != 0) {
      crc1 = gf2_times(even, crc1);
    }
    len2 >>= 1;
    if (len2 == 0) {
      break;
    }
    gf2_square(odd, even);
    if ((len2 & 1u) != 0) {
      crc1 = gf2_times(odd, crc1);
    }
    len2 >>= 1;
  } while (len2 != 0);
  return crc1 ^ crc2;
}

// Worker threads of the codec: streams of a whole shard are gigabytes (2.8 GB for 10M documents), small ones are not
// worth a thread start. MGX_MGIX_THREADS pins the count (the tests run both paths on small inputs).
unsigned codec_threads(uint64_t work_items) {
  if (const char* env = std::getenv("MGX_MGIX_THREADS")) {
    return static_cast<unsigned>(std::max(1, std::min(64, std::atoi(env))));
  }
  if (work_items < (1u << 20)) {
    return 1;
  }
  return std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
}

// fn(first, last, worker) over [0, n) cut into `threads` contiguous ranges.
template <class Fn>
void parallel_ranges(uint64_t n, unsigned threads, Fn&& fn) {
  if (threads <= 1 || n < threads) {
    fn(static_cast<uint64_t>(0), n, 0u);
    return;
  }
  std::vector<std::thread> pool;
  for (unsigned w = 0; w < threads; ++w) {
    pool.emplace_back([&, w]() { fn(n * w / threads, n * (w + 1) / threads, w); });
  }
  for (auto& t : pool) {
    t.join();
  }
}

// Range boundaries over [0, n) with about equal weight per range; weight_before(t) = total weight of items [0, t),
// non-decreasing (posting lists are Zipf-sized: equal item counts would leave one thread with most of the postings).
template <class W>
std::vector<uint64_t> balanced_bounds(uint64_t n, unsigned threads, W&& weight_before) {
  std::vector<uint64_t> bounds(threads + 1, n);
  bounds[0] = 0;
  const uint64_t total = weight_before(n);
  for (unsigned w = 1; w < threads; ++w) {
    const uint64_t target = total / threads * w;
    uint64_t lo = bounds[w - 1];
    uint64_t hi = n;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) / 2;
      if (weight_before(mid) < target) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    bounds[w] = lo;
  }
  return bounds;
}

template <class Fn>
void parallel_bounds(const std::vector<uint64_t>& bounds, Fn&& fn) {
  const unsigned threads = static_cast<unsigned>(bounds.size() - 1);
  if (threads <= 1) {
    fn(bounds.front(), bounds.back(), 0u);
    return;
  }
  std::vector<std::thread> pool;
  for (unsigned w = 0; w < threads; ++w) {
    pool.emplace_back([&, w]() { fn(bounds[w], bounds[w + 1], w); });
  }
  for (auto& t : pool) {
    t.join();
  }
}

uint32_t crc32_parallel(const uint8_t* p, uint64_t n, unsigned threads) {
  if (threads <= 1 || n < (1u << 22)) {
    return crc32(p, n);
  }
  std::vector<uint32_t> part(threads, 0);
  std::vector<uint64_t> len(threads, 0);
  parallel_ranges(n, threads, [&](uint64_t a, uint64_t b, unsigned w) {
    part[w] = crc32(p + a, b - a);
    len[w] = b - a;
  });
  uint32_t c = part[0];
  for (unsigned w = 1; w < threads; ++w) {
    c = crc32_combine(c, part[w], len[w]);
  }
  return c;
}

constexpr uint32_t kArrayMax = 4096;            // a Roaring array container holds at most this many values
constexpr uint32_t kBitsetBytes = 8192;         // 65536 bits
constexpr uint32_t kCookieNoRuns = 12346;       // SERIAL_COOKIE_NO_RUNCONTAINER
constexpr uint32_t kCookieRuns = 12347;         // SERIAL_COOKIE
constexpr uint32_t kNoOffsetThreshold = 4;      // with runs, fewer containers carry no offset header
constexpr uint32_t kAutoRoaringEntries = 4096;  // posting_list.cpp:21
constexpr uint32_t kMaxTermLength = 10000;      // index_serialization.cpp:521
constexpr uint64_t kMaxPostingBytes = 100000000ULL;  // :547

// Bounded writer: counts always, stores while the bytes fit.
struct Writer {
  uint8_t* out;
  uint64_t cap;
  uint64_t pos = 0;
  void bytes(const void* src, uint64_t n) {
    if (out != nullptr && pos + n <= cap) {
      std::memcpy(out + pos, src, n);
    }
    pos += n;
  }
  void u8(uint8_t v) { bytes(&v, 1); }
  void u16(uint16_t v) {
    const uint8_t b[2] = {static_cast<uint8_t>(v), static_cast<uint8_t>(v >> 8)};
    bytes(b, 2);
  }
  void u32(uint32_t v) {
    const uint8_t b[4] = {static_cast<uint8_t>(v), static_cast<uint8_t>(v >> 8), static_cast<uint8_t>(v >> 16),
                          static_cast<uint8_t>(v >> 24)};
    bytes(b, 4);
  }
  void u64(uint64_t v) {
    u32(static_cast<uint32_t>(v));
    u32(static_cast<uint32_t>(v >> 32));
  }
};

// Size of the Roaring portable form of one ascending list (no run containers).
uint64_t roaring_size(const uint32_t* ids, uint64_t n, uint32_t* n_containers) {
  uint64_t bytes = 8;
  uint32_t containers = 0;
  for (uint64_t i = 0; i < n;) {
    uint64_t j = i;
    const uint32_t high = ids[i] >> 16;
    while (j < n && (ids[j] >> 16) == high) {
      ++j;
    }
    const uint64_t card = j - i;
    bytes += 8 + (card > kArrayMax ? kBitsetBytes : card * 2);
    ++containers;
    i = j;
  }
  *n_containers = containers;
  return bytes;
}

void write_roaring(Writer& w, const uint32_t* ids, uint64_t n, uint32_t n_containers) {
  w.u32(kCookieNoRuns);
  w.u32(n_containers);
  for (uint64_t i = 0; i < n;) {  // descriptive header: key, cardinality - 1
    uint64_t j = i;
    const uint32_t high = ids[i] >> 16;
    while (j < n && (ids[j] >> 16) == high) {
      ++j;
    }
    w.u16(static_cast<uint16_t>(high));
    w.u16(static_cast<uint16_t>(j - i - 1));
    i = j;
  }
  uint32_t offset = 8 + n_containers * 8;
  for (uint64_t i = 0; i < n;) {  // offset header: where each container starts, from the cookie
    uint64_t j = i;
    const uint32_t high = ids[i] >> 16;
    while (j < n && (ids[j] >> 16) == high) {
      ++j;
    }
    w.u32(offset);
    offset += (j - i) > kArrayMax ? kBitsetBytes : static_cast<uint32_t>(j - i) * 2;
    i = j;
  }
  std::vector<uint64_t> words;
  for (uint64_t i = 0; i < n;) {
    uint64_t j = i;
    const uint32_t high = ids[i] >> 16;
    while (j < n && (ids[j] >> 16) == high) {
      ++j;
    }
    if (j - i > kArrayMax) {
      words.assign(1024, 0);
      for (uint64_t k = i; k < j; ++k) {
        const uint32_t low = ids[k] & 0xFFFFu;
        words[low >> 6] |= 1ULL << (low & 63);
      }
      for (uint64_t word : words) {
        w.u64(word);
      }
    } else if (w.out != nullptr && w.pos + (j - i) * 2 <= w.cap) {
      uint8_t* dst = w.out + w.pos;  // checked once for the container
      for (uint64_t k = i; k < j; ++k) {
        dst[0] = static_cast<uint8_t>(ids[k]);
        dst[1] = static_cast<uint8_t>(ids[k] >> 8);
        dst += 2;
      }
      w.pos += (j - i) * 2;
    } else {
      for (uint64_t k = i; k < j; ++k) {
        w.u16(static_cast<uint16_t>(ids[k] & 0xFFFFu));
      }
    }
    i = j;
  }
}

struct Reader {
  const uint8_t* p;
  uint64_t n;
  uint64_t pos = 0;
  bool has(uint64_t k) const { return n - pos >= k; }
  uint8_t u8() { return p[pos++]; }
  uint16_t u16() {
    const uint16_t v = static_cast<uint16_t>(p[pos] | (p[pos + 1] << 8));
    pos += 2;
    return v;
  }
  uint32_t u32() {
    const uint32_t v = static_cast<uint32_t>(p[pos]) | (static_cast<uint32_t>(p[pos + 1]) << 8) |
                       (static_cast<uint32_t>(p[pos + 2]) << 16) | (static_cast<uint32_t>(p[pos + 3]) << 24);
    pos += 4;
    return v;
  }
  uint64_t u64() {
    const uint64_t lo = u32();
    const uint64_t hi = u32();
    return lo | (hi << 32);
  }
};

int reject(const char* reference_code, const std::string& what) {
  set_last_error(std::string("MGIX stream rejected (") + reference_code + "): " + what);
  return MGX_ERR_FORMAT;
}

// One Roaring portable bitmap -> ascending ids appended to `out` (counted only when out == nullptr). Enforces what
// roaring_bitmap_portable_deserialize_safe + roaring_bitmap_internal_validate do (posting_list.cpp:1082-1090).
bool read_roaring(Reader r, uint64_t* count, uint32_t* out) {
  if (!r.has(4)) {
    return false;
  }
  const uint32_t cookie = r.u32();
  uint32_t n = 0;
  const uint8_t* run_flags = nullptr;
  if ((cookie & 0xFFFFu) == kCookieRuns) {
    n = (cookie >> 16) + 1;
    const uint64_t flag_bytes = (n + 7) / 8;
    if (!r.has(flag_bytes)) {
      return false;
    }
    run_flags = r.p + r.pos;
    r.pos += flag_bytes;
  } else if (cookie == kCookieNoRuns) {
    if (!r.has(4)) {
      return false;
    }
    n = r.u32();
    if (n > 65536) {
      return false;
    }
  } else {
    return false;
  }
  if (!r.has(static_cast<uint64_t>(n) * 4)) {
    return false;
  }
  Reader header = r;
  r.pos += static_cast<uint64_t>(n) * 4;
  if (run_flags == nullptr || n >= kNoOffsetThreshold) {
    if (!r.has(static_cast<uint64_t>(n) * 4)) {
      return false;
    }
    r.pos += static_cast<uint64_t>(n) * 4;  // offsets serve random access only
  }
  uint64_t total = 0;
  int64_t prev_key = -1;
  for (uint32_t c = 0; c < n; ++c) {
    const uint32_t key = header.u16();
    const uint32_t card = static_cast<uint32_t>(header.u16()) + 1;
    if (static_cast<int64_t>(key) <= prev_key) {
      return false;  // keys strictly ascending
    }
    prev_key = key;
    const uint32_t base = key << 16;
    const bool is_run = run_flags != nullptr && ((run_flags[c / 8] >> (c % 8)) & 1) != 0;
    if (is_run) {
      if (!r.has(2)) {
        return false;
      }
      const uint32_t n_runs = r.u16();
      if (n_runs == 0 || !r.has(static_cast<uint64_t>(n_runs) * 4)) {
        return false;
      }
      uint32_t seen = 0;
      int64_t last_end = -2;
      for (uint32_t i = 0; i < n_runs; ++i) {
        const uint32_t start = r.u16();
        const uint32_t len = r.u16();
        if (static_cast<int64_t>(start) <= last_end + 1 || start + len > 65535) {
          return false;  // runs ascending, disjoint and not adjacent
        }
        if (out != nullptr) {
          for (uint32_t v = start; v <= start + len; ++v) {
            out[total + seen + (v - start)] = base | v;
          }
        }
        seen += len + 1;
        last_end = start + len;
      }
      if (seen != card) {
        return false;
      }
    } else if (card > kArrayMax) {
      if (!r.has(kBitsetBytes)) {
        return false;
      }
      uint32_t seen = 0;
      for (uint32_t wi = 0; wi < 1024; ++wi) {
        uint64_t word = r.u64();
        while (word != 0) {
          const uint32_t bit = static_cast<uint32_t>(__builtin_ctzll(word));
          if (out != nullptr && seen < card) {
            out[total + seen] = base | (wi << 6) | bit;
          }
          ++seen;
          word &= word - 1;
        }
      }
      if (seen != card) {
        return false;
      }
    } else {
      if (!r.has(static_cast<uint64_t>(card) * 2)) {
        return false;
      }
      int64_t prev = -1;
      for (uint32_t i = 0; i < card; ++i) {
        const uint32_t v = r.u16();
        if (static_cast<int64_t>(v) <= prev) {
          return false;  // values strictly ascending
        }
        prev = v;
        if (out != nullptr) {
          out[total + i] = base | v;
        }
      }
    }
    total += card;
  }
  *count = total;
  return true;
}

// PostingList::Deserialize (posting_list.cpp:1023-1102) of one body.
bool read_posting(const uint8_t* body, uint64_t len, uint64_t* count, uint32_t* out) {
  Reader r{body, len};
  if (!r.has(1)) {
    return false;
  }
  const uint8_t strategy = r.u8();
  if (strategy > 1 || !r.has(4)) {
    return false;
  }
  const uint32_t size = r.u32();
  if (strategy == 0) {
    if (size > (len - r.pos) / 4) {
      return false;
    }
    uint64_t cumulative = 0;
    for (uint32_t i = 0; i < size; ++i) {  // IsValidDeltaEncoding, posting_list.cpp:132-148
      const uint32_t v = r.u32();
      if (i > 0 && v == 0) {
        return false;
      }
      cumulative += v;
      if (cumulative > 0xFFFFFFFFULL) {
        return false;
      }
      if (out != nullptr) {
        out[i] = static_cast<uint32_t>(cumulative);
      }
    }
    *count = size;
    return true;
  }
  if (size > len - r.pos) {
    return false;
  }
  return read_roaring(Reader{body + r.pos, size}, count, out);
}

}  // namespace
}  // namespace mgx

using namespace mgx;

int mgx_mgix_encode(const mgx_mgix_info_t* info, const uint8_t* term_bytes, const uint64_t* term_offsets,
                    const uint64_t* posting_offsets, const uint32_t* postings, double roaring_min_len, uint8_t* out,
                    uint64_t cap, uint64_t* out_len) {
  if (info == nullptr || out_len == nullptr ||
      (info->n_terms > 0 && (term_bytes == nullptr || term_offsets == nullptr || posting_offsets == nullptr))) {
    set_last_error("null argument");
    return MGX_ERR_INVALID_ARGUMENT;
  }
  *out_len = 0;
  const size_t width_len = strnlen(info->normalize_width, sizeof(info->normalize_width));
  Writer w{out, cap};
  w.bytes("MGIX", 4);
  w.u32(4);
  w.u32(static_cast<uint32_t>(info->ngram_size));
  w.u32(static_cast<uint32_t>(info->kanji_ngram_size));
  w.u8(info->cross_boundary != 0 ? 1 : 0);
  w.u8(info->normalize_nfkc != 0 ? 1 : 0);
  w.u32(static_cast<uint32_t>(width_len));
  w.bytes(info->normalize_width, width_len);
  w.u8(info->normalize_lower != 0 ? 1 : 0);
  w.u64(info->n_terms);
  const uint64_t header_len = w.pos;
  const uint64_t T = info->n_terms;
  const uint64_t P = T > 0 ? posting_offsets[T] : 0;
  const unsigned threads = codec_threads(P + T);
  // pass 1 (parallel over terms): validate the lists and size every record
  //   record = u32 term length | term | u64 body length | u8 strategy | u32 size | data
  std::vector<uint64_t> rec_off(T + 1, 0);
  std::atomic<int> failure{0};  // 1 = list not ascending, 2 = roaring body above 4 GiB
  const auto is_roaring = [&](uint64_t n) {
    // the representation the reference's list would be in: Roaring above 4096 entries since insertion
    // (posting_list.cpp:21,917-922), by density after Index::Optimize (:800-834)
    return n > kAutoRoaringEntries || (roaring_min_len > 0.0 && static_cast<double>(n) >= roaring_min_len);
  };
  const std::vector<uint64_t> bounds = balanced_bounds(
      T, threads, [&](uint64_t t) { return (t > 0 ? posting_offsets[t] - posting_offsets[0] : 0) + 8 * t; });
  parallel_bounds(bounds, [&](uint64_t t0, uint64_t t1, unsigned) {
    for (uint64_t t = t0; t < t1; ++t) {
      const uint32_t* ids = postings + posting_offsets[t];
      const uint64_t n = posting_offsets[t + 1] - posting_offsets[t];
      for (uint64_t i = 1; i < n; ++i) {
        if (ids[i] <= ids[i - 1]) {
          failure.store(1);
          return;
        }
      }
      uint64_t body = 5 + n * 4;
      if (is_roaring(n)) {
        uint32_t n_containers = 0;
        const uint64_t rb = roaring_size(ids, n, &n_containers);
        if (rb > 0xFFFFFFFFULL) {
          failure.store(2);
          return;
        }
        body = 5 + rb;
      }
      rec_off[t + 1] = 4 + (term_offsets[t + 1] - term_offsets[t]) + 8 + body;
    }
  });
  if (failure.load() == 1) {
    set_last_error("posting list is not strictly ascending");
    return MGX_ERR_INVALID_ARGUMENT;
  }
  if (failure.load() == 2) {
    set_last_error("roaring bitmap larger than 4 GiB (posting_list.cpp:1001-1008)");
    return MGX_ERR_UNSUPPORTED;
  }
  rec_off[0] = header_len;
  for (uint64_t t = 0; t < T; ++t) {
    rec_off[t + 1] += rec_off[t];
  }
  const uint64_t payload = rec_off[T];
  *out_len = payload + 4;
  if (out == nullptr || *out_len > cap) {
    set_last_error("output capacity too small");
    return MGX_ERR_CAPACITY;
  }
  // pass 2 (parallel): every record is written at its own offset
  parallel_bounds(bounds, [&](uint64_t t0, uint64_t t1, unsigned) {
    for (uint64_t t = t0; t < t1; ++t) {
      const uint64_t tl = term_offsets[t + 1] - term_offsets[t];
      const uint32_t* ids = postings + posting_offsets[t];
      const uint64_t n = posting_offsets[t + 1] - posting_offsets[t];
      Writer rw{out + rec_off[t], rec_off[t + 1] - rec_off[t]};
      rw.u32(static_cast<uint32_t>(tl));
      rw.bytes(term_bytes + term_offsets[t], tl);
      if (!is_roaring(n)) {
        rw.u64(5 + n * 4);
        rw.u8(0);
        rw.u32(static_cast<uint32_t>(n));
        uint8_t* dst = rw.out + rw.pos;  // the slice was sized for exactly these n words
        uint32_t prev = 0;
        for (uint64_t i = 0; i < n; ++i) {
          const uint32_t gap = ids[i] - prev;  // the first word is the doc id itself
          prev = ids[i];
          dst[0] = static_cast<uint8_t>(gap);
          dst[1] = static_cast<uint8_t>(gap >> 8);
          dst[2] = static_cast<uint8_t>(gap >> 16);
          dst[3] = static_cast<uint8_t>(gap >> 24);
          dst += 4;
        }
        rw.pos += n * 4;
      } else {
        uint32_t n_containers = 0;
        const uint64_t rb = roaring_size(ids, n, &n_containers);
        rw.u64(5 + rb);
        rw.u8(1);
        rw.u32(static_cast<uint32_t>(rb));
        write_roaring(rw, ids, n, n_containers);
      }
    }
  });
  const uint32_t crc = crc32_parallel(out, payload, threads);
  Writer tail{out + payload, 4};
  tail.u32(crc);
  return MGX_OK;
}

int mgx_mgix_decode(const uint8_t* data, uint64_t len, mgx_mgix_info_t* info, uint8_t* term_bytes,
                    uint64_t* term_offsets, uint64_t* posting_offsets, uint32_t* postings) {
  if (data == nullptr || info == nullptr) {
    set_last_error("null argument");
    return MGX_ERR_INVALID_ARGUMENT;
  }
  const bool fill = term_bytes != nullptr && term_offsets != nullptr && posting_offsets != nullptr &&
                    postings != nullptr;
  const mgx_mgix_info_t sized = *info;  // the sizes the caller allocated for (second call)
  std::memset(info, 0, sizeof(*info));
  // index_serialization.cpp:279-360
  if (len < 20) {
    return reject("kStorageInvalidFormat", "index data too short to be valid");
  }
  if (std::memcmp(data, "MGIX", 4) != 0) {
    return reject("kStorageInvalidFormat", "invalid magic number");
  }
  Reader r{data, len};
  r.pos = 4;
  const uint32_t version = r.u32();
  if (version < 1 || version > 4) {
    return reject("kStorageVersionMismatch", "unsupported index format version " + std::to_string(version));
  }
  uint64_t data_size = len;
  if (version >= 2) {
    const uint64_t min_header = version == 4 ? 31 : (version == 3 ? 25 : 20);
    if (len < min_header + 4) {
      return reject("kStorageInvalidFormat", "index data missing CRC32 trailer");
    }
    data_size = len - 4;
    Reader trailer{data, len};
    trailer.pos = data_size;
    if (trailer.u32() != crc32_parallel(data, data_size, codec_threads(data_size / 4))) {
      return reject("kStorageCRCMismatch", "CRC32 checksum mismatch in index data");
    }
  }
  r.n = data_size;
  info->version = version;
  info->ngram_size = static_cast<int32_t>(r.u32());
  info->kanji_ngram_size = info->ngram_size;
  info->cross_boundary = 1;
  if (version >= 3) {
    info->kanji_ngram_size = static_cast<int32_t>(r.u32());
    info->cross_boundary = r.u8() != 0 ? 1 : 0;
    if (version == 4) {
      info->normalize_nfkc = r.u8() != 0 ? 1 : 0;
      const uint32_t width_len = r.u32();
      if (width_len > data_size - r.pos - 1) {
        return reject("kStorageInvalidFormat", "normalize_width length exceeds payload size");
      }
      if (width_len >= sizeof(info->normalize_width)) {
        set_last_error("normalize_width longer than 31 bytes is not supported");
        return MGX_ERR_UNSUPPORTED;
      }
      std::memcpy(info->normalize_width, data + r.pos, width_len);
      r.pos += width_len;
      info->normalize_lower = r.u8() != 0 ? 1 : 0;
    }
  }
  if (r.pos > data_size || data_size - r.pos < 8) {
    return reject("kStorageInvalidFormat", "index data is truncated before term count");
  }
  const uint64_t term_count = r.u64();
  struct Record {
    std::string_view term;
    const uint8_t* body;
    uint64_t body_len;
    uint64_t count;
  };
  std::vector<Record> records;
  records.reserve(static_cast<size_t>(std::min<uint64_t>(term_count, (data_size - r.pos) / 17 + 1)));
  uint64_t total_term_bytes = 0;
  uint64_t total_postings = 0;
  // pass 1 (serial, headers only): record boundaries. A structural defect stops the scan; it is reported only if every
  // posting list in front of it deserialises (the reference loads record by record, :497-575)
  const char* structural_code = nullptr;
  const char* structural_what = nullptr;
  for (uint64_t i = 0; i < term_count; ++i) {
    if (!r.has(4)) {
      structural_code = "kStorageCorrupted";
      structural_what = "truncated index data at term header";
      break;
    }
    const uint32_t term_len = r.u32();
    if (term_len > kMaxTermLength) {
      structural_code = "kStorageCorrupted";
      structural_what = "term length exceeds maximum allowed size";
      break;
    }
    if (!r.has(term_len)) {
      structural_code = "kStorageCorrupted";
      structural_what = "truncated index data at term string";
      break;
    }
    Record rec;
    rec.term = std::string_view(reinterpret_cast<const char*>(data) + r.pos, term_len);
    r.pos += term_len;
    if (!r.has(8)) {
      structural_code = "kStorageCorrupted";
      structural_what = "truncated index data at posting list header";
      break;
    }
    rec.body_len = r.u64();
    if (rec.body_len > kMaxPostingBytes) {
      structural_code = "kStorageCorrupted";
      structural_what = "posting list size exceeds maximum allowed size";
      break;
    }
    if (!r.has(rec.body_len)) {
      structural_code = "kStorageCorrupted";
      structural_what = "truncated index data at posting list body";
      break;
    }
    rec.body = data + r.pos;
    r.pos += rec.body_len;
    rec.count = 0;
    records.push_back(rec);
  }
  // pass 2 (parallel, balanced by body bytes): PostingList::Deserialize's checks and the list sizes
  const unsigned threads = codec_threads(data_size / 4);
  {
    const uint8_t* base = records.empty() ? data : records.front().body;
    const std::vector<uint64_t> bounds = balanced_bounds(records.size(), threads, [&](uint64_t t) {
      return (t < records.size() ? static_cast<uint64_t>(records[t].body - base) : static_cast<uint64_t>(data + r.pos - base));
    });
    std::atomic<uint64_t> first_bad{UINT64_MAX};
    parallel_bounds(bounds, [&](uint64_t t0, uint64_t t1, unsigned) {
      for (uint64_t t = t0; t < t1; ++t) {
        if (!read_posting(records[t].body, records[t].body_len, &records[t].count, nullptr)) {
          uint64_t seen = first_bad.load();
          while (t < seen && !first_bad.compare_exchange_weak(seen, t)) {
          }
          return;
        }
      }
    });
    if (first_bad.load() != UINT64_MAX) {
      return reject("kIndexDeserializationFailed",
                    "failed to deserialize posting list of term " + std::string(records[first_bad.load()].term));
    }
  }
  if (structural_code != nullptr) {
    return reject(structural_code, structural_what);
  }
  for (const Record& rec : records) {
    total_term_bytes += rec.term.size();
    total_postings += rec.count;
  }
  // canonical order: ascending term bytes (the reference writes its hash map's order); a repeated term keeps its
  // LAST record, as new_postings[term] = ... does (:573)
  std::vector<uint32_t> order(records.size());
  std::iota(order.begin(), order.end(), 0u);
  const auto by_term = [&](uint32_t a, uint32_t b) { return records[a].term < records[b].term; };
  if (!std::is_sorted(order.begin(), order.end(), by_term)) {  // streams of mgx_mgix_encode already are
    std::stable_sort(order.begin(), order.end(), by_term);
  }
  std::vector<uint32_t> kept;
  kept.reserve(order.size());
  for (size_t i = 0; i < order.size(); ++i) {
    if (i + 1 < order.size() && records[order[i + 1]].term == records[order[i]].term) {
      total_term_bytes -= records[order[i]].term.size();
      total_postings -= records[order[i]].count;
      continue;
    }
    kept.push_back(order[i]);
  }
  info->n_terms = kept.size();
  info->n_postings = total_postings;
  info->term_bytes = total_term_bytes;
  if (!fill) {
    return MGX_OK;
  }
  if (sized.n_terms < info->n_terms || sized.n_postings < info->n_postings || sized.term_bytes < info->term_bytes) {
    set_last_error("output capacity too small (call once without outputs for the sizes)");
    return MGX_ERR_CAPACITY;
  }
  uint64_t tb = 0;
  uint64_t pp = 0;
  for (size_t i = 0; i < kept.size(); ++i) {
    const Record& rec = records[kept[i]];
    term_offsets[i] = tb;
    posting_offsets[i] = pp;
    tb += rec.term.size();
    pp += rec.count;
  }
  term_offsets[kept.size()] = tb;
  posting_offsets[kept.size()] = pp;
  // pass 3 (parallel, balanced by postings): every list decodes into its own slice
  const std::vector<uint64_t> fill_bounds =
      balanced_bounds(kept.size(), threads, [&](uint64_t t) { return posting_offsets[t] + 8 * t; });
  parallel_bounds(fill_bounds, [&](uint64_t t0, uint64_t t1, unsigned) {
    for (uint64_t i = t0; i < t1; ++i) {
      const Record& rec = records[kept[i]];
      std::memcpy(term_bytes + term_offsets[i], rec.term.data(), rec.term.size());
      uint64_t count = 0;
      read_posting(rec.body, rec.body_len, &count, postings + posting_offsets[i]);
    }
  });
  return MGX_OK;
}
