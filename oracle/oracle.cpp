// oracle/oracle.cpp — TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the reference's search core: tokenizer, inverted index,
// posting-list set algebra, BM25 and score sort, and the regular path of the
// search pipeline. Every function cites the reference file:line it follows
// (paths relative to /root/reference/). It is the checker for the CUDA path
// and the "port" CPU baseline; it is never linked into or called by libmgx.so.
//
// Deliberate differences from the reference, none of which can change a result:
//   * posting lists are plain sorted std::vector<uint32_t> (the reference keeps
//     the same set as fixed-width deltas or a Roaring bitmap,
//     posting_list.cpp:20-21,242-324 -- an encoding, not a semantic);
//   * no locks / RCU snapshots (single writer, readers after build);
//   * logging dropped.

#include "oracle.h"

#include <algorithm>
#include <charconv>
#include <atomic>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <queue>
#include <string>
#include <string_view>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// Tokenizer (src/utils/string_utils.cpp)
// ---------------------------------------------------------------------------

// TryParseUtf8Char, string_utils.cpp:92-164.
int TryParseUtf8Char(const unsigned char* data, size_t available, uint32_t* out_cp) {
  if (available == 0) {
    return -1;
  }
  const unsigned char b0 = data[0];
  if ((b0 & 0x80) == 0) {
    *out_cp = b0;
    return 1;
  }
  if ((b0 & 0xE0) == 0xC0) {
    if (b0 < 0xC2 || available < 2) {  // :105-110 overlong / truncated
      return -1;
    }
    if ((data[1] & 0xC0) != 0x80) {
      return -1;
    }
    *out_cp = (static_cast<uint32_t>(b0 & 0x1F) << 6) | (data[1] & 0x3F);
    return 2;
  }
  if ((b0 & 0xF0) == 0xE0) {
    if (available < 3) {
      return -1;
    }
    if ((data[1] & 0xC0) != 0x80 || (data[2] & 0xC0) != 0x80) {
      return -1;
    }
    const uint32_t cp =
        (static_cast<uint32_t>(b0 & 0x0F) << 12) | (static_cast<uint32_t>(data[1] & 0x3F) << 6) | (data[2] & 0x3F);
    if (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF)) {  // :131 overlong or surrogate
      return -1;
    }
    *out_cp = cp;
    return 3;
  }
  if ((b0 & 0xF8) == 0xF0) {
    if (b0 > 0xF4 || available < 4) {  // :139-144
      return -1;
    }
    if ((data[1] & 0xC0) != 0x80 || (data[2] & 0xC0) != 0x80 || (data[3] & 0xC0) != 0x80) {
      return -1;
    }
    const uint32_t cp = (static_cast<uint32_t>(b0 & 0x07) << 18) | (static_cast<uint32_t>(data[1] & 0x3F) << 12) |
                        (static_cast<uint32_t>(data[2] & 0x3F) << 6) | (data[3] & 0x3F);
    if (cp < 0x10000 || cp > 0x10FFFF) {  // :156
      return -1;
    }
    *out_cp = cp;
    return 4;
  }
  return -1;  // :162 invalid start byte
}

// Utf8ToCodepoints, string_utils.cpp:200-219: invalid byte => skip one byte.
std::vector<uint32_t> Utf8ToCodepoints(std::string_view text) {
  std::vector<uint32_t> cps;
  cps.reserve(text.size() / 2 + 1);
  const auto* data = reinterpret_cast<const unsigned char*>(text.data());
  size_t i = 0;
  while (i < text.size()) {
    uint32_t cp = 0;
    const int len = TryParseUtf8Char(data + i, text.size() - i, &cp);
    if (len > 0) {
      cps.push_back(cp);
      i += static_cast<size_t>(len);
    } else {
      ++i;
    }
  }
  return cps;
}

// CodepointsToUtf8, string_utils.cpp:243-272 (surrogates / >0x10FFFF skipped).
void AppendUtf8(std::string& out, const uint32_t* begin, const uint32_t* end) {
  for (const uint32_t* it = begin; it != end; ++it) {
    const uint32_t cp = *it;
    if ((cp >= 0xD800 && cp <= 0xDFFF) || cp > 0x10FFFF) {
      continue;
    }
    if (cp <= 0x7F) {
      out += static_cast<char>(cp);
    } else if (cp <= 0x7FF) {
      out += static_cast<char>(0xC0 | (cp >> 6));
      out += static_cast<char>(0x80 | (cp & 0x3F));
    } else if (cp <= 0xFFFF) {
      out += static_cast<char>(0xE0 | (cp >> 12));
      out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
      out += static_cast<char>(0x80 | (cp & 0x3F));
    } else {
      out += static_cast<char>(0xF0 | (cp >> 18));
      out += static_cast<char>(0x80 | ((cp >> 12) & 0x3F));
      out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
      out += static_cast<char>(0x80 | (cp & 0x3F));
    }
  }
}

std::string CodepointsToUtf8(const uint32_t* begin, const uint32_t* end) {
  std::string out;
  out.reserve(static_cast<size_t>(end - begin) * 3);
  AppendUtf8(out, begin, end);
  return out;
}

// CountCodePoints, string_utils.cpp:655-669.
size_t CountCodePoints(std::string_view text) {
  size_t count = 0;
  const auto* data = reinterpret_cast<const unsigned char*>(text.data());
  for (size_t i = 0; i < text.size();) {
    uint32_t cp = 0;
    const int len = TryParseUtf8Char(data + i, text.size() - i, &cp);
    if (len < 0) {
      ++i;
      continue;
    }
    i += static_cast<size_t>(len);
    ++count;
  }
  return count;
}

// IsCJKIdeograph, string_utils.cpp:441-448 (ranges :176-187).
bool IsCJKIdeograph(uint32_t cp) {
  return (cp >= 0x4E00 && cp <= 0x9FFF) || (cp >= 0x3400 && cp <= 0x4DBF) || (cp >= 0x20000 && cp <= 0x2A6DF) ||
         (cp >= 0x2A700 && cp <= 0x2B73F) || (cp >= 0x2B740 && cp <= 0x2B81F) || (cp >= 0xF900 && cp <= 0xFAFF);
}

// GenerateNgrams, string_utils.cpp:382-423.
std::vector<std::string> GenerateNgrams(std::string_view text, int n) {
  std::vector<std::string> ngrams;
  const std::vector<uint32_t> cps = Utf8ToCodepoints(text);
  const size_t cp_count = cps.size();
  if (cp_count == 0 || n <= 0) {
    return ngrams;
  }
  if (n == 1) {
    ngrams.reserve(cp_count);
    for (size_t i = 0; i < cp_count; ++i) {
      ngrams.push_back(CodepointsToUtf8(cps.data() + i, cps.data() + i + 1));
    }
    return ngrams;
  }
  if (cp_count < static_cast<size_t>(n)) {
    return ngrams;
  }
  ngrams.reserve(cp_count - static_cast<size_t>(n) + 1);
  for (size_t i = 0; i <= cp_count - static_cast<size_t>(n); ++i) {
    ngrams.push_back(CodepointsToUtf8(cps.data() + i, cps.data() + i + n));
  }
  return ngrams;
}

// GenerateHybridNgrams, string_utils.cpp:452-509.
std::vector<std::string> GenerateHybridNgrams(std::string_view text, int ascii_n, int kanji_n, bool cross_boundary) {
  std::vector<std::string> ngrams;
  if (ascii_n <= 0 || kanji_n <= 0) {  // :456
    return ngrams;
  }
  const std::vector<uint32_t> cps = Utf8ToCodepoints(text);
  const size_t cp_count = cps.size();
  if (cp_count == 0) {
    return ngrams;
  }
  ngrams.reserve(cp_count);
  for (size_t i = 0; i < cp_count; ++i) {
    const bool start_is_cjk = IsCJKIdeograph(cps[i]);  // :484 size chosen by the START code point
    const int size = start_is_cjk ? kanji_n : ascii_n;
    if (i + static_cast<size_t>(size) > cp_count) {  // :487
      continue;
    }
    if (!cross_boundary) {  // :491-503 legacy: reject windows mixing CJK / non-CJK
      bool crossed = false;
      for (int j = 1; j < size; ++j) {
        if (IsCJKIdeograph(cps[i + static_cast<size_t>(j)]) != start_is_cjk) {
          crossed = true;
          break;
        }
      }
      if (crossed) {
        continue;
      }
    }
    ngrams.push_back(CodepointsToUtf8(cps.data() + i, cps.data() + i + size));
  }
  return ngrams;
}

// GenerateQueryNgrams, string_utils.cpp:639-653.
std::vector<std::string> GenerateQueryNgrams(std::string_view normalized, int ngram_size, int kanji_ngram_size,
                                             bool cross_boundary) {
  if (kanji_ngram_size > 0) {
    const int effective = (ngram_size > 0) ? ngram_size : 2;
    return GenerateHybridNgrams(normalized, effective, kanji_ngram_size, cross_boundary);
  }
  if (ngram_size == 0) {
    return GenerateHybridNgrams(normalized, 2, 1, true);  // header defaults, string_utils.h
  }
  return GenerateNgrams(normalized, ngram_size);
}

// DeduplicateSorted, string_utils.h:192-196.
template <typename T>
void DeduplicateSorted(std::vector<T>& vec) {
  std::sort(vec.begin(), vec.end());
  vec.erase(std::unique(vec.begin(), vec.end()), vec.end());
}

// ---------------------------------------------------------------------------
// Index (src/index/index.cpp) over sorted-vector posting lists
// ---------------------------------------------------------------------------

using DocId = uint32_t;
using Posting = std::vector<DocId>;

// Read-only view of one posting list (either a hash-map entry or a CSR slice).
struct Span {
  const DocId* p = nullptr;
  size_t n = 0;
  bool found = false;
  const DocId* begin() const { return p; }
  const DocId* end() const { return p + n; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
};

struct SvHash {
  using is_transparent = void;
  size_t operator()(std::string_view s) const { return std::hash<std::string_view>{}(s); }
  size_t operator()(const std::string& s) const { return std::hash<std::string_view>{}(s); }
};
struct SvEq {
  using is_transparent = void;
  bool operator()(std::string_view a, std::string_view b) const { return a == b; }
};

}  // namespace

struct orc_index {
  int ngram_size = 2;
  int kanji_ngram_size = 2;  // effective, index.cpp:32
  bool cross_boundary = true;
  std::unordered_map<std::string, Posting, SvHash, SvEq> postings;
  // Bulk-built indexes (orc_index_build_bulk) keep the same term -> sorted doc ids mapping as flat
  // CSR arrays keyed by the packed n-gram instead of 10^7 heap-allocated map entries.
  bool csr = false;
  int csr_width = 0;
  std::vector<uint64_t> csr_keys;     // ascending
  std::vector<uint64_t> csr_offsets;  // [terms + 1]
  std::vector<DocId> csr_postings;

  // DocumentStore stand-in: normalised text per doc id (document_store.cpp:144-147).
  // Two backings: owned strings for incremental adds, or a borrowed arena with
  // ascending doc ids for bulk builds.
  std::unordered_map<DocId, std::string> texts;
  std::unordered_set<DocId> known_docs;  // every doc id ever added (DocumentStore::GetAllDocIds)
  const uint8_t* arena = nullptr;
  const uint64_t* arena_offsets = nullptr;
  const uint32_t* arena_doc_ids = nullptr;
  uint64_t arena_docs = 0;
  bool arena_sequential = false;  // doc_ids[i] == doc_ids[0] + i

  // BM25Stats, server_types.h:157-220.
  uint64_t total_doc_length = 0;
  uint64_t doc_count = 0;

  Span Find(std::string_view term) const;

  // text pointer or nullptr (VisitNormalizedTextsFor, document_store_retrieval.cpp:289-322)
  bool GetText(DocId doc, std::string_view* out) const {
    if (arena != nullptr) {
      uint64_t pos;
      if (arena_sequential) {
        if (doc < arena_doc_ids[0] || doc - arena_doc_ids[0] >= arena_docs) {
          return false;
        }
        pos = doc - arena_doc_ids[0];
      } else {
        const uint32_t* it = std::lower_bound(arena_doc_ids, arena_doc_ids + arena_docs, doc);
        if (it == arena_doc_ids + arena_docs || *it != doc) {
          return false;
        }
        pos = static_cast<uint64_t>(it - arena_doc_ids);
      }
      const uint64_t b = arena_offsets[pos];
      const uint64_t e = arena_offsets[pos + 1];
      if (e == b) {
        return false;  // empty text is never stored
      }
      *out = std::string_view(reinterpret_cast<const char*>(arena) + b, e - b);
      return true;
    }
    auto it = texts.find(doc);
    if (it == texts.end()) {
      return false;
    }
    *out = it->second;
    return true;
  }
};

namespace {

std::vector<std::string> IndexNgrams(const orc_index& idx, std::string_view text) {
  // index.cpp:41-42,88: always the hybrid generator with the effective kanji size.
  auto ngrams = GenerateHybridNgrams(text, idx.ngram_size, idx.kanji_ngram_size, idx.cross_boundary);
  DeduplicateSorted(ngrams);
  return ngrams;
}

// PostingList::Add, posting_list.cpp:242-289 (set insert keeping order).
void PostingAdd(Posting& list, DocId doc) {
  if (list.empty() || list.back() < doc) {
    list.push_back(doc);
    return;
  }
  auto pos = std::lower_bound(list.begin(), list.end(), doc);
  if (pos == list.end() || *pos != doc) {
    list.insert(pos, doc);
  }
}

// PostingList::AddBatch, posting_list.cpp:291-324 (set_union with sorted batch).
void PostingAddBatch(Posting& list, const std::vector<DocId>& sorted_docs) {
  if (list.empty()) {
    list = sorted_docs;
    list.erase(std::unique(list.begin(), list.end()), list.end());
    return;
  }
  if (!sorted_docs.empty() && list.back() < sorted_docs.front()) {
    size_t old = list.size();
    list.insert(list.end(), sorted_docs.begin(), sorted_docs.end());
    list.erase(std::unique(list.begin() + static_cast<std::ptrdiff_t>(old) - 1, list.end()), list.end());
    return;
  }
  Posting merged;
  merged.reserve(list.size() + sorted_docs.size());
  std::set_union(list.begin(), list.end(), sorted_docs.begin(), sorted_docs.end(), std::back_inserter(merged));
  list.swap(merged);
}

void StoreTextAndStats(orc_index& idx, DocId doc, std::string_view text) {
  idx.known_docs.insert(doc);
  if (!text.empty()) {
    idx.texts[doc] = std::string(text);
    // binlog_event_processor.cpp:98-100 / server_orchestrator.cpp:758-772
    idx.total_doc_length += CountCodePoints(text);
    idx.doc_count += 1;
  }
}

// TakePostingSnapshots, index.cpp:728-747.
std::vector<Span> Snapshots(const orc_index& idx, const std::vector<std::string_view>& terms) {
  std::vector<Span> out;
  out.reserve(terms.size());
  for (auto term : terms) {
    out.push_back(idx.Find(term));
  }
  return out;
}

// Index::SearchAnd, index.cpp:199-368. The planner branches (:221-332) pick a
// cheaper evaluation order but compute the same set; the standard path is
// restated, then limit/reverse is applied as :356-366 (and GetTopN :476-545).
std::vector<DocId> SearchAnd(const orc_index& idx, const std::vector<std::string_view>& terms, size_t limit,
                             bool reverse) {
  if (terms.empty()) {
    return {};
  }
  auto snaps = Snapshots(idx, terms);
  for (const auto& s : snaps) {
    if (!s.found) {
      return {};
    }
  }
  std::vector<DocId> result(snaps[0].begin(), snaps[0].end());  // GetAll() materialises a copy (:338)
  for (size_t i = 1; i < snaps.size(); ++i) {
    const std::vector<DocId> term_docs(snaps[i].begin(), snaps[i].end());  // GetAll() (:342)
    std::vector<DocId> inter;
    std::set_intersection(result.begin(), result.end(), term_docs.begin(), term_docs.end(),
                          std::back_inserter(inter));
    result = std::move(inter);
    if (result.empty()) {
      break;
    }
  }
  if (limit > 0 && result.size() > limit) {
    if (reverse) {
      result.erase(result.begin(), result.begin() + static_cast<std::ptrdiff_t>(result.size() - limit));
      std::reverse(result.begin(), result.end());
    } else {
      result.resize(limit);
    }
  } else if (reverse) {
    std::reverse(result.begin(), result.end());
  }
  return result;
}

// PostingList::RetainPresent, posting_list.cpp:432-474: keep sorted candidates present in the list.
std::vector<DocId> RetainPresent(const Span& list, const std::vector<DocId>& sorted_candidates) {
  std::vector<DocId> out;
  size_t j = 0;
  for (DocId c : sorted_candidates) {
    while (j < list.size() && list.p[j] < c) {
      ++j;
    }
    if (j < list.size() && list.p[j] == c) {
      out.push_back(c);  // duplicates in the candidate list are each retained
    }
  }
  return out;
}

// Index::FilterByNgrams, index.cpp:370-416.
std::vector<DocId> FilterByNgrams(const orc_index& idx, const std::vector<DocId>& candidates,
                                  const std::vector<std::string_view>& terms) {
  if (candidates.empty()) {
    return {};
  }
  auto snaps = Snapshots(idx, terms);
  if (snaps.empty()) {
    return candidates;
  }
  for (const auto& s : snaps) {
    if (!s.found) {
      return {};
    }
  }
  const bool ascending = std::is_sorted(candidates.begin(), candidates.end());
  std::vector<DocId> sorted_candidates;
  if (!ascending) {
    sorted_candidates = candidates;
    std::sort(sorted_candidates.begin(), sorted_candidates.end());
  }
  std::vector<DocId> retained = RetainPresent(snaps[0], ascending ? candidates : sorted_candidates);
  for (size_t i = 1; i < snaps.size() && !retained.empty(); ++i) {
    retained = RetainPresent(snaps[i], retained);
  }
  if (ascending || retained.empty()) {
    return retained;
  }
  const std::unordered_set<DocId> kept(retained.begin(), retained.end());
  std::vector<DocId> ordered;
  for (DocId d : candidates) {
    if (kept.find(d) != kept.end()) {
      ordered.push_back(d);
    }
  }
  return ordered;
}

// Index::SearchOr, index.cpp:418-448.
std::vector<DocId> SearchOr(const orc_index& idx, const std::vector<std::string_view>& terms) {
  if (terms.empty()) {
    return {};
  }
  std::vector<DocId> result;
  std::vector<DocId> temp;
  for (const auto& s : Snapshots(idx, terms)) {
    if (s.found) {
      temp.clear();
      std::set_union(result.begin(), result.end(), s.begin(), s.end(), std::back_inserter(temp));
      result.swap(temp);
    }
  }
  return result;
}

// Index::SearchNot, index.cpp:450-486.
std::vector<DocId> SearchNot(const orc_index& idx, const std::vector<DocId>& all_docs,
                             const std::vector<std::string_view>& terms) {
  if (terms.empty()) {
    return all_docs;
  }
  const std::vector<DocId> excluded = SearchOr(idx, terms);
  std::vector<DocId> result;
  std::set_difference(all_docs.begin(), all_docs.end(), excluded.begin(), excluded.end(), std::back_inserter(result));
  return result;
}

// Index::SearchByThreshold, index.cpp:488-578.
std::vector<DocId> SearchByThreshold(const orc_index& idx, const std::vector<std::string_view>& terms,
                                     size_t threshold) {
  if (terms.empty() || threshold == 0) {
    return {};
  }
  std::vector<std::string_view> unique_terms = terms;
  DeduplicateSorted(unique_terms);
  if (threshold > unique_terms.size()) {
    return {};
  }
  if (threshold == unique_terms.size()) {
    return SearchAnd(idx, unique_terms, 0, false);
  }
  std::vector<Span> valid;
  for (const auto& s : Snapshots(idx, unique_terms)) {
    if (s.found) {
      valid.push_back(s);
    }
  }
  if (valid.size() < threshold) {
    return {};
  }
  using HeapEntry = std::tuple<DocId, size_t, size_t>;
  std::priority_queue<HeapEntry, std::vector<HeapEntry>, std::greater<HeapEntry>> heap;
  for (size_t i = 0; i < valid.size(); ++i) {
    if (!valid[i].empty()) {
      heap.emplace(valid[i].p[0], i, 0);
    }
  }
  std::vector<DocId> result;
  DocId current = 0;
  size_t count = 0;
  bool has_current = false;
  while (!heap.empty()) {
    auto [doc, li, pos] = heap.top();
    heap.pop();
    if (!has_current || doc != current) {
      if (has_current && count >= threshold) {
        result.push_back(current);
      }
      current = doc;
      count = 1;
      has_current = true;
    } else {
      ++count;
    }
    if (pos + 1 < valid[li].size()) {
      heap.emplace(valid[li].p[pos + 1], li, pos + 1);
    }
  }
  if (has_current && count >= threshold) {
    result.push_back(current);
  }
  return result;
}

// ---------------------------------------------------------------------------
// BM25 (src/index/bm25_scorer.cpp) and score sort (src/query/result_sorter.cpp)
// ---------------------------------------------------------------------------

// BM25Scorer::ComputeIDF, bm25_scorer.cpp:14-25.
double ComputeIDF(uint64_t total_docs, uint64_t doc_freq) {
  if (total_docs == 0) {
    return 0.0;
  }
  if (doc_freq > total_docs) {
    doc_freq = total_docs;
  }
  const auto n = static_cast<double>(total_docs);
  const auto df = static_cast<double>(doc_freq);
  return std::log((n - df + 0.5) / (df + 0.5) + 1.0);
}

// BM25Scorer::CountTermOccurrences, bm25_scorer.cpp:27-45 (bytes, non-overlapping).
uint32_t CountTermOccurrences(std::string_view text, std::string_view term) {
  if (text.empty() || term.empty() || term.size() > text.size()) {
    return 0;
  }
  uint32_t count = 0;
  size_t pos = 0;
  while (pos <= text.size() - term.size()) {
    const auto found = text.find(term, pos);
    if (found == std::string_view::npos) {
      break;
    }
    ++count;
    pos = found + term.size();
  }
  return count;
}

// BM25Scorer::ScoreDocuments, bm25_scorer.cpp:47-99. The arithmetic is written
// operation by operation so that -ffp-contract=off reproduces the reference's
// unfused evaluation order.
double ScoreOne(std::string_view text, const std::vector<std::string_view>& terms, const std::vector<double>& idfs,
                double avg_doc_length, double k1, double b) {
  double score = 0.0;
  const auto doc_length = static_cast<double>(CountCodePoints(text));
  for (size_t i = 0; i < terms.size(); ++i) {
    const auto tf = static_cast<double>(CountTermOccurrences(text, terms[i]));
    if (tf > 0.0) {
      const double length_norm = 1.0 - b + b * doc_length / std::max(avg_doc_length, 1.0);
      const double numerator = tf * (k1 + 1.0);
      const double denominator = tf + k1 * length_norm;
      score += idfs[i] * numerator / denominator;
    }
  }
  return score;
}

void ScoreDocuments(const orc_index& idx, const DocId* candidates, size_t n, const std::vector<std::string_view>& terms,
                    const uint64_t* dfs, uint64_t total_docs, double avg_doc_length, double k1, double b,
                    double* out) {
  std::vector<double> idfs;
  idfs.reserve(terms.size());
  for (size_t i = 0; i < terms.size(); ++i) {
    idfs.push_back(ComputeIDF(total_docs, dfs[i]));
  }
  for (size_t c = 0; c < n; ++c) {
    std::string_view text;
    double score = 0.0;
    if (idx.GetText(candidates[c], &text) && !text.empty()) {
      score = ScoreOne(text, terms, idfs, avg_doc_length, k1, b);
    }
    out[c] = score;
  }
}

// ResultSorter::SortByScore, result_sorter.cpp:661-716 (kPartialSortThreshold = 0.5).
std::vector<DocId> SortByScore(const DocId* results, const double* scores, size_t n, bool descending, uint32_t limit,
                               uint32_t offset) {
  if (n == 0) {
    return {};
  }
  struct Entry {
    size_t index;
    DocId doc;
    double score;
  };
  std::vector<Entry> entries;
  entries.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    entries.push_back({i, results[i], scores[i]});
  }
  auto cmp = [descending](const Entry& l, const Entry& r) {
    if (l.score == r.score) {
      return descending ? (l.doc > r.doc) : (l.doc < r.doc);  // :681-686 tie => same direction on doc_id
    }
    return descending ? (l.score > r.score) : (l.score < r.score);
  };
  const uint64_t needed64 = static_cast<uint64_t>(offset) + static_cast<uint64_t>(limit);
  const size_t needed = (limit == 0 || needed64 > entries.size()) ? entries.size() : static_cast<size_t>(needed64);
  const bool partial =
      needed < entries.size() && static_cast<double>(needed) < static_cast<double>(entries.size()) * 0.5;
  if (partial) {
    std::partial_sort(entries.begin(), entries.begin() + static_cast<std::ptrdiff_t>(needed), entries.end(), cmp);
  } else {
    std::sort(entries.begin(), entries.end(), cmp);
  }
  const size_t start = std::min(static_cast<size_t>(offset), entries.size());
  const size_t end = (limit == 0) ? entries.size() : std::min(start + static_cast<size_t>(limit), entries.size());
  std::vector<DocId> out;
  out.reserve(end - start);
  for (size_t i = start; i < end; ++i) {
    out.push_back(results[entries[i].index]);
  }
  return out;
}

// ---------------------------------------------------------------------------
// Search pipeline, regular path (src/server/search_pipeline.cpp)
// ---------------------------------------------------------------------------

struct TermInfo {  // SearchTermInfo, search_pipeline.h
  std::vector<std::string> ngrams;
  size_t estimated_size = 0;
  uint64_t df = 0;
  std::string normalized;
  size_t source_index = 0;  // position in the caller's term table (for out_df)
};

std::vector<std::string_view> Views(const std::vector<std::string>& v) {
  return std::vector<std::string_view>(v.begin(), v.end());
}

// query::SearchNormalizedSubstring, query/substring_search.h:24-42: every stored
// text in ascending doc-id order.
std::vector<DocId> SearchNormalizedSubstring(const orc_index& idx, std::string_view term) {
  std::vector<DocId> matches;
  if (term.empty()) {
    return matches;
  }
  if (idx.arena != nullptr) {
    for (uint64_t i = 0; i < idx.arena_docs; ++i) {
      const uint64_t b = idx.arena_offsets[i];
      const uint64_t e = idx.arena_offsets[i + 1];
      if (e > b && std::string_view(reinterpret_cast<const char*>(idx.arena) + b, e - b).find(term) !=
                       std::string_view::npos) {
        matches.push_back(idx.arena_doc_ids[i]);
      }
    }
    if (!idx.arena_sequential) {
      std::sort(matches.begin(), matches.end());
    }
    return matches;
  }
  for (const auto& [doc, text] : idx.texts) {
    if (text.find(term) != std::string::npos) {
      matches.push_back(doc);
    }
  }
  std::sort(matches.begin(), matches.end());
  return matches;
}

// SearchTermDocuments, search_pipeline.cpp:438-446.
std::vector<DocId> SearchTermDocuments(const orc_index& idx, const TermInfo& ti) {
  if (ti.ngrams.empty()) {
    return SearchNormalizedSubstring(idx, ti.normalized);
  }
  return SearchAnd(idx, Views(ti.ngrams), 0, false);
}

// GenerateTermInfos + PopulateTermDocumentFrequency, search_pipeline.cpp:569-603, 542-565.
// Index::NormalizeText is the identity here: the path's contract is that text
// and terms arrive normalised (index.h:83-84; ICU stays on the host).
TermInfo MakeTermInfo(const orc_index& idx, std::string_view term, const orc_query_params_t& p, bool compute_df) {
  TermInfo ti;
  ti.normalized = std::string(term);
  ti.ngrams = GenerateQueryNgrams(ti.normalized, p.ngram_size, p.kanji_ngram_size, p.cross_boundary != 0);
  DeduplicateSorted(ti.ngrams);
  size_t min_size = std::numeric_limits<size_t>::max();
  for (const auto& g : ti.ngrams) {
    const uint64_t size = idx.Find(g).size();  // EstimatePostingSize, index.cpp:756-759
    if (size > 0) {
      min_size = std::min(min_size, static_cast<size_t>(size));
    } else {
      min_size = 0;
      break;
    }
  }
  ti.estimated_size = min_size;
  ti.df = 0;
  if (compute_df && !ti.ngrams.empty() && ti.estimated_size != 0 &&
      ti.estimated_size != std::numeric_limits<size_t>::max()) {
    const auto candidates = SearchAnd(idx, Views(ti.ngrams), 0, false);
    uint64_t matching = 0;
    for (DocId d : candidates) {
      std::string_view text;
      if (idx.GetText(d, &text) && text.find(ti.normalized) != std::string_view::npos) {
        ++matching;
      }
    }
    ti.df = matching;
  }
  return ti;
}

// Private CJK classifier of the pipeline, search_pipeline.cpp:73-78 (wider than the
// tokenizer's: it also covers 2B820-2CEAF).
bool PipelineIsCjk(uint32_t cp) {
  return (cp >= 0x4E00 && cp <= 0x9FFF) || (cp >= 0x3400 && cp <= 0x4DBF) || (cp >= 0x20000 && cp <= 0x2A6DF) ||
         (cp >= 0x2A700 && cp <= 0x2B73F) || (cp >= 0x2B740 && cp <= 0x2B81F) || (cp >= 0x2B820 && cp <= 0x2CEAF) ||
         (cp >= 0xF900 && cp <= 0xFAFF);
}

// HasUncoveredHybridFragment, search_pipeline.cpp:80-136.
bool HasUncoveredHybridFragment(std::string_view term, int ngram_size, int kanji_ngram_size, bool cross_boundary) {
  if (term.empty() || kanji_ngram_size <= 0) {
    return false;
  }
  const int ascii_n = ngram_size > 0 ? ngram_size : 2;
  const auto cps = Utf8ToCodepoints(term);
  if (cps.size() < 2) {
    return false;
  }
  bool has_cjk = false;
  bool has_non = false;
  for (uint32_t cp : cps) {
    (PipelineIsCjk(cp) ? has_cjk : has_non) = true;
  }
  if (!has_cjk || !has_non) {
    return false;
  }
  std::vector<bool> covered(cps.size(), false);
  for (size_t i = 0; i < cps.size(); ++i) {
    const bool start_cjk = PipelineIsCjk(cps[i]);
    const int size = start_cjk ? kanji_ngram_size : ascii_n;
    if (size <= 0 || i + static_cast<size_t>(size) > cps.size()) {
      continue;
    }
    if (!cross_boundary) {
      bool crossed = false;
      for (int j = 1; j < size; ++j) {
        if (PipelineIsCjk(cps[i + static_cast<size_t>(j)]) != start_cjk) {
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

// ShouldApplyVerifyText, search_pipeline.cpp:48-66.
bool ShouldApplyVerifyText(int mode, const std::vector<std::string_view>& terms) {
  if (mode == 1) {
    return true;
  }
  if (mode == 2) {
    for (auto t : terms) {
      for (unsigned char ch : t) {
        if (ch >= 0x80) {
          return false;
        }
      }
    }
    return true;
  }
  return false;
}

// PostFilterByText / RetainCandidatesMatchingText, search_pipeline.cpp:1239-1246, 386-402:
// a candidate without stored text is kept.
std::vector<DocId> PostFilterByText(const orc_index& idx, const std::vector<DocId>& candidates,
                                    const std::vector<std::string_view>& terms) {
  std::vector<DocId> kept;
  for (DocId d : candidates) {
    std::string_view text;
    bool ok = true;
    if (idx.GetText(d, &text)) {
      for (auto t : terms) {
        if (text.find(t) == std::string_view::npos) {
          ok = false;
          break;
        }
      }
    }
    if (ok) {
      kept.push_back(d);
    }
  }
  return kept;
}

struct QueryOutput {
  std::vector<DocId> results;  // ascending result set
  std::vector<TermInfo> term_infos;
};

// ExecuteFullPipeline regular path (:2002-2033) + Execute (:795-869) + ApplyNotFilter (:871-932).
QueryOutput RunQuery(const orc_index& idx, const orc_query_params_t& p, const std::vector<std::string_view>& terms,
                     size_t first_term_index, const std::vector<std::string_view>& not_terms) {
  QueryOutput out;
  for (size_t i = 0; i < terms.size(); ++i) {
    out.term_infos.push_back(MakeTermInfo(idx, terms[i], p, p.compute_score != 0));
    out.term_infos.back().source_index = first_term_index + i;
  }
  // :2012-2014 std::sort by estimated_size. libstdc++ runs a plain insertion
  // sort for <= 16 elements, i.e. equal keys keep their query order; a stable
  // sort reproduces that (queries are limited to 64 terms, query_ast.h:184-185;
  // beyond 16 the reference's own order for equal sizes is unspecified).
  std::stable_sort(out.term_infos.begin(), out.term_infos.end(),
                   [](const TermInfo& l, const TermInfo& r) { return l.estimated_size < r.estimated_size; });

  // Execute :804-810 early exit
  for (const auto& ti : out.term_infos) {
    if ((ti.estimated_size == 0 || ti.estimated_size == std::numeric_limits<size_t>::max()) &&
        (!ti.ngrams.empty() || ti.normalized.empty())) {
      return out;  // empty_term_detected => results cleared
    }
  }
  std::vector<DocId> results;
  if (!out.term_infos.empty()) {
    results = SearchTermDocuments(idx, out.term_infos[0]);
    for (size_t i = 1; i < out.term_infos.size() && !results.empty(); ++i) {
      const auto& ti = out.term_infos[i];
      if (ti.ngrams.empty() || results.size() > p.filter_threshold) {
        const auto and_results = SearchTermDocuments(idx, ti);
        std::vector<DocId> inter;
        std::set_intersection(results.begin(), results.end(), and_results.begin(), and_results.end(),
                              std::back_inserter(inter));
        results = std::move(inter);
      } else {
        results = FilterByNgrams(idx, results, Views(ti.ngrams));  // :826-828
      }
    }
  }
  // ApplyNotFilter :871-932
  if (!results.empty() && !not_terms.empty()) {
    std::vector<DocId> excluded;
    std::vector<DocId> temp;
    for (auto nt : not_terms) {
      const TermInfo ti = MakeTermInfo(idx, nt, p, false);
      const std::vector<DocId> term_docs = SearchTermDocuments(idx, ti);
      temp.clear();
      std::set_union(excluded.begin(), excluded.end(), term_docs.begin(), term_docs.end(), std::back_inserter(temp));
      excluded.swap(temp);
    }
    if (!excluded.empty()) {
      std::vector<DocId> filtered;
      std::set_difference(results.begin(), results.end(), excluded.begin(), excluded.end(),
                          std::back_inserter(filtered));
      results = std::move(filtered);
    }
  }
  // ApplyVerifyTextFilter :1248-1266, then the hybrid-fragment exact-text filter :858-866
  if (!results.empty() && ShouldApplyVerifyText(p.verify_text, terms)) {
    results = PostFilterByText(idx, results, terms);
  }
  bool hybrid_exact = false;
  for (auto t : terms) {
    if (HasUncoveredHybridFragment(t, p.ngram_size, p.kanji_ngram_size, p.cross_boundary != 0)) {
      hybrid_exact = true;
      break;
    }
  }
  if (hybrid_exact) {
    results = PostFilterByText(idx, results, terms);
  }
  out.results = std::move(results);
  return out;
}

// ApplyNotFilter, search_pipeline.cpp:871-932, over NOT terms whose synonym expansion (if any) the caller has
// already flattened: what is excluded is the union over all of them.
std::vector<DocId> ApplyNotTerms(const orc_index& idx, const orc_query_params_t& p, std::vector<DocId> results,
                                 const std::vector<std::string_view>& not_terms) {
  if (results.empty() || not_terms.empty()) {
    return results;
  }
  std::vector<DocId> excluded;
  std::vector<DocId> temp;
  for (auto nt : not_terms) {
    const TermInfo ti = MakeTermInfo(idx, nt, p, false);
    const std::vector<DocId> term_docs = SearchTermDocuments(idx, ti);
    temp.clear();
    std::set_union(excluded.begin(), excluded.end(), term_docs.begin(), term_docs.end(), std::back_inserter(temp));
    excluded.swap(temp);
  }
  if (excluded.empty()) {
    return results;
  }
  std::vector<DocId> filtered;
  std::set_difference(results.begin(), results.end(), excluded.begin(), excluded.end(), std::back_inserter(filtered));
  return filtered;
}

// IntersectSorted, search_pipeline.cpp:420-436.
void IntersectSorted(std::vector<DocId>& acc, std::vector<DocId>&& fresh, bool& is_first) {
  if (is_first) {
    acc = std::move(fresh);
    is_first = false;
  } else if (!acc.empty() && !fresh.empty()) {
    std::vector<DocId> inter;
    std::set_intersection(acc.begin(), acc.end(), fresh.begin(), fresh.end(), std::back_inserter(inter));
    acc = std::move(inter);
  } else {
    acc.clear();
  }
}

// ---- utils/edit_distance.cpp ----
// ComputeDistanceImpl (:62-113): threshold Levenshtein over two rows folded into one, shorter sequence as the row.
uint32_t ComputeDistance(const uint32_t* a, uint32_t a_len, const uint32_t* b, uint32_t b_len, uint32_t max_distance) {
  if (a_len > b_len) {
    std::swap(a, b);
    std::swap(a_len, b_len);
  }
  if (b_len - a_len > max_distance) {
    return max_distance + 1;
  }
  if (a_len == 0) {
    return b_len;
  }
  std::vector<uint32_t> dp(a_len + 1);
  for (uint32_t j = 0; j <= a_len; ++j) {
    dp[j] = j;
  }
  for (uint32_t i = 0; i < b_len; ++i) {
    uint32_t prev = dp[0];
    dp[0] = i + 1;
    uint32_t row_min = dp[0];
    for (uint32_t j = 0; j < a_len; ++j) {
      const uint32_t cost = a[j] == b[i] ? 0 : 1;
      const uint32_t v = std::min({dp[j + 1] + 1, dp[j] + 1, prev + cost});
      prev = dp[j + 1];
      dp[j + 1] = v;
      row_min = std::min(row_min, v);
    }
    if (row_min > max_distance) {
      return max_distance + 1;
    }
  }
  return dp[a_len] <= max_distance ? dp[a_len] : max_distance + 1;
}

// ContainsFuzzyCodepointWindow (:115-135).
bool ContainsFuzzyWindow(const uint32_t* word, uint32_t word_len, const uint32_t* term, uint32_t term_len,
                         uint32_t max_distance) {
  const uint64_t min_len = term_len > max_distance ? term_len - max_distance : 1;
  const uint64_t max_len = std::min<uint64_t>(word_len, static_cast<uint64_t>(term_len) + max_distance);
  if (min_len > max_len) {
    return false;
  }
  for (uint32_t start = 0; start < word_len; ++start) {
    const uint64_t last = std::min<uint64_t>(max_len, word_len - start);
    for (uint64_t len = min_len; len <= last; ++len) {
      if (ComputeDistance(word + start, static_cast<uint32_t>(len), term, term_len, max_distance) <= max_distance) {
        return true;
      }
    }
  }
  return false;
}

bool IsAsciiOnly(std::string_view s) {
  return std::all_of(s.begin(), s.end(), [](char c) { return static_cast<unsigned char>(c) < 0x80; });
}

// ContainsFuzzyMatch (:168-255), after NormalizeUnicodeWhitespace (:25-49: U+3000 and U+00A0 become ' ').
bool ContainsFuzzyMatch(std::string_view text, std::string_view term, uint32_t max_distance) {
  if (term.empty()) {
    return true;
  }
  if (text.empty()) {
    return false;
  }
  std::string norm;
  norm.reserve(text.size());
  for (size_t i = 0; i < text.size();) {
    const auto byte = static_cast<unsigned char>(text[i]);
    if (byte == 0xE3 && i + 2 < text.size() && static_cast<unsigned char>(text[i + 1]) == 0x80 &&
        static_cast<unsigned char>(text[i + 2]) == 0x80) {
      norm += ' ';
      i += 3;
    } else if (byte == 0xC2 && i + 1 < text.size() && static_cast<unsigned char>(text[i + 1]) == 0xA0) {
      norm += ' ';
      i += 2;
    } else {
      norm += text[i];
      ++i;
    }
  }
  const std::vector<uint32_t> term_cps =
      IsAsciiOnly(term) ? std::vector<uint32_t>(term.begin(), term.end()) : Utf8ToCodepoints(term);
  const auto term_len = static_cast<uint32_t>(term_cps.size());
  const std::string_view view = norm;
  size_t pos = 0;
  while (pos < view.size()) {
    const size_t word_start = view.find_first_not_of(" \t\r\n", pos);
    if (word_start == std::string_view::npos) {
      break;
    }
    size_t word_end = view.find_first_of(" \t\r\n", word_start);
    if (word_end == std::string_view::npos) {
      word_end = view.size();
    }
    const std::string_view word = view.substr(word_start, word_end - word_start);
    if (IsAsciiOnly(word)) {
      // whole word against the term (both branches of :201-243 for an ASCII word reduce to this)
      std::vector<uint32_t> word_cps;
      for (char c : word) {
        word_cps.push_back(static_cast<unsigned char>(c));
      }
      const auto word_len = static_cast<uint32_t>(word_cps.size());
      const uint32_t diff = word_len > term_len ? word_len - term_len : term_len - word_len;
      if (diff <= max_distance &&
          ComputeDistance(word_cps.data(), word_len, term_cps.data(), term_len, max_distance) <= max_distance) {
        return true;
      }
    } else {
      const std::vector<uint32_t> word_cps = Utf8ToCodepoints(word);
      if (ContainsFuzzyWindow(word_cps.data(), static_cast<uint32_t>(word_cps.size()), term_cps.data(), term_len,
                              max_distance)) {
        return true;
      }
    }
    pos = word_end;
  }
  return false;
}

// ExecuteWithFuzzy, search_pipeline.cpp:1659-1740, with PostFilterByFuzzyText (:1742-1752) when verify_text applies
// to the terms. *empty_term = the reference's empty_term_detected, under which ExecuteFullPipeline clears the
// results (:1933-1937).
std::vector<DocId> RunFuzzy(const orc_index& idx, const orc_query_params_t& p,
                            const std::vector<std::string_view>& terms, uint32_t max_distance,
                            const std::vector<std::string_view>& not_terms, bool* empty_term) {
  *empty_term = false;
  std::vector<DocId> results;
  if (terms.empty()) {
    *empty_term = true;
    return results;
  }
  bool first_term = true;
  for (auto term : terms) {
    const TermInfo ti = MakeTermInfo(idx, term, p, false);
    if (ti.ngrams.empty()) {
      *empty_term = true;  // :1674-1680
      return {};
    }
    int eff = p.ngram_size > 0 ? p.ngram_size : 2;  // :1682-1695
    if (p.kanji_ngram_size > 0) {
      size_t short_count = 0;
      for (const auto& g : ti.ngrams) {
        short_count += g.size() <= 3 ? 1 : 0;
      }
      if (short_count > ti.ngrams.size() / 2) {
        eff = p.kanji_ngram_size;
      }
    }
    const size_t n = ti.ngrams.size();
    const size_t drop = static_cast<size_t>(max_distance) * static_cast<size_t>(eff);
    const size_t threshold = n > drop ? n - drop : 1;  // :1697-1700
    IntersectSorted(results, SearchByThreshold(idx, Views(ti.ngrams), threshold), first_term);
  }
  results = ApplyNotTerms(idx, p, std::move(results), not_terms);
  if (!results.empty() && ShouldApplyVerifyText(p.verify_text, terms)) {  // :1714-1726
    std::vector<DocId> kept;
    for (DocId d : results) {
      std::string_view text;
      bool ok = true;
      if (idx.GetText(d, &text)) {  // a candidate without stored text is kept (:386-402)
        for (auto t : terms) {
          if (text.find(t) == std::string_view::npos && !ContainsFuzzyMatch(text, t, max_distance)) {
            ok = false;
            break;
          }
        }
      }
      if (ok) {
        kept.push_back(d);
      }
    }
    results = std::move(kept);
  }
  for (auto t : terms) {  // RequiresExactTextForHybridFragments, :1728-1737
    if (HasUncoveredHybridFragment(t, p.ngram_size, p.kanji_ngram_size, p.cross_boundary != 0)) {
      results = PostFilterByText(idx, results, terms);
      break;
    }
  }
  return results;
}

// ExecuteWithSynonyms + PostFilterByTextWithSynonyms, search_pipeline.cpp:1580-1657. A group is the list of its
// variants (ExpandNormalizedTermWithSynonyms :1360-1388: normalized_terms == the variants' terms).
std::vector<DocId> RunSynonyms(const orc_index& idx, const orc_query_params_t& p,
                               const std::vector<std::vector<std::string_view>>& groups,
                               const std::vector<std::string_view>& not_terms, bool* empty_term) {
  *empty_term = false;
  std::vector<DocId> results;
  bool first_group = true;
  for (const auto& group : groups) {
    std::vector<DocId> group_results;
    for (auto variant : group) {
      const TermInfo ti = MakeTermInfo(idx, variant, p, false);
      if (ti.normalized.empty() || (!ti.ngrams.empty() && ti.estimated_size == 0)) {
        continue;  // :1593-1595
      }
      auto var_results = SearchTermDocuments(idx, ti);
      if (group_results.empty()) {
        group_results = std::move(var_results);
      } else {
        std::vector<DocId> merged;
        std::set_union(group_results.begin(), group_results.end(), var_results.begin(), var_results.end(),
                       std::back_inserter(merged));
        group_results = std::move(merged);
      }
    }
    IntersectSorted(results, std::move(group_results), first_group);
  }
  if (first_group) {
    *empty_term = true;  // :1618-1621
    return {};
  }
  results = ApplyNotTerms(idx, p, std::move(results), not_terms);
  bool verify = p.verify_text == 1;  // ShouldApplyVerifyTextSynonyms :154-172
  if (p.verify_text == 2) {
    verify = true;
    for (const auto& group : groups) {
      verify = verify && ShouldApplyVerifyText(2, group);
    }
  }
  if (verify && !results.empty()) {
    std::vector<DocId> kept;
    for (DocId d : results) {
      std::string_view text;
      bool ok = true;
      if (idx.GetText(d, &text)) {  // a candidate without stored text is kept (:386-402)
        for (const auto& group : groups) {
          bool any = false;
          for (auto v : group) {
            any = any || text.find(v) != std::string_view::npos;
          }
          ok = ok && any;
        }
      }
      if (ok) {
        kept.push_back(d);
      }
    }
    results = std::move(kept);
  }
  return results;
}

std::vector<std::string_view> TermList(const uint8_t* bytes, const uint64_t* offsets, uint64_t begin, uint64_t end) {
  std::vector<std::string_view> out;
  out.reserve(end - begin);
  for (uint64_t i = begin; i < end; ++i) {
    out.emplace_back(reinterpret_cast<const char*>(bytes) + offsets[i], offsets[i + 1] - offsets[i]);
  }
  return out;
}

uint64_t CopyOut(const std::vector<DocId>& v, uint32_t* out, uint64_t cap) {
  if (out != nullptr) {
    std::memcpy(out, v.data(), std::min<uint64_t>(v.size(), cap) * sizeof(uint32_t));
  }
  return v.size();
}

// Packed key used only by the bulk builder: (cp+1) in 21-bit fields, first
// code point in the most significant field, so that integer order == bytewise
// order of the UTF-8 n-gram strings.
inline uint64_t PackKey(const uint32_t* cps, int n, int width) {
  uint64_t key = 0;
  for (int j = 0; j < width; ++j) {
    key = (key << 21) | (j < n ? static_cast<uint64_t>(cps[j]) + 1 : 0);
  }
  return key;
}

std::string UnpackKey(uint64_t key, int width) {
  uint32_t cps[3];
  int n = 0;
  for (int j = width - 1; j >= 0; --j) {
    const uint64_t f = (key >> (21 * j)) & 0x1FFFFF;
    if (f != 0) {
      cps[n++] = static_cast<uint32_t>(f - 1);
    }
  }
  // CodepointsToUtf8 would drop surrogates; keys only ever hold decoded (valid) code points.
  return CodepointsToUtf8(cps, cps + n);
}

struct Pair {
  uint64_t key;
  uint32_t doc;
};

// n-gram string -> packed key; false if it is not the UTF-8 encoding of 1..width code points
// (then it cannot be a dictionary term).
bool NgramToKey(std::string_view term, int width, uint64_t* key) {
  uint32_t cps[3];
  int n = 0;
  const auto* data = reinterpret_cast<const unsigned char*>(term.data());
  size_t i = 0;
  while (i < term.size()) {
    uint32_t cp = 0;
    const int len = TryParseUtf8Char(data + i, term.size() - i, &cp);
    if (len <= 0 || n >= width) {
      return false;
    }
    cps[n++] = cp;
    i += static_cast<size_t>(len);
  }
  if (n == 0) {
    return false;
  }
  *key = PackKey(cps, n, width);
  return true;
}

}  // namespace

Span orc_index::Find(std::string_view term) const {
  Span out;
  if (csr) {
    uint64_t key = 0;
    if (!NgramToKey(term, csr_width, &key)) {
      return out;
    }
    const auto it = std::lower_bound(csr_keys.begin(), csr_keys.end(), key);
    if (it == csr_keys.end() || *it != key) {
      return out;
    }
    const size_t t = static_cast<size_t>(it - csr_keys.begin());
    out.p = csr_postings.data() + csr_offsets[t];
    out.n = static_cast<size_t>(csr_offsets[t + 1] - csr_offsets[t]);
    out.found = true;
    return out;
  }
  auto it = postings.find(term);
  if (it != postings.end()) {
    out.p = it->second.data();
    out.n = it->second.size();
    out.found = true;
  }
  return out;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------

extern "C" {

uint64_t orc_utf8_to_codepoints(const uint8_t* text, uint64_t len, uint32_t* out, uint64_t cap) {
  const auto cps = Utf8ToCodepoints(std::string_view(reinterpret_cast<const char*>(text), len));
  if (out != nullptr) {
    std::memcpy(out, cps.data(), std::min<uint64_t>(cps.size(), cap) * sizeof(uint32_t));
  }
  return cps.size();
}

uint64_t orc_codepoints_to_utf8(const uint32_t* cps, uint64_t n, uint8_t* out) {
  const std::string s = CodepointsToUtf8(cps, cps + n);
  std::memcpy(out, s.data(), s.size());
  return s.size();
}

uint64_t orc_count_code_points(const uint8_t* text, uint64_t len) {
  return CountCodePoints(std::string_view(reinterpret_cast<const char*>(text), len));
}

int orc_is_cjk_ideograph(uint32_t cp) { return IsCJKIdeograph(cp) ? 1 : 0; }

int64_t orc_ngrams(int mode, const uint8_t* text, uint64_t len, int a, int k, int cross, uint8_t* out_bytes,
                   uint64_t cap_bytes, uint64_t* out_offsets, uint64_t cap_ngrams) {
  const std::string_view sv(reinterpret_cast<const char*>(text), len);
  std::vector<std::string> ngrams;
  if (mode == 0) {
    ngrams = GenerateNgrams(sv, a);
  } else if (mode == 1) {
    ngrams = GenerateHybridNgrams(sv, a, k, cross != 0);
  } else {
    ngrams = GenerateQueryNgrams(sv, a, k, cross != 0);
  }
  uint64_t total = 0;
  for (const auto& g : ngrams) {
    total += g.size();
  }
  if (ngrams.size() > cap_ngrams || total > cap_bytes) {
    return -static_cast<int64_t>(std::max<uint64_t>(ngrams.size(), total) + 1);
  }
  uint64_t pos = 0;
  for (size_t i = 0; i < ngrams.size(); ++i) {
    out_offsets[i] = pos;
    std::memcpy(out_bytes + pos, ngrams[i].data(), ngrams[i].size());
    pos += ngrams[i].size();
  }
  out_offsets[ngrams.size()] = pos;
  return static_cast<int64_t>(ngrams.size());
}

orc_index_t* orc_index_create(int ngram_size, int kanji_ngram_size, int cross_boundary) {
  auto* idx = new orc_index();
  idx->ngram_size = ngram_size;
  idx->kanji_ngram_size = kanji_ngram_size > 0 ? kanji_ngram_size : ngram_size;  // index.cpp:32
  idx->cross_boundary = cross_boundary != 0;
  return idx;
}

void orc_index_destroy(orc_index_t* idx) { delete idx; }

int orc_index_add_document(orc_index_t* idx, uint32_t doc_id, const uint8_t* text, uint64_t len) {
  const std::string_view sv(reinterpret_cast<const char*>(text), len);
  StoreTextAndStats(*idx, doc_id, sv);
  const auto ngrams = IndexNgrams(*idx, sv);
  if (ngrams.empty()) {
    return 0;  // index.cpp:49-57
  }
  for (const auto& g : ngrams) {
    PostingAdd(idx->postings[g], doc_id);
  }
  return 1;
}

void orc_index_add_batch(orc_index_t* idx, const uint32_t* doc_ids, const uint8_t* text, const uint64_t* offsets,
                         uint64_t n_docs, uint64_t batch) {
  if (batch == 0) {
    batch = 1000;  // initial_loader.cpp:41
  }
  for (uint64_t begin = 0; begin < n_docs; begin += batch) {
    const uint64_t end = std::min(n_docs, begin + batch);
    std::unordered_map<std::string, std::vector<DocId>> term_to_docs;  // index.cpp:83
    for (uint64_t d = begin; d < end; ++d) {
      const std::string_view sv(reinterpret_cast<const char*>(text) + offsets[d], offsets[d + 1] - offsets[d]);
      StoreTextAndStats(*idx, doc_ids[d], sv);
      const auto ngrams = IndexNgrams(*idx, sv);
      for (const auto& g : ngrams) {  // empty => skipped (:93-101)
        term_to_docs[g].push_back(doc_ids[d]);
      }
    }
    for (auto& [term, docs] : term_to_docs) {
      std::sort(docs.begin(), docs.end());  // :110-112
      PostingAddBatch(idx->postings[term], docs);
    }
  }
}

int orc_index_build_bulk(orc_index_t* idx, const uint32_t* doc_ids, const uint8_t* text, const uint64_t* offsets,
                         uint64_t n_docs, int n_threads) {
  const int width = std::max(idx->ngram_size, idx->kanji_ngram_size);
  if (width > 3 || idx->ngram_size <= 0) {
    return -1;
  }
  if (n_threads <= 0) {
    n_threads = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
  }
  n_threads = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(n_threads), std::max<uint64_t>(1, n_docs)));
  idx->arena = text;
  idx->arena_offsets = offsets;
  idx->arena_doc_ids = doc_ids;
  idx->arena_docs = n_docs;
  idx->arena_sequential = true;
  for (uint64_t i = 1; i < n_docs; ++i) {
    if (doc_ids[i] != doc_ids[0] + i) {
      idx->arena_sequential = false;
      break;
    }
  }
  if (!idx->arena_sequential && !std::is_sorted(doc_ids, doc_ids + n_docs)) {
    return -2;
  }

  // Phase 1: tokenise per thread (doc ranges), per-doc sort+unique like index.cpp:88-91.
  std::vector<std::vector<Pair>> parts(static_cast<size_t>(n_threads));
  std::vector<uint64_t> len_sum(static_cast<size_t>(n_threads), 0);
  std::vector<uint64_t> cnt_sum(static_cast<size_t>(n_threads), 0);
  auto tokenize = [&](int t) {
    const uint64_t begin = n_docs * static_cast<uint64_t>(t) / static_cast<uint64_t>(n_threads);
    const uint64_t end = n_docs * static_cast<uint64_t>(t + 1) / static_cast<uint64_t>(n_threads);
    auto& out = parts[static_cast<size_t>(t)];
    std::vector<uint64_t> keys;
    for (uint64_t d = begin; d < end; ++d) {
      const std::string_view sv(reinterpret_cast<const char*>(text) + offsets[d], offsets[d + 1] - offsets[d]);
      if (sv.empty()) {
        continue;
      }
      const auto cps = Utf8ToCodepoints(sv);
      len_sum[static_cast<size_t>(t)] += cps.size();
      cnt_sum[static_cast<size_t>(t)] += 1;
      keys.clear();
      for (size_t i = 0; i < cps.size(); ++i) {
        const bool cjk = IsCJKIdeograph(cps[i]);
        const int size = cjk ? idx->kanji_ngram_size : idx->ngram_size;
        if (i + static_cast<size_t>(size) > cps.size()) {
          continue;
        }
        if (!idx->cross_boundary) {
          bool crossed = false;
          for (int j = 1; j < size; ++j) {
            if (IsCJKIdeograph(cps[i + static_cast<size_t>(j)]) != cjk) {
              crossed = true;
              break;
            }
          }
          if (crossed) {
            continue;
          }
        }
        keys.push_back(PackKey(cps.data() + i, size, width));
      }
      std::sort(keys.begin(), keys.end());
      keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
      for (uint64_t k : keys) {
        out.push_back({k, doc_ids[d]});
      }
    }
  };
  {
    std::vector<std::thread> threads;
    for (int t = 0; t < n_threads; ++t) {
      threads.emplace_back(tokenize, t);
    }
    for (auto& th : threads) {
      th.join();
    }
  }
  for (int t = 0; t < n_threads; ++t) {
    idx->total_doc_length += len_sum[static_cast<size_t>(t)];
    idx->doc_count += cnt_sum[static_cast<size_t>(t)];
  }

  // Phase 2: sample sort. Splitters are drawn from the data so that buckets are balanced even
  // though n-gram frequencies are Zipf-skewed; every bucket is then sorted independently.
  uint64_t total = 0;
  for (const auto& p : parts) {
    total += p.size();
  }
  const size_t n_buckets = total < (1u << 16) ? 1 : 4096;
  std::vector<uint64_t> splitters;  // bucket b holds keys in (splitters[b-1], splitters[b]]
  if (n_buckets > 1) {
    std::vector<uint64_t> sample;
    uint64_t state = 0x9E3779B97F4A7C15ULL;
    const size_t per_part = (n_buckets * 16) / parts.size() + 1;
    for (const auto& p : parts) {
      for (size_t k = 0; k < per_part && !p.empty(); ++k) {
        state = state * 6364136223846793005ULL + 1442695040888963407ULL;
        sample.push_back(p[(state >> 16) % p.size()].key);
      }
    }
    std::sort(sample.begin(), sample.end());
    for (size_t b = 1; b < n_buckets; ++b) {
      splitters.push_back(sample[b * sample.size() / n_buckets]);
    }
    splitters.erase(std::unique(splitters.begin(), splitters.end()), splitters.end());
  }
  const size_t nb = splitters.size() + 1;
  auto bucket_of = [&](uint64_t key) {
    return static_cast<size_t>(std::lower_bound(splitters.begin(), splitters.end(), key) - splitters.begin());
  };
  // per-part bucket counts -> write cursors (part-major inside a bucket keeps documents ascending)
  std::vector<std::vector<uint64_t>> counts(parts.size(), std::vector<uint64_t>(nb, 0));
  {
    std::vector<std::thread> threads;
    for (size_t t = 0; t < parts.size(); ++t) {
      threads.emplace_back([&, t]() {
        for (const auto& pr : parts[t]) {
          counts[t][bucket_of(pr.key)]++;
        }
      });
    }
    for (auto& th : threads) {
      th.join();
    }
  }
  std::vector<uint64_t> bucket_begin(nb + 1, 0);
  for (size_t b = 0; b < nb; ++b) {
    uint64_t c = 0;
    for (size_t t = 0; t < parts.size(); ++t) {
      const uint64_t v = counts[t][b];
      counts[t][b] = bucket_begin[b] + c;  // becomes this part's write cursor in bucket b
      c += v;
    }
    bucket_begin[b + 1] = bucket_begin[b] + c;
  }
  std::vector<Pair> all(total);
  {
    std::vector<std::thread> threads;
    for (size_t t = 0; t < parts.size(); ++t) {
      threads.emplace_back([&, t]() {
        auto& cursor = counts[t];
        for (const auto& pr : parts[t]) {
          all[cursor[bucket_of(pr.key)]++] = pr;
        }
        std::vector<Pair>().swap(parts[t]);
      });
    }
    for (auto& th : threads) {
      th.join();
    }
  }
  std::atomic<size_t> next_bucket{0};
  auto sort_buckets = [&]() {
    for (;;) {
      const size_t bkt = next_bucket.fetch_add(1);
      if (bkt >= nb) {
        break;
      }
      std::stable_sort(all.begin() + static_cast<std::ptrdiff_t>(bucket_begin[bkt]),
                       all.begin() + static_cast<std::ptrdiff_t>(bucket_begin[bkt + 1]),
                       [](const Pair& l, const Pair& r) { return l.key < r.key; });
    }
  };
  {
    std::vector<std::thread> threads;
    for (int t = 0; t < n_threads; ++t) {
      threads.emplace_back(sort_buckets);
    }
    for (auto& th : threads) {
      th.join();
    }
  }

  // Phase 3: CSR — one posting list per distinct key (docs ascending and unique inside a key).
  idx->csr = true;
  idx->csr_width = width;
  idx->csr_postings.resize(total);
  idx->csr_keys.clear();
  idx->csr_offsets.clear();
  for (uint64_t i = 0; i < total; ++i) {
    if (i == 0 || all[i].key != all[i - 1].key) {
      idx->csr_keys.push_back(all[i].key);
      idx->csr_offsets.push_back(i);
    }
    idx->csr_postings[i] = all[i].doc;
  }
  idx->csr_offsets.push_back(total);
  return 0;
}

void orc_index_remove_document(orc_index_t* idx, uint32_t doc_id, const uint8_t* text, uint64_t len) {
  const std::string_view sv(reinterpret_cast<const char*>(text), len);
  for (const auto& g : IndexNgrams(*idx, sv)) {  // index.cpp:175-189
    auto it = idx->postings.find(g);
    if (it == idx->postings.end()) {
      continue;
    }
    auto pos = std::lower_bound(it->second.begin(), it->second.end(), doc_id);
    if (pos != it->second.end() && *pos == doc_id) {
      it->second.erase(pos);
    }
    if (it->second.empty()) {
      idx->postings.erase(it);  // RemoveFromPostingList erases empty lists, index.cpp:693-717
    }
  }
  auto t = idx->texts.find(doc_id);
  if (t != idx->texts.end()) {
    // BM25Stats::RemoveDocument saturating subtract, server_types.h:204-218
    const uint64_t dl = CountCodePoints(t->second);
    idx->total_doc_length = idx->total_doc_length > dl ? idx->total_doc_length - dl : 0;
    idx->doc_count = idx->doc_count > 1 ? idx->doc_count - 1 : 0;
    idx->texts.erase(t);
  }
  idx->known_docs.erase(doc_id);
}

void orc_index_update_document(orc_index_t* idx, uint32_t doc_id, const uint8_t* old_text, uint64_t old_len,
                               const uint8_t* new_text, uint64_t new_len) {
  const std::string_view old_sv(reinterpret_cast<const char*>(old_text), old_len);
  const std::string_view new_sv(reinterpret_cast<const char*>(new_text), new_len);
  const auto old_ngrams = IndexNgrams(*idx, old_sv);  // index.cpp:123-130
  const auto new_ngrams = IndexNgrams(*idx, new_sv);
  std::vector<std::string> to_remove;
  std::vector<std::string> to_add;
  std::set_difference(old_ngrams.begin(), old_ngrams.end(), new_ngrams.begin(), new_ngrams.end(),
                      std::back_inserter(to_remove));
  std::set_difference(new_ngrams.begin(), new_ngrams.end(), old_ngrams.begin(), old_ngrams.end(),
                      std::back_inserter(to_add));
  for (const auto& g : to_remove) {
    auto it = idx->postings.find(g);
    if (it == idx->postings.end()) {
      continue;
    }
    auto pos = std::lower_bound(it->second.begin(), it->second.end(), doc_id);
    if (pos != it->second.end() && *pos == doc_id) {
      it->second.erase(pos);
    }
    if (it->second.empty()) {
      idx->postings.erase(it);
    }
  }
  for (const auto& g : to_add) {
    PostingAdd(idx->postings[g], doc_id);
  }
  // text + stats follow binlog_event_processor.cpp:236-243 (remove old length, add new)
  auto t = idx->texts.find(doc_id);
  if (t != idx->texts.end()) {
    const uint64_t dl = CountCodePoints(t->second);
    idx->total_doc_length = idx->total_doc_length > dl ? idx->total_doc_length - dl : 0;
    idx->doc_count = idx->doc_count > 1 ? idx->doc_count - 1 : 0;
    idx->texts.erase(t);
  }
  StoreTextAndStats(*idx, doc_id, new_sv);
}

uint64_t orc_index_term_count(const orc_index_t* idx) { return idx->csr ? idx->csr_keys.size() : idx->postings.size(); }

uint64_t orc_index_posting_size(const orc_index_t* idx, const uint8_t* term, uint64_t len) {
  return idx->Find(std::string_view(reinterpret_cast<const char*>(term), len)).size();
}

uint64_t orc_index_total_postings(const orc_index_t* idx) {
  if (idx->csr) {
    return idx->csr_postings.size();
  }
  uint64_t total = 0;
  for (const auto& [term, list] : idx->postings) {
    total += list.size();
  }
  return total;
}

uint64_t orc_index_get_postings(const orc_index_t* idx, const uint8_t* term, uint64_t len, uint32_t* out,
                                uint64_t cap) {
  const Span list = idx->Find(std::string_view(reinterpret_cast<const char*>(term), len));
  if (out != nullptr) {
    std::memcpy(out, list.p, std::min<uint64_t>(list.n, cap) * sizeof(uint32_t));
  }
  return list.n;
}

uint64_t orc_index_export(const orc_index_t* idx, uint8_t* term_bytes_out, uint64_t* term_offsets_out,
                          uint64_t* posting_offsets_out, uint32_t* postings_out, uint64_t* total_term_bytes) {
  if (idx->csr) {
    uint64_t bytes = 0;
    for (size_t t = 0; t < idx->csr_keys.size(); ++t) {
      const std::string term = UnpackKey(idx->csr_keys[t], idx->csr_width);
      if (term_bytes_out != nullptr) {
        term_offsets_out[t] = bytes;
        std::memcpy(term_bytes_out + bytes, term.data(), term.size());
      }
      bytes += term.size();
    }
    if (total_term_bytes != nullptr) {
      *total_term_bytes = bytes;
    }
    if (term_bytes_out != nullptr) {
      term_offsets_out[idx->csr_keys.size()] = bytes;
      std::memcpy(posting_offsets_out, idx->csr_offsets.data(), idx->csr_offsets.size() * sizeof(uint64_t));
      std::memcpy(postings_out, idx->csr_postings.data(), idx->csr_postings.size() * sizeof(uint32_t));
    }
    return idx->csr_keys.size();
  }
  std::vector<const std::pair<const std::string, Posting>*> order;
  order.reserve(idx->postings.size());
  uint64_t bytes = 0;
  for (const auto& kv : idx->postings) {
    order.push_back(&kv);
    bytes += kv.first.size();
  }
  if (total_term_bytes != nullptr) {
    *total_term_bytes = bytes;
  }
  if (term_bytes_out == nullptr) {
    return order.size();
  }
  std::sort(order.begin(), order.end(), [](const auto* l, const auto* r) { return l->first < r->first; });
  uint64_t tb = 0;
  uint64_t pb = 0;
  for (size_t i = 0; i < order.size(); ++i) {
    term_offsets_out[i] = tb;
    posting_offsets_out[i] = pb;
    std::memcpy(term_bytes_out + tb, order[i]->first.data(), order[i]->first.size());
    tb += order[i]->first.size();
    std::memcpy(postings_out + pb, order[i]->second.data(), order[i]->second.size() * sizeof(uint32_t));
    pb += order[i]->second.size();
  }
  term_offsets_out[order.size()] = tb;
  posting_offsets_out[order.size()] = pb;
  return order.size();
}

void orc_index_bm25_stats(const orc_index_t* idx, uint64_t* total_doc_length, uint64_t* doc_count) {
  *total_doc_length = idx->total_doc_length;
  *doc_count = idx->doc_count;
}

uint64_t orc_search_and(const orc_index_t* idx, const uint8_t* term_bytes, const uint64_t* term_offsets,
                        uint64_t n_terms, uint64_t limit, int reverse, uint32_t* out, uint64_t cap) {
  return CopyOut(SearchAnd(*idx, TermList(term_bytes, term_offsets, 0, n_terms), limit, reverse != 0), out, cap);
}

uint64_t orc_search_or(const orc_index_t* idx, const uint8_t* term_bytes, const uint64_t* term_offsets,
                       uint64_t n_terms, uint32_t* out, uint64_t cap) {
  return CopyOut(SearchOr(*idx, TermList(term_bytes, term_offsets, 0, n_terms)), out, cap);
}

uint64_t orc_search_not(const orc_index_t* idx, const uint32_t* all_docs, uint64_t n_all, const uint8_t* term_bytes,
                        const uint64_t* term_offsets, uint64_t n_terms, uint32_t* out, uint64_t cap) {
  const std::vector<DocId> all(all_docs, all_docs + n_all);
  return CopyOut(SearchNot(*idx, all, TermList(term_bytes, term_offsets, 0, n_terms)), out, cap);
}

uint64_t orc_filter_by_ngrams(const orc_index_t* idx, const uint32_t* candidates, uint64_t n_candidates,
                              const uint8_t* term_bytes, const uint64_t* term_offsets, uint64_t n_terms,
                              uint32_t* out, uint64_t cap) {
  const std::vector<DocId> cands(candidates, candidates + n_candidates);
  return CopyOut(FilterByNgrams(*idx, cands, TermList(term_bytes, term_offsets, 0, n_terms)), out, cap);
}

uint64_t orc_search_by_threshold(const orc_index_t* idx, const uint8_t* term_bytes, const uint64_t* term_offsets,
                                 uint64_t n_terms, uint64_t threshold, uint32_t* out, uint64_t cap) {
  return CopyOut(SearchByThreshold(*idx, TermList(term_bytes, term_offsets, 0, n_terms), threshold), out, cap);
}

double orc_compute_idf(uint64_t total_docs, uint64_t doc_freq) { return ComputeIDF(total_docs, doc_freq); }

uint32_t orc_count_term_occurrences(const uint8_t* text, uint64_t text_len, const uint8_t* term, uint64_t term_len) {
  return CountTermOccurrences(std::string_view(reinterpret_cast<const char*>(text), text_len),
                              std::string_view(reinterpret_cast<const char*>(term), term_len));
}

void orc_score_documents(const orc_index_t* idx, const uint32_t* candidates, uint64_t n_candidates,
                         const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* term_doc_freqs,
                         uint64_t n_terms, uint64_t total_docs, double avg_doc_length, double k1, double b,
                         double* out_scores) {
  ScoreDocuments(*idx, candidates, n_candidates, TermList(term_bytes, term_offsets, 0, n_terms), term_doc_freqs,
                 total_docs, avg_doc_length, k1, b, out_scores);
}

uint64_t orc_sort_by_score(const uint32_t* results, const double* scores, uint64_t n, int descending, uint32_t limit,
                           uint32_t offset, uint32_t* out) {
  const auto sorted = SortByScore(results, scores, n, descending != 0, limit, offset);
  std::memcpy(out, sorted.data(), sorted.size() * sizeof(uint32_t));
  return sorted.size();
}

int orc_query_batch(const orc_index_t* idx, const orc_query_params_t* params, uint64_t n_queries,
                    const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                    const uint8_t* not_bytes, const uint64_t* not_offsets, const uint64_t* q_not_begin,
                    uint64_t stride, uint32_t* out_ids, double* out_scores, uint32_t* out_count,
                    uint64_t* out_total, uint64_t* out_df, uint32_t* out_sets, uint64_t sets_cap,
                    uint64_t* out_sets_offsets, int n_threads) {
  const orc_query_params_t p = *params;
  uint64_t total_docs = p.total_docs_override != 0 ? p.total_docs_override : idx->doc_count;
  uint64_t total_len = p.total_docs_override != 0 ? p.total_len_override : idx->total_doc_length;
  const double avgdl =
      total_docs > 0 ? static_cast<double>(total_len) / static_cast<double>(total_docs) : 0.0;  // server_types.h:182-187
  if (n_threads <= 0) {
    n_threads = 1;
  }
  std::vector<std::vector<DocId>> sets;
  if (out_sets_offsets != nullptr) {
    sets.resize(n_queries);
  }
  std::atomic<uint64_t> next{0};
  auto worker = [&]() {
    for (;;) {
      const uint64_t q = next.fetch_add(1);
      if (q >= n_queries) {
        break;
      }
      const auto terms = TermList(term_bytes, term_offsets, q_term_begin[q], q_term_begin[q + 1]);
      std::vector<std::string_view> not_terms;
      if (q_not_begin != nullptr) {
        not_terms = TermList(not_bytes, not_offsets, q_not_begin[q], q_not_begin[q + 1]);
      }
      QueryOutput qo = RunQuery(*idx, p, terms, q_term_begin[q], not_terms);
      out_total[q] = qo.results.size();
      if (out_df != nullptr) {
        for (const auto& ti : qo.term_infos) {
          out_df[ti.source_index] = ti.df;
        }
      }
      uint32_t* ids = out_ids + q * stride;
      uint64_t written = 0;
      if (p.compute_score != 0) {
        // handlers/search_handler.cpp:436-470: terms and dfs in term_infos (size-sorted) order
        std::vector<std::string_view> norm_terms;
        std::vector<uint64_t> dfs;
        for (const auto& ti : qo.term_infos) {
          norm_terms.emplace_back(ti.normalized);
          dfs.push_back(ti.df);
        }
        std::vector<double> scores(qo.results.size());
        ScoreDocuments(*idx, qo.results.data(), qo.results.size(), norm_terms, dfs.data(), total_docs, avgdl, p.k1,
                       p.b, scores.data());
        const auto sorted =
            SortByScore(qo.results.data(), scores.data(), qo.results.size(), p.descending != 0, p.limit, p.offset);
        written = std::min<uint64_t>(sorted.size(), stride);
        for (uint64_t i = 0; i < written; ++i) {
          ids[i] = sorted[i];
          if (out_scores != nullptr) {
            // results are ascending => locate the score by binary search
            const auto pos = std::lower_bound(qo.results.begin(), qo.results.end(), sorted[i]) - qo.results.begin();
            out_scores[q * stride + i] = scores[static_cast<size_t>(pos)];
          }
        }
      } else {
        const size_t start = std::min<size_t>(p.offset, qo.results.size());
        const size_t end = p.limit == 0 ? qo.results.size() : std::min<size_t>(start + p.limit, qo.results.size());
        written = std::min<uint64_t>(end - start, stride);
        std::memcpy(ids, qo.results.data() + start, written * sizeof(uint32_t));
      }
      out_count[q] = static_cast<uint32_t>(written);
      if (out_sets_offsets != nullptr) {
        sets[q] = std::move(qo.results);
      }
    }
  };
  std::vector<std::thread> threads;
  for (int t = 1; t < n_threads; ++t) {
    threads.emplace_back(worker);
  }
  worker();
  for (auto& th : threads) {
    th.join();
  }
  if (out_sets_offsets != nullptr) {
    uint64_t pos = 0;
    for (uint64_t q = 0; q < n_queries; ++q) {
      out_sets_offsets[q] = pos;
      if (pos + sets[q].size() <= sets_cap && out_sets != nullptr) {
        std::memcpy(out_sets + pos, sets[q].data(), sets[q].size() * sizeof(uint32_t));
      }
      pos += sets[q].size();
    }
    out_sets_offsets[n_queries] = pos;
  }
  return 0;
}

uint64_t orc_eval_boolean(const orc_index_t* idx, const int32_t* ops, const int32_t* args, uint64_t n_ops,
                          const uint8_t* term_bytes, const uint64_t* term_offsets, uint32_t* out, uint64_t cap) {
  // QueryNode::Evaluate, query_ast.cpp:67-161, over a postfix encoding.
  orc_query_params_t p{};
  p.ngram_size = idx->ngram_size;
  p.kanji_ngram_size = idx->kanji_ngram_size;  // Evaluate reads Index::GetKanjiNgramSize() (effective value)
  p.cross_boundary = idx->cross_boundary ? 1 : 0;
  std::vector<DocId> all_docs;  // DocumentStore::GetAllDocIds (document_store_retrieval.cpp:242-258), lazily
  bool have_all = false;
  auto get_all = [&]() -> const std::vector<DocId>& {
    if (!have_all) {
      if (idx->arena != nullptr) {
        all_docs.assign(idx->arena_doc_ids, idx->arena_doc_ids + idx->arena_docs);
      } else {
        all_docs.assign(idx->known_docs.begin(), idx->known_docs.end());
      }
      std::sort(all_docs.begin(), all_docs.end());
      have_all = true;
    }
    return all_docs;
  };
  std::vector<std::vector<DocId>> stack;
  for (uint64_t i = 0; i < n_ops; ++i) {
    const int op = ops[i];
    if (op == 0) {
      const auto t = static_cast<uint64_t>(args[i]);
      const std::string_view term(reinterpret_cast<const char*>(term_bytes) + term_offsets[t],
                                  term_offsets[t + 1] - term_offsets[t]);
      const TermInfo ti = MakeTermInfo(*idx, term, p, false);
      stack.push_back(SearchTermDocuments(*idx, ti));  // :76-93
    } else if (op == 1 || op == 2) {
      const auto n = static_cast<size_t>(args[i]);
      if (n > stack.size()) {
        return 0;
      }
      std::vector<DocId> result;
      const size_t base = stack.size() - n;
      if (op == 1) {  // AND :96-123 (early break cannot change the value)
        for (size_t c = 0; c < n; ++c) {
          if (c == 0) {
            result = stack[base];
          } else {
            std::vector<DocId> inter;
            std::set_intersection(result.begin(), result.end(), stack[base + c].begin(), stack[base + c].end(),
                                  std::back_inserter(inter));
            result = std::move(inter);
          }
        }
      } else {  // OR :125-136 concat + sort + unique
        for (size_t c = 0; c < n; ++c) {
          result.insert(result.end(), stack[base + c].begin(), stack[base + c].end());
        }
        std::sort(result.begin(), result.end());
        result.erase(std::unique(result.begin(), result.end()), result.end());
      }
      stack.resize(base);
      stack.push_back(std::move(result));
    } else if (op == 3) {  // NOT :138-157
      if (stack.empty()) {
        return 0;
      }
      const auto& all = get_all();
      std::vector<DocId> result;
      std::set_difference(all.begin(), all.end(), stack.back().begin(), stack.back().end(),
                          std::back_inserter(result));
      stack.back() = std::move(result);
    }
  }
  if (stack.empty()) {
    return 0;
  }
  return CopyOut(stack.back(), out, cap);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- column filters
extern "C" int orc_contains_fuzzy_match(const uint8_t* text, uint64_t text_len, const uint8_t* term, uint64_t term_len,
                                        uint32_t max_distance) {
  return ContainsFuzzyMatch(std::string_view(reinterpret_cast<const char*>(text), text_len),
                            std::string_view(reinterpret_cast<const char*>(term), term_len), max_distance)
             ? 1
             : 0;
}

extern "C" uint64_t orc_search_fuzzy(const orc_index_t* idx, const orc_query_params_t* params,
                                     const uint8_t* term_bytes, const uint64_t* term_offsets, uint64_t n_terms,
                                     uint32_t max_distance, const uint8_t* not_bytes, const uint64_t* not_offsets,
                                     uint64_t n_not, uint32_t* out, uint64_t cap, int32_t* empty_term_detected) {
  const auto terms = TermList(term_bytes, term_offsets, 0, n_terms);
  const auto not_terms = n_not > 0 ? TermList(not_bytes, not_offsets, 0, n_not) : std::vector<std::string_view>{};
  bool empty_term = false;
  const auto r = RunFuzzy(*idx, *params, terms, max_distance, not_terms, &empty_term);
  if (empty_term_detected != nullptr) {
    *empty_term_detected = empty_term ? 1 : 0;
  }
  return CopyOut(r, out, cap);
}

extern "C" uint64_t orc_search_synonyms(const orc_index_t* idx, const orc_query_params_t* params,
                                        const uint8_t* variant_bytes, const uint64_t* variant_offsets,
                                        const uint64_t* group_begin, uint64_t n_groups, const uint8_t* not_bytes,
                                        const uint64_t* not_offsets, uint64_t n_not, uint32_t* out, uint64_t cap,
                                        int32_t* empty_term_detected) {
  std::vector<std::vector<std::string_view>> groups;
  for (uint64_t g = 0; g < n_groups; ++g) {
    groups.push_back(TermList(variant_bytes, variant_offsets, group_begin[g], group_begin[g + 1]));
  }
  const auto not_terms = n_not > 0 ? TermList(not_bytes, not_offsets, 0, n_not) : std::vector<std::string_view>{};
  bool empty_term = false;
  const auto r = RunSynonyms(*idx, *params, groups, not_terms, &empty_term);
  if (empty_term_detected != nullptr) {
    *empty_term_detected = empty_term ? 1 : 0;
  }
  return CopyOut(r, out, cap);
}

namespace {

struct ParsedLiteral {  // ParseFilterValue, search_pipeline.cpp:943-993
  std::string text;
  bool bool_val = false;
  double double_val = 0.0;
  int64_t int64_val = 0;
  uint64_t uint64_val = 0;
  bool double_valid = false;
  bool int64_valid = false;
  bool uint64_valid = false;
};

ParsedLiteral ParseLiteral(std::string_view value) {
  ParsedLiteral p;
  p.text.assign(value);
  p.bool_val = (value == "1" || value == "true");
  const char* b = value.data();
  const char* e = value.data() + value.size();
  {
    double r = 0.0;
    auto [ptr, ec] = std::from_chars(b, e, r);
    if (ec == std::errc() && ptr == e) {
      p.double_val = r;
      p.double_valid = true;
    }
  }
  {
    int64_t r = 0;
    auto [ptr, ec] = std::from_chars(b, e, r);
    if (ec == std::errc() && ptr == e) {
      p.int64_val = r;
      p.int64_valid = true;
    }
  }
  {
    uint64_t r = 0;
    auto [ptr, ec] = std::from_chars(b, e, r);
    if (ec == std::errc() && ptr == e) {
      p.uint64_val = r;
      p.uint64_valid = true;
    }
  }
  return p;
}

template <typename T>
bool CompareOp(const T& lhs, const T& rhs, int op) {  // utils/comparison_utils.h:28-43
  switch (op) {
    case 0: return lhs == rhs;
    case 1: return lhs != rhs;
    case 2: return lhs > rhs;
    case 3: return lhs >= rhs;
    case 4: return lhs < rhs;
    case 5: return lhs <= rhs;
  }
  return false;
}

bool CompareDoubleOp(double lhs, double rhs, int op) {  // comparison_utils.h:56-70, epsilon constants.h:104
  constexpr double kEps = 1e-9;
  if (op == 0) return std::abs(lhs - rhs) < kEps;
  if (op == 1) return std::abs(lhs - rhs) >= kEps;
  return CompareOp(lhs, rhs, op);
}

bool IsSignedType(int t) { return t == 2 || t == 4 || t == 6 || t == 8; }
bool IsUnsignedType(int t) { return t == 3 || t == 5 || t == 7 || t == 9; }

// Does the stored value equal some type interpretation of the literal? (BuildTypeUnionBitmap :1021-1094: the
// bitmap of a value is keyed by SerializeFilterValue, filter_index.cpp:177-255 = type tag + little-endian bytes)
bool BitmapEqMatch(int type, uint64_t bits, std::string_view str, const ParsedLiteral& lit) {
  if (type == 1) {
    if (lit.text == "1" || lit.text == "true") return bits != 0;
    if (lit.text == "0" || lit.text == "false") return bits == 0;
    return false;
  }
  if (IsSignedType(type) || type == 10) {  // int8..int64 (tried when the value fits the width), TIME seconds
    return lit.int64_valid && static_cast<int64_t>(bits) == lit.int64_val;
  }
  if (IsUnsignedType(type)) {
    return lit.uint64_valid && bits == lit.uint64_val;
  }
  if (type == 11) {
    return str == lit.text;
  }
  if (type == 12) {
    uint64_t lb = 0;
    std::memcpy(&lb, &lit.double_val, sizeof(lb));
    return lit.double_valid && bits == lb;  // serialised bytes compare: bit pattern equality
  }
  return false;
}

// ApplyFilters' per-document visitor, :1134-1175 (value is not NULL)
bool TypedMatch(int type, uint64_t bits, std::string_view str, const ParsedLiteral& lit, int op) {
  if (type == 11) {
    return CompareOp(std::string(str), lit.text, op);
  }
  if (type == 1) {
    return CompareOp(bits != 0, lit.bool_val, op);
  }
  if (type == 12) {
    if (!lit.double_valid) return false;
    double v = 0.0;
    std::memcpy(&v, &bits, sizeof(v));
    return CompareDoubleOp(v, lit.double_val, op);
  }
  if (IsUnsignedType(type)) {
    if (!lit.uint64_valid) return false;
    return CompareOp(bits, lit.uint64_val, op);
  }
  if (!lit.int64_valid) return false;  // signed integers and TIME
  return CompareOp(static_cast<int64_t>(bits), lit.int64_val, op);
}

}  // namespace

extern "C" uint64_t orc_apply_filters(uint64_t n_docs, uint32_t first_doc_id, uint32_t n_cols, const int32_t* col_type,
                                      const uint64_t* col_values, const uint8_t* col_null, const uint8_t* str_bytes,
                                      const uint64_t* str_offsets, uint32_t n_filters, const uint32_t* filter_col,
                                      const uint8_t* filter_op, const uint8_t* lit_bytes, const uint64_t* lit_offsets,
                                      const uint32_t* results, uint64_t n_results, uint32_t* out) {
  std::vector<ParsedLiteral> lits;
  bool all_bitmap = true;  // AllFiltersHaveBitmapSupport :995-1003
  for (uint32_t f = 0; f < n_filters; ++f) {
    lits.push_back(ParseLiteral(std::string_view(reinterpret_cast<const char*>(lit_bytes) + lit_offsets[f],
                                                 lit_offsets[f + 1] - lit_offsets[f])));
    all_bitmap = all_bitmap && (filter_op[f] == 0 || filter_op[f] == 1);
  }
  uint64_t n_out = 0;
  for (uint64_t i = 0; i < n_results; ++i) {
    const uint32_t id = results[i];
    bool keep = id >= first_doc_id && static_cast<uint64_t>(id - first_doc_id) < n_docs;
    const uint64_t row = keep ? id - first_doc_id : 0;
    for (uint32_t f = 0; f < n_filters && keep; ++f) {
      const uint32_t c = filter_col[f];
      if (c >= n_cols) {
        // unknown column: no bitmap / no stored value => EQ matches nothing, NE everything; typed: NULL rule
        keep = filter_op[f] == 1;
        continue;
      }
      const int type = col_type[c];
      const uint64_t bits = col_values[static_cast<uint64_t>(c) * n_docs + row];
      const bool is_null = col_null[static_cast<uint64_t>(c) * n_docs + row] != 0;
      std::string_view str;
      if (type == 11 && !is_null) {
        str = std::string_view(reinterpret_cast<const char*>(str_bytes) + str_offsets[bits],
                               str_offsets[bits + 1] - str_offsets[bits]);
      }
      if (all_bitmap) {
        const bool eq = !is_null && BitmapEqMatch(type, bits, str, lits[f]);  // only non-NULL values are indexed
        keep = filter_op[f] == 0 ? eq : !eq;                                   // :1216-1229 and / andnot
      } else if (is_null) {
        keep = filter_op[f] == 1;  // :1126-1132
      } else {
        keep = TypedMatch(type, bits, str, lits[f], filter_op[f]);
      }
    }
    if (keep) {
      out[n_out++] = id;
    }
  }
  return n_out;
}
