// mygram_adapter.h — C++17 adapter with the reference's class signatures over the C ABI (include/mgx.h).
//
// A MygramDB maintainer swaps `#include "index/index.h"` for this header (or links it behind the same names)
// at the call sites of the hot path; everything above stays host C++. Signatures mirror, with file:line of the
// reference declaration they replace:
//   Index                    src/index/index.h:49-413      (ctor :58-60, AddDocumentBatch :75-100, SearchAnd :127,
//                                                           FilterByNgrams :138, SearchOr :147, SearchNot :156,
//                                                           PostingSize/Count :183-188, TermCount :193,
//                                                           SaveToStream :265)
//   search_pipeline::ExecuteWithFuzzy / ExecuteWithSynonyms  src/server/search_pipeline.h:271-311
//   BM25Scorer               src/index/bm25_scorer.h:43-83 (ScoreDocuments :79-82, BM25Params :23-26)
//   ResultSorter::SortByScore src/query/result_sorter.h:75
// Results are returned by value exactly as the reference does; errors follow the reference's conventions:
// searches have no error channel (a failing device call throws std::runtime_error instead of returning wrong
// data), ScoreDocuments reports kInvalidArgument / kInternalError through the result object.
#pragma once

#include <algorithm>
#include <iterator>
#include <cstdint>
#include <stdexcept>
#include <memory>
#include <ostream>
#include <string>
#include <string_view>
#include <vector>

#include "../../include/mgx.h"

namespace mygramdb_b200 {

using DocId = uint32_t;  // src/types/doc_id.h:31

enum class ErrorCode { kOk = 0, kInvalidArgument, kInternalError };  // subset of utils/error.h:35-130
enum class SortOrder : uint8_t { ASC, DESC };                        // query/query_parser.h

namespace detail {
struct Flat {
  std::vector<uint8_t> bytes;
  std::vector<uint64_t> offsets{0};
  void add(std::string_view s) {
    bytes.insert(bytes.end(), s.begin(), s.end());
    offsets.push_back(bytes.size());
  }
  const uint8_t* data() const {
    static const uint8_t kEmpty[1] = {0};
    return bytes.empty() ? kEmpty : bytes.data();
  }
};
inline void check(int rc) {
  if (rc != MGX_OK) {
    throw std::runtime_error(std::string("mgx: ") + mgx_last_error());
  }
}
template <typename Call>
std::vector<DocId> grow_call(Call&& call, uint64_t cap = 4096) {
  for (;;) {
    std::vector<DocId> out(cap);
    uint64_t n = 0;
    const int rc = call(out.data(), cap, &n);
    if (rc == MGX_ERR_CAPACITY) {
      cap = n > cap ? n : cap * 2;
      continue;
    }
    check(rc);
    out.resize(n);
    return out;
  }
}
}  // namespace detail

class Index {
 public:
  struct DocumentItem {  // index.h:75-78
    DocId doc_id;
    std::string text;
  };

  explicit Index(int ngram_size = 2, int kanji_ngram_size = 0, double /*roaring_threshold*/ = 0.18,
                 bool cross_boundary_ngrams = true, bool /*normalize_nfkc*/ = true,
                 const std::string& /*normalize_width*/ = "keep", bool /*normalize_lower*/ = true, int device = 0)
      : ngram_size_(ngram_size), kanji_ngram_size_(kanji_ngram_size > 0 ? kanji_ngram_size : ngram_size),
        cross_boundary_ngrams_(cross_boundary_ngrams) {
    mgx_index_config_t cfg{};
    cfg.ngram_size = ngram_size;
    cfg.kanji_ngram_size = kanji_ngram_size;
    cfg.cross_boundary_ngrams = cross_boundary_ngrams ? 1 : 0;
    cfg.device = device;
    detail::check(mgx_index_create(&cfg, &handle_));
  }
  ~Index() { mgx_index_destroy(handle_); }
  Index(const Index&) = delete;
  Index& operator=(const Index&) = delete;

  // Index::AddDocumentBatch (index.h:75-100, index.cpp:76-119): ADDITIVE, ids in any order, exactly as
  // InitialLoader::FlushBatch calls it for every 1000 documents (initial_loader.cpp:450-512). The batch is journaled and
  // folded in (one device merge + build) before the next read.
  void AddDocumentBatch(const std::vector<DocumentItem>& documents) {
    detail::Flat flat;
    std::vector<DocId> ids;
    ids.reserve(documents.size());
    for (const auto& d : documents) {
      ids.push_back(d.doc_id);
      flat.add(d.text);
    }
    detail::check(mgx_index_add_document_batch(handle_, ids.data(), flat.data(), flat.offsets.data(), documents.size(),
                                               nullptr));
  }
  // Not in the reference: the whole snapshot at once (ids strictly ascending), replacing any previous content -- the
  // direct form for a loader that has collected the snapshot anyway (no journal, no merge).
  void BuildFromSnapshot(const std::vector<DocumentItem>& documents) {
    detail::Flat flat;
    std::vector<DocId> ids;
    ids.reserve(documents.size());
    for (const auto& d : documents) {
      ids.push_back(d.doc_id);
      flat.add(d.text);
    }
    detail::check(mgx_index_build(handle_, ids.data(), flat.data(), flat.offsets.data(), documents.size()));
  }

  [[nodiscard]] std::vector<DocId> SearchAnd(const std::vector<std::string>& terms, size_t limit = 0,
                                             bool reverse = false) const {
    detail::Flat f = Pack(terms);
    return detail::grow_call([&](DocId* out, uint64_t cap, uint64_t* n) {
      return mgx_search_and(handle_, f.data(), f.offsets.data(), terms.size(), limit, reverse ? 1 : 0, out, cap, n);
    });
  }
  [[nodiscard]] std::vector<DocId> FilterByNgrams(const std::vector<DocId>& candidates,
                                                  const std::vector<std::string>& terms) const {
    detail::Flat f = Pack(terms);
    static const DocId kNone[1] = {0};
    return detail::grow_call(
        [&](DocId* out, uint64_t cap, uint64_t* n) {
          return mgx_filter_by_ngrams(handle_, candidates.empty() ? kNone : candidates.data(), candidates.size(),
                                      f.data(), f.offsets.data(), terms.size(), out, cap, n);
        },
        candidates.size() + 1);
  }
  [[nodiscard]] std::vector<DocId> SearchOr(const std::vector<std::string>& terms) const {
    detail::Flat f = Pack(terms);
    return detail::grow_call([&](DocId* out, uint64_t cap, uint64_t* n) {
      return mgx_search_or(handle_, f.data(), f.offsets.data(), terms.size(), out, cap, n);
    });
  }
  [[nodiscard]] std::vector<DocId> SearchNot(const std::vector<DocId>& all_docs,
                                             const std::vector<std::string>& terms) const {
    detail::Flat f = Pack(terms);
    static const DocId kNone[1] = {0};
    return detail::grow_call(
        [&](DocId* out, uint64_t cap, uint64_t* n) {
          return mgx_search_not(handle_, all_docs.empty() ? kNone : all_docs.data(), all_docs.size(), f.data(),
                                f.offsets.data(), terms.size(), out, cap, n);
        },
        all_docs.size() + 1);
  }
  // index.cpp:488-578
  [[nodiscard]] std::vector<DocId> SearchByThreshold(const std::vector<std::string>& terms, size_t threshold) const {
    detail::Flat f = Pack(terms);
    return detail::grow_call([&](DocId* out, uint64_t cap, uint64_t* n) {
      return mgx_search_by_threshold(handle_, f.data(), f.offsets.data(), terms.size(), threshold, out, cap, n);
    });
  }
  // Incremental mutations (index.cpp:39-74, 121-173, 175-197): journaled, visible to the next read.
  bool AddDocument(DocId doc_id, std::string_view text) {
    int32_t indexed = 0;
    detail::check(mgx_index_add_document(handle_, doc_id, reinterpret_cast<const uint8_t*>(text.data()), text.size(),
                                         &indexed));
    return indexed != 0;
  }
  void UpdateDocument(DocId doc_id, std::string_view old_text, std::string_view new_text) {
    detail::check(mgx_index_update_document(handle_, doc_id, reinterpret_cast<const uint8_t*>(old_text.data()),
                                            old_text.size(), reinterpret_cast<const uint8_t*>(new_text.data()),
                                            new_text.size()));
  }
  void RemoveDocument(DocId doc_id, std::string_view text) {
    detail::check(mgx_index_remove_document(handle_, doc_id, reinterpret_cast<const uint8_t*>(text.data()),
                                            text.size()));
  }
  // Index::GetStatistics / Optimize / Clear (index.cpp:604-641, index_optimization.cpp:36-120)
  struct IndexStatistics {
    size_t total_terms = 0;
    size_t total_postings = 0;
    size_t delta_encoded_lists = 0;
    size_t roaring_bitmap_lists = 0;
    size_t memory_usage_bytes = 0;
  };
  [[nodiscard]] IndexStatistics GetStatistics() const {
    mgx_index_statistics_t s{};
    detail::check(mgx_index_get_statistics(handle_, &s));
    return {s.total_terms, s.total_postings, s.delta_encoded_lists, s.roaring_bitmap_lists, s.memory_usage_bytes};
  }
  void Optimize(uint64_t total_docs) { detail::check(mgx_index_optimize(handle_, total_docs)); }
  void Clear() { detail::check(mgx_index_clear(handle_)); }
  // Not in the reference: with overlapped commits searches neither fold journaled mutations in nor wait for a commit
  // in progress (they answer from the current generation); the mutating thread publishes its calls with Commit()
  // (what a server with an asynchronous binlog applier wants; mgx.h).
  void SetOverlappedCommits(bool on) { detail::check(mgx_index_set_commit_mode(handle_, on ? 1 : 0)); }
  void Commit() { detail::check(mgx_index_commit(handle_)); }
  [[nodiscard]] uint64_t PostingSize(std::string_view term) const {
    uint64_t n = 0;
    detail::check(mgx_index_posting_size(handle_, reinterpret_cast<const uint8_t*>(term.data()), term.size(), &n));
    return n;
  }
  [[nodiscard]] uint64_t Count(std::string_view term) const { return PostingSize(term); }
  [[nodiscard]] uint64_t EstimatePostingSize(std::string_view term) const { return PostingSize(term); }
  [[nodiscard]] size_t TermCount() const {
    mgx_index_stats_t s{};
    detail::check(mgx_index_get_stats(handle_, &s));
    return s.n_terms;
  }
  // Index::SaveToStream (index.h:265, index_serialization.cpp:111-224): the MGIX v4 stream of the device index, what
  // DUMP SAVE / SYNC write for the table (storage/dump_format_v2.h:13-41). Returns false where the reference
  // returns an Error (kIndexSerializationFailed). The normalisation triple is the table's (index.h:58-60).
  bool SaveToStream(std::ostream& output_stream, bool normalize_nfkc = true, const std::string& normalize_width = "keep",
                    bool normalize_lower = true) const {
    uint64_t len = 0;
    int rc = mgx_index_save_mgix(handle_, normalize_nfkc ? 1 : 0, normalize_width.c_str(), normalize_lower ? 1 : 0,
                                 nullptr, 0, &len);
    if (rc != MGX_ERR_CAPACITY && rc != MGX_OK) {
      return false;
    }
    std::vector<uint8_t> bytes(len);
    rc = mgx_index_save_mgix(handle_, normalize_nfkc ? 1 : 0, normalize_width.c_str(), normalize_lower ? 1 : 0,
                             bytes.data(), bytes.size(), &len);
    if (rc != MGX_OK) {
      return false;
    }
    output_stream.write(reinterpret_cast<const char*>(bytes.data()), static_cast<std::streamsize>(len));
    return output_stream.good();
  }
  // Index::LoadFromStream (index.h:279, index_serialization.cpp:279-613): the stream's posting lists replace the index
  // content. false where the reference returns an Error (kStorage* / kIndexDeserializationFailed; mgx_last_error()
  // names it). The device shard then holds no document text, like the reference's Index with an empty DocumentStore.
  bool LoadFromStream(std::istream& input_stream) {
    const std::vector<char> bytes((std::istreambuf_iterator<char>(input_stream)), std::istreambuf_iterator<char>());
    return mgx_index_load_mgix(handle_, reinterpret_cast<const uint8_t*>(bytes.data()), bytes.size()) == MGX_OK;
  }
  [[nodiscard]] int GetNgramSize() const { return ngram_size_; }
  [[nodiscard]] int GetKanjiNgramSize() const { return kanji_ngram_size_; }
  [[nodiscard]] bool GetCrossBoundaryNgrams() const { return cross_boundary_ngrams_; }
  // Index::NormalizeText stays on the host (ICU); the path's contract is pre-normalised text (index.h:83-84).
  [[nodiscard]] mgx_index_t* handle() const { return handle_; }

 private:
  static detail::Flat Pack(const std::vector<std::string>& terms) {
    detail::Flat f;
    for (const auto& t : terms) {
      f.add(t);
    }
    return f;
  }
  int ngram_size_;
  int kanji_ngram_size_;
  bool cross_boundary_ngrams_;
  mgx_index_t* handle_ = nullptr;
};

struct BM25Params {  // bm25_scorer.h:23-26
  double k1 = 1.2;
  double b = 0.75;
};
struct ScoredDoc {  // bm25_scorer.h:31-34
  DocId doc_id;
  double score;
};
struct ScoreResult {  // stands in for Expected<std::vector<ScoredDoc>, Error>
  ErrorCode code = ErrorCode::kOk;
  std::string message;
  std::vector<ScoredDoc> value;
  explicit operator bool() const { return code == ErrorCode::kOk; }
};

class BM25Scorer {
 public:
  // The document store argument of the reference (bm25_scorer.h:79-82) is the index's device-resident text mirror.
  static ScoreResult ScoreDocuments(const std::vector<DocId>& candidates, const std::vector<std::string>& search_terms,
                                    const std::vector<uint64_t>& term_doc_freqs, const Index& index,
                                    uint64_t total_docs, double avg_doc_length, const BM25Params& params = {}) {
    ScoreResult r;
    if (search_terms.size() != term_doc_freqs.size()) {  // bm25_scorer.cpp:51-55
      r.code = ErrorCode::kInvalidArgument;
      r.message = "BM25 search_terms and term_doc_freqs must have identical lengths";
      return r;
    }
    detail::Flat f;
    for (const auto& t : search_terms) {
      f.add(t);
    }
    std::vector<double> scores(candidates.size());
    const int rc = mgx_score_documents(index.handle(), candidates.data(), candidates.size(), f.data(),
                                       f.offsets.data(), term_doc_freqs.data(), search_terms.size(), total_docs,
                                       avg_doc_length, params.k1, params.b, scores.data());
    if (rc != MGX_OK) {
      r.code = ErrorCode::kInternalError;  // bm25_scorer.cpp:93-96
      r.message = mgx_last_error();
      return r;
    }
    r.value.reserve(candidates.size());
    for (size_t i = 0; i < candidates.size(); ++i) {
      r.value.push_back({candidates[i], scores[i]});
    }
    return r;
  }
};

// QueryNode (query/query_ast.h:59-95): the boolean AST, evaluated on the device as one postfix program.
enum class NodeType { TERM, AND, OR, NOT };
struct QueryNode {
  NodeType type = NodeType::TERM;
  std::string term;
  std::vector<std::unique_ptr<QueryNode>> children;

  // QueryNode::Evaluate(index, doc_store, all_docs) — query_ast.cpp:67-161. The document store and the all-docs
  // list of the reference are the index's device-resident mirror.
  [[nodiscard]] std::vector<DocId> Evaluate(const Index& index) const {
    std::vector<int32_t> ops;
    std::vector<int32_t> args;
    detail::Flat terms;
    int32_t n_terms = 0;
    Emit(&ops, &args, &terms, &n_terms);
    return detail::grow_call([&](DocId* out, uint64_t cap, uint64_t* n) {
      return mgx_eval_boolean(index.handle(), ops.data(), args.data(), ops.size(), terms.data(), terms.offsets.data(),
                              static_cast<uint64_t>(n_terms), out, cap, n);
    });
  }

 private:
  void Emit(std::vector<int32_t>* ops, std::vector<int32_t>* args, detail::Flat* terms, int32_t* n_terms) const {
    switch (type) {
      case NodeType::TERM:
        terms->add(term);
        ops->push_back(0);
        args->push_back((*n_terms)++);
        return;
      case NodeType::AND:
      case NodeType::OR: {
        int32_t n = 0;
        for (const auto& c : children) {
          if (c != nullptr) {
            c->Emit(ops, args, terms, n_terms);
            ++n;
          }
        }
        ops->push_back(type == NodeType::AND ? 1 : 2);
        args->push_back(n);
        return;
      }
      case NodeType::NOT:
        if (children.empty() || children[0] == nullptr) {
          ops->push_back(2);  // NOT without a child evaluates to the empty set (query_ast.cpp:139-141): OR of nothing
          args->push_back(0);
          return;
        }
        children[0]->Emit(ops, args, terms, n_terms);  // only the first child is used (:150)
        ops->push_back(3);
        args->push_back(0);
        return;
    }
  }
};

class ResultSorter {
 public:
  static std::vector<DocId> SortByScore(const Index& index, const std::vector<DocId>& results,
                                        const std::vector<double>& scores, SortOrder order, uint32_t limit,
                                        uint32_t offset) {
    if (results.empty()) {
      return {};  // result_sorter.cpp:663-665
    }
    // limit 0 = everything after offset (result_sorter.cpp:689-710); the window never exceeds min(limit, size)
    std::vector<DocId> out(limit == 0 ? results.size() : std::min<size_t>(limit, results.size()));
    uint64_t n = 0;
    detail::check(mgx_sort_by_score(index.handle(), results.data(), scores.data(), results.size(),
                                    order == SortOrder::DESC ? 1 : 0, limit, offset, out.data(), &n));
    out.resize(n);
    return out;
  }
};

// search_pipeline::ExecuteWithFuzzy / ExecuteWithSynonyms (server/search_pipeline.cpp:1659-1740, 1580-1657) on the
// device. The query's NOT terms and column conditions ride along the way ApplyNotAndFilters applies them (:470-485).
namespace search_pipeline {

struct FilterCondition {  // query::FilterCondition, query/query_parser.h:93-110 (column already resolved to its id)
  uint32_t column = 0;
  uint8_t op = 0;  // 0 EQ, 1 NE, 2 GT, 3 GTE, 4 LT, 5 LTE
  std::string value;
};

struct ExpandedQuery {
  int ngram_size = 2;        // RAW table configuration (search_pipeline.cpp:578)
  int kanji_ngram_size = 0;
  bool cross_boundary = true;
  int verify_text = 0;       // memory.verify_text: 0 "off", 1 "all", 2 "ascii"
  std::vector<std::string> not_terms;
  std::vector<FilterCondition> filters;
};

namespace detail_sp {
struct Packed {
  detail::Flat nots, lits;
  std::vector<uint32_t> cols;
  std::vector<uint8_t> ops;
  mgx_expanded_query_t c{};
  explicit Packed(const ExpandedQuery& q) {
    for (const auto& t : q.not_terms) nots.add(t);
    for (const auto& f : q.filters) {
      cols.push_back(f.column);
      ops.push_back(f.op);
      lits.add(f.value);
    }
    c.ngram_size = q.ngram_size;
    c.kanji_ngram_size = q.kanji_ngram_size;
    c.cross_boundary = q.cross_boundary ? 1 : 0;
    c.verify_text = q.verify_text;
    c.not_bytes = nots.data();
    c.not_offsets = nots.offsets.data();
    c.n_not = q.not_terms.size();
    c.filter_col = cols.data();
    c.filter_op = ops.data();
    c.filter_bytes = lits.data();
    c.filter_offsets = lits.offsets.data();
    c.n_filters = q.filters.size();
  }
};
}  // namespace detail_sp

// `terms`: the query's normalised search terms (all_search_terms). Result: SearchPipelineResult::results.
inline std::vector<DocId> ExecuteWithFuzzy(const Index& index, const ExpandedQuery& query,
                                           const std::vector<std::string>& terms, uint32_t max_distance) {
  detail_sp::Packed p(query);
  detail::Flat f;
  for (const auto& t : terms) f.add(t);
  return detail::grow_call([&](DocId* out, uint64_t cap, uint64_t* n) {
    return mgx_search_fuzzy(index.handle(), &p.c, f.data(), f.offsets.data(), terms.size(), max_distance, out, cap, n);
  });
}

// `groups`: SynonymTermGroup::normalized_terms of every group (ExpandTermsWithSynonyms stays on the host); with a
// synonym dictionary in play `query.not_terms` holds every synonym of every NOT term (ApplyNotFilter :886-899).
inline std::vector<DocId> ExecuteWithSynonyms(const Index& index, const ExpandedQuery& query,
                                              const std::vector<std::vector<std::string>>& groups) {
  detail_sp::Packed p(query);
  detail::Flat f;
  std::vector<uint64_t> begin{0};
  for (const auto& g : groups) {
    for (const auto& v : g) f.add(v);
    begin.push_back(f.offsets.size() - 1);
  }
  return detail::grow_call([&](DocId* out, uint64_t cap, uint64_t* n) {
    return mgx_search_synonyms(index.handle(), &p.c, f.data(), f.offsets.data(), begin.data(), groups.size(), out, cap,
                               n);
  });
}

// Batch forms (no counterpart in the reference, whose pipeline runs one request per worker thread): what a batcher
// thread calls with the requests it has collected, when they share the table configuration, NOT terms and filters.
// result[i] = ExecuteWithFuzzy(index, query, queries[i], max_distance), computed in one device batch.
inline std::vector<std::vector<DocId>> ExecuteWithFuzzyBatch(const Index& index, const ExpandedQuery& query,
                                                             const std::vector<std::vector<std::string>>& queries,
                                                             uint32_t max_distance) {
  detail_sp::Packed p(query);
  detail::Flat f;
  std::vector<uint64_t> qbegin{0};
  for (const auto& q : queries) {
    for (const auto& t : q) f.add(t);
    qbegin.push_back(f.offsets.size() - 1);
  }
  std::vector<uint64_t> off(queries.size() + 1, 0);
  std::vector<DocId> flat(1 << 16);
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int rc = mgx_search_fuzzy_batch(index.handle(), &p.c, queries.size(), f.data(), f.offsets.data(), qbegin.data(),
                                          max_distance, flat.data(), flat.size(), off.data());
    if (rc == MGX_ERR_CAPACITY && attempt == 0) {
      flat.resize(off.back());
      continue;
    }
    detail::check(rc);
    break;
  }
  std::vector<std::vector<DocId>> out(queries.size());
  for (size_t i = 0; i < queries.size(); ++i) {
    out[i].assign(flat.begin() + static_cast<std::ptrdiff_t>(off[i]), flat.begin() + static_cast<std::ptrdiff_t>(off[i + 1]));
  }
  return out;
}

inline std::vector<std::vector<DocId>> ExecuteWithSynonymsBatch(
    const Index& index, const ExpandedQuery& query, const std::vector<std::vector<std::vector<std::string>>>& queries) {
  detail_sp::Packed p(query);
  detail::Flat f;
  std::vector<uint64_t> gbegin{0};
  std::vector<uint64_t> qbegin{0};
  for (const auto& groups : queries) {
    for (const auto& g : groups) {
      for (const auto& v : g) f.add(v);
      gbegin.push_back(f.offsets.size() - 1);
    }
    qbegin.push_back(gbegin.size() - 1);
  }
  std::vector<uint64_t> off(queries.size() + 1, 0);
  std::vector<DocId> flat(1 << 16);
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int rc = mgx_search_synonyms_batch(index.handle(), &p.c, queries.size(), f.data(), f.offsets.data(), gbegin.data(),
                                             qbegin.data(), flat.data(), flat.size(), off.data());
    if (rc == MGX_ERR_CAPACITY && attempt == 0) {
      flat.resize(off.back());
      continue;
    }
    detail::check(rc);
    break;
  }
  std::vector<std::vector<DocId>> out(queries.size());
  for (size_t i = 0; i < queries.size(); ++i) {
    out[i].assign(flat.begin() + static_cast<std::ptrdiff_t>(off[i]), flat.begin() + static_cast<std::ptrdiff_t>(off[i + 1]));
  }
  return out;
}

}  // namespace search_pipeline

}  // namespace mygramdb_b200
