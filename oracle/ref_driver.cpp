// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY.
//
// Thin C-ABI driver (same entry points as oracle.h) over the REFERENCE'S OWN,
// UNMODIFIED sources, which oracle/Makefile compiles where they lie under
// /root/reference/src with the three shims in oracle/shim/ (abseil container,
// spdlog, CRoaring -- all pinned-but-unvendored network dependencies, see
// third_party/CMakeLists.txt:17-181) and without USE_ICU (inputs are
// pre-normalised; string_utils.cpp:370-379 degrades to ASCII tolower).
// Nothing from the reference is copied into this repository: this file only
// *calls* the reference's public classes the way its own callers do:
//   InitialLoader::FlushBatch      src/loader/initial_loader.cpp:450-512
//   SearchHandler::HandleSearch    src/server/handlers/search_handler.cpp:405-470
//   search_pipeline::ExecuteFullPipeline  src/server/search_pipeline.cpp:1757-2059
// It is used to (1) validate oracle.cpp, (2) generate tests/golden/ref_*.json
// (oracle/gen_golden.py) and (3) optionally serve as the "reference" CPU
// baseline (labelled "reference sources + Roaring shim").

#include <algorithm>
#include <atomic>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "config/config.h"
#include "index/bm25_scorer.h"
#include "index/index.h"
#include "query/query_ast.h"
#include "query/query_parser.h"
#include "query/result_sorter.h"
#include "server/search_pipeline.h"
#include "server/server_types.h"
#include "storage/document_store.h"
#include "utils/edit_distance.h"
#include "utils/string_utils.h"

#include "oracle.h"

namespace ref = mygramdb;

// DocumentStore assigns DocIds itself (sequential from 1, document_store.h:520).
// SetNextDocId is a protected hook of the reference class; exposing it through
// a subclass lets tests index documents under caller-chosen ids (the
// reference's own index tests use ids such as 100..500) without touching the
// reference's sources.
struct OpenDocumentStore : ref::storage::DocumentStore {
  using ref::storage::DocumentStore::SetNextDocId;
};

struct orc_index {
  std::unique_ptr<ref::index::Index> index;
  std::unique_ptr<OpenDocumentStore> store;
  ref::server::BM25Stats stats;
  ref::config::Config config;
  int ngram_size = 2;
  int kanji_ngram_size = 0;  // raw config value as given
  bool cross_boundary = true;
};

namespace {

std::vector<std::string> TermList(const uint8_t* bytes, const uint64_t* offsets, uint64_t begin, uint64_t end) {
  std::vector<std::string> out;
  out.reserve(end - begin);
  for (uint64_t i = begin; i < end; ++i) {
    out.emplace_back(reinterpret_cast<const char*>(bytes) + offsets[i], offsets[i + 1] - offsets[i]);
  }
  return out;
}

uint64_t CopyOut(const std::vector<uint32_t>& v, uint32_t* out, uint64_t cap) {
  if (out != nullptr) {
    std::memcpy(out, v.data(), std::min<uint64_t>(v.size(), cap) * sizeof(uint32_t));
  }
  return v.size();
}

int64_t PackNgrams(const std::vector<std::string>& ngrams, uint8_t* out_bytes, uint64_t cap_bytes,
                   uint64_t* out_offsets, uint64_t cap_ngrams) {
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

// One document through DocumentStore + Index + BM25Stats the way the binlog
// INSERT path does (binlog_event_processor.cpp:83-100).
int AddOne(orc_index& h, uint32_t doc_id, std::string_view text) {
  h.store->SetNextDocId(doc_id);
  auto added = h.store->AddDocument(std::to_string(doc_id), {}, text, "");
  if (!added || *added != doc_id) {
    return -1;
  }
  const bool indexed = h.index->AddDocument(doc_id, text);
  if (!text.empty()) {
    h.stats.AddDocument(static_cast<uint32_t>(ref::utils::CountCodePoints(text)));
  }
  return indexed ? 1 : 0;
}

}  // namespace

extern "C" {

uint64_t orc_utf8_to_codepoints(const uint8_t* text, uint64_t len, uint32_t* out, uint64_t cap) {
  const auto cps = ref::utils::Utf8ToCodepoints(std::string_view(reinterpret_cast<const char*>(text), len));
  if (out != nullptr) {
    std::memcpy(out, cps.data(), std::min<uint64_t>(cps.size(), cap) * sizeof(uint32_t));
  }
  return cps.size();
}

uint64_t orc_codepoints_to_utf8(const uint32_t* cps, uint64_t n, uint8_t* out) {
  const std::string s = ref::utils::CodepointsToUtf8(cps, cps + n);
  std::memcpy(out, s.data(), s.size());
  return s.size();
}

uint64_t orc_count_code_points(const uint8_t* text, uint64_t len) {
  return ref::utils::CountCodePoints(std::string_view(reinterpret_cast<const char*>(text), len));
}

int orc_is_cjk_ideograph(uint32_t cp) {
  // IsCJKIdeograph is file-local in the reference; observe it through the
  // hybrid generator: with (ascii=2, kanji=1) a lone code point yields an
  // n-gram iff it is classified as a CJK ideograph.
  const std::string s = ref::utils::CodepointsToUtf8(&cp, &cp + 1);
  return ref::utils::GenerateHybridNgrams(s, 2, 1, true).size() == 1 ? 1 : 0;
}

int64_t orc_ngrams(int mode, const uint8_t* text, uint64_t len, int a, int k, int cross, uint8_t* out_bytes,
                   uint64_t cap_bytes, uint64_t* out_offsets, uint64_t cap_ngrams) {
  const std::string_view sv(reinterpret_cast<const char*>(text), len);
  std::vector<std::string> ngrams;
  if (mode == 0) {
    ngrams = ref::utils::GenerateNgrams(sv, a);
  } else if (mode == 1) {
    ngrams = ref::utils::GenerateHybridNgrams(sv, a, k, cross != 0);
  } else {
    ngrams = ref::utils::GenerateQueryNgrams(sv, a, k, cross != 0);
  }
  return PackNgrams(ngrams, out_bytes, cap_bytes, out_offsets, cap_ngrams);
}

orc_index_t* orc_index_create(int ngram_size, int kanji_ngram_size, int cross_boundary) {
  auto* h = new orc_index();
  h->ngram_size = ngram_size;
  h->kanji_ngram_size = kanji_ngram_size;
  h->cross_boundary = cross_boundary != 0;
  // server_orchestrator.cpp:378-382 (normalisation flags are irrelevant without ICU)
  h->index = std::make_unique<ref::index::Index>(ngram_size, kanji_ngram_size, 0.18, cross_boundary != 0);
  h->store = std::make_unique<OpenDocumentStore>();
  h->config.bm25.enable = true;
  h->config.memory.verify_text = "off";  // config.h:329 default
  return h;
}

void orc_index_destroy(orc_index_t* idx) { delete idx; }

int orc_index_add_document(orc_index_t* idx, uint32_t doc_id, const uint8_t* text, uint64_t len) {
  return AddOne(*idx, doc_id, std::string_view(reinterpret_cast<const char*>(text), len));
}

void orc_index_add_batch(orc_index_t* idx, const uint32_t* doc_ids, const uint8_t* text, const uint64_t* offsets,
                         uint64_t n_docs, uint64_t batch) {
  if (batch == 0) {
    batch = 1000;  // initial_loader.cpp:41
  }
  for (uint64_t begin = 0; begin < n_docs; begin += batch) {
    const uint64_t end = std::min(n_docs, begin + batch);
    bool consecutive = true;
    for (uint64_t d = begin + 1; d < end; ++d) {
      if (doc_ids[d] != doc_ids[d - 1] + 1) {
        consecutive = false;
        break;
      }
    }
    if (!consecutive) {
      for (uint64_t d = begin; d < end; ++d) {
        (void)AddOne(*idx, doc_ids[d],
                     std::string_view(reinterpret_cast<const char*>(text) + offsets[d], offsets[d + 1] - offsets[d]));
      }
      continue;
    }
    // InitialLoader::FlushBatch, initial_loader.cpp:450-512
    std::vector<ref::storage::DocumentStore::DocumentItem> doc_batch;
    std::vector<ref::index::Index::DocumentItem> index_batch;
    doc_batch.reserve(end - begin);
    index_batch.reserve(end - begin);
    for (uint64_t d = begin; d < end; ++d) {
      std::string t(reinterpret_cast<const char*>(text) + offsets[d], offsets[d + 1] - offsets[d]);
      ref::storage::DocumentStore::DocumentItem item;
      item.primary_key = std::to_string(doc_ids[d]);
      item.normalized_text = t;
      doc_batch.push_back(std::move(item));
      index_batch.push_back({doc_ids[d], std::move(t)});
    }
    idx->store->SetNextDocId(doc_ids[begin]);
    auto ids = idx->store->AddDocumentBatch(doc_batch);
    if (!ids) {
      continue;
    }
    for (size_t i = 0; i < ids->size(); ++i) {
      index_batch[i].doc_id = (*ids)[i];
    }
    idx->index->AddDocumentBatch(index_batch);
    for (const auto& item : index_batch) {  // server_orchestrator.cpp:758-772
      if (!item.text.empty()) {
        idx->stats.AddDocument(static_cast<uint32_t>(ref::utils::CountCodePoints(item.text)));
      }
    }
  }
}

int orc_index_build_bulk(orc_index_t* idx, const uint32_t* doc_ids, const uint8_t* text, const uint64_t* offsets,
                         uint64_t n_docs, int /*n_threads*/) {
  // The reference has exactly one build path (single-threaded, 1000-document
  // batches, initial_loader.cpp:296-385) followed by Index::Optimize(total).
  orc_index_add_batch(idx, doc_ids, text, offsets, n_docs, 1000);
  idx->index->Optimize(n_docs);
  return 0;
}

void orc_index_remove_document(orc_index_t* idx, uint32_t doc_id, const uint8_t* text, uint64_t len) {
  const std::string_view sv(reinterpret_cast<const char*>(text), len);
  idx->index->RemoveDocument(doc_id, sv);  // binlog_event_processor.cpp:283
  if (!sv.empty()) {
    idx->stats.RemoveDocument(static_cast<uint32_t>(ref::utils::CountCodePoints(sv)));
  }
  idx->store->RemoveDocument(doc_id);
}

void orc_index_update_document(orc_index_t* idx, uint32_t doc_id, const uint8_t* old_text, uint64_t old_len,
                               const uint8_t* new_text, uint64_t new_len) {
  const std::string_view old_sv(reinterpret_cast<const char*>(old_text), old_len);
  const std::string_view new_sv(reinterpret_cast<const char*>(new_text), new_len);
  idx->index->UpdateDocument(doc_id, old_sv, new_sv);  // binlog_event_processor.cpp:234
  if (!old_sv.empty()) {
    idx->stats.RemoveDocument(static_cast<uint32_t>(ref::utils::CountCodePoints(old_sv)));
  }
  if (!new_sv.empty()) {
    idx->stats.AddDocument(static_cast<uint32_t>(ref::utils::CountCodePoints(new_sv)));
  }
  idx->store->SetNormalizedText(doc_id, new_sv);
}

uint64_t orc_index_term_count(const orc_index_t* idx) { return idx->index->TermCount(); }

uint64_t orc_index_posting_size(const orc_index_t* idx, const uint8_t* term, uint64_t len) {
  return idx->index->PostingSize(std::string_view(reinterpret_cast<const char*>(term), len));
}

uint64_t orc_index_total_postings(const orc_index_t* idx) { return idx->index->GetStatistics().total_postings; }

uint64_t orc_index_get_postings(const orc_index_t* idx, const uint8_t* term, uint64_t len, uint32_t* out,
                                uint64_t cap) {
  // PostingList::GetAll through the public API: a one-term SearchAnd returns it (index.cpp:338).
  const std::vector<std::string> terms{std::string(reinterpret_cast<const char*>(term), len)};
  return CopyOut(idx->index->SearchAnd(terms), out, cap);
}

uint64_t orc_index_export(const orc_index_t* /*idx*/, uint8_t* /*term_bytes_out*/, uint64_t* /*term_offsets_out*/,
                          uint64_t* /*posting_offsets_out*/, uint32_t* /*postings_out*/,
                          uint64_t* /*total_term_bytes*/) {
  return UINT64_MAX;  // the reference has no term enumeration API outside DUMP I/O (out of scope)
}

void orc_index_bm25_stats(const orc_index_t* idx, uint64_t* total_doc_length, uint64_t* doc_count) {
  *total_doc_length = idx->stats.total_doc_length.load();
  *doc_count = idx->stats.doc_count.load();
}

uint64_t orc_search_and(const orc_index_t* idx, const uint8_t* term_bytes, const uint64_t* term_offsets,
                        uint64_t n_terms, uint64_t limit, int reverse, uint32_t* out, uint64_t cap) {
  return CopyOut(idx->index->SearchAnd(TermList(term_bytes, term_offsets, 0, n_terms), limit, reverse != 0), out, cap);
}

uint64_t orc_search_or(const orc_index_t* idx, const uint8_t* term_bytes, const uint64_t* term_offsets,
                       uint64_t n_terms, uint32_t* out, uint64_t cap) {
  return CopyOut(idx->index->SearchOr(TermList(term_bytes, term_offsets, 0, n_terms)), out, cap);
}

uint64_t orc_search_not(const orc_index_t* idx, const uint32_t* all_docs, uint64_t n_all, const uint8_t* term_bytes,
                        const uint64_t* term_offsets, uint64_t n_terms, uint32_t* out, uint64_t cap) {
  const std::vector<uint32_t> all(all_docs, all_docs + n_all);
  return CopyOut(idx->index->SearchNot(all, TermList(term_bytes, term_offsets, 0, n_terms)), out, cap);
}

uint64_t orc_filter_by_ngrams(const orc_index_t* idx, const uint32_t* candidates, uint64_t n_candidates,
                              const uint8_t* term_bytes, const uint64_t* term_offsets, uint64_t n_terms,
                              uint32_t* out, uint64_t cap) {
  const std::vector<uint32_t> cands(candidates, candidates + n_candidates);
  return CopyOut(idx->index->FilterByNgrams(cands, TermList(term_bytes, term_offsets, 0, n_terms)), out, cap);
}

uint64_t orc_search_by_threshold(const orc_index_t* idx, const uint8_t* term_bytes, const uint64_t* term_offsets,
                                 uint64_t n_terms, uint64_t threshold, uint32_t* out, uint64_t cap) {
  return CopyOut(idx->index->SearchByThreshold(TermList(term_bytes, term_offsets, 0, n_terms), threshold), out, cap);
}

double orc_compute_idf(uint64_t total_docs, uint64_t doc_freq) {
  return ref::index::BM25Scorer::ComputeIDF(total_docs, doc_freq);
}

uint32_t orc_count_term_occurrences(const uint8_t* text, uint64_t text_len, const uint8_t* term, uint64_t term_len) {
  return ref::index::BM25Scorer::CountTermOccurrences(
      std::string_view(reinterpret_cast<const char*>(text), text_len),
      std::string_view(reinterpret_cast<const char*>(term), term_len));
}

void orc_score_documents(const orc_index_t* idx, const uint32_t* candidates, uint64_t n_candidates,
                         const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* term_doc_freqs,
                         uint64_t n_terms, uint64_t total_docs, double avg_doc_length, double k1, double b,
                         double* out_scores) {
  const std::vector<uint32_t> cands(candidates, candidates + n_candidates);
  const std::vector<uint64_t> dfs(term_doc_freqs, term_doc_freqs + n_terms);
  auto scored = ref::index::BM25Scorer::ScoreDocuments(cands, TermList(term_bytes, term_offsets, 0, n_terms), dfs,
                                                       *idx->store, total_docs, avg_doc_length, {k1, b});
  if (!scored) {
    return;
  }
  for (size_t i = 0; i < scored->size(); ++i) {
    out_scores[i] = (*scored)[i].score;
  }
}

uint64_t orc_sort_by_score(const uint32_t* results, const double* scores, uint64_t n, int descending, uint32_t limit,
                           uint32_t offset, uint32_t* out) {
  const std::vector<uint32_t> r(results, results + n);
  const std::vector<double> s(scores, scores + n);
  const auto sorted = ref::query::ResultSorter::SortByScore(
      r, s, descending != 0 ? ref::query::SortOrder::DESC : ref::query::SortOrder::ASC, limit, offset);
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
  ref::config::Config config = idx->config;
  config.bm25.enable = true;
  config.bm25.k1 = p.k1;
  config.bm25.b = p.b;
  config.memory.verify_text = p.verify_text == 1 ? "all" : (p.verify_text == 2 ? "ascii" : "off");
  const uint64_t total_docs = p.total_docs_override != 0 ? p.total_docs_override : idx->stats.doc_count.load();
  double avgdl = idx->stats.avg_doc_length();
  if (p.total_docs_override != 0) {
    avgdl = static_cast<double>(p.total_len_override) / static_cast<double>(p.total_docs_override);
  }
  if (n_threads <= 0) {
    n_threads = 1;
  }
  std::vector<std::vector<uint32_t>> sets;
  if (out_sets_offsets != nullptr) {
    sets.resize(n_queries);
  }
  std::atomic<uint64_t> next{0};
  std::atomic<int> failures{0};
  auto worker = [&]() {
    for (;;) {
      const uint64_t q = next.fetch_add(1);
      if (q >= n_queries) {
        break;
      }
      const auto terms = TermList(term_bytes, term_offsets, q_term_begin[q], q_term_begin[q + 1]);
      ref::query::Query query;
      query.type = ref::query::QueryType::SEARCH;
      query.table = "t";
      if (!terms.empty()) {
        query.search_text = terms[0];
        query.and_terms.assign(terms.begin() + 1, terms.end());
      }
      if (q_not_begin != nullptr) {
        query.not_terms = TermList(not_bytes, not_offsets, q_not_begin[q], q_not_begin[q + 1]);
      }
      query.limit = p.limit;
      query.offset = p.offset;
      if (p.compute_score != 0) {
        ref::query::OrderByClause order;
        order.column = "_score";
        order.order = p.descending != 0 ? ref::query::SortOrder::DESC : ref::query::SortOrder::ASC;
        query.order_by = order;
      }
      ref::server::search_pipeline::FullPipelineParams fp;
      fp.current_index = idx->index.get();
      fp.current_doc_store = idx->store.get();
      fp.full_config = &config;
      fp.cache_manager = nullptr;  // benchmarks run cache-off (docs/releases/v1.3.5.md:208)
      fp.ngram_size = p.ngram_size;
      fp.kanji_ngram_size = p.kanji_ngram_size;
      fp.cross_boundary_ngrams = p.cross_boundary != 0;
      fp.filter_threshold = p.filter_threshold;
      fp.bm25_stats = &idx->stats;
      auto output = ref::server::search_pipeline::ExecuteFullPipeline(query, fp);
      uint32_t* ids = out_ids + q * stride;
      if (!output) {
        failures.fetch_add(1);
        out_total[q] = 0;
        out_count[q] = 0;
        continue;
      }
      out_total[q] = output->results.size();
      if (out_df != nullptr) {
        // term_infos are size-sorted; map each back to its term slot by text (first unused match)
        std::vector<bool> used(terms.size(), false);
        for (const auto& ti : output->term_infos) {
          for (size_t t = 0; t < terms.size(); ++t) {
            if (!used[t] && terms[t] == ti.normalized_term) {
              used[t] = true;
              out_df[q_term_begin[q] + t] = ti.term_doc_freq;
              break;
            }
          }
        }
      }
      uint64_t written = 0;
      if (p.compute_score != 0) {
        // SearchHandler::HandleSearch, handlers/search_handler.cpp:436-470
        std::vector<std::string> normalized_terms;
        std::vector<uint64_t> term_dfs;
        for (const auto& ti : output->term_infos) {
          normalized_terms.push_back(ti.normalized_term);
          term_dfs.push_back(ti.term_doc_freq);
        }
        auto scored = ref::index::BM25Scorer::ScoreDocuments(output->results, normalized_terms, term_dfs, *idx->store,
                                                             total_docs, avgdl, {p.k1, p.b});
        if (!scored) {
          failures.fetch_add(1);
          out_count[q] = 0;
          continue;
        }
        std::vector<double> scores;
        scores.reserve(scored->size());
        for (const auto& sd : *scored) {
          scores.push_back(sd.score);
        }
        const auto sorted = ref::query::ResultSorter::SortByScore(
            output->results, scores, p.descending != 0 ? ref::query::SortOrder::DESC : ref::query::SortOrder::ASC,
            p.limit, p.offset);
        written = std::min<uint64_t>(sorted.size(), stride);
        for (uint64_t i = 0; i < written; ++i) {
          ids[i] = sorted[i];
          if (out_scores != nullptr) {
            const auto pos =
                std::lower_bound(output->results.begin(), output->results.end(), sorted[i]) - output->results.begin();
            out_scores[q * stride + i] = scores[static_cast<size_t>(pos)];
          }
        }
      } else {
        const size_t start = std::min<size_t>(p.offset, output->results.size());
        const size_t end =
            p.limit == 0 ? output->results.size() : std::min<size_t>(start + p.limit, output->results.size());
        written = std::min<uint64_t>(end - start, stride);
        std::memcpy(ids, output->results.data() + start, written * sizeof(uint32_t));
      }
      out_count[q] = static_cast<uint32_t>(written);
      if (out_sets_offsets != nullptr) {
        sets[q] = std::move(output->results);
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
  return failures.load() == 0 ? 0 : -failures.load();
}

uint64_t orc_eval_boolean(const orc_index_t* idx, const int32_t* ops, const int32_t* args, uint64_t n_ops,
                          const uint8_t* term_bytes, const uint64_t* term_offsets, uint32_t* out, uint64_t cap) {
  using ref::query::NodeType;
  using ref::query::QueryNode;
  std::vector<std::unique_ptr<QueryNode>> stack;
  for (uint64_t i = 0; i < n_ops; ++i) {
    if (ops[i] == 0) {
      const auto t = static_cast<uint64_t>(args[i]);
      stack.push_back(std::make_unique<QueryNode>(
          std::string(reinterpret_cast<const char*>(term_bytes) + term_offsets[t], term_offsets[t + 1] - term_offsets[t])));
    } else if (ops[i] == 1 || ops[i] == 2) {
      const auto n = static_cast<size_t>(args[i]);
      if (n > stack.size()) {
        return 0;
      }
      auto node = std::make_unique<QueryNode>(ops[i] == 1 ? NodeType::AND : NodeType::OR);
      for (size_t c = stack.size() - n; c < stack.size(); ++c) {
        node->children.push_back(std::move(stack[c]));
      }
      stack.resize(stack.size() - n);
      stack.push_back(std::move(node));
    } else if (ops[i] == 3) {
      if (stack.empty()) {
        return 0;
      }
      auto node = std::make_unique<QueryNode>(NodeType::NOT);
      node->children.push_back(std::move(stack.back()));
      stack.back() = std::move(node);
    }
  }
  if (stack.empty()) {
    return 0;
  }
  return CopyOut(stack.back()->Evaluate(*idx->index, *idx->store), out, cap);
}

int orc_contains_fuzzy_match(const uint8_t* text, uint64_t text_len, const uint8_t* term, uint64_t term_len,
                             uint32_t max_distance) {
  return mygram::utils::ContainsFuzzyMatch(std::string_view(reinterpret_cast<const char*>(text), text_len),
                                           std::string_view(reinterpret_cast<const char*>(term), term_len),
                                           max_distance)
             ? 1
             : 0;
}

// The reference's own ExecuteWithFuzzy (search_pipeline.cpp:1659-1740), called the way ExecuteFullPipeline does
// (:1908-1937): term infos from GenerateTermInfos, results cleared under empty_term_detected.
uint64_t orc_search_fuzzy(const orc_index_t* idx, const orc_query_params_t* params, const uint8_t* term_bytes,
                          const uint64_t* term_offsets, uint64_t n_terms, uint32_t max_distance,
                          const uint8_t* not_bytes, const uint64_t* not_offsets, uint64_t n_not, uint32_t* out,
                          uint64_t cap, int32_t* empty_term_detected) {
  namespace sp = ref::server::search_pipeline;
  const orc_query_params_t p = *params;
  ref::config::Config config = idx->config;
  config.memory.verify_text = p.verify_text == 1 ? "all" : (p.verify_text == 2 ? "ascii" : "off");
  const auto terms = TermList(term_bytes, term_offsets, 0, n_terms);
  ref::query::Query query;
  query.type = ref::query::QueryType::SEARCH;
  query.table = "t";
  if (n_not > 0) {
    query.not_terms = TermList(not_bytes, not_offsets, 0, n_not);
  }
  query.fuzzy_max_distance = max_distance;
  const auto term_infos =
      sp::GenerateTermInfos(terms, idx->index.get(), p.ngram_size, p.kanji_ngram_size, p.cross_boundary != 0);
  auto result = sp::ExecuteWithFuzzy(query, term_infos, terms, max_distance, idx->index.get(), idx->store.get(),
                                     &config, p.ngram_size, p.kanji_ngram_size, p.cross_boundary != 0,
                                     p.filter_threshold);
  if (empty_term_detected != nullptr) {
    *empty_term_detected = result.empty_term_detected ? 1 : 0;
  }
  if (result.empty_term_detected) {
    result.results.clear();
  }
  return CopyOut(result.results, out, cap);
}

// The reference's own ExecuteWithSynonyms (search_pipeline.cpp:1580-1631) over groups built the way
// ExpandNormalizedTermWithSynonyms does (:1360-1388; the SynonymDictionary itself only loads from a file): one
// SearchTermInfo per variant from GenerateTermInfos, normalized_terms = the variants.
uint64_t orc_search_synonyms(const orc_index_t* idx, const orc_query_params_t* params, const uint8_t* variant_bytes,
                             const uint64_t* variant_offsets, const uint64_t* group_begin, uint64_t n_groups,
                             const uint8_t* not_bytes, const uint64_t* not_offsets, uint64_t n_not, uint32_t* out,
                             uint64_t cap, int32_t* empty_term_detected) {
  namespace sp = ref::server::search_pipeline;
  const orc_query_params_t p = *params;
  ref::config::Config config = idx->config;
  config.memory.verify_text = p.verify_text == 1 ? "all" : (p.verify_text == 2 ? "ascii" : "off");
  std::vector<sp::SynonymTermGroup> groups;
  for (uint64_t g = 0; g < n_groups; ++g) {
    sp::SynonymTermGroup group;
    group.normalized_terms = TermList(variant_bytes, variant_offsets, group_begin[g], group_begin[g + 1]);
    group.variants = sp::GenerateTermInfos(group.normalized_terms, idx->index.get(), p.ngram_size,
                                           p.kanji_ngram_size, p.cross_boundary != 0);
    groups.push_back(std::move(group));
  }
  ref::query::Query query;
  query.type = ref::query::QueryType::SEARCH;
  query.table = "t";
  if (n_not > 0) {
    query.not_terms = TermList(not_bytes, not_offsets, 0, n_not);
  }
  auto result = sp::ExecuteWithSynonyms(query, groups, idx->index.get(), idx->store.get(), &config, p.ngram_size,
                                        p.kanji_ngram_size, p.cross_boundary != 0, p.filter_threshold, nullptr);
  if (empty_term_detected != nullptr) {
    *empty_term_detected = result.empty_term_detected ? 1 : 0;
  }
  if (result.empty_term_detected) {
    result.results.clear();
  }
  return CopyOut(result.results, out, cap);
}

// Index::SaveToStream / LoadFromStream (index_serialization.cpp:111-224, 279-613): the MGIX v4 stream of the index
// (reference-only entry points: the MGIX codec of the product is checked against the reference itself).
// save: returns the stream length (bytes copied if it fits `cap`), 0 on failure.
uint64_t ref_index_save_stream(const orc_index_t* idx, uint8_t* out, uint64_t cap) {
  std::ostringstream stream(std::ios::binary);
  if (!idx->index->SaveToStream(stream)) {
    return 0;
  }
  const std::string bytes = stream.str();
  if (out != nullptr && bytes.size() <= cap) {
    std::memcpy(out, bytes.data(), bytes.size());
  }
  return bytes.size();
}
// load: replaces the index content; returns 0 on success, the reference's ErrorCode otherwise.
int ref_index_load_stream(orc_index_t* idx, const uint8_t* data, uint64_t len) {
  std::istringstream stream(std::string(reinterpret_cast<const char*>(data), len), std::ios::binary);
  auto loaded = idx->index->LoadFromStream(stream);
  return loaded ? 0 : static_cast<int>(loaded.error().code());
}

}  // extern "C"

// Column filters through the reference's own DocumentStore / FilterIndex / ApplyFiltersWithBitmap
// (search_pipeline.cpp:1196-1237). A fresh store is filled with one document per row carrying the typed values.
extern "C" uint64_t orc_apply_filters(uint64_t n_docs, uint32_t first_doc_id, uint32_t n_cols, const int32_t* col_type,
                                      const uint64_t* col_values, const uint8_t* col_null, const uint8_t* str_bytes,
                                      const uint64_t* str_offsets, uint32_t n_filters, const uint32_t* filter_col,
                                      const uint8_t* filter_op, const uint8_t* lit_bytes, const uint64_t* lit_offsets,
                                      const uint32_t* results, uint64_t n_results, uint32_t* out) {
  namespace st = ref::storage;
  OpenDocumentStore store;
  store.SetNextDocId(first_doc_id);
  for (uint64_t row = 0; row < n_docs; ++row) {
    st::FilterMap filters;
    for (uint32_t c = 0; c < n_cols; ++c) {
      const uint64_t bits = col_values[static_cast<uint64_t>(c) * n_docs + row];
      const std::string name = "c" + std::to_string(c);
      if (col_null[static_cast<uint64_t>(c) * n_docs + row] != 0) {
        filters[name] = st::FilterValue{std::monostate{}};
        continue;
      }
      switch (col_type[c]) {
        case 1: filters[name] = st::FilterValue{bits != 0}; break;
        case 2: filters[name] = st::FilterValue{static_cast<int8_t>(bits)}; break;
        case 3: filters[name] = st::FilterValue{static_cast<uint8_t>(bits)}; break;
        case 4: filters[name] = st::FilterValue{static_cast<int16_t>(bits)}; break;
        case 5: filters[name] = st::FilterValue{static_cast<uint16_t>(bits)}; break;
        case 6: filters[name] = st::FilterValue{static_cast<int32_t>(bits)}; break;
        case 7: filters[name] = st::FilterValue{static_cast<uint32_t>(bits)}; break;
        case 8: filters[name] = st::FilterValue{static_cast<int64_t>(bits)}; break;
        case 9: filters[name] = st::FilterValue{static_cast<uint64_t>(bits)}; break;
        case 10: filters[name] = st::FilterValue{st::TimeValue{static_cast<int64_t>(bits)}}; break;
        case 11:
          filters[name] = st::FilterValue{std::string(reinterpret_cast<const char*>(str_bytes) + str_offsets[bits],
                                                      str_offsets[bits + 1] - str_offsets[bits])};
          break;
        case 12: {
          double v = 0.0;
          std::memcpy(&v, &bits, sizeof(v));
          filters[name] = st::FilterValue{v};
          break;
        }
        default: break;
      }
    }
    auto added = store.AddDocument("pk" + std::to_string(row), filters, "", "");
    if (!added) {
      return 0;
    }
  }
  std::vector<ref::query::FilterCondition> conds;
  for (uint32_t f = 0; f < n_filters; ++f) {
    ref::query::FilterCondition fc;
    fc.column = "c" + std::to_string(filter_col[f]);
    fc.op = static_cast<ref::query::FilterOp>(filter_op[f]);
    fc.value.assign(reinterpret_cast<const char*>(lit_bytes) + lit_offsets[f], lit_offsets[f + 1] - lit_offsets[f]);
    conds.push_back(std::move(fc));
  }
  const std::vector<ref::storage::DocId> in(results, results + n_results);
  const auto kept = ref::server::search_pipeline::ApplyFiltersWithBitmap(in, conds, &store);
  std::copy(kept.begin(), kept.end(), out);
  return kept.size();
}
