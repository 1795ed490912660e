/* oracle/oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * C ABI of the CPU restatement ("port") of the reference's search core. Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (libmgx.so) never does.
 *
 * Parity status: PINNED. The restatement is checked (tests/test_oracle_*.py)
 * against (1) the known-answer vectors transcribed from the reference's own
 * unit tests (tests/golden/reference_kat.json, each entry citing test file:line)
 * and (2) outputs of the reference's unmodified sources compiled into
 * oracle/_ref/libmygram_ref.so (fixtures in tests/golden/ref_*.json produced by
 * oracle/gen_golden.py, which is committed).
 *
 * All citations are relative to /root/reference/.
 */
#ifndef ORACLE_H_
#define ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- tokenizer: src/utils/string_utils.cpp ---- */

/* Utf8ToCodepoints (string_utils.cpp:200-219). Returns the number of code
 * points; writes at most `cap` of them. */
uint64_t orc_utf8_to_codepoints(const uint8_t* text, uint64_t len, uint32_t* out, uint64_t cap);
/* CodepointsToUtf8 (string_utils.cpp:243-272). Returns bytes written (<= 4*n). */
uint64_t orc_codepoints_to_utf8(const uint32_t* cps, uint64_t n, uint8_t* out);
/* CountCodePoints (string_utils.cpp:655-669). */
uint64_t orc_count_code_points(const uint8_t* text, uint64_t len);
/* IsCJKIdeograph (string_utils.cpp:441-448). */
int orc_is_cjk_ideograph(uint32_t cp);

/* N-gram generators. The n-grams are returned concatenated in `out_bytes`
 * with `out_offsets[i]..out_offsets[i+1]` delimiting n-gram i (so
 * out_offsets needs count+1 slots). Return value = number of n-grams, or
 * -(needed) if a capacity was too small (nothing useful written).
 *   mode 0: GenerateNgrams(text, a)                      string_utils.cpp:382-423
 *   mode 1: GenerateHybridNgrams(text, a, k, cross)      string_utils.cpp:452-509
 *   mode 2: GenerateQueryNgrams(text, a, k, cross)       string_utils.cpp:639-653 */
int64_t orc_ngrams(int mode, const uint8_t* text, uint64_t len, int a, int k, int cross, uint8_t* out_bytes,
                   uint64_t cap_bytes, uint64_t* out_offsets, uint64_t cap_ngrams);

/* ---- index: src/index/index.cpp + posting_list.cpp (set semantics only) ---- */

typedef struct orc_index orc_index_t;

/* Index::Index (index.cpp:29-37): kanji_ngram_size <= 0 means "use ngram_size". */
orc_index_t* orc_index_create(int ngram_size, int kanji_ngram_size, int cross_boundary);
void orc_index_destroy(orc_index_t* idx);

/* Index::AddDocument (index.cpp:39-74). Returns 1 if indexed, 0 if the text
 * produced no n-grams. Also records the normalised text in the oracle's
 * document store (DocumentStore::AddDocument, document_store.cpp:144-147: empty
 * text is not stored) and the BM25 corpus stats (server_types.h:189-193). */
int orc_index_add_document(orc_index_t* idx, uint32_t doc_id, const uint8_t* text, uint64_t len);
/* Index::AddDocumentBatch (index.cpp:76-119) fed in `batch` sized groups the
 * way InitialLoader::FlushBatch does (initial_loader.cpp:41,450-512). */
void orc_index_add_batch(orc_index_t* idx, const uint32_t* doc_ids, const uint8_t* text, const uint64_t* offsets,
                         uint64_t n_docs, uint64_t batch);
/* Same resulting index as orc_index_add_batch, built by a multi-threaded
 * tokenise + sort of packed (n-gram, doc) pairs. Exists so that a 10M-document
 * CPU index can be prepared in seconds for the timed CPU baseline; equality
 * with orc_index_add_batch is tested. Requires ngram sizes <= 3. Text is
 * referenced, not copied: the caller keeps `text`/`offsets`/`doc_ids` alive. */
int orc_index_build_bulk(orc_index_t* idx, const uint32_t* doc_ids, const uint8_t* text, const uint64_t* offsets,
                         uint64_t n_docs, int n_threads);
/* Index::RemoveDocument (index.cpp:175-197), Index::UpdateDocument (:121-173). */
void orc_index_remove_document(orc_index_t* idx, uint32_t doc_id, const uint8_t* text, uint64_t len);
void orc_index_update_document(orc_index_t* idx, uint32_t doc_id, const uint8_t* old_text, uint64_t old_len,
                               const uint8_t* new_text, uint64_t new_len);

uint64_t orc_index_term_count(const orc_index_t* idx);                                       /* index.h:193 */
uint64_t orc_index_posting_size(const orc_index_t* idx, const uint8_t* term, uint64_t len); /* index.cpp:580-588 */
uint64_t orc_index_total_postings(const orc_index_t* idx);
/* PostingList::GetAll (posting_list.cpp:421-430) for one term. Returns size. */
uint64_t orc_index_get_postings(const orc_index_t* idx, const uint8_t* term, uint64_t len, uint32_t* out,
                                uint64_t cap);
/* Whole index in term-byte order: term strings (concatenated + offsets),
 * posting offsets and postings. Pass NULL outputs to size the buffers:
 * returns term count, *term_bytes = total term bytes. */
uint64_t orc_index_export(const orc_index_t* idx, uint8_t* term_bytes_out, uint64_t* term_offsets_out,
                          uint64_t* posting_offsets_out, uint32_t* postings_out, uint64_t* total_term_bytes);

/* BM25Stats (server_types.h:157-220) as maintained by the add/remove calls. */
void orc_index_bm25_stats(const orc_index_t* idx, uint64_t* total_doc_length, uint64_t* doc_count);

/* Term lists for the search calls below: `n_terms` strings concatenated in
 * `term_bytes`, string i = term_bytes[term_offsets[i]..term_offsets[i+1]). */

/* Index::SearchAnd (index.cpp:199-368). Returns result size (may exceed cap). */
uint64_t orc_search_and(const orc_index_t* idx, const uint8_t* term_bytes, const uint64_t* term_offsets,
                        uint64_t n_terms, uint64_t limit, int reverse, uint32_t* out, uint64_t cap);
/* Index::SearchOr (index.cpp:418-448). */
uint64_t orc_search_or(const orc_index_t* idx, const uint8_t* term_bytes, const uint64_t* term_offsets,
                       uint64_t n_terms, uint32_t* out, uint64_t cap);
/* Index::SearchNot (index.cpp:450-486). */
uint64_t orc_search_not(const orc_index_t* idx, const uint32_t* all_docs, uint64_t n_all, const uint8_t* term_bytes,
                        const uint64_t* term_offsets, uint64_t n_terms, uint32_t* out, uint64_t cap);
/* Index::FilterByNgrams (index.cpp:370-416). */
uint64_t orc_filter_by_ngrams(const orc_index_t* idx, const uint32_t* candidates, uint64_t n_candidates,
                              const uint8_t* term_bytes, const uint64_t* term_offsets, uint64_t n_terms,
                              uint32_t* out, uint64_t cap);
/* Index::SearchByThreshold (index.cpp:488-578). */
uint64_t orc_search_by_threshold(const orc_index_t* idx, const uint8_t* term_bytes, const uint64_t* term_offsets,
                                 uint64_t n_terms, uint64_t threshold, uint32_t* out, uint64_t cap);

/* ---- BM25: src/index/bm25_scorer.cpp, src/query/result_sorter.cpp ---- */

double orc_compute_idf(uint64_t total_docs, uint64_t doc_freq);                       /* bm25_scorer.cpp:14-25 */
uint32_t orc_count_term_occurrences(const uint8_t* text, uint64_t text_len, const uint8_t* term,
                                    uint64_t term_len);                                /* bm25_scorer.cpp:27-45 */
/* BM25Scorer::ScoreDocuments (bm25_scorer.cpp:47-99): one score per candidate,
 * in candidate order; a candidate with no stored text scores 0.0. */
void orc_score_documents(const orc_index_t* idx, const uint32_t* candidates, uint64_t n_candidates,
                         const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* term_doc_freqs,
                         uint64_t n_terms, uint64_t total_docs, double avg_doc_length, double k1, double b,
                         double* out_scores);
/* ResultSorter::SortByScore (result_sorter.cpp:661-716). descending != 0 is
 * SortOrder::DESC. Returns the number of ids written to `out`. */
uint64_t orc_sort_by_score(const uint32_t* results, const double* scores, uint64_t n, int descending, uint32_t limit,
                           uint32_t offset, uint32_t* out);

/* ---- search pipeline, regular path (src/server/search_pipeline.cpp) ---- */

typedef struct {
  int32_t ngram_size;          /* raw table config values, as passed by the   */
  int32_t kanji_ngram_size;    /* pipeline to GenerateQueryNgrams (:578)      */
  int32_t cross_boundary;
  int32_t compute_score;       /* SORT _score: df + ScoreDocuments + SortByScore */
  int32_t descending;          /* SortOrder for _score                        */
  uint32_t limit;              /* query_parser.h:217 default 100              */
  uint32_t offset;
  uint32_t filter_threshold;   /* search_pipeline.h:331 default 1000          */
  double k1;                   /* bm25_scorer.h:23-26                         */
  double b;
  /* corpus stats override for shard-parallel runs; 0 => use the index's own */
  uint64_t total_docs_override;
  uint64_t total_len_override;
  int32_t verify_text;         /* memory.verify_text: 0 "off" (config.h:329), 1 "all", 2 "ascii" */
  int32_t reserved;
} orc_query_params_t;

/* One query = GenerateTermInfos (:569-603, df via PopulateTermDocumentFrequency
 * :542-565 when compute_score) -> sort by estimated_size (:2012-2014) ->
 * Execute (:795-869: early exit, AND smallest-first with FilterByNgrams below
 * filter_threshold, ApplyNotFilter :871-932, ApplyVerifyTextFilter :1248-1266,
 * hybrid-fragment PostFilterByText :858-866) -> [ScoreDocuments -> SortByScore
 * as in handlers/search_handler.cpp:405-470] else ascending ids cut to
 * [offset, offset+limit).
 *
 * Queries are given as ranges into a flat term table:
 *   query q uses terms  [q_term_begin[q], q_term_begin[q+1])  as search terms and
 *                       [q_not_begin[q],  q_not_begin[q+1])   (into the NOT table) as NOT terms.
 * Outputs per query q: out_total[q] = |result set|, out_count[q] = ids written
 * at out_ids[q*stride ...], scores at out_scores[q*stride ...] (if compute_score),
 * and per search term its verified document frequency in out_df (same
 * indexing as the term table; may be NULL).
 * If out_sets_offsets != NULL the full ascending result set of every query is
 * appended to out_sets (capacity sets_cap) with out_sets_offsets[q..q+1].
 * Runs `n_threads` worker threads, one query at a time per thread (the
 * reference's only parallelism, src/server/thread_pool.cpp). Returns 0. */
int orc_query_batch(const orc_index_t* idx, const orc_query_params_t* params, uint64_t n_queries,
                    const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                    const uint8_t* not_bytes, const uint64_t* not_offsets, const uint64_t* q_not_begin,
                    uint64_t stride, uint32_t* out_ids, double* out_scores, uint32_t* out_count,
                    uint64_t* out_total, uint64_t* out_df, uint32_t* out_sets, uint64_t sets_cap,
                    uint64_t* out_sets_offsets, int n_threads);

/* ---- boolean AST (src/query/query_ast.cpp:67-161) ----
 * Postfix program: op 0 = TERM(arg = index into term table), 1 = AND(arg = n
 * children), 2 = OR(arg = n children), 3 = NOT(one child). Children are
 * evaluated left to right exactly as QueryNode::Evaluate does. */
uint64_t orc_eval_boolean(const orc_index_t* idx, const int32_t* ops, const int32_t* args, uint64_t n_ops,
                          const uint8_t* term_bytes, const uint64_t* term_offsets, uint32_t* out, uint64_t cap);

/* ---- fuzzy and synonym execution paths (src/server/search_pipeline.cpp:1580-1752) ----
 * ExecuteWithFuzzy (:1659-1740) over normalised terms: per term Index::SearchByThreshold(ngrams,
 * max(1, |ngrams| - max_distance * effective n-gram size)), AND-ed; NOT terms; PostFilterByFuzzyText (:1742-1752,
 * ContainsFuzzyMatch of utils/edit_distance.cpp) when verify_text applies; hybrid-fragment exact text filter.
 * Column filters are applied by the caller with orc_apply_filters (they are per-document predicates).
 * Returns the result size (ids ascending). *empty_term_detected (may be NULL) as in SearchPipelineResult. */
/* ContainsFuzzyMatch (src/utils/edit_distance.cpp:168-255). */
int orc_contains_fuzzy_match(const uint8_t* text, uint64_t text_len, const uint8_t* term, uint64_t term_len,
                             uint32_t max_distance);
uint64_t orc_search_fuzzy(const orc_index_t* idx, const orc_query_params_t* params, const uint8_t* term_bytes,
                          const uint64_t* term_offsets, uint64_t n_terms, uint32_t max_distance,
                          const uint8_t* not_bytes, const uint64_t* not_offsets, uint64_t n_not, uint32_t* out,
                          uint64_t cap, int32_t* empty_term_detected);
/* ExecuteWithSynonyms (:1580-1631) + PostFilterByTextWithSynonyms (:1633-1657) over expanded groups: group g =
 * variants [group_begin[g], group_begin[g+1]) (the term and its synonyms, ExpandTermsWithSynonyms :1392-1406).
 * NOT terms: pass every synonym of every NOT term. */
uint64_t orc_search_synonyms(const orc_index_t* idx, const orc_query_params_t* params, const uint8_t* variant_bytes,
                             const uint64_t* variant_offsets, const uint64_t* group_begin, uint64_t n_groups,
                             const uint8_t* not_bytes, const uint64_t* not_offsets, uint64_t n_not, uint32_t* out,
                             uint64_t cap, int32_t* empty_term_detected);

/* ---- column filters (src/server/search_pipeline.cpp:1021-1237) ----
 * ApplyFiltersWithBitmap (:1196-1237): when every condition is EQ / NE the FilterIndex bitmaps decide (a value
 * matches when its serialisation equals that of some type interpretation of the literal, BuildTypeUnionBitmap
 * :1021-1094); otherwise ApplyFilters (:1098-1194) compares every document's typed value with the pre-parsed
 * literal (ParseFilterValue :954-993; NULL matches only NE).
 * Stand-alone form: documents are rows 0..n_docs-1, doc id = first_doc_id + row. Column c has type
 * col_type[c] (the FilterValue variant index, document_store.h:73-86: 1 bool, 2 int8, 3 uint8, 4 int16, 5 uint16,
 * 6 int32, 7 uint32, 8 int64, 9 uint64, 10 TIME seconds, 11 string, 12 double); its value for a row is
 * col_values[c * n_docs + row] (integer / bool / seconds value, the bits of the double, or for strings an index
 * into the string table str_bytes / str_offsets) and col_null[c * n_docs + row] != 0 marks NULL.
 * Filter f: column filter_col[f], op filter_op[f] (query_parser.h:93-100: 0 EQ, 1 NE, 2 GT, 3 GTE, 4 LT, 5 LTE),
 * literal lit_bytes[lit_offsets[f] .. lit_offsets[f+1]). `results` are ascending doc ids. Returns the number of ids
 * written to `out` (capacity n_results). */
uint64_t orc_apply_filters(uint64_t n_docs, uint32_t first_doc_id, uint32_t n_cols, const int32_t* col_type,
                           const uint64_t* col_values, const uint8_t* col_null, const uint8_t* str_bytes,
                           const uint64_t* str_offsets, uint32_t n_filters, const uint32_t* filter_col,
                           const uint8_t* filter_op, const uint8_t* lit_bytes, const uint64_t* lit_offsets,
                           const uint32_t* results, uint64_t n_results, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* ORACLE_H_ */
