/* include/mgx.h — C ABI of the B200-native MygramDB search core (libmgx.so).
 *
 * This is the drop-in boundary. The reference has no FFI layer for this path:
 * the path sits behind in-process C++ classes (mygramdb_index, mygramdb_query;
 * src/index/CMakeLists.txt:1-22). Each entry point below names the reference
 * interface it replaces (paths relative to the reference tree). The C++17
 * adapter in mygram-db_b200/adapter/mygram_adapter.h re-creates the reference's
 * class signatures on top of this ABI; INTEGRATION.md shows the binding.
 *
 * Conventions (modelled on the reference's own C client, src/client/mygramclient_c.h):
 * opaque handles, `int` status (0 = ok, negative = error), caller-owned output
 * buffers, mgx_last_error() for a thread-local message. No torch / C++ types.
 * Unless a name ends in `_device`, every pointer is a HOST pointer; the library
 * does its own H2D/D2H copies (pinned host memory makes them asynchronous).
 *
 * Document ids are the reference's DocId = uint32_t (src/types/doc_id.h:31).
 * Text is normalised UTF-8 (ICU normalisation stays on the host, index.h:83-84).
 * Strings are passed flattened: `bytes` + `offsets[n+1]` (uint64), string i is
 * bytes[offsets[i] .. offsets[i+1]).
 *
 * Threading contract (the reference calls this path from its worker pool, server/thread_pool.cpp:33, with one
 * binlog-apply writer beside it; Index guards itself with shared_mutexes, index.h:343-359):
 *  - every entry point may be called from any thread, concurrently, on the same handle;
 *  - reading calls (searches, batches, stats, export) share the index; concurrent single calls run side by side on
 *    the device, each on its own stream with its own workspace (up to 8 at a time, further callers wait);
 *  - mgx_index_build*, mgx_index_set_filter_column, mgx_index_clear / trim / optimize take the index exclusively:
 *    they wait for the readers in flight and hold new ones back meanwhile. The commit of journaled mutations builds
 *    the next generation of the shard beside the current one while readers go on, and is exclusive only for the
 *    exchange of the two; writers never overlap each other;
 *  - add / update / remove_document only append to a journal (never block on readers or on a commit); a reading call
 *    that STARTS after they return sees them (it commits the journal first, or waits for the commit that is doing
 *    so); with mgx_index_set_commit_mode(index, 1) reading calls never commit or wait, and mgx_index_commit
 *    publishes the mutations;
 *  - a staged batch (mgx_batch_prepare* .. mgx_batch_destroy) is a reader for its whole lifetime. A thread must not
 *    call a committing entry point of the same index while it holds such a batch and mutations are pending: the
 *    commit would wait for the batch. Stages of ONE batch must not be issued concurrently.
 */
#ifndef MGX_H_
#define MGX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGX_OK 0
#define MGX_ERR_INVALID_ARGUMENT (-1) /* reference: ErrorCode::kInvalidArgument (utils/error.h) */
#define MGX_ERR_CUDA (-2)             /* reference: ErrorCode::kInternalError                  */
#define MGX_ERR_UNSUPPORTED (-3)      /* a configuration this build refuses rather than guesses */
#define MGX_ERR_CAPACITY (-4)         /* a caller buffer was too small; nothing partial is valid */
#define MGX_ERR_NO_DEVICE (-5)        /* no CUDA device: the library has no CPU fallback        */
#define MGX_ERR_FORMAT (-6)           /* an MGIX stream was rejected; mgx_last_error() names the reference's
                                         ErrorCode (kStorageInvalidFormat / CRCMismatch / Corrupted / ...) */
#define MGX_ERR_TIMEOUT (-7)          /* mgx_share_*: the other side did not arrive within timeout_ms        */

/* Thread-local description of the last failure on this thread ("" if none). */
const char* mgx_last_error(void);
/* Library / build identification, e.g. "mgx 0.1 sm_100a". */
const char* mgx_version(void);
/* Number of kernels this library has launched in this process so far. */
uint64_t mgx_kernel_launch_count(void);

/* ------------------------------------------------------------------ index */

typedef struct mgx_index mgx_index_t;

/* Mirrors Index::Index(ngram_size, kanji_ngram_size, roaring_threshold,
 * cross_boundary_ngrams, ...) — src/index/index.h:58-60, index.cpp:29-37.
 * kanji_ngram_size <= 0 means "use ngram_size" (index.cpp:32). */
typedef struct {
  int32_t ngram_size;            /* 1..10 (config-schema.json:279-285); sizes above 3 use wide keys */
  int32_t kanji_ngram_size;      /* 0..10                                                   */
  int32_t cross_boundary_ngrams; /* config.h:211-213 default true                           */
  int32_t device;                /* CUDA device ordinal                                     */
  double dense_threshold;        /* posting density at which a list ALSO gets a doc bitmap; */
                                 /* <= 0 selects 1/128 (4x below the size break-even 4*|P| = N/8: a bit probe is */
                                 /* one memory round trip, a range search several; measured 5 % faster batches)  */
  uint64_t max_dense_bytes;      /* cap for all dense bitmaps together; 0 = 8 GiB           */
  uint64_t scratch_bytes;        /* per-batch query scratch; 0 = 4 GiB                      */
  double roaring_threshold;      /* Index's roaring_threshold (index.h:58, default 0.18): only used to REPORT the    */
                                 /* reference's delta / Roaring list counts (mgx_index_get_statistics); <= 0 = 0.18 */
} mgx_index_config_t;

int mgx_index_create(const mgx_index_config_t* config, mgx_index_t** out);
void mgx_index_destroy(mgx_index_t* index);

/* Bulk build of one shard: replaces the loop of DocumentStore::AddDocumentBatch +
 * Index::AddDocumentBatch over 1000-document batches (loader/initial_loader.cpp:450-512,
 * index.cpp:76-119) and the BM25Stats rebuild (app/server_orchestrator.cpp:758-772).
 * doc_ids must be strictly ascending. Replaces any previous content.
 * The text arena, the code-point lengths and the doc ids stay resident on the
 * device as the mirror of DocumentStore's normalised text (document_store.h:380-426). */
int mgx_index_build(mgx_index_t* index, const uint32_t* doc_ids, const uint8_t* text, const uint64_t* text_offsets,
                    uint64_t n_docs);
/* Same, but all three arrays are already DEVICE pointers on index's device
 * (kernel-only timing: no H2D inside). */
int mgx_index_build_device(mgx_index_t* index, const uint32_t* d_doc_ids, const uint8_t* d_text,
                           const uint64_t* d_text_offsets, uint64_t n_docs);

/* Incremental mutations — Index::AddDocument (index.cpp:39-74), UpdateDocument (:121-173), RemoveDocument
 * (:175-197), as the binlog applier calls them (mysql/binlog_event_processor.cpp:66-324) together with the
 * matching DocumentStore calls: the document of `doc_id` (text + postings) is added, replaced or removed. The
 * calls are journaled on the host and take effect before the NEXT read of the index (any search / stats / export
 * call, or mgx_index_commit): the resident corpus is merged with the journal on the device and the shard is
 * rebuilt (~0.1 s per 10M documents), so a burst of mutations costs one rebuild. The new generation of the shard is
 * built BESIDE the current one: calls that are reading the index (staged batches included) go on during the build, and
 * the exchange of the two generations is the only exclusive moment. If that commit fails (device out of memory,
 * more than 2^32 n-gram occurrences) the reading call returns the error, the shard stays at the generation before the
 * commit and the journal is discarded: rebuild from the DocumentStore with mgx_index_build. `old_text` / `text` of update /
 * remove must be the text the document currently has (what the reference requires to find its n-grams); the
 * resident copy is what is actually used. *out_indexed (may be NULL) = AddDocument's return value: 0 when the text
 * yields no n-gram (the document is stored but never matches). */
int mgx_index_add_document(mgx_index_t* index, uint32_t doc_id, const uint8_t* text, uint64_t text_len,
                           int32_t* out_indexed);
/* Index::AddDocumentBatch (index.cpp:76-119) as InitialLoader::FlushBatch calls it for every 1000 documents
 * (loader/initial_loader.cpp:450-512): ADDITIVE, ids in any order, the index keeps what it has. The batch is appended to
 * the same journal as the single-document calls (three memcpys) and folded in before the next read -- a loader's ten
 * thousand batches cost ONE merge + build at the first search. *out_indexed (may be NULL) = documents whose text yields
 * at least one n-gram (the reference skips the others silently, :93-101; here they are stored and never match).
 * When the whole snapshot is at hand, mgx_index_build is the direct (and faster) form of the same thing. */
int mgx_index_add_document_batch(mgx_index_t* index, const uint32_t* doc_ids, const uint8_t* text,
                                 const uint64_t* text_offsets, uint64_t n_docs, uint64_t* out_indexed);
int mgx_index_update_document(mgx_index_t* index, uint32_t doc_id, const uint8_t* old_text, uint64_t old_len,
                              const uint8_t* new_text, uint64_t new_len);
int mgx_index_remove_document(mgx_index_t* index, uint32_t doc_id, const uint8_t* text, uint64_t text_len);
int mgx_index_commit(mgx_index_t* index);
/* Who commits, and who waits. overlapped == 0 (default): every reading call commits the journal first (or waits for
 * the commit that is doing so), so it sees every mutation that returned before it began (the reference applies a
 * mutation before Index::AddDocument returns). overlapped != 0: reading calls neither commit nor wait -- they answer
 * from the current generation, and mutations become visible when the mgx_index_commit of the mutating thread (the
 * binlog applier, binlog_event_processor.cpp) returns. This is the mode for servers whose search threads keep
 * staged batches in flight: a commit only ever waits for batches to END, never the other way round. */
int mgx_index_set_commit_mode(mgx_index_t* index, int overlapped);

typedef struct {
  uint64_t n_docs;
  uint64_t n_terms;          /* Index::TermCount, index.h:193                       */
  uint64_t n_postings;       /* IndexStatistics::total_postings, index.cpp:604-633  */
  uint64_t n_dense_terms;    /* lists that also have a bitmap                        */
  uint64_t text_bytes;
  uint64_t total_doc_length; /* BM25Stats::total_doc_length, server_types.h:158     */
  uint64_t doc_count;        /* BM25Stats::doc_count (docs with non-empty text)     */
  uint64_t device_bytes;     /* resident device memory of this index                */
  uint64_t n_pair_slots;     /* (n-gram, doc) slots sorted by the last build        */
  int32_t all_valid_utf8;    /* 1 if no document contained an invalid byte          */
  int32_t key_width;         /* code points per packed key = max(ngram, kanji)      */
  double last_build_ms;      /* device time of the last build (CUDA events)         */
} mgx_index_stats_t;

int mgx_index_get_stats(const mgx_index_t* index, mgx_index_stats_t* out);

/* Index::GetStatistics (index.cpp:604-633) / Index::Optimize(total_docs) (index_optimization.cpp:36-120,
 * posting_list.cpp:799-834) / Index::Clear (index.cpp:635-641). The device index keeps every list as sorted
 * uint32 plus a bitmap for dense lists (chosen for speed, see mgx_index_config_t.dense_threshold); the reference's
 * representation only shows in these counters, which are reproduced by rule: a list is "Roaring" when it has more
 * than 4096 entries (posting_list.cpp:21,917-922: converted on insert) or, once Optimize(total_docs) has run, when
 * size / total_docs >= roaring_threshold; every other list is "delta". memory_usage_bytes is the resident device
 * memory of this index (the reference reports host bytes of its own containers). */
typedef struct {
  uint64_t total_terms;
  uint64_t total_postings;
  uint64_t delta_encoded_lists;
  uint64_t roaring_bitmap_lists;
  uint64_t memory_usage_bytes;
} mgx_index_statistics_t;
int mgx_index_get_statistics(const mgx_index_t* index, mgx_index_statistics_t* out);
/* Releases the build workspace the index keeps between (re)builds (about 25 bytes per n-gram occurrence of the
 * last build: the double-buffered (key, doc) pairs and the sort scratch). The next build allocates it again. */
int mgx_index_trim(mgx_index_t* index);
int mgx_index_optimize(mgx_index_t* index, uint64_t total_docs);
int mgx_index_clear(mgx_index_t* index);

/* Index::PostingSize / Count (index.cpp:580-588): `term` is one n-gram. */
int mgx_index_posting_size(const mgx_index_t* index, const uint8_t* term, uint64_t term_len, uint64_t* out);
/* PostingList::GetAll (posting_list.cpp:421-430): ascending doc ids of one n-gram. */
int mgx_index_get_postings(const mgx_index_t* index, const uint8_t* term, uint64_t term_len, uint32_t* out,
                           uint64_t cap, uint64_t* out_count);
/* The payload the device index keeps next to every posting of one n-gram (key width <= 2; this is the library's own
 * extension of PostingList, used by the verified-df kernels -- search_pipeline.cpp:542-565 -- and exposed for tests and
 * diagnostics). docs[i]: LOCAL index of the document (rank of its id in the shard); first[i] / second[i]: bits 0..14 =
 * byte offset of the n-gram's first / second occurrence in the document's text (0x7FFF = not recorded), bit 15 =
 * a further occurrence exists, bits 16..22 / 24..30 = signature of the character after / before that occurrence
 * (0 = none; a signature of b bits is max(1, ((cp * 0x9E3779B1) >> 25) >> (7 - b))). layout[3] = {offset bits,
 * next-signature bits, prev-signature bits} of this shard (all 0 when the index carries no payload). out_count = the
 * list's length; at most cap entries are written; any output pointer may be NULL. */
int mgx_index_get_posting_payload(const mgx_index_t* index, const uint8_t* term, uint64_t term_len, uint32_t* docs,
                                  uint32_t* first, uint32_t* second, uint64_t cap, uint64_t* out_count, int* layout);
/* Whole index as CSR in ascending term (UTF-8 byte) order. keys are packed
 * n-grams (mgx_key_to_utf8 decodes them). Sizes come from mgx_index_get_stats:
 * keys[n_terms], offsets[n_terms+1], postings[n_postings] (global doc ids).
 * With n-gram sizes 4..10 (key_width > 3) keys[t] is t + 1, the rank of the term: the n-grams themselves come from
 * mgx_index_export_terms, which works for every width. */
int mgx_index_export(const mgx_index_t* index, uint64_t* keys, uint64_t* offsets, uint32_t* postings);
/* The terms of the index (Index::term_postings_ keys, index.h:343-359) as UTF-8 strings in ascending byte order:
 * term t = term_bytes[term_offsets[t] .. term_offsets[t+1]); term_offsets[n_terms+1]. *out_bytes receives the total
 * length; term_bytes == NULL only measures. At most 4 * key_width bytes per term. */
int mgx_index_export_terms(const mgx_index_t* index, uint8_t* term_bytes, uint64_t cap_bytes, uint64_t* term_offsets,
                           uint64_t* out_bytes);
/* Per-document code-point lengths (CountCodePoints, string_utils.cpp:655-669). out[n_docs]. */
int mgx_index_doc_lengths(const mgx_index_t* index, uint32_t* out);

/* Packed key -> UTF-8 n-gram. `out` needs 4*width bytes. Returns the byte length. width 1..3. */
int mgx_key_to_utf8(uint64_t key, int32_t width, uint8_t* out);
/* Words per key for a key width (max of the two n-gram sizes, config-schema.json:279-285 allows 1..10): 1 for
 * width <= 3, ceil(width / 3) otherwise (21 bits per code point, first code point in the most significant field of
 * word 0, keys compare word by word). */
int mgx_key_words(int32_t width);
/* The same for a key of mgx_key_words(width) words. `out` needs 4*width bytes. */
int mgx_wide_key_to_utf8(const uint64_t* words, int32_t width, uint8_t* out);

/* --------------------------------------------------------------- tokenizer */

/* GenerateHybridNgrams (utils/string_utils.cpp:452-509) for a batch of documents
 * on the GPU, in generation order (before the per-document sort+unique of
 * index.cpp:88-91). Output: packed key + index of the document in the batch.
 * out_keys/out_doc need room for one entry per code point; *out_count receives
 * the number of n-grams. n-gram sizes 1..10 (string_utils.cpp:382-423 has no upper bound; the configuration schema
 * stops at 10): with a key width above 3 every n-gram takes mgx_key_words(width) consecutive words of out_keys. */
int mgx_tokenize_batch(const mgx_index_config_t* config, const uint8_t* text, const uint64_t* text_offsets,
                       uint64_t n_docs, uint64_t* out_keys, uint32_t* out_doc, uint64_t cap, uint64_t* out_count);

/* ------------------------------------------------- MGIX index stream (DUMP / SYNC side of the index) */

/* Header and sizes of an MGIX stream: what Index::SaveToStream writes in front of the term records
 * (index/index_serialization.cpp:111-160; kanji_ngram_size is the EFFECTIVE value the reference Index holds,
 * index.cpp:29-37). normalize_width is NUL-terminated ("keep", "narrow", "wide"). */
typedef struct {
  uint32_t version;          /* decode: 1..4 as found; encode always writes 4 */
  int32_t ngram_size;
  int32_t kanji_ngram_size;
  int32_t cross_boundary;
  int32_t normalize_nfkc;
  int32_t normalize_lower;
  char normalize_width[32];
  uint64_t n_terms;
  uint64_t n_postings;
  uint64_t term_bytes;
} mgx_mgix_info_t;

/* Index::SaveToStream (index_serialization.cpp:111-224) + PostingList::Serialize (posting_list.cpp:973-1021) of an
 * index given as CSR: term t = term_bytes[term_offsets[t] .. term_offsets[t+1]) with the ascending doc ids
 * postings[posting_offsets[t] .. posting_offsets[t+1]); info->n_terms terms. A list is written in the representation
 * the reference's PostingList would hold: Roaring (portable interchange format) above 4096 entries, or when
 * roaring_min_len > 0 (= roaring_threshold x the total_docs of the last Index::Optimize) and the list has at least
 * that many entries; fixed-width delta otherwise. Host-only (no device needed). *out_len = stream length;
 * MGX_ERR_CAPACITY when it exceeds cap (call with out = NULL to size the buffer). */
int mgx_mgix_encode(const mgx_mgix_info_t* info, const uint8_t* term_bytes, const uint64_t* term_offsets,
                    const uint64_t* posting_offsets, const uint32_t* postings, double roaring_min_len, uint8_t* out,
                    uint64_t cap, uint64_t* out_len);

/* Index::LoadFromStream (index_serialization.cpp:279-613) + PostingList::Deserialize (posting_list.cpp:1023-1102):
 * validates the stream exactly as the reference does (magic, version 1..4, CRC32 trailer, header, size guards,
 * delta validity, Roaring structure incl. run containers) and returns the index as CSR in ascending term order.
 * First call with NULL outputs fills *info (header + sizes); the second, with info's sizes as the capacities
 * (term_offsets / posting_offsets hold n_terms + 1 entries), fills the arrays. Rejected streams: MGX_ERR_FORMAT.
 * The configuration check of LoadFromData (:371-447) is the caller's: compare *info with the target index. */
int mgx_mgix_decode(const uint8_t* data, uint64_t len, mgx_mgix_info_t* info, uint8_t* term_bytes,
                    uint64_t* term_offsets, uint64_t* posting_offsets, uint32_t* postings);

/* Index::SaveToStream of the device index: postings are read back once and encoded as above, with the
 * representation rule fed by the last mgx_index_optimize call. */
int mgx_index_save_mgix(const mgx_index_t* index, int32_t normalize_nfkc, const char* normalize_width,
                        int32_t normalize_lower, uint8_t* out, uint64_t cap, uint64_t* out_len);

/* Index::LoadFromStream (index_serialization.cpp:279-613) into the device index: the stream is validated and decoded
 * as mgx_mgix_decode does, its n-gram configuration must equal the index's (LoadFromData :371-447, else
 * MGX_ERR_FORMAT), and the posting lists replace the index content. A stream carries posting lists only, so the shard
 * then holds NO document text -- exactly the state of the reference's Index after LoadFromStream with an empty
 * DocumentStore: Index::Search* / FilterByNgrams / SearchByThreshold / GetStatistics answer from the lists, documents
 * are the ids that occur in them, text-dependent paths (verify_text, _score, substring terms) see documents without
 * stored text, BM25 statistics are zero. Single-document mutations are refused (MGX_ERR_UNSUPPORTED) until the
 * documents arrive through mgx_index_build, which recomputes the same lists from the texts. */
int mgx_index_load_mgix(mgx_index_t* index, const uint8_t* data, uint64_t len);

/* ------------------------------------------------- Index set-algebra calls */

/* Index::SearchAnd(terms, limit, reverse) — index.cpp:199-368. `terms` are
 * n-gram strings. Unknown term or empty list => empty. *out_count = result
 * size after limit; MGX_ERR_CAPACITY if it exceeds cap. */
int mgx_search_and(const mgx_index_t* index, const uint8_t* term_bytes, const uint64_t* term_offsets,
                   uint64_t n_terms, uint64_t limit, int32_t reverse, uint32_t* out, uint64_t cap,
                   uint64_t* out_count);
/* Index::SearchOr — index.cpp:418-448 (unknown terms ignored). */
int mgx_search_or(const mgx_index_t* index, const uint8_t* term_bytes, const uint64_t* term_offsets,
                  uint64_t n_terms, uint32_t* out, uint64_t cap, uint64_t* out_count);
/* Index::SearchNot(all_docs, terms) — index.cpp:450-486. all_docs ascending. */
int mgx_search_not(const mgx_index_t* index, const uint32_t* all_docs, uint64_t n_all, const uint8_t* term_bytes,
                   const uint64_t* term_offsets, uint64_t n_terms, uint32_t* out, uint64_t cap, uint64_t* out_count);
/* Index::FilterByNgrams(candidates, terms) — index.cpp:370-416: keeps caller
 * order and duplicates; no terms => candidates; unknown term => empty. */
int mgx_filter_by_ngrams(const mgx_index_t* index, const uint32_t* candidates, uint64_t n_candidates,
                         const uint8_t* term_bytes, const uint64_t* term_offsets, uint64_t n_terms, uint32_t* out,
                         uint64_t cap, uint64_t* out_count);

/* Index::SearchByThreshold(terms, threshold) — index.cpp:488-578: documents present in at least `threshold` of
 * the (de-duplicated) n-gram lists, ascending. threshold == 0 or > #distinct terms => empty; == #distinct =>
 * SearchAnd; terms without a posting list do not count. */
int mgx_search_by_threshold(const mgx_index_t* index, const uint8_t* term_bytes, const uint64_t* term_offsets,
                            uint64_t n_terms, uint64_t threshold, uint32_t* out, uint64_t cap, uint64_t* out_count);

/* QueryNode::Evaluate(index, doc_store, all_docs) — query/query_ast.cpp:67-161, for a boolean expression given
 * as a postfix program over `terms` (search TERMS, not n-grams; they are tokenised with the index's own n-gram
 * configuration exactly as Evaluate does, :80-84):
 *   op 0 TERM  arg = index into the term table   -> SearchAnd(n-grams); no n-grams -> substring scan of the texts
 *   op 1 AND   arg = number of children          (0 children -> empty)
 *   op 2 OR    arg = number of children          (0 children -> empty)
 *   op 3 NOT   one child                         -> all documents of the index minus the child
 *   op 4 ATLEAST arg = children | (t << 16)      -> documents of at least t of the children (t = 0 -> empty);
 *              an extension with no AST counterpart: Index::SearchByThreshold (index.cpp:488-578) as a node
 * Output: the ascending doc ids of the expression. all_docs is the set of documents given to mgx_index_build
 * (DocumentStore::GetAllDocIds). */
int mgx_eval_boolean(const mgx_index_t* index, const int32_t* ops, const int32_t* args, uint64_t n_ops,
                     const uint8_t* term_bytes, const uint64_t* term_offsets, uint64_t n_terms, uint32_t* out,
                     uint64_t cap, uint64_t* out_count);

/* ------------------------------------------------------- batched pipeline */

/* One batch of SEARCH queries through the regular path of
 * search_pipeline::ExecuteFullPipeline (server/search_pipeline.cpp:1757-2059):
 * GenerateTermInfos (:569-603) with verified document frequencies (:542-565),
 * terms ordered by estimated size (:2012-2014), Execute (:795-869: early exit,
 * AND, ApplyNotFilter :871-932), then — for SORT _score —
 * BM25Scorer::ScoreDocuments (index/bm25_scorer.cpp:47-99) and
 * ResultSorter::SortByScore (query/result_sorter.cpp:661-716) exactly as
 * SearchHandler::HandleSearch chains them (handlers/search_handler.cpp:405-470). */
typedef struct {
  int32_t ngram_size;        /* RAW table config values handed to GenerateQueryNgrams   */
  int32_t kanji_ngram_size;  /* (search_pipeline.cpp:578); 0 is meaningful there         */
  int32_t cross_boundary;
  int32_t compute_score;     /* 1: SORT _score; 0: ascending doc ids                      */
  int32_t descending;        /* SortOrder::DESC (default) / ASC for _score                */
  uint32_t limit;            /* query_parser.h:217 default 100; any offset; 0 = everything from offset on (single shard) */
  uint32_t offset;
  int32_t verify_text;       /* 0 = "off" (config.h:329): n-gram AND incl. false positives; */
                             /* 1 = "all": PostFilterByText (search_pipeline.cpp:1239-1246) */
  double k1;                 /* BM25Params, bm25_scorer.h:23-26 (1.2)                     */
  double b;                  /* (0.75)                                                    */
  /* Corpus statistics for scoring. 0/0 => this index's own BM25Stats. A sharded
   * deployment passes the GLOBAL totals so every shard scores identically. */
  uint64_t total_docs;
  uint64_t total_doc_length;
} mgx_query_params_t;

/* Queries are ranges into flat term tables:
 *   search terms of query q: terms [q_term_begin[q], q_term_begin[q+1])
 *   NOT terms of query q:    not-terms [q_not_begin[q], q_not_begin[q+1])  (q_not_begin may be NULL)
 * Outputs (host): for query q, out_count[q] ids at out_ids[q*stride..] (+ scores
 * at out_scores[q*stride..] when compute_score), out_total[q] = size of the full
 * result set (SearchHandler's total_results), out_df[t] = verified document
 * frequency of search term t (may be NULL). stride >= min(limit, ...) entries. */
int mgx_query_batch(mgx_index_t* index, const mgx_query_params_t* params, uint64_t n_queries,
                    const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                    const uint8_t* not_bytes, const uint64_t* not_offsets, const uint64_t* q_not_begin,
                    uint64_t stride, uint32_t* out_ids, double* out_scores, uint32_t* out_count, uint64_t* out_total,
                    uint64_t* out_df);

/* ----------------------------------------------- column filters + boolean programs in a batch */

/* Device mirror of one DocumentStore filter column (storage/document_store.h:73-92, filter_index.h:38-123):
 * `type` is the FilterValue variant index (1 bool, 2 int8, 3 uint8, 4 int16, 5 uint16, 6 int32, 7 uint32, 8 int64,
 * 9 uint64, 10 TIME seconds, 11 string, 12 double); values[i] belongs to the i-th document of the last build
 * (integer / bool / seconds value, the bits of the double, or for strings an index into the string table
 * str_bytes / str_offsets[n_strings + 1]); nulls[i] != 0 marks NULL (may be NULL = no NULLs). `column` is a
 * caller-chosen id < 64 (the host keeps DocumentStore::ResolveFilterColumnName). Columns follow the documents
 * through committed mutations (add / update / remove): a surviving document keeps its values, a document the journal
 * added or replaced is NULL in every column until the column is set again (the reference's FilterIndex is fed by the
 * same DocumentStore calls, storage/filter_index.h:38-123). A bulk build (mgx_index_build*, mgx_index_clear) starts
 * a new corpus and drops them. A filter that names a column the index does not hold: only != holds
 * (no stored value, as for a document the reference has no value for). */
int mgx_index_set_filter_column(mgx_index_t* index, uint32_t column, int32_t type, const uint64_t* values,
                                const uint8_t* nulls, uint64_t n_docs, const uint8_t* str_bytes,
                                const uint64_t* str_offsets, uint64_t n_strings);

/* Optional per-query extensions of mgx_query_batch:
 *  - boolean programs (QueryNode::Evaluate, query_ast.cpp:67-161): when q_prog_begin[q+1] > q_prog_begin[q] the
 *    query's terms [q_term_begin[q], q_term_begin[q+1]) are the TERM operands of the postfix program
 *    ops/args[q_prog_begin[q] ..) (TERM arg = index inside the query's own term range) instead of being AND-ed.
 *    With compute_score != 0 the results of a program are scored the way the reference scores every result shape
 *    (search_handler.cpp:405-470): BM25 over the TERM operands that are not below a NOT, left to right, duplicates
 *    kept (CollectAstScoringTerms, search_pipeline.cpp:232-254), each with its verified document frequency, then
 *    SortByScore; a scoring term that does not occur in a result's text adds nothing;
 *  - filter conditions (ApplyFiltersWithBitmap / ApplyFilters, search_pipeline.cpp:1098-1237): query q keeps a
 *    document only if it passes filters [q_filter_begin[q], q_filter_begin[q+1]): column id, op (query_parser.h:
 *    93-100: 0 EQ, 1 NE, 2 GT, 3 GTE, 4 LT, 5 LTE) and the literal as written in the query. As in the reference,
 *    a query whose conditions are all EQ / NE follows the FilterIndex bitmap semantics (exact match with any type
 *    interpretation of the literal), any other mix the typed per-document comparison. */
typedef struct {
  const int32_t* prog_ops;
  const int32_t* prog_args;
  const uint64_t* q_prog_begin;    /* [n_queries + 1] or NULL */
  const uint32_t* filter_col;
  const uint8_t* filter_op;
  const uint8_t* filter_bytes;
  const uint64_t* filter_offsets;  /* [n_filters + 1] */
  const uint64_t* q_filter_begin;  /* [n_queries + 1] or NULL */
} mgx_query_ext_t;

int mgx_query_batch_ex(mgx_index_t* index, const mgx_query_params_t* params, uint64_t n_queries,
                       const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                       const uint8_t* not_bytes, const uint64_t* not_offsets, const uint64_t* q_not_begin,
                       const mgx_query_ext_t* ext, uint64_t stride, uint32_t* out_ids, double* out_scores,
                       uint32_t* out_count, uint64_t* out_total, uint64_t* out_df);

/* ------------------------------------------------- fuzzy and synonym execution paths (single query) */

/* What search_pipeline::ExecuteWithFuzzy / ExecuteWithSynonyms read besides the terms: the RAW table n-gram
 * configuration (search_pipeline.cpp:578), memory.verify_text (0 "off", 1 "all", 2 "ascii"), the query's NOT terms
 * (ApplyNotFilter :871-932 — on the synonym path pass every synonym of every NOT term, their union is what is
 * excluded) and its column conditions (ApplyFiltersWithBitmap :1196-1237; columns as set by
 * mgx_index_set_filter_column, ops as in mgx_query_ext_t). */
typedef struct {
  int32_t ngram_size;
  int32_t kanji_ngram_size;
  int32_t cross_boundary;
  int32_t verify_text;
  const uint8_t* not_bytes;
  const uint64_t* not_offsets;     /* [n_not + 1] */
  uint64_t n_not;
  const uint32_t* filter_col;
  const uint8_t* filter_op;
  const uint8_t* filter_bytes;
  const uint64_t* filter_offsets;  /* [n_filters + 1] */
  uint64_t n_filters;
} mgx_expanded_query_t;

/* search_pipeline::ExecuteWithFuzzy (server/search_pipeline.cpp:1659-1740): per (normalised) search term the
 * documents holding at least max(1, |ngrams| - max_distance * effective_ngram_size) of its n-grams
 * (Index::SearchByThreshold), AND-ed over the terms; then NOT terms, column conditions and — when a term has a
 * hybrid fragment no n-gram covers — the exact text match of every term (:1728-1737). When verify_text applies to
 * the terms, PostFilterByFuzzyText (:1742-1752) runs on the device as well: a document stays if every term occurs
 * in its text or a whitespace-delimited word of the text is within max_distance edits of it (ContainsFuzzyMatch,
 * utils/edit_distance.cpp; terms of at most 64 code points, else MGX_ERR_UNSUPPORTED). A term too short for an
 * n-gram, or no term at all, gives the empty result (empty_term_detected). Output: ascending doc ids. */
int mgx_search_fuzzy(const mgx_index_t* index, const mgx_expanded_query_t* query, const uint8_t* term_bytes,
                     const uint64_t* term_offsets, uint64_t n_terms, uint32_t max_distance, uint32_t* out,
                     uint64_t cap, uint64_t* out_count);

/* search_pipeline::ExecuteWithSynonyms (server/search_pipeline.cpp:1580-1631) over already expanded groups
 * (ExpandTermsWithSynonyms :1392-1406 stays on the host: group g = the normalised term and its synonyms =
 * variants [group_begin[g], group_begin[g+1])): OR of SearchTermDocuments within a group (n-gram AND, or the
 * substring test for a variant shorter than an n-gram), AND across groups, NOT terms, column conditions, and with
 * verify_text the synonym-aware text check of PostFilterByTextWithSynonyms (:1633-1657). No group => empty. */
int mgx_search_synonyms(const mgx_index_t* index, const mgx_expanded_query_t* query, const uint8_t* variant_bytes,
                        const uint64_t* variant_offsets, const uint64_t* group_begin, uint64_t n_groups,
                        uint32_t* out, uint64_t cap, uint64_t* out_count);

/* Batch forms: n_queries fuzzy (synonym) queries that share `query` (NOT terms, filters, verification mode, n-gram
 * sizes) and, for the fuzzy form, max_distance, answered in ONE batch on the device -- one upload, one planning
 * pass, the queries' tiles side by side, one download -- instead of one small batch per call. q_term_begin[q .. q+1)
 * are query q's terms in term_offsets (q_group_begin[q .. q+1) its groups in group_begin, whose entries index
 * variant_offsets globally). Output: the queries' ascending doc ids back to back in `out`, out_offsets[q .. q+1) =
 * query q (out_offsets has n_queries + 1 entries). Every answer equals the single call's. MGX_ERR_CAPACITY when the
 * sum exceeds cap: out_offsets then holds the sizes. */
int mgx_search_fuzzy_batch(const mgx_index_t* index, const mgx_expanded_query_t* query, uint64_t n_queries,
                           const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                           uint32_t max_distance, uint32_t* out, uint64_t cap, uint64_t* out_offsets);
int mgx_search_synonyms_batch(const mgx_index_t* index, const mgx_expanded_query_t* query, uint64_t n_queries,
                              const uint8_t* variant_bytes, const uint64_t* variant_offsets, const uint64_t* group_begin,
                              const uint64_t* q_group_begin, uint32_t* out, uint64_t cap, uint64_t* out_offsets);

/* Timing / accounting of one batch (device times from CUDA events recorded on the launch
 * stream around the named kernels; bytes as defined in SURVEY.md §8(d) and DESIGN.md). */
typedef struct {
  double ms_plan;             /* lookup + term/query planning kernels (+ the size read-back)      */
  double ms_df_kernel;        /* df_tile_kernel launches                                          */
  double ms_and_kernel;       /* and_tile_kernel launches (intersection + fused BM25 epilogue)    */
  double ms_topk_kernel;      /* topk_kernel launches                                             */
  double ms_total;            /* first event .. last event of the batch                           */
  uint64_t launches;          /* kernels launched for the batch                                   */
  uint64_t n_df_tiles;        /* CTAs of df_tile_kernel                                           */
  uint64_t n_and_tiles;       /* CTAs of and_tile_kernel                                          */
  uint64_t algo_bytes_intersect; /* B_intersect = sum_q [ sum_lists min(4|P|, ceil(N/8)) + 4|R| ]  */
  uint64_t algo_bytes_score;     /* B_score = sum_q [ sum_{d in R}(text_bytes(d) + 4) + 12 k ]      */
  uint64_t algo_bytes_df;        /* B_df = sum over scanned terms, sum_{d in C_t} text_bytes(d)    */
  uint64_t algo_bytes_df_lists;  /* min(4|P|, ceil(N/8)) over the lists of the scanned terms       */
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
  uint64_t driver_entries;    /* posting entries walked by and_tile_kernel                        */
  uint64_t result_docs;       /* sum |R|                                                          */
  uint64_t df_candidates;     /* documents whose text was scanned for df                          */
  uint64_t unique_terms;
  double ms_df_stream_kernel; /* df_stream_kernel launch (one pass over the text arena), 0 if not chosen */
  uint64_t df_stream_terms;   /* terms whose df came from the streaming pass                          */
  uint64_t df_stream_bytes;   /* text bytes that pass read (= the shard's arena), its algorithmic bytes */
  uint64_t df_stream_hits;    /* (term, document) pairs it counted                                    */
  uint64_t df_scanned_docs;   /* df candidates whose whole text was scanned; the others were decided by one */
                              /* comparison at the recorded first occurrence of the driver n-gram           */
} mgx_batch_stats_t;

/* Staged form of the same call, for doc-range sharded deployments (one process
 * per GPU): the two exchange points of SURVEY.md §8(e) sit between the stages.
 *   prepare : host compile + H2D of the compiled batch
 *   plan    : dictionary lookup + per-term / per-query planning (device; one size read-back)
 *   df      : per-shard verified document frequencies -> d_df (DEVICE, n_terms uint64)
 *             [caller all-reduces d_df (SUM) across shards]
 *   search  : AND + NOT + BM25 + top-k with the GLOBAL df -> DEVICE outputs
 *             [caller all-gathers the per-shard top-k records]
 *   merge   : mgx_merge_topk_device over the gathered runs
 * All stages are enqueued on `stream` (a cudaStream_t passed as void*; NULL =
 * the legacy default stream) and do not synchronise the host except where a
 * size has to be read back (documented in DESIGN.md). */
typedef struct mgx_batch mgx_batch_t;

int mgx_batch_prepare(mgx_index_t* index, const mgx_query_params_t* params, uint64_t n_queries,
                      const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                      const uint8_t* not_bytes, const uint64_t* not_offsets, const uint64_t* q_not_begin,
                      void* stream, mgx_batch_t** out);
/* Same with the optional per-query extensions of mgx_query_batch_ex (boolean programs, column filters). */
int mgx_batch_prepare_ex(mgx_index_t* index, const mgx_query_params_t* params, uint64_t n_queries,
                         const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* q_term_begin,
                         const uint8_t* not_bytes, const uint64_t* not_offsets, const uint64_t* q_not_begin,
                         const mgx_query_ext_t* ext, void* stream, mgx_batch_t** out);
/* Planning stage of a prepared batch (dictionary lookup, per-term and per-query plans). Called
 * implicitly by the df / search stages if it has not run yet. */
int mgx_batch_plan_device(mgx_batch_t* batch);
/* Streamed form of the staged calls: planning leaves the work sizes on the device and the df / search stages run as
 * persistent kernels that pull their work from device-side counters, so the whole batch is enqueued without a host
 * synchronisation. The price is a fixed-size per-stream workspace: a batch that does not fit computes nothing and
 * reports it (mgx_batch_overflowed, after the search stage; packed records carry the flag in their status block);
 * mgx_batch_reset returns it to its uploaded state, and the repeat runs in the synchronous form, whose workspace
 * grows with the batch. Call before the planning stage. mgx_query_batch* and mgx_sharded_batch_* do all of this
 * themselves. */
int mgx_batch_set_streamed(mgx_batch_t* batch, int32_t enable);
int mgx_batch_overflowed(mgx_batch_t* batch, int32_t* out_overflowed); /* waits for the batch's enqueued work */
int mgx_batch_reset(mgx_batch_t* batch);
/* Stats of a staged batch; synchronises the batch's stream. */
int mgx_batch_get_stats(mgx_batch_t* batch, mgx_batch_stats_t* out);
/* Number of search-term slots (= length of d_df / out_df). */
uint64_t mgx_batch_term_slots(const mgx_batch_t* batch);
int mgx_batch_df_device(mgx_batch_t* batch, uint64_t* d_df);
int mgx_batch_search_device(mgx_batch_t* batch, const uint64_t* d_df, uint64_t stride, uint32_t* d_ids,
                            double* d_scores, uint32_t* d_count, uint64_t* d_total);
void mgx_batch_destroy(mgx_batch_t* batch);

/* Merge of per-shard top-k runs (already in final order within each shard):
 * inputs are [n_shards][n_queries][stride] DEVICE arrays as produced by an
 * all-gather of mgx_batch_search_device outputs. Order = SortByScore's
 * comparator (result_sorter.cpp:681-686). For compute_score == 0 the runs are
 * ascending id lists of disjoint ascending ranges and are concatenated. */
int mgx_merge_topk_device(int32_t device, void* stream, const mgx_query_params_t* params, uint32_t n_shards,
                          uint64_t n_queries, uint64_t stride, const uint32_t* d_ids_all, const double* d_scores_all,
                          const uint32_t* d_count_all, const uint64_t* d_total_all, uint32_t* d_ids_out,
                          double* d_scores_out, uint32_t* d_count_out, uint64_t* d_total_out);

/* Packed form of the same exchange: ONE buffer per shard, so that the per-shard top-k runs of a batch travel in
 * a single all-gather (BASELINE north_star: "merged with a single NCCL all-gather over NVLink"). A record holds
 * [scores f64 Q*S][total u64 Q][ids u32 Q*S][count u32 Q], padded to 16 bytes; mgx_shard_record_layout gives the
 * byte offsets. d_records = [n_shards][layout.bytes] as gathered; d_record_out = one record (the merged answer,
 * one device-to-host copy away from the caller). The record ends with a 16-byte status block (uint32[4]): word 0 is
 * non-zero when the shard's streamed batch did not fit its workspace (see mgx_batch_set_streamed); the merged
 * record carries the OR over the shards. stride must be >= limit + offset: a shard returns its best limit + offset
 * records un-offset and the merge skips the offset (MGX_ERR_INVALID_ARGUMENT otherwise). */
typedef struct mgx_shard_record_layout {
  uint64_t scores_offset;
  uint64_t total_offset;
  uint64_t ids_offset;
  uint64_t count_offset;
  uint64_t status_offset;
  uint64_t bytes;
} mgx_shard_record_layout_t;
int mgx_shard_record_layout(uint64_t n_queries, uint64_t stride, mgx_shard_record_layout_t* out);
int mgx_batch_search_packed_device(mgx_batch_t* batch, const uint64_t* d_df, uint64_t stride, void* d_record);
int mgx_merge_topk_packed_device(int32_t device, void* stream, const mgx_query_params_t* params, uint32_t n_shards,
                                 uint64_t n_queries, uint64_t stride, const void* d_records, void* d_record_out);

/* ------------------------------------------------- sharded pipeline: one process per GPU, one doc-id range each
 *
 * The multi-GPU protocol of SURVEY.md §8(e) behind the C ABI, for a C++ host (the reference's server is C++): the
 * library issues the two NCCL exchanges itself, on a highest-priority stream per communicator lane, ordered with
 * the batch's own stream by events -- so a batch is ONE uninterrupted enqueue and several batches can be in flight
 * (one per lane) without an exchange queueing behind another batch's grid-filling kernels.
 *   enqueue: plan -> df -> ncclAllReduce(SUM) of the per-term df -> search (GLOBAL df) -> ncclAllGather of the packed
 *            per-shard records -> merge -> optional device-to-host copy of the merged record
 *   finish : waits for the batch; if ANY shard's streamed batch did not fit its workspace (status block of the
 *            merged record, identical on every rank) all ranks repeat it together in the synchronous form.
 * Every rank must prepare the SAME batch (same queries, same order) with the GLOBAL corpus statistics in
 * mgx_query_params_t. NCCL is bound at run time (dlopen of libnccl.so.2, MGX_NCCL_LIB overrides); without it the
 * calls below fail with MGX_ERR_UNSUPPORTED and comm == NULL (a single shard, no exchange) still works.
 * Bootstrap: rank 0 calls mgx_comm_unique_id once per lane and hands the ids to every rank by any means (the
 * benchmarks hand them round with the host framework's own broadcast); every rank then calls mgx_comm_create with the same ids. */
#define MGX_COMM_ID_BYTES 128
#define MGX_COMM_MAX_LANES 4
typedef struct mgx_shard_comm mgx_shard_comm_t;
int mgx_comm_unique_id(uint8_t* out_id /* MGX_COMM_ID_BYTES */);
int mgx_comm_create(const uint8_t* ids /* n_lanes x MGX_COMM_ID_BYTES */, int32_t n_lanes, int32_t n_ranks,
                    int32_t rank, int32_t device, mgx_shard_comm_t** out);
void mgx_comm_destroy(mgx_shard_comm_t* comm);
int mgx_comm_info(const mgx_shard_comm_t* comm, int32_t* n_ranks, int32_t* rank, int32_t* n_lanes,
                  int32_t* nccl_version);
/* batch: freshly prepared (mgx_batch_prepare*), on the stream given there. h_record_out (may be NULL): pinned host
 * buffer of mgx_shard_record_layout(n_queries, stride).bytes that receives the merged record. */
int mgx_sharded_batch_enqueue(mgx_shard_comm_t* comm, int32_t lane, mgx_batch_t* batch, uint64_t stride,
                              void* h_record_out);
/* *d_record_out (may be NULL) = the merged record on the device, valid until the batch is destroyed;
 * *out_repeated (may be NULL) = 1 when the batch had to be repeated. Same comm / lane / stride / h_record_out as
 * the enqueue. */
int mgx_sharded_batch_finish(mgx_shard_comm_t* comm, int32_t lane, mgx_batch_t* batch, uint64_t stride,
                             void* h_record_out, const void** d_record_out, int32_t* out_repeated);

/* --------------------------------------- compiled batches shared between the shard processes of one node
 *
 * Every shard answers every query, so without this every rank compiles every batch (the per-request host work of
 * QueryParser / GenerateQueryNgrams, server/search_pipeline.cpp:569-603, repeated N times per node). The channel is a
 * POSIX shared-memory ring of n_slots slots of slot_bytes payload each: ONE rank compiles batch `seq` with
 * mgx_batch_prepare* and publishes it, every other rank imports it (a copy into its own pinned staging buffer + the
 * same single H2D transfer a local compile ends with) -- the sequence numbers, not the call order, pair the two sides.
 * rank 0 creates the segment (name as for shm_open, e.g. "/mgx_<port>"), the others attach. A batch with column
 * conditions or an explicit driver set, or one larger than a slot, is refused by publish AND by every import of that
 * seq (same status): the ranks then compile it locally. timeout_ms < 0 waits for ever, otherwise MGX_ERR_TIMEOUT. The importing index must have
 * the compiling index's n-gram configuration; params carries the same GLOBAL corpus statistics on every rank. */
typedef struct mgx_share mgx_share_t;
int mgx_share_open(const char* name, int32_t n_ranks, int32_t rank, int32_t n_slots, uint64_t slot_bytes,
                   mgx_share_t** out);
void mgx_share_close(mgx_share_t* share);
int mgx_share_publish(mgx_share_t* share, uint64_t seq, const mgx_batch_t* batch, int32_t timeout_ms);
int mgx_share_import(mgx_share_t* share, uint64_t seq, mgx_index_t* index, const mgx_query_params_t* params,
                     void* stream, int32_t timeout_ms, mgx_batch_t** out);

/* Stats of the last mgx_query_batch on this index. */
int mgx_index_last_batch_stats(const mgx_index_t* index, mgx_batch_stats_t* out);

/* ---------------------------------------------------------------- scoring */

/* BM25Scorer::ScoreDocuments (index/bm25_scorer.cpp:47-99) for explicit
 * candidates, terms and document frequencies: one score per candidate in
 * candidate order, 0.0 for a candidate without stored text. */
int mgx_score_documents(mgx_index_t* index, const uint32_t* candidates, uint64_t n_candidates,
                        const uint8_t* term_bytes, const uint64_t* term_offsets, const uint64_t* term_doc_freqs,
                        uint64_t n_terms, uint64_t total_docs, double avg_doc_length, double k1, double b,
                        double* out_scores);
/* ResultSorter::SortByScore (query/result_sorter.cpp:661-716): any offset; limit 0 = everything from offset on
 * (:689-710). out holds min(limit, n) entries, n when limit is 0. */
int mgx_sort_by_score(mgx_index_t* index, const uint32_t* results, const double* scores, uint64_t n,
                      int32_t descending, uint32_t limit, uint32_t offset, uint32_t* out, uint64_t* out_count);

#ifdef __cplusplus
}
#endif
#endif /* MGX_H_ */
