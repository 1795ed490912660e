// query.cuh — batch object shared by query.cu (kernels) and api.cu (C ABI).
#pragma once

#include <string>
#include <unordered_map>
#include <vector>

#include "mgx_internal.cuh"

namespace mgx {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint64_t kEstNone = ~0ULL;   // "term produced no n-grams" (SIZE_MAX in search_pipeline.cpp:583)
constexpr int kTile = 1024;            // driver entries per CTA
constexpr int kTileThreads = 256;
constexpr int kTileItems = kTile / kTileThreads;
constexpr uint32_t kMaxTermBytes = 256;
constexpr uint32_t kMaxTopK = 1024;

// ---- streamed batches (no host read-back between planning and the tile kernels) --------------------------------
// The planning tail leaves the work sizes in a small device block; persistent kernels pull their work items from the
// counters next to them. The block follows the accounting counters, so one memset clears both.
#ifndef MGX_DF_UNIT
#define MGX_DF_UNIT 2048
#endif
constexpr uint32_t kDfUnit = MGX_DF_UNIT;          // driver entries per df work unit of a streamed batch (kTile in the other form)
enum LaunchSlot : int {
  kLaunchDfUnits = 0,    // df work units of the batch
  kLaunchAndTiles = 1,   // intersect tiles of the batch
  kLaunchGroups = 2,     // top-k pre-reduction groups
  kLaunchOverflow = 3,   // != 0: the batch does not fit the stream's workspace; every later kernel does nothing
  kLaunchDfNext = 4,     // work counters of the persistent kernels
  kLaunchAndNext = 5,
  kLaunchGroupNext = 6,
  kLaunchTicket = 7,     // blocks of the planning kernel that have finished
  kLaunchNeedTiles = 8,  // what the batch would have needed (reported with an overflow)
  kLaunchCount = 16
};

// ---- streaming document-frequency pass (df_stream_kernel) ------------------------------------------------
// For very large batches the verified document frequencies of the multi-n-gram terms are computed by ONE pass
// over the shard's text arena that matches all of them at once, instead of visiting every candidate document
// per term. Two levels, both built on the host per batch:
//   stage 1  a 128 Ki-bit filter in shared memory over the first min(len, 8) bytes of every eligible term,
//            probed at every character start of the text;
//   stage 2  a bucket table over the first min(len, 12) bytes, probed only at the positions that passed stage 1,
//            followed by the comparison of the remaining bytes, the attribution to a document and the
//            per-(term, document) de-duplication.
constexpr uint32_t kStreamKeyBytes = 12;         // bytes of a term held in a table entry
constexpr uint32_t kStreamFilterBytes = 8;       // bytes hashed by the stage-1 filter
constexpr uint32_t kStreamMinTermBytes = 3;
__host__ __device__ inline uint32_t stream_hash(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t len) {
  uint32_t h = w0 * 0x9E3779B1u ^ (w1 * 0x85EBCA77u) ^ (w2 * 0xC2B2AE3Du) ^ (len * 0x27D4EB2Fu);
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 13;
  return h;
}
constexpr uint32_t kStreamBloomWords = 4096;     // 128 Ki bits
__host__ __device__ inline uint32_t stream_filter_bit(uint32_t w0, uint32_t w1, uint32_t len8) {
  uint32_t h = w0 * 0x85EBCA6Bu ^ (w1 * 0xC2B2AE35u) ^ (len8 * 0x165667B1u);
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  return h & (kStreamBloomWords * 32 - 1);
}
// bytes [0, n) of a little-endian word, n >= 4 keeps everything
__host__ __device__ inline uint32_t low_bytes_mask(uint32_t n) { return n >= 4 ? 0xFFFFFFFFu : ((1u << (8 * n)) - 1u); }
// One eligible term in bucket order: its first min(len, 12) bytes (zero padded), that length and its term id.
struct StreamEntry {
  uint32_t w0, w1, w2;
  uint32_t len_term;  // key bytes (low 8 bits) | unique term id << 8
};
constexpr uint32_t kStreamMaxTerms = 1u << 19;   // entry index field of a (term, document) record
constexpr uint32_t kStreamMaxBucket = 255;       // entries per bucket (count field of a slot)

// query flag bits (q_flags)
constexpr uint32_t kQEmpty = 1u;        // early exit / no terms: result is empty
constexpr uint32_t kQVerify = 2u;       // every search term must also occur in the text (verify_text / hybrid fragment)
constexpr uint32_t kQAnyMode = 4u;      // OR semantics over the lists (Index::SearchOr)
constexpr uint32_t kQDriverAll = 8u;    // driver = every document of the shard
constexpr uint32_t kQDriverExplicit = 16u;  // driver = caller supplied candidate ids (query 0 only)
constexpr uint32_t kQProgram = 32u;     // membership = a boolean postfix program over terms (QueryNode::Evaluate)
// kQProgram whose text leaves (substring-only terms, FUZZYTEXT) all occur positively: the program is monotone in them,
// so a first evaluation with those leaves taken as true rules documents out without reading their text, and only the
// documents that pass it pay for the real one
constexpr uint32_t kQOptimistic = 64u;
constexpr uint32_t kMaxProgramDepth = 64;  // evaluation stack = one 64-bit word

// postfix program ops (same encoding as the oracle's orc_eval_boolean)
constexpr uint8_t kOpTerm = 0;  // arg = unique term id
constexpr uint8_t kOpAnd = 1;   // arg = number of children
constexpr uint8_t kOpOr = 2;    // arg = number of children
constexpr uint8_t kOpNot = 3;   // one child
// at least t of n children (Index::SearchByThreshold as a node: the per-term candidate rule of
// search_pipeline::ExecuteWithFuzzy, search_pipeline.cpp:1696-1702): arg = n | (t << 16)
constexpr uint8_t kOpAtLeast = 4;
// the stored text holds a whitespace-delimited word (or, in a word with non-ASCII characters, a window of it) within
// edit distance d of the term (utils/edit_distance.cpp ContainsFuzzyMatch; PostFilterByFuzzyText,
// search_pipeline.cpp:1742-1752): arg = unique term id | (d << 24)
constexpr uint8_t kOpFuzzyText = 5;
constexpr uint32_t kFuzzyMaxTermCps = 64;  // code points of a term the edit-distance rows are sized for

// driver kinds for single-call set APIs
struct ExplicitDriver {
  const uint32_t* d_ids = nullptr;  // global doc ids (device)
  uint64_t n = 0;
};

// One filter condition resolved against the index's columns (device form, 32 bytes).
struct FilterPred {
  const uint64_t* values;
  const uint8_t* nulls;
  uint64_t c;         // literal in the column's class (rank for strings, bits for doubles)
  uint32_t cls;       // FilterClass; kFcNone = the index has no such column
  uint32_t op_flags;  // op (bits 0-2) | kFilterBitmapMode | kFilterValid | kFilterFound
};
constexpr uint32_t kFilterBitmapMode = 8;   // EQ / NE by FilterIndex semantics (every condition of the query is EQ / NE)
constexpr uint32_t kFilterValid = 16;       // the literal parses in the column's class
constexpr uint32_t kFilterFound = 32;       // string literal present in the column's dictionary

struct HostFilter {
  uint32_t col = 0;
  uint8_t op = 0;
  std::string literal;
};

// Posting list of one looked-up key, resolved once per batch (instead of key -> dictionary slot -> offsets / bitmap
// slot in every tile): 32 bytes, two 16-byte loads.
struct KeyRef {
  const uint32_t* p;   // sorted local doc indices (nullptr if the key is not in the dictionary)
  const uint32_t* bm;  // dense bitmap or nullptr
  uint32_t len;
  uint32_t pad[3];
};
// One df tile (one CTA of df_tile_kernel), resolved once per batch.
struct DfTileDesc {
  const uint32_t* drv_p;  // the term's shortest list
  uint32_t drv_len;
  uint32_t term;
  uint32_t k0;            // key range of the term
  uint32_t k1;
  uint32_t tile;          // tile index inside the term
  uint32_t pad;
};

// One compiled query term.
struct HostTerm {
  std::string bytes;
  KeyVec keys;             // sorted unique packed n-grams
  TermOffsetVec key_toff;  // per key: byte offset in the term if the n-gram occurs once in it (else kNoTermOffset)
  bool raw = false;            // keys given directly (Index::SearchAnd style): no text semantics
  bool exact_single = false;   // the term IS its single n-gram, so df == posting size when all text is valid UTF-8
  bool streamable = false;     // valid UTF-8 (>= 3 bytes) that needs a text check: may use the streaming df pass
  bool payload_tf = false;     // the term IS its single n-gram, valid UTF-8, and both tokenisers agree: its tf in a
                               // document follows from the occurrences recorded with the posting
  uint64_t hash = 0;           // hash of `bytes` (host query compiler)
};

// Host-built bucket table of the streamable terms.
struct HostStreamTable {
  std::vector<uint32_t> slots;
  std::vector<StreamEntry> entries;
  std::vector<uint32_t> bloom;
  uint32_t len8_mask = 0;
  uint32_t len12_mask = 0;
  // builder scratch, kept with the table so that a pooled batch re-uses the memory
  struct Item {
    uint32_t slot;
    StreamEntry e;
  };
  std::vector<Item> items;
  std::vector<uint32_t> count;
};

struct HostQuery {
  TermIdVec terms;                  // unique-term ids in query order
  TermIdVec not_terms;
  uint32_t flags = 0;               // kQVerify / kQAnyMode / kQDriverAll / kQDriverExplicit / kQProgram
  uint32_t threshold = 1;           // kQAnyMode: lists that must hold the document (Index::SearchByThreshold); 1 = OR
  std::vector<uint8_t> prog_ops;    // kQProgram: postfix program
  std::vector<uint32_t> prog_args;
  std::vector<uint32_t> conjuncts;  // kQProgram: terms every result must satisfy (driver candidates)
  std::vector<HostFilter> filters;  // column conditions, AND-ed
};

// Counts that fix the layout of a compiled batch in its staging buffer (every array padded to 256 bytes, in the
// order stage_offsets() lays them out). Plain data: it travels with the staging bytes when a batch compiled by one
// process is handed to others (mgx_share_*).
struct StageLayout {
  uint64_t n_terms = 0, n_queries = 0, n_out_queries = 0, n_slots = 0, n_bytes = 0, n_keys = 0, n_tids = 0, n_ntids = 0;
  uint64_t n_prog = 0, n_conj = 0, n_preds = 0, n_sslots = 0, n_sentries = 0, n_sbloom = 0, n_xoff = 0, list_cap = 0;
  uint32_t stream_len8_mask = 0, stream_len12_mask = 0;
  uint32_t assumed_all_valid_utf8 = 0;  // the compiling shard held no invalid UTF-8 (term shortcuts were decided with that)
  uint32_t wide_words = 0;  // words per wide key (index key width > 3): n_keys * wide_words words follow the other arrays
};
struct StageOffsets {
  size_t i_bytes, i_boff, i_koff, i_keys, i_raw, i_toff, i_tids, i_tids0, i_noff, i_ntids, i_loff, i_hflags, i_slot, i_ktoff,
      i_thr, i_poff, i_pops, i_pargs, i_coff, i_conj, i_foff, i_preds, i_sslots, i_sentries, i_sbloom, i_xoff, i_wide, total;
};
StageOffsets stage_offsets(const StageLayout& L);

struct Batch {
  Index* ix = nullptr;
  StageLayout layout;
  SearchScratch* sc = nullptr;  // the (index, stream) workspace; the d_tile_* / d_rec_* members below are views into it
  uint64_t serial = 0;          // unique per use of the batch object
  mgx_query_params_t params{};
  cudaStream_t stream = nullptr;
  bool owns_stream = false;

  // ---- host-side compiled form
  uint32_t n_queries = 0;    // queries the kernels run (an OR-rooted program of the caller is several of them)
  uint32_t n_out_queries = 0;  // queries of the caller = rows of the outputs
  uint32_t n_terms = 0;      // unique terms (search + not)
  uint32_t n_keys = 0;
  uint32_t n_slots = 0;      // caller's search-term slots (for df output)
  std::vector<uint32_t> h_slot_tid;  // slot -> unique term id
  // compile workspace of the staged / batched entry points: kept with the pooled batch object, so that a steady
  // stream of batches neither allocates nor touches fresh pages on the host
  std::vector<HostTerm> h_terms;
  std::vector<HostQuery> h_queries;
  std::vector<uint32_t> h_slot_scratch;
  HostStreamTable h_stream_table;

  // ---- device: terms
  DevBuf<uint8_t> d_term_bytes;
  DevBuf<uint32_t> d_term_boff;   // [T+1]
  DevBuf<uint32_t> d_term_koff;   // [T+1]
  DevBuf<uint64_t> d_keys;        // [K]
  DevBuf<uint64_t> d_wide;        // [K * wide_words] words of the wide keys (null for packed keys)
  DevBuf<uint32_t> d_key_list;    // [K] dictionary term index or kNone; sorted by length inside a term
  DevBuf<uint32_t> d_key_len;     // [K]
  DevBuf<uint32_t> d_key_toff;    // [K] byte offset of the n-gram inside its term, kNoTermOffset if unusable
  DevBuf<KeyRef> d_key_ref;       // [K] resolved lists, in the per-term order term_plan_kernel leaves
  DevBuf<uint64_t> d_t_est;       // [T]
  DevBuf<uint32_t> d_t_df_tiles;  // [T]
  DevBuf<uint64_t> d_t_df_tile_off;  // [T+1]
  DevBuf<uint64_t> d_t_df;        // [T]
  DevBuf<uint64_t> d_key_glen;    // [K] posting size per key in upload order; directly behind d_t_df (one exchange)
  bool global_order = false;      // sharded pipeline: order terms / compute IDFs from the exchanged global values
  DevBuf<uint32_t> d_slot_tid;    // [S]

  // ---- device: queries
  DevBuf<uint32_t> d_q_toff;      // [Q+1] search terms
  DevBuf<uint32_t> d_q_tids;      // [sum] unique term ids, re-ordered by estimated size by the planner
  DevBuf<uint32_t> d_q_tids0;     // [sum] the same in query order, never modified
  DevBuf<uint32_t> d_q_noff;      // [Q+1] NOT terms
  DevBuf<uint32_t> d_q_ntids;
  DevBuf<uint32_t> d_q_loff;      // [Q+1] capacity ranges for the merged list table
  DevBuf<uint32_t> d_q_list;      // dictionary term index per merged list (ascending length)
  DevBuf<uint32_t> d_q_list_len;
  DevBuf<uint32_t> d_q_nlists;    // [Q]
  DevBuf<uint32_t> d_q_flags;     // [Q]
  DevBuf<uint32_t> d_q_driver_len;  // [Q]
  DevBuf<uint32_t> d_q_ntiles;    // [Q]
  DevBuf<uint64_t> d_q_tile_off;  // [Q+1]
  DevBuf<uint64_t> d_q_group_off; // [Q+1] top-k pre-reduction groups per query (streamed batches)
  DevBuf<double> d_q_idf;         // [sum search terms], in planner order
  DevBuf<uint32_t> d_q_host_flags;  // [Q] flags decided on the host (verify etc.)
  DevBuf<uint32_t> d_q_threshold;   // [Q] any-mode threshold
  DevBuf<uint32_t> d_q_poff;        // [Q+1] program ranges
  DevBuf<uint8_t> d_prog_op;
  DevBuf<uint32_t> d_prog_arg;
  DevBuf<uint32_t> d_q_coff;        // [Q+1] conjunct ranges
  DevBuf<uint32_t> d_q_conj;
  DevBuf<uint32_t> d_q_foff;        // [Q+1] filter ranges
  DevBuf<FilterPred> d_filters;

  // ---- device: per-tile results
  DevBuf<uint32_t> d_tile_count;  // [tiles in flight] records written by the tile
  DevBuf<uint32_t> d_tile_total;  // [tiles in flight] survivors of the tile (>= records when pruned to top-k)
  DevBuf<uint32_t> d_rec_doc;     // survivors (global doc ids), tile k of query q at rec_off[q] + k*kTile
  DevBuf<double> d_rec_score;

  // host mirror read back after planning (not for streamed batches)
  std::vector<uint64_t> h_q_tile_off;
  // streamed batches: sizes and work counters stay on the device (LaunchSlot); h_launch receives them with the results
  DevBuf<uint32_t> d_launch;
  PinBuf<uint32_t> h_launch;
  bool streamed = false;       // planned without a host read-back
  bool allow_streamed = false; // set by the entry points that can repeat an overflowed batch
  uint32_t rec_slot = kTile;   // record slots per tile in the (index, stream) workspace
  uint8_t* in_base = nullptr;  // device copy of the compiled batch (for a repeat after an overflow)
  size_t in_total = 0;

  DevBuf<uint8_t> d_term_flags;   // [T] bit0 raw, bit1 exact_single, bit2 eligible for the streaming df pass
  DevBuf<uint32_t> d_stream_slots;      // [n_stream_slots] (first entry << 8) | entries in the bucket
  DevBuf<StreamEntry> d_stream_entries; // [n_stream_terms] in bucket order
  DevBuf<uint32_t> d_stream_bloom;      // [kStreamBloomWords] stage-1 filter
  uint32_t stream_len8_mask = 0;        // bit m set: some eligible term has min(len, 8) == m   (stage-1 classes)
  uint32_t stream_len12_mask = 0;       // bit m set: some eligible term has min(len, 12) == m  (stage-2 classes)
  uint32_t n_stream_slots = 0;          // power of two, 0 = no eligible term
  uint32_t n_stream_terms = 0;
  DevBuf<uint32_t> d_df_mode;           // [2] [0] = 1 when the streaming pass was chosen (device decision)
  // result sets of the Index::Search* style calls (batch_search_sets / merge_disjoint_runs), grow-only with the object
  DevArena sets_arena;
  DevBuf<uint32_t> sets_buf;
  DevBuf<uint32_t> union_buf;
  DevBuf<uint64_t> run_off_buf;
  DevBuf<uint32_t> driver_buf;  // explicit driver ids of SearchNot / FilterByNgrams
  DevBuf<uint64_t> len_buf;     // lookup_list_lengths
  DevBuf<unsigned long long> d_q_thr;   // [Q + 1] running thresholds of the per-tile top-k pruning, cleared with the counters
  int h_df_mode = 0;                    // host copy, valid after planning
  // Everything above is a view into one of these two grow-only arenas: `in_arena` receives the compiled batch in ONE
  // host-to-device copy from the pinned `staging` buffer; `work_arena` holds the device-only planning arrays.
  DevArena in_arena;
  DevArena work_arena;
  PinBuf<uint8_t> staging;
  uint32_t n_qterms = 0;  // total search-term slots over all queries
  DevBuf<uint64_t> d_scan_scratch;  // block sums of the planning scans
  DevBuf<DfTileDesc> d_df_tile_desc;  // [df tiles]
  DevBuf<uint32_t> d_tile_query;    // [and tiles]
  DevBuf<unsigned long long> d_stats;  // device-side accounting, see StatSlot

  // sharded pipeline (mgx_sharded_batch_*): the gathered per-shard records, the merged record, and the events that
  // order the batch's stream with the communicator's stream; all recycled with the workspace
  DevBuf<uint8_t> o_gather;
  DevBuf<uint8_t> o_merged;
  cudaEvent_t ev_x[4] = {nullptr, nullptr, nullptr, nullptr};
  PinBuf<uint32_t> h_status;  // status block of the merged record
  bool sharded_enqueued = false;
  // device-side result buffers of the host-buffer call (mgx_query_batch), recycled with the workspace
  DevBuf<uint32_t> o_ids;
  DevBuf<double> o_scores;
  DevBuf<uint32_t> o_count;
  DevBuf<uint64_t> o_total;
  DevBuf<uint64_t> o_df;

  // Driver expansion of OR-rooted boolean programs (api.cu expand_or_roots): caller query q is run as the internal
  // queries [h_xoff[q], h_xoff[q+1]) with pairwise disjoint answers, folded back by fold_expanded_kernel.
  std::vector<uint32_t> h_xoff;       // [n_out_queries + 1], empty = no expansion in this batch
  DevBuf<uint32_t> d_xoff;
  DevBuf<uint32_t> x_ids;             // answers of the internal queries, [n_queries][x_stride]
  DevBuf<uint32_t> x_count;
  DevBuf<uint64_t> x_total;
  ExplicitDriver explicit_driver;
  uint64_t h2d_bytes = 0;
  uint64_t d2h_bytes = 0;
  uint64_t launches_at_start = 0;
  uint64_t n_df_tiles = 0;
  uint64_t n_and_tiles = 0;
  uint64_t driver_entries = 0;
  bool planned = false;
  bool df_done = false;
  bool status_copied = false;  // h_launch is being / has been filled for this batch
  bool searched = false;  // batch_search has enqueued the batch's last kernels (ev_last follows them)

  // CUDA-event timing of the named kernels on the launch stream
  struct Timed {
    cudaEvent_t a;
    cudaEvent_t b;
    int kind;  // 0 plan, 1 df kernel, 2 and kernel, 3 topk kernel, 4 streaming df kernel
  };
  std::vector<Timed> timed;
  cudaEvent_t ev_first = nullptr;
  cudaEvent_t ev_last = nullptr;
  void time_begin(int kind);
  void time_end();
  void mark_last();
  void collect_stats(mgx_batch_stats_t* out);  // synchronises the stream
  void recycle();  // forget per-batch state, keep the allocations
  ~Batch();
};

enum StatSlot : int {
  kStatIntersectLists = 0,  // sum_q sum_lists min(4|P|, ceil(N/8))
  kStatResultDocs = 1,      // sum |R|
  kStatScoreBytes = 2,      // sum_{d in R} (text_bytes + 4)
  kStatDfBytes = 3,         // text bytes scanned for df
  kStatDfCandidates = 4,
  kStatDfLists = 5,
  kStatStreamEntries = 6,   // shortest-list entries of the terms eligible for the streaming df pass
  kStatStreamHits = 7,      // verified (term, document) pairs counted by the streaming pass
  kStatDfScanned = 8,       // df candidates whose whole text had to be scanned (no usable first-occurrence position)
  kStatDriverEntries = 9,   // driver entries of all queries (what the intersect tiles walk)
  kStatCount = 10
};
constexpr int kStatStripes = 64;  // each counter is striped over 64 words to spread the atomics

// query.cu
// `terms` is not const: the stream-table builder clears `streamable` of terms that do not fit a bucket.
// xoff (optional): expansion table of the caller's queries, see Batch::h_xoff.
void batch_upload(Batch& b, std::vector<HostTerm>& terms, const std::vector<HostQuery>& queries,
                  const std::vector<uint32_t>& slot_tid, const std::vector<uint32_t>* xoff = nullptr);
// Bucket table over the terms with `streamable` set (clears the flag of terms that do not fit a bucket).
void build_stream_table(std::vector<HostTerm>& terms, HostStreamTable* out);
void batch_bind(Batch& b);
// A batch compiled by another process for an index of the same configuration: `bytes` = its staging buffer
// (layout.total bytes). Fix-ups for this shard (term shortcuts that assumed an all-valid-UTF-8 corpus) are applied to
// the copy. Batches with column filters cannot travel (their predicates hold device pointers of the compiling shard).
void batch_import(Batch& b, const StageLayout& layout, const uint8_t* bytes);
void batch_plan(Batch& b);
void batch_clear_counters(Batch& b);
void batch_df(Batch& b);
// Streamed batches only: true when the batch did not fit the stream's workspace (call after its stream work has
// completed). The batch is then reset to its uploaded state and must be run again; it will take the synchronous form.
bool batch_overflowed(Batch& b);
// after_overflow: the repeat must take the synchronous form (and the stream's next batches too, for a while).
void batch_reset_for_repeat(Batch& b, bool after_overflow);
// Runs the intersect/score kernels and the per-query output kernel. Outputs are DEVICE pointers.
// set_mode: 0 = top-k by score or first ids (params.limit/offset), 1 = full ascending sets
// written to d_sets (offsets d_set_off[Q+1] computed here), 2 = last `limit` ids descending (reverse)
void batch_search(Batch& b, const uint64_t* d_df_global, uint64_t stride, uint32_t* d_ids, double* d_scores,
                  uint32_t* d_count, uint64_t* d_total);
// Full ascending result sets of every query: fills h_totals and returns a device buffer of all sets back to back.
void batch_search_sets(Batch& b, std::vector<uint64_t>* h_set_off, DevBuf<uint32_t>* d_sets);
// Union of the ascending, pairwise disjoint runs d_in[run_off[r] .. run_off[r+1]) into one ascending array.
void lookup_list_lengths(Batch& b, const uint64_t* h_keys, uint32_t n, uint32_t* h_lens);
// Grouped form: runs [group_begin[g], group_begin[g + 1]) are merged among themselves; the union of group g takes the
// place of its runs in the output (same offsets as the input's concatenation).
void merge_grouped_runs(Batch& b, const uint32_t* d_in, const std::vector<uint64_t>& run_off,
                        const std::vector<uint32_t>& group_begin, DevBuf<uint32_t>* d_out);
void merge_disjoint_runs(Batch& b, const uint32_t* d_in, const std::vector<uint64_t>& run_off,
                         DevBuf<uint32_t>* d_out);
void launch_merge_topk(cudaStream_t stream, const mgx_query_params_t& params, uint32_t n_shards, uint64_t n_queries,
                       uint64_t stride, const uint32_t* d_ids_all, const double* d_scores_all,
                       const uint32_t* d_count_all, const uint64_t* d_total_all, uint64_t shard_pitch_bytes,
                       uint32_t* d_ids_out, double* d_scores_out, uint32_t* d_count_out, uint64_t* d_total_out);
// OR of the 16-byte status blocks of n_shards packed records (first block at d_status_first, one every pitch bytes).
void launch_or_status(cudaStream_t stream, const uint8_t* d_status_first, uint64_t shard_pitch_bytes, uint32_t n_shards,
                      uint8_t* d_status_out);
void batch_df_to_slots(Batch& b, uint64_t* d_df_slots);
void launch_score_documents(Index& ix, cudaStream_t stream, const uint32_t* d_cands, uint64_t n_cands,
                            const uint8_t* d_term_bytes, const uint32_t* d_term_boff, const uint64_t* d_dfs,
                            uint32_t n_terms, uint64_t total_docs, double avgdl, double k1, double b, double* d_scores);
void launch_sort_by_score(cudaStream_t stream, const uint32_t* d_docs, const double* d_scores, uint64_t n,
                          bool descending, uint32_t limit, uint32_t offset, uint32_t* d_out, uint32_t* d_out_count);

}  // namespace mgx
