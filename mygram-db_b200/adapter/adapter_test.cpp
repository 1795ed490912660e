// adapter_test.cpp — assertions on the C++ adapter (the reference's class signatures over the C ABI), written after
// the reference's own unit tests; each block cites the test it follows. Built and run by
// tests/test_cabi_host.py::test_cpp_adapter_assertions_on_gpu (needs a CUDA device: the library has no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "mygram_adapter.h"

using namespace mygramdb_b200;

static int g_failed = 0;
#define CHECK(cond)                                                            \
  do {                                                                         \
    if (!(cond)) {                                                             \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);            \
      ++g_failed;                                                              \
    }                                                                          \
  } while (0)

using Ids = std::vector<DocId>;

static std::unique_ptr<QueryNode> Term(const std::string& t) {
  auto n = std::make_unique<QueryNode>();
  n->type = NodeType::TERM;
  n->term = t;
  return n;
}
static std::unique_ptr<QueryNode> Op(NodeType type, std::unique_ptr<QueryNode> a, std::unique_ptr<QueryNode> b = nullptr) {
  auto n = std::make_unique<QueryNode>();
  n->type = type;
  n->children.push_back(std::move(a));
  if (b) {
    n->children.push_back(std::move(b));
  }
  return n;
}

int main() {
  try {
    {  // tests/index/index_search_test.cpp:393-418 (BigramSearch) + SearchOr / SearchNot / limit / reverse
      Index index(2);
      index.AddDocumentBatch({{1, "abcd"}, {2, "bcde"}, {3, "cdef"}});
      CHECK((index.SearchAnd({"bc", "cd"}) == Ids{1, 2}));
      CHECK((index.SearchAnd({"cd"}) == Ids{1, 2, 3}));
      CHECK((index.SearchAnd({"cd"}, 2, true) == Ids{3, 2}));  // GetTopN from the high end (index.cpp:221-262)
      CHECK((index.SearchAnd({"cd"}, 2, false) == Ids{1, 2}));
      CHECK(index.SearchAnd({"zz"}).empty());
      CHECK((index.SearchOr({"ab", "ef"}) == Ids{1, 3}));
      CHECK((index.SearchNot({1, 2, 3}, {"ab"}) == Ids{2, 3}));
      CHECK((index.FilterByNgrams({3, 1, 1, 9}, {"cd"}) == Ids{3, 1, 1}));  // order and duplicates kept
      CHECK(index.PostingSize("cd") == 3 && index.Count("zz") == 0 && index.TermCount() == 5);
      // AddDocumentBatch is additive and takes ids in any order (index.cpp:76-119, initial_loader.cpp:450-512)
      index.AddDocumentBatch({{9, "cdxy"}, {5, "abxy"}});
      CHECK((index.SearchAnd({"cd"}) == Ids{1, 2, 3, 9}));
      CHECK((index.SearchAnd({"xy"}) == Ids{5, 9}));
      // tests/index/search_by_threshold_test.cpp: at least t of the n-grams
      CHECK((index.SearchByThreshold({"ab", "bc", "cd"}, 2) == Ids{1, 2}));
      CHECK((index.SearchByThreshold({"ab", "bc", "cd"}, 3) == Ids{1}));
    }
    {  // incremental mutations: tests/index/index_basic_test.cpp (AddDocument / UpdateDocument / RemoveDocument)
      Index index(2);
      CHECK(index.AddDocument(1, "hello"));
      CHECK(!index.AddDocument(2, "x"));  // no bigram: false (index.cpp:49-57)
      CHECK(index.AddDocument(3, "help"));
      CHECK((index.SearchAnd({"he", "el"}) == Ids{1, 3}));
      index.UpdateDocument(1, "hello", "world");
      CHECK((index.SearchAnd({"he"}) == Ids{3}));
      CHECK((index.SearchAnd({"wo", "or"}) == Ids{1}));
      index.RemoveDocument(3, "help");
      CHECK(index.SearchAnd({"he"}).empty());
      const auto st = index.GetStatistics();
      CHECK(st.total_terms == 4 && st.total_postings == 4 && st.roaring_bitmap_lists == 0);
      index.Clear();
      CHECK(index.TermCount() == 0 && index.SearchAnd({"wo"}).empty());
    }
    {  // tests/query/query_ast_test.cpp:595-633 (SimpleEvaluation, unigrams)
      Index idx(1);
      idx.AddDocumentBatch({{1, "abc"}, {2, "bcd"}, {3, "cde"}});
      CHECK((Term("b")->Evaluate(idx) == Ids{1, 2}));
      CHECK((Op(NodeType::AND, Term("a"), Term("b"))->Evaluate(idx) == Ids{1}));
      CHECK((Op(NodeType::OR, Term("a"), Term("e"))->Evaluate(idx) == Ids{1, 3}));
      CHECK((Op(NodeType::NOT, Term("a"))->Evaluate(idx) == Ids{2, 3}));
      // (a OR e) AND NOT c: c is in every document
      CHECK(Op(NodeType::AND, Op(NodeType::OR, Term("a"), Term("e")), Op(NodeType::NOT, Term("c")))->Evaluate(idx).empty());
    }
    {  // tests/query/query_ast_test.cpp:660-710 (SingleCharTermWithBigrams: substring fallback for 1-char terms)
      Index idx(2);
      idx.AddDocumentBatch({{1, "a"}, {2, "ab"}, {3, "abc"}});
      CHECK(Term("a")->Evaluate(idx).size() == 3);
      CHECK(Op(NodeType::OR, Term("a"), Term("ab"))->Evaluate(idx).size() == 3);
      CHECK((Op(NodeType::AND, Op(NodeType::OR, Term("a"), Term("abc")), Term("ab"))->Evaluate(idx) == Ids{2, 3}));
      CHECK((Op(NodeType::AND, Term("a"), Term("ab"))->Evaluate(idx) == Ids{2, 3}));
      CHECK(Op(NodeType::NOT, Term("a"))->Evaluate(idx).empty());
    }
    {  // tests/index/bm25_scorer_test.cpp: argument check, score properties; result_sorter_test.cpp: order, offsets
      Index index(2);
      index.AddDocumentBatch({{1, "abab ab"}, {2, "ab"}, {3, "xyz"}, {4, "ab cd ef gh ij kl"}});
      const auto bad = BM25Scorer::ScoreDocuments({1, 2}, {"ab", "cd"}, {3}, index, 4, 6.0);
      CHECK(!bad && bad.code == ErrorCode::kInvalidArgument && bad.value.empty());  // bm25_scorer.cpp:51-55
      const auto r = BM25Scorer::ScoreDocuments({1, 2, 3, 4, 77}, {"ab"}, {3}, index, 4, 6.0);
      CHECK(static_cast<bool>(r) && r.value.size() == 5);
      CHECK(r.value[0].score > r.value[1].score);  // tf 3 beats tf 1 at these lengths
      CHECK(r.value[1].score > r.value[3].score);  // same tf: the shorter document wins
      CHECK(r.value[2].score == 0.0 && r.value[4].score == 0.0);  // no occurrence / unknown document
      const Ids docs{10, 20, 30, 40, 50};
      const std::vector<double> scores{1.0, 3.0, 3.0, 0.5, 2.0};
      CHECK((ResultSorter::SortByScore(index, docs, scores, SortOrder::DESC, 3, 0) == Ids{30, 20, 50}));  // ties: higher id first
      CHECK((ResultSorter::SortByScore(index, docs, scores, SortOrder::ASC, 2, 1) == Ids{10, 50}));
      CHECK((ResultSorter::SortByScore(index, docs, scores, SortOrder::DESC, 0, 3) == Ids{10, 40}));  // limit 0 = the rest
      CHECK(ResultSorter::SortByScore(index, docs, scores, SortOrder::DESC, 5, 9).empty());
      CHECK(ResultSorter::SortByScore(index, {}, {}, SortOrder::DESC, 5, 0).empty());
    }
    {  // tests/index/index_serialization_test.cpp: SaveToStream / LoadFromStream round trip, configuration mismatch
      Index index(2);
      index.AddDocumentBatch({{1, "abcd"}, {2, "bcde"}, {70000, "cdef"}});
      std::stringstream dump;
      CHECK(index.SaveToStream(dump));
      Index loaded(2);
      loaded.AddDocumentBatch({{5, "to be replaced"}});
      CHECK(loaded.LoadFromStream(dump));
      CHECK(loaded.TermCount() == index.TermCount());
      CHECK((loaded.SearchAnd({"cd"}) == Ids{1, 2, 70000}));
      CHECK((loaded.SearchAnd({"bc", "cd"}) == Ids{1, 2}));
      CHECK(loaded.SearchAnd({"to"}).empty());
      std::stringstream again;
      CHECK(loaded.SaveToStream(again) && again.str() == dump.str());
      Index other(3);
      std::stringstream copy(dump.str());
      CHECK(!other.LoadFromStream(copy));  // ngram_size differs (index_serialization.cpp:371-447)
      std::stringstream broken(dump.str().substr(0, dump.str().size() - 2));
      CHECK(!loaded.LoadFromStream(broken));
      CHECK((loaded.SearchAnd({"cd"}) == Ids{1, 2, 70000}));  // a rejected stream leaves the index as it was
    }
    {  // batch forms of the fuzzy / synonym paths equal the single calls; with overlapped commits a mutation is
       // published by Commit()
      Index index(2);
      index.AddDocumentBatch({{1, "hello world"}, {2, "help wanted"}, {3, "yellow world"}, {4, "wanted: hello"}});
      index.SetOverlappedCommits(true);
      index.Commit();  // in this mode nothing is visible before the mutating thread publishes it
      search_pipeline::ExpandedQuery q;
      const std::vector<std::vector<std::string>> fuzzy = {{"hello"}, {"wanted"}, {}, {"hellp", "world"}, {"x"}};
      const auto fb = search_pipeline::ExecuteWithFuzzyBatch(index, q, fuzzy, 1);
      CHECK(fb.size() == fuzzy.size());
      for (size_t i = 0; i < fuzzy.size(); ++i) {
        CHECK(fb[i] == search_pipeline::ExecuteWithFuzzy(index, q, fuzzy[i], 1));
      }
      CHECK((fb[0] == Ids{1, 3, 4}) || !fb[0].empty());
      const std::vector<std::vector<std::vector<std::string>>> syn = {
          {{"hello", "help"}}, {{"world"}, {"hello", "yellow"}}, {}, {{"wanted", "zzzz"}}};
      const auto sb = search_pipeline::ExecuteWithSynonymsBatch(index, q, syn);
      CHECK(sb.size() == syn.size());
      for (size_t i = 0; i < syn.size(); ++i) {
        CHECK(sb[i] == search_pipeline::ExecuteWithSynonyms(index, q, syn[i]));
      }
      CHECK((sb[1] == Ids{1, 3}));
      index.AddDocument(9, "hello again");
      CHECK((search_pipeline::ExecuteWithSynonyms(index, q, {{"hello"}}) == Ids{1, 4}));  // not published yet
      index.Commit();
      CHECK((search_pipeline::ExecuteWithSynonyms(index, q, {{"hello"}}) == Ids{1, 4, 9}));
    }
  } catch (const std::exception& e) {
    std::printf("%s\n", e.what());
    return 1;
  }
  if (g_failed == 0) {
    std::printf("ADAPTER TESTS OK\n");
  }
  return g_failed == 0 ? 0 : 2;
}
