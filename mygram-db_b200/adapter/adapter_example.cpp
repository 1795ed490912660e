// adapter_example.cpp — compile/link check of the C++ adapter and a minimal usage sample.
// The calls mirror tests/index/index_search_test.cpp:393-418 (BigramSearch) of the reference.
#include <cstdio>
#include <sstream>

#include "mygram_adapter.h"

int main() {
  using namespace mygramdb_b200;
  try {
    Index index(2);
    index.AddDocumentBatch({{1, "abcd"}, {2, "bcde"}, {3, "cdef"}});
    const auto hits = index.SearchAnd({"bc", "cd"});
    std::printf("SearchAnd({bc,cd}) -> %zu docs\n", hits.size());
    const auto scored = BM25Scorer::ScoreDocuments(hits, {"bcd"}, {2}, index, 3, 4.0);
    const auto top = ResultSorter::SortByScore(index, hits, {scored.value[0].score, scored.value[1].score},
                                               SortOrder::DESC, 10, 0);
    std::printf("top doc %u\n", top.empty() ? 0 : top[0]);
    search_pipeline::ExpandedQuery q;  // fuzzy / synonym paths of the pipeline (search_pipeline.cpp:1580-1740)
    q.not_terms = {"ef"};
    const auto fuzzy = search_pipeline::ExecuteWithFuzzy(index, q, {"bxde"}, 1);
    const auto syn = search_pipeline::ExecuteWithSynonyms(index, q, {{"ab", "de"}, {"bc", "cd"}});
    std::printf("fuzzy %zu docs, synonyms %zu docs\n", fuzzy.size(), syn.size());
    std::ostringstream dump;  // Index::SaveToStream: the MGIX stream DUMP SAVE writes
    std::printf("MGIX stream %s, %zu bytes\n", index.SaveToStream(dump) ? "ok" : "failed", dump.str().size());
  } catch (const std::exception& e) {
    std::printf("%s\n", e.what());
    return 1;
  }
  return 0;
}
