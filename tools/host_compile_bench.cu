// tools/host_compile_bench.cu — host-only timing of the query compiler (no GPU needed): the translation unit of
// the C ABI is included so that its internal compile_batch() can be driven directly on a C2-like batch
// (4096 queries x 3 terms of 2-4 CJK code points, Zipf-ish over 8192 ideographs).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false tools/host_compile_bench.cu \
//        mygram-db_b200/build/{primitives,build,query,mgix}.o -o /tmp/hcb && /tmp/hcb
#include "../mygram-db_b200/csrc/api.cu"

#include <random>

int main(int argc, char** argv) {
  const uint64_t Q = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 4096;
  const int reps = argc > 2 ? std::atoi(argv[2]) : 50;
  std::mt19937_64 rng(42);
  std::uniform_real_distribution<double> uni(0.0, 1.0);
  std::vector<uint8_t> bytes;
  std::vector<uint64_t> offs{0};
  std::vector<uint64_t> qbeg{0};
  for (uint64_t q = 0; q < Q; ++q) {
    for (int t = 0; t < 3; ++t) {
      const int n_cp = 2 + static_cast<int>(rng() % 3);
      for (int c = 0; c < n_cp; ++c) {
        const uint32_t cp = 0x4E00 + static_cast<uint32_t>(std::pow(8192.0, uni(rng))) - 1;
        bytes.push_back(static_cast<uint8_t>(0xE0 | (cp >> 12)));
        bytes.push_back(static_cast<uint8_t>(0x80 | ((cp >> 6) & 0x3F)));
        bytes.push_back(static_cast<uint8_t>(0x80 | (cp & 0x3F)));
      }
      offs.push_back(bytes.size());
    }
    qbeg.push_back(offs.size() - 1);
  }
  mgx::Index ix;
  ix.ngram = 2;
  ix.kanji = 2;
  ix.cross = true;
  ix.width = 2;
  ix.all_valid_utf8 = true;
  mgx_query_params_t p{};
  p.ngram_size = 2;
  p.kanji_ngram_size = 0;
  p.cross_boundary = 1;
  p.compute_score = 1;
  p.limit = 100;
  {  // the tokeniser alone, every term slot, into reused vectors
    mgx::KeyVec keys;
    mgx::TermOffsetVec toff;
    double tb = 1e9;
    for (int r = 0; r < reps; ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      uint64_t acc = 0;
      for (size_t s = 0; s + 1 < offs.size(); ++s) {
        mgx::host_query_keys(bytes.data() + offs[s], offs[s + 1] - offs[s], 2, 0, true, 2, &keys, &toff);
        acc += keys.size();
      }
      const auto t1 = std::chrono::steady_clock::now();
      tb = std::min(tb, std::chrono::duration<double, std::milli>(t1 - t0).count());
      if (acc == 0) std::printf("?");
    }
    std::printf("host_query_keys over %zu term slots: best %.3f ms\n", offs.size() - 1, tb);
  }
  double best = 1e9, sum = 0;
  size_t n_terms = 0;
  std::vector<mgx::HostTerm> terms;  // kept across batches, as the pooled batch object keeps them
  std::vector<mgx::HostQuery> queries;
  std::vector<uint32_t> slot_tid;
  mgx::HostStreamTable table;
  for (int r = 0; r < reps; ++r) {
    terms.clear();
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = compile_batch(ix, p, Q, bytes.data(), offs.data(), qbeg.data(), nullptr, nullptr, nullptr, nullptr,
                                 &terms, &queries, &slot_tid);
    const auto t1 = std::chrono::steady_clock::now();
    mgx::build_stream_table(terms, &table);
    const auto t2 = std::chrono::steady_clock::now();
    if (rc != MGX_OK) {
      std::fprintf(stderr, "compile failed: %s\n", mgx_last_error());
      return 1;
    }
    const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (r == reps - 1) {
      std::printf("last rep: compile %.3f ms, stream table %.3f ms\n", ms,
                  std::chrono::duration<double, std::milli>(t2 - t1).count());
    }
    best = std::min(best, ms);
    sum += ms;
    n_terms = terms.size();
  }
  std::printf("compile_batch: %llu queries, %zu unique terms, threads=%u: best %.3f ms, mean %.3f ms\n",
              static_cast<unsigned long long>(Q), n_terms, compile_threads(Q), best, sum / reps);
  return 0;
}
