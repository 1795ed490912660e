"""Known-answer vectors transcribed from the reference's own unit tests
(tests/golden/reference_kat.json; every entry cites test file:line) applied to

  * the CPU oracle (oracle/liboracle.so)                      -- pins the oracle        [CPU]
  * the reference's own sources (oracle/_ref/...so), if built -- sanity of the transcription [CPU]
  * the CUDA path through the C ABI                           -- parity                  [gpu]
"""
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "reference_kat.json"), encoding="utf-8"))


@pytest.fixture(params=["oracle", "reference"])
def cpu(request):
    if request.param == "oracle":
        return request.getfixturevalue("oracle")
    return request.getfixturevalue("reflib")


def b(s):
    return s.encode("utf-8")


def test_kat_tokenizer_cpu(cpu):
    for v in KAT["utf8_to_codepoints"]:
        assert cpu.utf8_to_codepoints(v["text"]) == v["cps"], v["src"]
        assert cpu.codepoints_to_utf8(v["cps"]) == b(v["text"]), v["src"]
    for v in KAT["roundtrip"]:
        assert cpu.codepoints_to_utf8(cpu.utf8_to_codepoints(v["text"])) == b(v["text"]), v["src"]
    for v in KAT["count_code_points"]:
        assert cpu.count_code_points(v["text"]) == v["expect"], v["src"]
    for v in KAT["ngrams"]:
        got = cpu.ngrams(v["mode"], v["text"], v["a"], v["k"], v["cross"])
        assert got == [b(x) for x in v["expect"]], (v["src"], v)
    for v in KAT["ngram_equalities"]:
        l, r = v["lhs"], v["rhs"]
        assert cpu.ngrams(l["mode"], l["text"], l["a"], l["k"], l["cross"]) == \
            cpu.ngrams(r["mode"], r["text"], r["a"], r["k"], r["cross"]), v["src"]


def _docs_of(case):
    if "docs_repeat" in case:
        r = case["docs_repeat"]
        return [(r["first"] + i, r["text"]) for i in range(r["count"])]
    return [(d, t) for d, t in case["docs"]]


def _run_index_checks(idx, case):
    for c in case["checks"]:
        if c["op"] == "and":
            got = idx.search_and([b(t) for t in c["terms"]], c.get("limit", 0), c.get("reverse", False))
            assert got.tolist() == c["expect"], (case["src"], c)
        elif c["op"] == "or":
            assert idx.search_or([b(t) for t in c["terms"]]).tolist() == c["expect"], (case["src"], c)
        elif c["op"] == "term_count":
            assert idx.term_count() == c["expect"], (case["src"], c)
        elif c["op"] == "count":
            assert idx.posting_size(b(c["term"])) == c["expect"], (case["src"], c)


def test_kat_index_cpu(cpu):
    for case in KAT["index"]:
        idx = cpu.index(case["ngram"], case["kanji"], True)
        for d, t in _docs_of(case):  # the reference tests add documents one by one (Index::AddDocument)
            idx.add_document(d, t)
        _run_index_checks(idx, case)


def test_kat_bm25_cpu(cpu):
    env = {"log": math.log}
    for v in KAT["idf"]:
        assert cpu.compute_idf(v["n"], v["df"]) == pytest.approx(eval(v["expect"], env), abs=1e-10), v["src"]
    for v in KAT["tf"]:
        assert cpu.count_term_occurrences(v["text"], v["term"]) == v["expect"], v["src"]
    for v in KAT["sort"]:
        got = cpu.sort_by_score(np.array(v["results"], np.uint32), np.array(v["scores"], np.float64), v["desc"],
                                v["limit"], v["offset"])
        assert got.tolist() == v["expect"], v["src"]
    for v in KAT["score_properties"]:
        idx = cpu.index(2, 1, True)
        for d, t in v["docs"]:
            idx.add_document(d, t)
        s = idx.score_documents(np.array(v["cands"], np.uint32), [b(t) for t in v["terms"]], v["dfs"], v["n"],
                                v["avgdl"], v["k1"], v["b"])
        _check_score_property(s, v)


def _check_score_property(s, v):
    p = v["property"]
    if p == "all_positive":
        assert (s > 0).all(), v["src"]
    elif p == "ascending":
        assert (np.diff(s) > 0).all(), v["src"]
    elif p == "all_zero":
        assert (s == 0).all(), v["src"]
    elif p == "all_equal":
        assert np.allclose(s, s[0], atol=1e-10), v["src"]
    elif p == "first_positive_rest_zero":
        assert s[0] > 0 and (s[1:] == 0).all(), v["src"]


def test_kat_df_cpu(cpu):
    for v in KAT["df"]:
        idx = cpu.index(v["ngram"], v["kanji"], True)
        for d, t in v["docs"]:
            idx.add_document(d, t)
        r = idx.query_batch([[b(t)] for t in v["terms"]], score=True)
        assert r.df.tolist() == v["expect_df"], v["src"]


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_kat_tokenizer_gpu(mgx):
    for v in KAT["ngrams"]:
        if v["mode"] != "hybrid" or not (1 <= v["a"] <= 3 and 1 <= v["k"] <= 3):
            continue  # the device tokenizer is the index-side generator (GenerateHybridNgrams)
        got = mgx.tokenize_batch([v["text"]], v["a"], v["k"], v["cross"])[0]
        assert got == [b(x) for x in v["expect"]], (v["src"], v)
    for v in KAT["count_code_points"]:
        idx = mgx.Index(2, 0, True)
        idx.add_document_batch([1], [v["text"]])
        assert idx.doc_lengths().tolist() == [v["expect"]], v["src"]


@pytest.mark.gpu
def test_kat_index_gpu(mgx):
    for case in KAT["index"]:
        docs = _docs_of(case)
        idx = mgx.Index(case["ngram"], case["kanji"], True)
        idx.add_document_batch([d for d, _ in docs], [t for _, t in docs])
        _run_index_checks(idx, case)


@pytest.mark.gpu
def test_kat_bm25_gpu(mgx):
    for v in KAT["sort"]:
        limit = v["limit"]  # limit 0 == "all" (result_sorter.cpp:689-710)
        idx = mgx.Index(2, 0, True)
        idx.add_document_batch([1], ["ab"])
        got = mgx.ResultSorter.sort_by_score(idx, np.array(v["results"], np.uint32), np.array(v["scores"], np.float64),
                                             v["desc"], limit, v["offset"])
        assert got.tolist() == v["expect"], v["src"]
    for v in KAT["score_properties"]:
        idx = mgx.Index(2, 1, True)
        idx.add_document_batch([d for d, _ in v["docs"]], [t for _, t in v["docs"]])
        s = mgx.BM25Scorer.score_documents(idx, np.array(v["cands"], np.uint32), [b(t) for t in v["terms"]], v["dfs"],
                                           v["n"], v["avgdl"], v["k1"], v["b"])
        _check_score_property(s, v)
    # tf through a one-term scored query: score > 0 iff tf > 0, and df counts the containing docs
    for v in KAT["tf"]:
        if not v["text"] or not v["term"]:
            continue
        idx = mgx.Index(2, 0, True)
        idx.add_document_batch([1], [v["text"]])
        s = mgx.BM25Scorer.score_documents(idx, np.array([1], np.uint32), [b(v["term"])], [1], 1, 5.0, 1.2, 0.0)
        # with b = 0: score = idf * tf * 2.2 / (tf + 1.2)  =>  recover tf exactly
        idf = math.log((1 - 1 + 0.5) / (1 + 0.5) + 1.0)
        tf = v["expect"]
        want = idf * (tf * 2.2) / (tf + 1.2) if tf else 0.0
        assert s[0] == pytest.approx(want, rel=1e-12), v["src"]


@pytest.mark.gpu
def test_kat_df_gpu(mgx):
    for v in KAT["df"]:
        idx = mgx.Index(v["ngram"], v["kanji"], True)
        idx.add_document_batch([d for d, _ in v["docs"]], [t for _, t in v["docs"]])
        r = idx.query_batch([[b(t)] for t in v["terms"]], score=True)
        assert r.df.tolist() == v["expect_df"], v["src"]
