"""GPU parity of the fuzzy and synonym execution paths (search_pipeline::ExecuteWithFuzzy / ExecuteWithSynonyms,
SURVEY §8f-3) through the C ABI (mgx_search_fuzzy / mgx_search_synonyms) against the CPU oracle, whose restatement
tests/test_oracle_expanded.py pins to the reference's own functions. Bit-exact doc-id sets."""
import random

import numpy as np
import pytest

import corpus as corpus_mod
from test_gpu_parity import assert_same_index, build_pair, make_docs
from test_oracle_bulk import _docs
from test_oracle_expanded import expanded_cases, spaced_docs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", [(2, 0, True), (2, 1, True), (2, 1, False), (3, 2, False), (1, 1, True), (3, 0, True)])
def test_fuzzy_and_synonyms_match_oracle(mgx, oracle, cfg):
    rnd = random.Random(0x51 + (hash(cfg) & 0xFFF))
    docs = spaced_docs(rnd, 3000)
    ids = np.arange(1, len(docs) + 1, dtype=np.uint32)
    gi, oi = build_pair(mgx, oracle, docs, ids, cfg)
    nonempty = verified = 0
    for fuzzy_terms, groups, nots, dist in expanded_cases(rnd, docs, 120):
        for vt in (0, 1, 2):  # 1 / 2: PostFilterByFuzzyText (edit-distance verification) on the device
            want, _ = oi.search_fuzzy(fuzzy_terms, dist, nots, verify_text=vt)
            got = gi.search_fuzzy(fuzzy_terms, dist, nots, verify_text=vt)
            assert np.array_equal(got, want), (cfg, fuzzy_terms, dist, nots, vt, got[:10], want[:10])
            nonempty += want.size > 0
            verified += vt == 1 and want.size > 0
        for vt in (0, 1, 2):
            want, _ = oi.search_synonyms(groups, nots, verify_text=vt)
            got = gi.search_synonyms(groups, nots, verify_text=vt)
            assert np.array_equal(got, want), (cfg, groups, nots, vt, got[:10], want[:10])
            nonempty += want.size > 0
    assert nonempty > 40 and verified > 5
    assert gi.search_fuzzy([], 1).size == 0 and gi.search_synonyms([]).size == 0
    # terms beyond the edit-distance rows of the device function are refused, never approximated
    with pytest.raises(mgx.MgxError):
        gi.search_fuzzy(["ab" * 40], 1, verify_text=1)


@pytest.mark.parametrize("cfg", [(2, 0, True), (2, 1, True), (3, 2, False)])
def test_fuzzy_and_synonym_batches_equal_the_single_calls(mgx, oracle, cfg):
    """mgx_search_fuzzy_batch / mgx_search_synonyms_batch answer many queries in one batch on the device (programs
    side by side, driver-expanded queries united per query by the grouped merge): every answer must be the single
    call's, which the test above pins to the oracle; queries the reference answers with the empty set without touching
    a list (no terms, a term shorter than an n-gram, no groups) take no kernel query at all."""
    rnd = random.Random(0x77 + (hash(cfg) & 0xFFF))
    docs = spaced_docs(rnd, 3000)
    ids = np.arange(1, len(docs) + 1, dtype=np.uint32)
    gi, oi = build_pair(mgx, oracle, docs, ids, cfg)
    cases = list(expanded_cases(rnd, docs, 150))
    nots = cases[0][2]
    for vt in (0, 1):
        for dist in (1, 2):
            fq = [c[0] for c in cases] + [[], ["a"]]
            got = gi.search_fuzzy_batch(fq, dist, nots, verify_text=vt)
            assert len(got) == len(fq)
            nonempty = 0
            for q, g in zip(fq, got):
                want = gi.search_fuzzy(q, dist, nots, verify_text=vt)
                assert np.array_equal(g, want), (cfg, q, dist, vt, g[:10], want[:10])
                nonempty += want.size > 0
            assert nonempty > 5
        sq = [c[1] for c in cases] + [[], [["a"]], [[]]]
        got = gi.search_synonyms_batch(sq, nots, verify_text=vt)
        nonempty = 0
        for q, g in zip(sq, got):
            want = gi.search_synonyms(q, nots, verify_text=vt)
            assert np.array_equal(g, want), (cfg, q, vt, g[:10], want[:10])
            nonempty += want.size > 0
        assert nonempty > 5
    # against the oracle directly, and the capacity protocol (a first buffer that is too small)
    want = [oi.search_fuzzy(c[0], 1, nots)[0] for c in cases[:40]]
    got = gi.search_fuzzy_batch([c[0] for c in cases[:40]], 1, nots)
    assert all(np.array_equal(g, w) for g, w in zip(got, want))
    small = gi._grouped_call
    gi._grouped_call = lambda fn, n, cap=1: small(fn, n, cap=1)
    try:
        again = gi.search_fuzzy_batch([c[0] for c in cases[:40]], 1, nots)
    finally:
        gi._grouped_call = small
    assert all(np.array_equal(g, w) for g, w in zip(again, want))
    assert gi.search_fuzzy_batch([], 1) == [] and gi.search_synonyms_batch([]) == []


def test_fuzzy_and_synonyms_with_filters_dense_lists_and_invalid_utf8(mgx, oracle):
    """Zipf corpus with dense-bitmap lists, column conditions applied after the NOT terms (ApplyNotAndFilters,
    search_pipeline.cpp:470-485), and a corpus with invalid UTF-8."""
    rnd = random.Random(77)
    c = corpus_mod.generate("cjk", 20000, 0xF3, alphabet=96, min_len=6, max_len=40)
    gi = mgx.Index(2, 0, True, dense_threshold=0.02)
    gi.build(c.doc_ids, c.arena, c.offsets)
    oi = oracle.index(2, 0, True)
    oi.build_bulk(c.doc_ids, c.arena, c.offsets, 8)
    assert gi.stats().n_dense_terms > 0
    n = c.n_docs
    status = [None if rnd.random() < 0.1 else rnd.choice([1, 2, 3]) for _ in range(n)]
    category = [None if rnd.random() < 0.1 else rnd.choice([b"news", b"blog", b"wiki"]) for _ in range(n)]
    columns = [(8, status), (11, category)]
    for ci, (typ, vals) in enumerate(columns):
        gi.set_filter_column(ci, typ, vals)
    first = int(c.doc_ids[0])
    filter_pool = [[(0, 0, "1")], [(1, 1, "news"), (0, 0, "3")], [(0, 3, "2")], [(1, 0, "blog")], []]

    def piece(lo, hi):
        t = c.text(rnd.randrange(n)).decode()
        ln = rnd.randint(lo, min(hi, len(t)))
        st = rnd.randrange(0, len(t) - ln + 1)
        return t[st:st + ln]

    def misspell(s):
        i = rnd.randrange(len(s))
        return s[:i] + chr(0x4E00 + rnd.randrange(96)) + s[i + 1:]

    sizes = []
    for _ in range(60):
        fl = rnd.choice(filter_pool)
        nots = [piece(2, 2)] if rnd.random() < 0.3 else []
        terms = [misspell(piece(3, 8)) for _ in range(rnd.randint(1, 2))]
        dist = rnd.randint(1, 2)
        want, _ = oi.search_fuzzy(terms, dist, nots)
        want = oracle.apply_filters(n, first, columns, fl, want) if fl else want
        got = gi.search_fuzzy(terms, dist, nots, filters=fl)
        assert np.array_equal(got, want), (terms, dist, nots, fl, got.size, want.size)
        sizes.append(want.size)
        groups = [[piece(2, 4) for _ in range(rnd.randint(1, 3))] for _ in range(rnd.randint(1, 2))]
        want, _ = oi.search_synonyms(groups, nots)
        want = oracle.apply_filters(n, first, columns, fl, want) if fl else want
        got = gi.search_synonyms(groups, nots, filters=fl)
        assert np.array_equal(got, want), (groups, nots, fl, got.size, want.size)
        sizes.append(want.size)
    assert max(sizes) > 100 and sum(1 for s in sizes if s > 0) > 30
    # invalid UTF-8 in documents and terms
    docs = make_docs(5, 2000, 30, bad=True)
    ids = np.arange(10, 10 + len(docs), dtype=np.uint32)
    for cfg in [(2, 1, True), (2, 0, True)]:
        g2, o2 = build_pair(mgx, oracle, docs, ids, cfg)
        for _ in range(80):
            d = docs[rnd.randrange(len(docs))]
            if len(d) < 4:
                continue
            st = rnd.randrange(0, len(d) - 3)
            term = d[st:st + rnd.randint(2, 9)]
            for vt in (0, 1):
                want, _ = o2.search_fuzzy([term], 1, verify_text=vt)
                assert np.array_equal(g2.search_fuzzy([term], 1, verify_text=vt), want), (cfg, term, vt)
            want, _ = o2.search_synonyms([[term, d[:3]]], verify_text=1)
            assert np.array_equal(g2.search_synonyms([[term, d[:3]]], verify_text=1), want), (cfg, term)


def test_device_index_saves_an_mgix_stream(mgx, oracle):
    """mgx_index_save_mgix (Index::SaveToStream, index_serialization.cpp:111-224): the stream of a device-built index
    decodes to the oracle's CSR, follows the representation rule after Optimize, and — where the reference's own
    sources are built (oracle/_ref) — loads into the reference's Index::LoadFromStream and answers like it."""
    import os

    import pyoracle
    from test_mgix_codec import big_corpus, csr_of
    docs, ids = big_corpus(21)
    for cfg in [(2, 0, True), (2, 1, True)]:
        gi, oi = build_pair(mgx, oracle, docs, ids, cfg)
        terms, offs, posts = csr_of(oi)
        stream = gi.save_mgix()
        meta, t2, o2, p2 = mgx.mgix_decode(stream)
        assert t2 == terms and np.array_equal(o2, offs) and np.array_equal(p2, posts)
        assert (meta["ngram_size"], meta["kanji_ngram_size"]) == (cfg[0], cfg[1] if cfg[1] > 0 else cfg[0])
        assert stream == mgx.mgix_encode(terms, offs, posts, cfg[0], meta["kanji_ngram_size"], cfg[2])
        gi.optimize(500)  # lists of >= 0.18 * 500 entries become Roaring bodies (posting_list.cpp:800-834)
        optimized = gi.save_mgix()
        assert optimized == mgx.mgix_encode(terms, offs, posts, cfg[0], meta["kanji_ngram_size"], cfg[2],
                                            roaring_min_len=0.18 * 500)
        assert optimized != stream and np.array_equal(mgx.mgix_decode(optimized)[3], posts)
        if os.path.exists(pyoracle.REF_LIB):
            ri = pyoracle.OracleLib(pyoracle.REF_LIB).index(*cfg)
            assert ri.load_stream(stream) == 0 and ri.term_count() == len(terms)
            rnd = random.Random(3)
            for _ in range(40):
                pick = [terms[rnd.randrange(len(terms))] for _ in range(2)]
                assert np.array_equal(ri.search_and(pick), gi.search_and(pick))
    empty = mgx.Index(2, 0, True)
    meta, terms, offs, posts = mgx.mgix_decode(empty.save_mgix())
    assert meta["n_terms"] == 0 and posts.size == 0


def test_device_index_loads_an_mgix_stream(mgx, oracle):
    """mgx_index_load_mgix (Index::LoadFromStream, index_serialization.cpp:279-613): a stream -- written by the codec
    from the oracle's index, and the streams recorded from the reference's own SaveToStream -- replaces the device
    index; the set calls then answer like the index the stream came from (dense and sparse lists, sparse doc ids),
    the stream saved again is byte-identical, a configuration mismatch and mutations are refused, and a bulk build
    brings the documents back."""
    import base64
    import json

    from test_mgix_codec import GOLDEN, big_corpus, csr_of
    docs, ids = big_corpus(33)
    for cfg in [(2, 0, True), (2, 1, True), (3, 2, False)]:
        oi = oracle.index(*cfg)
        oi.add_texts(ids, docs)
        terms, offs, posts = csr_of(oi)
        eff_kanji = cfg[1] if cfg[1] > 0 else cfg[0]
        stream = mgx.mgix_encode(terms, offs, posts, cfg[0], eff_kanji, cfg[2])
        gi = mgx.Index(*cfg)
        gi.add_document_batch([7, 9], ["will be", "replaced"])
        gi.load_mgix(stream)
        st = gi.stats()
        assert (st.n_terms, st.n_postings) == (len(terms), posts.size) and st.n_docs == np.unique(posts).size
        t2, o2, p2 = gi.export()
        assert t2 == terms and np.array_equal(o2, offs) and np.array_equal(p2, posts)
        assert gi.save_mgix() == stream
        rnd = random.Random(4)
        big = [terms[i] for i in np.argsort(np.diff(offs.astype(np.int64)))[-6:]]
        for _ in range(60):
            pick = [rnd.choice(big) if rnd.random() < 0.4 else terms[rnd.randrange(len(terms))] for _ in range(rnd.randint(1, 3))]
            assert np.array_equal(gi.search_and(pick), oi.search_and(pick)), pick
            assert np.array_equal(gi.search_or(pick), oi.search_or(pick)), pick
            assert np.array_equal(gi.search_and(pick, 5, True), oi.search_and(pick, 5, True))
            assert gi.posting_size(pick[0]) == oi.posting_size(pick[0])
        all_ids = np.unique(posts)
        assert np.array_equal(gi.search_not(all_ids, [big[0]]), oi.search_not(all_ids, [big[0]]))
        # un-scored batch over the lists (no text needed), scored batch: documents without stored text score 0
        qs = [[rnd.choice(big)] for _ in range(8)]
        r = gi.query_batch(qs, score=False, limit=20)
        for q, terms_q in enumerate(qs):
            want = oi.search_and(terms_q)
            assert int(r.total[q]) == want.size and np.array_equal(r.ids[q, :int(r.count[q])], want[:20])
        with pytest.raises(mgx.MgxError):
            gi.add_document(123456, "no documents here")
        other = mgx.Index(cfg[0] + 1 if cfg[0] < 3 else 1, cfg[1], cfg[2])
        with pytest.raises(mgx.MgxError):
            other.load_mgix(stream)
        with pytest.raises(mgx.MgxError):
            gi.load_mgix(stream[:-3])  # CRC / truncation: the index is left as it was
        assert gi.stats().n_terms == len(terms)
        arena, doc_offs = mgx.pack_strings(docs)
        gi.build(ids, arena, doc_offs)  # the documents arrive: a full index again
        assert_same_index(gi, oi)
        gi.add_document(123456, "ab")
    for case in json.load(open(GOLDEN))["cases"]:
        gi = mgx.Index(case["config"][0], case["config"][1], bool(case["config"][2]))
        gi.load_mgix(base64.b64decode(case["stream_b64"]))
        t2, o2, p2 = gi.export()
        assert [t.hex() for t in t2] == case["terms_hex"] and o2.tolist() == case["posting_offsets"]
        assert p2.size == case["n_postings"]
