"""The oracle's multi-threaded bulk builder (used to prepare the 10M-document CPU baseline index) must produce the
same index and the same query results as the faithful AddDocumentBatch restatement; and, where the reference's own
sources were compiled (oracle/_ref), the restatement must agree with them on random inputs."""
import random

import numpy as np
import pytest

import corpus as corpus_mod


@pytest.mark.parametrize("kind,cfg", [("cjk", (2, 0, True)), ("ascii", (2, 0, True)), ("cjk", (1, 1, True)),
                                      ("ascii", (3, 2, False))])
def test_bulk_build_equals_add_batch(oracle, kind, cfg):
    c = corpus_mod.generate(kind, 20000, 99, alphabet=512 if kind == "cjk" else 300)
    a = oracle.index(*cfg)
    a.add_batch(c.doc_ids, c.arena, c.offsets)
    b = oracle.index(*cfg)
    b.build_bulk(c.doc_ids, c.arena, c.offsets, 4)
    ta, oa, pa = a.export()
    tb, ob, pb = b.export()
    assert ta == tb and np.array_equal(oa, ob) and np.array_equal(pa, pb)
    assert a.bm25_stats() == b.bm25_stats()
    qs = corpus_mod.sample_queries(c, 150, 3, n_terms=2)
    ra = a.query_batch(qs, want_sets=True)
    rb = b.query_batch(qs, want_sets=True, n_threads=4)
    assert np.array_equal(ra.ids, rb.ids) and np.array_equal(ra.df, rb.df) and np.array_equal(ra.scores, rb.scores)
    for x, y in zip(ra.sets, rb.sets):
        assert np.array_equal(x, y)


def _docs(rnd, n):
    words = [bytes(rnd.choice(b"abcdefgh") for _ in range(rnd.randint(2, 5))) for _ in range(40)]
    cj = [chr(0x4E00 + i).encode() for i in range(30)] + ["あ".encode(), "😀".encode()]
    docs = []
    for _ in range(n):
        parts = []
        for _ in range(rnd.randint(0, 6)):
            parts.append(rnd.choice(words) if rnd.random() < 0.5 else b"".join(rnd.choice(cj) for _ in range(rnd.randint(1, 4))))
        docs.append(b" ".join(parts) if rnd.random() < 0.5 else b"".join(parts))
    return docs


@pytest.mark.parametrize("cfg", [(2, 0, True), (2, 1, True), (2, 1, False), (3, 2, False), (1, 1, True)])
def test_oracle_agrees_with_reference_sources(oracle, reflib, cfg):
    rnd = random.Random(hash(cfg) & 0xFFFF)
    docs = _docs(rnd, 1500)
    ids = np.arange(1, len(docs) + 1, dtype=np.uint32)
    pi, ri = oracle.index(*cfg), reflib.index(*cfg)
    pi.add_texts(ids, docs)
    ri.add_texts(ids, docs)
    assert pi.term_count() == ri.term_count() and pi.total_postings() == ri.total_postings()
    assert pi.bm25_stats() == ri.bm25_stats()
    qs, nots = [], []
    while len(qs) < 200:
        t = docs[rnd.randrange(len(docs))].decode()
        if len(t) < 3:
            continue
        terms = []
        for _ in range(rnd.randint(1, 3)):
            ln = rnd.randint(1, 4)
            st = rnd.randrange(0, max(1, len(t) - ln + 1))
            terms.append(t[st:st + ln].strip() or "ab")
        qs.append(terms)
        t2 = docs[rnd.randrange(len(docs))].decode()
        nots.append([t2[:2]] if (len(t2.strip()) >= 2 and rnd.random() < 0.3) else [])
    for vt in (0, 1, 2):
        for score in (True, False):
            a = pi.query_batch(qs, not_terms=nots, want_sets=True, verify_text=vt, score=score, limit=20)
            b = ri.query_batch(qs, not_terms=nots, want_sets=True, verify_text=vt, score=score, limit=20)
            assert np.array_equal(a.total, b.total)
            assert np.array_equal(a.ids, b.ids) and np.array_equal(a.count, b.count)
            if score:
                assert np.array_equal(a.df, b.df) and np.array_equal(a.scores, b.scores)
            for x, y in zip(a.sets, b.sets):
                assert np.array_equal(x, y)
    # set algebra + boolean AST
    terms_all, _, _ = pi.export()
    for _ in range(40):
        terms = [terms_all[rnd.randrange(len(terms_all))] for _ in range(rnd.randint(1, 3))]
        assert np.array_equal(pi.search_and(terms), ri.search_and(terms))
        assert np.array_equal(pi.search_and(terms, 3, True), ri.search_and(terms, 3, True))
        assert np.array_equal(pi.search_or(terms), ri.search_or(terms))
        assert np.array_equal(pi.search_not(ids[::2], terms), ri.search_not(ids[::2], terms))
        assert np.array_equal(pi.search_by_threshold(terms + terms[:1], 2), ri.search_by_threshold(terms + terms[:1], 2))
        cands = np.array([ids[rnd.randrange(len(ids))] for _ in range(30)], dtype=np.uint32)
        assert np.array_equal(pi.filter_by_ngrams(cands, terms), ri.filter_by_ngrams(cands, terms))
    words = ["ab", "cd", "東", "e", "bcd"]
    # (a OR c) AND b ; NOT a ; a AND NOT (b OR e)
    progs = [([0, 0, 2, 0, 1], [0, 2, 2, 1, 2]), ([0, 3], [0, 0]), ([0, 0, 0, 2, 3, 1], [0, 1, 3, 2, 0, 2])]
    for ops, args in progs:
        assert np.array_equal(pi.eval_boolean(ops, args, words), ri.eval_boolean(ops, args, words))
