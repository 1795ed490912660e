"""GPU parity for n-gram sizes 4..10 (GenerateNgrams has no upper bound, string_utils.cpp:382-423; the
configuration schema allows 1..10, config/config-schema.json:279-285). Keys wider than three code points are
"wide keys" on the device: the same calls as tests/test_gpu_parity.py, through the C ABI, against the CPU
oracle, which works on strings and has no width limit. Bit-exact ids / postings / counts, scores within 1e-9."""
import random

import numpy as np
import pytest

from test_gpu_parity import (assert_batch_equal, assert_same_index, build_pair, make_docs, rand_text, random_program,
                             some_terms)

pytestmark = pytest.mark.gpu

# (ngram_size, kanji_ngram_size, cross_boundary): 2, 3 and 4 words per key, hybrid sizes on either side of the limit
WIDE_CONFIGS = [(4, 0, True), (5, 2, True), (4, 4, False), (2, 6, True), (7, 3, False), (10, 0, True), (3, 9, True),
                (6, 5, True)]

WORDS = ["東京都", "大阪", "検索エンジン", "全文検索", "データベース", "index", "search", "engine", "ngram", "B200",
         "東方Project", "漢字かな交じり", "한국어", "naïve", "日本語のテキスト", "abc", "de", "x", "高速", "並列処理",
         "\U00020000\U00020001\U0002a6df", "㐀㐁㐂㐃", "豈更車", "tokenizer", "posting list", "の", "は", "を"]


def phrase_docs(seed, n, max_words=12, bad=False):
    """Documents of words from a small vocabulary (so long n-grams repeat across documents), some random noise."""
    rnd = random.Random(seed)
    docs = []
    for i in range(n):
        parts = []
        for _ in range(rnd.randint(0, max_words)):
            r = rnd.random()
            if r < 0.8:
                parts.append(rnd.choice(WORDS).encode())
            elif r < 0.9:
                parts.append(b" ")
            else:
                parts.append(rand_text(rnd, 4, bad))
        docs.append(b"".join(parts))
    return docs


def sample_terms(docs, rnd, max_cp):
    t = docs[rnd.randrange(len(docs))].decode("utf-8", "ignore")
    if not t:
        return b"zz"
    ln = rnd.randint(1, max_cp)
    st = rnd.randrange(0, max(1, len(t) - ln + 1))
    return t[st:st + ln].encode()


@pytest.mark.parametrize("cfg", WIDE_CONFIGS)
def test_wide_tokenizer_matches_oracle(mgx, oracle, cfg):
    ng, kj, cross = cfg
    docs = [b"", b"a", "hello world".encode(), "東方Projectの検索".encode(), "漢字ABC漢字漢字漢字ABCDEFGHIJKL".encode(),
            b"\xe6\x9d\xb1\xff\xe6\x96\xb9abcdefghijk", b"\x80\x80", b"abcdefghij", b"abcdefghi", "一二三四五六七八九十".encode(),
            "一二三四五六七八九".encode()]
    docs += make_docs(11, 200, 40, bad=True, long_every=37) + phrase_docs(12, 200, 30, bad=True)
    got = mgx.tokenize_batch(docs, ng, kj, cross)
    eff_kj = kj if kj > 0 else ng
    bad = []
    for i, d in enumerate(docs):
        want = oracle.ngrams("hybrid", d, ng, eff_kj, cross)
        if got[i] != want:
            bad.append((i, d[:60], got[i][:4], want[:4], len(got[i]), len(want)))
    assert not bad, f"{len(bad)} docs differ, first: {bad[:3]}"


@pytest.mark.parametrize("cfg", WIDE_CONFIGS)
def test_wide_build_matches_oracle(mgx, oracle, cfg):
    docs = phrase_docs(5, 1500, 12, bad=True) + make_docs(6, 300, 30, bad=True, long_every=101)
    ids = np.arange(len(docs), dtype=np.uint32) * 3 + 100
    gi, oi = build_pair(mgx, oracle, docs, ids, cfg)
    assert_same_index(gi, oi)
    s = gi.stats()
    assert s.key_width == max(cfg[0], cfg[1] if cfg[1] > 0 else cfg[0])
    assert (s.total_doc_length, s.doc_count) == oi.bm25_stats()
    want_len = np.array([oracle.count_code_points(d) for d in docs], dtype=np.uint32)
    assert np.array_equal(gi.doc_lengths(), want_len)
    rnd = random.Random(1)
    for t in some_terms(oi, rnd, 30) + [b"ab", "東".encode(), b"", b"\xff", b"abcdefghijklmnop"]:
        assert gi.posting_size(t) == oi.posting_size(t), t
        assert np.array_equal(gi.postings(t), oi.search_and([t])), t


@pytest.mark.parametrize("cfg", [(4, 0, True), (5, 2, True), (2, 6, True), (10, 0, True)])
def test_wide_set_calls(mgx, oracle, cfg):
    docs = phrase_docs(21, 3000, 10)
    ids = np.arange(1, len(docs) + 1, dtype=np.uint32)
    gi, oi = build_pair(mgx, oracle, docs, ids, cfg, dense_threshold=0.02)
    rnd = random.Random(3)
    for it in range(50):
        terms = some_terms(oi, rnd, rnd.randint(1, 4))
        if rnd.random() < 0.2:
            terms.append("龘龘龘龘龘".encode())  # unknown n-gram
        if rnd.random() < 0.2:
            terms.append(terms[0])
        for limit, reverse in ((0, False), (5, False), (5, True)):
            assert np.array_equal(gi.search_and(terms, limit, reverse), oi.search_and(terms, limit, reverse)), terms
        assert np.array_equal(gi.search_or(terms), oi.search_or(terms)), ("or", terms)
        all_docs = np.concatenate([ids[::2], np.array([900000], np.uint32)])
        assert np.array_equal(gi.search_not(all_docs, terms), oi.search_not(all_docs, terms)), ("not", terms)
        cands = np.array([ids[rnd.randrange(len(ids))] for _ in range(rnd.randint(0, 50))], dtype=np.uint32)
        assert np.array_equal(gi.filter_by_ngrams(cands, terms), oi.filter_by_ngrams(cands, terms)), ("filter", terms)
    grams = some_terms(oi, rnd, 40)
    for _ in range(30):
        ts = [rnd.choice(grams) for _ in range(rnd.randint(1, 6))]
        if rnd.random() < 0.3:
            ts.append(ts[0])
        for thr in range(0, len(set(ts)) + 2):
            assert np.array_equal(gi.search_by_threshold(ts, thr), oi.search_by_threshold(ts, thr)), (ts, thr)
    for _ in range(80):
        n_terms = rnd.randint(1, 4)
        terms = [sample_terms(docs, rnd, 12) for _ in range(n_terms)]
        ops, args = random_program(rnd, n_terms)
        assert np.array_equal(gi.eval_boolean(ops, args, terms), oi.eval_boolean(ops, args, terms)), (ops, args, terms)


@pytest.mark.parametrize("cfg", [(4, 0, True), (5, 2, True), (4, 4, False), (2, 6, True), (7, 3, False), (10, 0, True)])
def test_wide_query_batch_matches_oracle(mgx, oracle, cfg):
    docs = phrase_docs(33, 4000, 10)
    ids = np.arange(1, len(docs) + 1, dtype=np.uint32)
    gi, oi = build_pair(mgx, oracle, docs, ids, cfg, dense_threshold=0.02)
    rnd = random.Random(9)
    qs = [[sample_terms(docs, rnd, 14) for _ in range(rnd.randint(1, 3))] for _ in range(300)]
    qs += [[b""], [b"a"], [b"zzzzzzzzzzzz"], [], ["東京都".encode(), b""], [b"search", b"search"]]
    nots = [[sample_terms(docs, rnd, 8)] if rnd.random() < 0.3 else [] for _ in qs]
    for kw in (dict(score=True, descending=True, limit=100, offset=0),
               dict(score=True, descending=False, limit=7, offset=3),
               dict(score=False, limit=20, offset=2),
               dict(score=True, descending=True, limit=10, offset=0, verify_text=1),
               dict(score=False, limit=50, offset=0, verify_text=2)):
        g = gi.query_batch(qs, not_terms=nots, **kw)
        o = oi.query_batch(qs, not_terms=nots, **kw)
        assert_batch_equal(g, o, qs)
    assert int(g.total.sum()) > 1000


def test_wide_large_batch_compiled_by_several_threads(mgx, oracle, monkeypatch):
    """Batches of >= 2048 queries may be compiled by several host threads, each with its own pool of wide keys."""
    monkeypatch.setenv("MGX_COMPILE_THREADS", "3")
    cfg = (5, 0, True)
    docs = phrase_docs(41, 3000, 10)
    ids = np.arange(1, len(docs) + 1, dtype=np.uint32)
    gi, oi = build_pair(mgx, oracle, docs, ids, cfg)
    rnd = random.Random(2)
    qs = [[sample_terms(docs, rnd, 12) for _ in range(rnd.randint(1, 2))] for _ in range(3000)]
    g = gi.query_batch(qs, score=True, limit=10)
    o = oi.query_batch(qs, score=True, limit=10)
    assert_batch_equal(g, o, qs)


@pytest.mark.parametrize("cfg", [(4, 0, True), (2, 6, True)])
def test_wide_mutations(mgx, oracle, cfg):
    rnd = random.Random(77)
    base = phrase_docs(88, 1200, 10)
    docs = {i: d for i, d in zip(range(5, 5 + 2 * len(base), 2), base)}
    gi = mgx.Index(*cfg)
    oi = oracle.index(*cfg)
    ids0 = np.asarray(sorted(docs), dtype=np.uint32)
    gi.add_document_batch(ids0, [docs[int(i)] for i in ids0])
    oi.add_texts(ids0, [docs[int(i)] for i in ids0])
    for round_ in range(2):
        for _ in range(150):
            r = rnd.random()
            if r < 0.4:
                new_id = rnd.choice([rnd.randrange(1, 5), 2 * rnd.randrange(3, 1200), 4000 + rnd.randrange(1000)])
                if new_id in docs:
                    continue
                text = phrase_docs(rnd.randrange(1 << 30), 1, 10)[0]
                docs[new_id] = text
                assert gi.add_document(new_id, text) == bool(oi.add_document(new_id, text))
            elif r < 0.7:
                doc_id = rnd.choice(sorted(docs))
                text = phrase_docs(rnd.randrange(1 << 30), 1, 10)[0]
                gi.update_document(doc_id, docs[doc_id], text)
                oi.update_document(doc_id, docs[doc_id], text)
                docs[doc_id] = text
            else:
                doc_id = rnd.choice(sorted(docs))
                gi.remove_document(doc_id, docs[doc_id])
                oi.remove_document(doc_id, docs[doc_id])
                del docs[doc_id]
        assert_same_index(gi, oi)
        live = list(docs.values())
        qs = [[sample_terms(live, rnd, 12)] for _ in range(100)]
        assert_batch_equal(gi.query_batch(qs, score=True, limit=20), oi.query_batch(qs, score=True, limit=20), qs)


@pytest.mark.parametrize("cfg", [(4, 0, True), (5, 2, True)])
def test_wide_fuzzy_synonyms_and_mgix(mgx, oracle, cfg):
    from test_oracle_expanded import expanded_cases, spaced_docs
    rnd = random.Random(0x77)
    docs = spaced_docs(rnd, 2000)
    ids = np.arange(1, len(docs) + 1, dtype=np.uint32)
    gi, oi = build_pair(mgx, oracle, docs, ids, cfg)
    nonempty = 0
    for fuzzy_terms, groups, nots, dist in expanded_cases(rnd, docs, 60):
        for vt in (0, 1):
            want, _ = oi.search_fuzzy(fuzzy_terms, dist, nots, verify_text=vt)
            assert np.array_equal(gi.search_fuzzy(fuzzy_terms, dist, nots, verify_text=vt), want), (fuzzy_terms, dist, vt)
            nonempty += want.size > 0
            want, _ = oi.search_synonyms(groups, nots, verify_text=vt)
            assert np.array_equal(gi.search_synonyms(groups, nots, verify_text=vt), want), (groups, nots, vt)
            nonempty += want.size > 0
    assert nonempty > 10
    # MGIX stream of a wide-key index: decodes to the oracle's CSR and loads back into a device index
    terms, offs, posts = oi.export()
    stream = gi.save_mgix()
    meta, t2, o2, p2 = mgx.mgix_decode(stream)
    assert t2 == terms and np.array_equal(o2, offs) and np.array_equal(p2, posts)
    g2 = mgx.Index(*cfg)
    g2.load_mgix(stream)
    t3, o3, p3 = g2.export()
    assert t3 == terms and np.array_equal(o3, offs) and np.array_equal(p3, posts)
    assert g2.save_mgix() == stream
    for _ in range(30):
        pick = [terms[rnd.randrange(len(terms))] for _ in range(rnd.randint(1, 3))]
        assert np.array_equal(g2.search_and(pick), oi.search_and(pick)), pick
        assert g2.posting_size(pick[0]) == oi.posting_size(pick[0])


def test_ngram_sizes_out_of_range_are_refused(mgx):
    for ng, kj in ((0, 0), (11, 0), (2, 11), (-1, 2)):
        with pytest.raises(mgx.MgxError):
            mgx.Index(ng, kj, True)
