#!/usr/bin/env python
"""tools/fuzz_gpu.py — randomized GPU-vs-oracle comparison over many small corpora, index configurations and
query shapes (run on a B200 box: `python tools/fuzz_gpu.py [seconds] [seed]`). Not part of the test-suite; it is
the shake-out used while developing the positional df shortcut, boolean programs, mutations, filters, the fuzzy /
synonym paths and the MGIX export."""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "support")):
    sys.path.insert(0, p)
import mgx_loader  # noqa: E402
import pyoracle  # noqa: E402
import test_gpu_parity as T  # noqa: E402  (document / query generators and comparators)

mgx = mgx_loader.load()
oracle = pyoracle.OracleLib(pyoracle.PORT_LIB)
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t_end = time.time() + budget
it = 0
while time.time() < t_end:
    seed = seed0 * 100003 + it
    rnd = random.Random(seed)
    cfg = rnd.choice(T.CONFIGS)
    bad = rnd.random() < 0.3
    n = rnd.choice([1, 7, 200, 3000])
    docs = T.make_docs(seed, n, rnd.choice([3, 12, 40]), bad=bad, long_every=rnd.choice([0, 0, 97]))
    if rnd.random() < 0.3:   # tiny alphabet: repeated n-grams inside documents and terms
        docs = [bytes(rnd.choice(b"ab") for _ in range(rnd.randint(0, 30))) for _ in range(n)]
    first = rnd.choice([1, 5, 1000])
    step = rnd.choice([1, 1, 3])
    ids = (np.arange(n, dtype=np.uint32) * step + first).astype(np.uint32)
    for mode in ("tiles", "stream", None):
        if mode is None:
            os.environ.pop("MGX_DF_MODE", None)
        else:
            os.environ["MGX_DF_MODE"] = mode
        gi, oi = T.build_pair(mgx, oracle, docs, ids, cfg, dense_threshold=rnd.choice([0.0, 0.02, 0.5]))
        T.assert_same_index(gi, oi)
        qs = T.sample_queries_from_docs([d for d in docs if len(d.decode("utf-8", "ignore")) >= 3] or [b"abc"], rnd, 60)
        qs += [[b""], [b"a"], [b"aaa"], [b"abab", b"ba"], []]
        nots = [[rnd.choice(docs)[:3]] if rnd.random() < 0.2 else [] for _ in qs]
        kw = rnd.choice([dict(score=True, limit=100), dict(score=True, descending=False, limit=7, offset=2),
                         dict(score=False, limit=30, offset=1), dict(score=True, limit=20, verify_text=1),
                         dict(score=True, limit=20, verify_text=2), dict(score=True, limit=1000, offset=1100),
                         dict(score=True, descending=False, limit=300, offset=2000)])
        try:
            T.assert_batch_equal(gi.query_batch(qs, not_terms=nots, **kw), oi.query_batch(qs, not_terms=nots, **kw), qs)
        except AssertionError:
            print("MISMATCH seed", seed, "cfg", cfg, "mode", mode, "kw", kw, "bad", bad, "n", n, flush=True)
            raise
    # boolean programs + threshold on the last index
    grams = T.some_terms(oi, rnd, 10) if oi.term_count() else []
    for _ in range(10):
        if grams:
            ts = [rnd.choice(grams) for _ in range(rnd.randint(1, 4))]
            thr = rnd.randint(0, len(ts) + 1)
            assert np.array_equal(gi.search_by_threshold(ts, thr), oi.search_by_threshold(ts, thr)), (seed, ts, thr)
            assert np.array_equal(gi.search_or(ts), oi.search_or(ts)), ("or", seed, ts)
        nt = rnd.randint(1, 3)
        terms = []
        for _t in range(nt):
            src = rnd.choice(docs).decode("utf-8", "ignore")
            ln = rnd.randint(1, 4)
            st = rnd.randrange(0, max(1, len(src) - ln + 1))
            terms.append(src[st:st + ln].encode() if src else b"q")
        ops, args = T.random_program(rnd, nt)
        g, o = gi.eval_boolean(ops, args, terms), oi.eval_boolean(ops, args, terms)
        assert np.array_equal(g, o), ("boolean", seed, cfg, ops, args, terms, g[:8], o[:8])
    # fuzzy / synonym execution paths (ExecuteWithFuzzy / ExecuteWithSynonyms) and the MGIX stream of the index
    import test_oracle_expanded as X
    for fuzzy_terms, groups, nots2, dist in X.expanded_cases(rnd, docs, 12):
        vt = rnd.randrange(3)
        g, o = gi.search_fuzzy(fuzzy_terms, dist, nots2, verify_text=vt), oi.search_fuzzy(fuzzy_terms, dist, nots2, verify_text=vt)[0]
        assert np.array_equal(g, o), ("fuzzy", seed, cfg, fuzzy_terms, dist, nots2, vt, g[:8], o[:8])
        g, o = gi.search_synonyms(groups, nots2, verify_text=vt), oi.search_synonyms(groups, nots2, verify_text=vt)[0]
        assert np.array_equal(g, o), ("synonyms", seed, cfg, groups, nots2, vt, g[:8], o[:8])
    meta, mt, mo, mp = mgx.mgix_decode(gi.save_mgix())
    ot, oo, op_ = oi.export()
    assert [bytes(t) for t in ot] == mt and np.array_equal(mo, oo) and np.array_equal(mp, op_), ("mgix", seed, cfg)
    # ... and loaded into a second device index (Index::LoadFromStream): same CSR, same set answers
    g2 = mgx.Index(*cfg)
    g2.load_mgix(gi.save_mgix())
    lt, lo_, lp = g2.export()
    assert lt == mt and np.array_equal(lo_, oo) and np.array_equal(lp, op_), ("mgix load", seed, cfg)
    for _ in range(5):
        if grams:
            ts = [rnd.choice(grams) for _ in range(rnd.randint(1, 3))]
            assert np.array_equal(g2.search_and(ts), oi.search_and(ts)), ("loaded and", seed, ts)
            assert np.array_equal(g2.search_or(ts), oi.search_or(ts)), ("loaded or", seed, ts)
    g2.close()
    # mutations: a burst of add / update / remove, then compare postings, stats and a batch
    if n >= 7 and not bad:
        os.environ.pop("MGX_DF_MODE", None)
        live = {int(i): d for i, d in zip(ids, docs)}
        for _ in range(rnd.randint(1, 40)):
            r = rnd.random()
            if r < 0.4:
                new_id = rnd.randrange(1, int(ids[-1]) + 50)
                if new_id in live:
                    continue
                text = T.rand_text(rnd, 20)
                live[new_id] = text
                assert gi.add_document(new_id, text) == bool(oi.add_document(new_id, text))
            elif r < 0.7 and live:
                d = rnd.choice(sorted(live))
                text = T.rand_text(rnd, 20)
                gi.update_document(d, live[d], text)
                oi.update_document(d, live[d], text)
                live[d] = text
            elif live:
                d = rnd.choice(sorted(live))
                gi.remove_document(d, live[d])
                oi.remove_document(d, live[d])
                del live[d]
        T.assert_same_index(gi, oi)
        st = gi.stats()
        assert (st.total_doc_length, st.doc_count) == oi.bm25_stats(), seed
        pool = [d for d in live.values() if len(d.decode("utf-8", "ignore")) >= 3]
        if pool:
            qs2 = T.sample_queries_from_docs(pool, rnd, 40)
            T.assert_batch_equal(gi.query_batch(qs2, score=True, limit=20), oi.query_batch(qs2, score=True, limit=20), qs2)
        # filters on the mutated index (columns follow the index's document order = ascending ids)
        order = sorted(live)
        ncur = len(order)
        if ncur:
            cols = [(8, [None if rnd.random() < 0.2 else rnd.randint(-3, 3) for _ in range(ncur)]),
                    (11, [None if rnd.random() < 0.2 else rnd.choice([b"a", b"b", b"ab", b""]) for _ in range(ncur)]),
                    (12, [None if rnd.random() < 0.2 else rnd.choice([0.0, 1.5, -2.0]) for _ in range(ncur)]),
                    (1, [None if rnd.random() < 0.2 else rnd.random() < 0.5 for _ in range(ncur)])]
            for ci, (typ, vals) in enumerate(cols):
                gi.set_filter_column(ci, typ, vals)
            lits = ["1", "-2", "a", "ab", "", "1.5", "true", "0", "x", "-2.0"]
            fl = [[(rnd.randrange(5), rnd.randrange(6) if rnd.random() < 0.5 else rnd.randrange(2), rnd.choice(lits))
                   for _ in range(rnd.randint(1, 2))] for _ in range(20)]
            grams2 = T.some_terms(oi, rnd, 20) if oi.term_count() else []
            if grams2:
                qf = [[rnd.choice(grams2)] for _ in fl]
                g = gi.query_batch(qf, filters=fl, score=False, limit=1000, raw_ngram=None)
                for qi, q in enumerate(qf):
                    full = oi.query_batch([q], score=False, limit=0, stride=1, want_sets=True).sets[0]
                    # rows of the filter columns are positions in the ascending id order: map ids -> rows
                    rows = np.searchsorted(np.asarray(order, dtype=np.uint32), full).astype(np.uint32)
                    keep = oracle.apply_filters(ncur, 0, cols, fl[qi], rows)
                    want = np.asarray(order, dtype=np.uint32)[keep]
                    assert int(g.total[qi]) == want.size, ("filter", seed, q, fl[qi], int(g.total[qi]), want.size)
                    k = min(1000, want.size)
                    assert np.array_equal(g.ids[qi, :k], want[:k]), ("filter ids", seed, q, fl[qi])
    it += 1
print(f"fuzz ok: {it} iterations in {budget:.0f} s (seed {seed0})")
