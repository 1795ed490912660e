#!/usr/bin/env python
"""tools/bench_expanded.py — measurement of the fuzzy / synonym execution paths and of the MGIX export on the C2 corpus
(SURVEY §8f-3 / f-4 widened to the same parity + measurement bar as the headline path).

  python tools/bench_expanded.py [--docs 10000000] [--queries 200] [--cpu-sample 24] [--out profiles/…json]

Per query class (fuzzy d=1 without / with edit-distance verification, synonym groups): queries/s through the public
single-query calls (mgx_search_fuzzy / mgx_search_synonyms: host buffers in, doc ids out, wall clock after a warm-up),
the CPU oracle timed beside it on a bounded sample of the same queries (one thread, the reference processes one query
per thread), and a parity check of that sample (bit-exact doc-id sets). MGIX: mgx_index_save_mgix on the whole shard,
bytes/s, decoded back and compared with the device CSR. Prints ONE JSON line. The oracle is the checker here, as in
bench.py's cpu_baseline leg; nothing in the product imports it."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "support"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def make_queries(c, rng, n):
    """fuzzy: 1-2 terms of 3-6 code points cut from a document, one character replaced; synonyms: two groups of two
    2-3 code point variants, the first of each group from the same document (non-empty AND of ORs)."""
    def piece(d, lo, hi):
        t = c.text(d).decode()
        ln = int(rng.integers(lo, min(hi, len(t)) + 1))
        st = int(rng.integers(0, len(t) - ln + 1))
        return t[st:st + ln]

    fuzzy, syn = [], []
    for _ in range(n):
        d = int(rng.integers(0, c.n_docs))
        terms = []
        for _ in range(int(rng.integers(1, 3))):
            t = piece(d, 3, 6)
            i = int(rng.integers(0, len(t)))
            terms.append(t[:i] + chr(0x4E00 + int(rng.integers(0, 8192))) + t[i + 1:])
        fuzzy.append(terms)
        syn.append([[piece(d, 2, 3), piece(int(rng.integers(0, c.n_docs)), 2, 3)] for _ in range(2)])
    return fuzzy, syn


def timed(fn, items, warmup=3):
    for it in items[:warmup]:
        fn(it)
    t0 = time.perf_counter()
    out = [fn(it) for it in items]
    return out, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=200)
    ap.add_argument("--cpu-sample", type=int, default=24)
    ap.add_argument("--out", default="")
    ap.add_argument("--gpu-only", action="store_true", help="query timings only: no MGIX export, no CPU oracle")
    args = ap.parse_args()
    import corpus as corpus_mod
    import mgx_loader
    import pyoracle
    m = mgx_loader.load()
    c = corpus_mod.generate("cjk", args.docs, 0xC2)
    gi = m.Index(2, 0, True, device=0)
    gi.build(c.doc_ids, c.arena, c.offsets)
    launches0 = m.lib().mgx_kernel_launch_count()
    rng = np.random.default_rng(0xE7)
    fuzzy, syn = make_queries(c, rng, args.queries)
    out = {"workload": f"C2 corpus, {args.docs} documents, bigram index; single-query calls, host buffers in / doc ids out",
           "classes": {}}

    def flush():
        line = json.dumps(out)
        if args.out:
            with open(args.out, "w") as f:
                f.write(line + "\n")
        return line

    # Index::SearchOr / SearchByThreshold: 3 bigrams of one document (OR) / 6 bigrams, at least 4 of them
    def grams(d, n):
        t = c.text(d).decode()
        return [t[i:i + 2] for i in range(min(n, len(t) - 1))]
    or_qs = [grams(int(rng.integers(0, c.n_docs)), 3) for _ in range(args.queries)]
    thr_qs = [grams(int(rng.integers(0, c.n_docs)), 6) for _ in range(args.queries)]

    def pinned(fn):  # the round-1 plan: one pass over every document of the shard
        def run(q):
            os.environ["MGX_NO_OR_EXPANSION"] = "1"
            try:
                return fn(q)
            finally:
                del os.environ["MGX_NO_OR_EXPANSION"]
        return run

    gpu_classes = [
        ("search_or_3", or_qs, lambda q: gi.search_or(q)),
        ("search_or_3_single_pass", or_qs, pinned(lambda q: gi.search_or(q))),
        ("search_by_threshold_4_of_6", thr_qs, lambda q: gi.search_by_threshold(q, 4)),
        ("search_by_threshold_4_of_6_single_pass", thr_qs, pinned(lambda q: gi.search_by_threshold(q, 4))),
        ("fuzzy_d1", fuzzy, lambda q: gi.search_fuzzy(q, 1)),
        ("fuzzy_d1_verify_all", fuzzy, lambda q: gi.search_fuzzy(q, 1, verify_text=1)),
        ("synonyms_2x2", syn, lambda q: gi.search_synonyms(q)),
        ("synonyms_2x2_verify_all", syn, lambda q: gi.search_synonyms(q, verify_text=1)),
        ("fuzzy_d1_second_pass", fuzzy, lambda q: gi.search_fuzzy(q, 1)),  # the call's buffers have grown to their size
    ]
    results = {}
    for name, qs, gpu_fn in gpu_classes:
        got, dt = timed(gpu_fn, qs)
        results[name] = got
        out["classes"][name] = {"gpu_queries_per_s": len(qs) / dt, "gpu_ms_per_query": 1e3 * dt / len(qs),
                                "result_docs_mean": float(np.mean([g.size for g in got]))}
        flush()
    # the batch forms: the same queries in ONE call each (answers must equal the single calls')
    for name, qs, single, batch_fn in (
            ("fuzzy_d1", fuzzy, "fuzzy_d1", lambda qs: gi.search_fuzzy_batch(qs, 1)),
            ("fuzzy_d1_verify_all", fuzzy, "fuzzy_d1_verify_all", lambda qs: gi.search_fuzzy_batch(qs, 1, verify_text=1)),
            ("synonyms_2x2", syn, "synonyms_2x2", lambda qs: gi.search_synonyms_batch(qs)),
            ("synonyms_2x2_verify_all", syn, "synonyms_2x2_verify_all",
             lambda qs: gi.search_synonyms_batch(qs, verify_text=1))):
        batch_fn(qs)  # (the pooled batch object grows its buffers to this size once)
        t0 = time.perf_counter()
        got = batch_fn(qs)
        dt = time.perf_counter() - t0
        same = all(np.array_equal(a, b) for a, b in zip(got, results[single]))
        out["classes"][name + "_batch"] = {"gpu_queries_per_s": len(qs) / dt, "gpu_ms_per_query": 1e3 * dt / len(qs),
                                            "queries_per_call": len(qs), "equal_to_single_calls": bool(same)}
        flush()
    out["gpu_launches"] = int(m.lib().mgx_kernel_launch_count() - launches0)
    if args.gpu_only:
        print(flush(), flush=True)
        return
    # MGIX export of the whole shard (raw buffers: tens of millions of terms are not turned into Python objects)
    C = m.C
    st = gi.stats()
    t0 = time.perf_counter()
    stream = gi.save_mgix()
    dt = time.perf_counter() - t0
    buf = np.frombuffer(stream, dtype=np.uint8)
    info = m.MgixInfo()
    u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    rc = m.lib().mgx_mgix_decode(buf.ctypes.data_as(u8p), buf.size, C.byref(info), None, None, None, None)
    assert rc == 0, m.lib().mgx_last_error()
    tb = np.zeros(max(1, info.term_bytes), dtype=np.uint8)
    to = np.zeros(info.n_terms + 1, dtype=np.uint64)
    po = np.zeros(info.n_terms + 1, dtype=np.uint64)
    pp = np.zeros(max(1, info.n_postings), dtype=np.uint32)
    t1 = time.perf_counter()
    rc = m.lib().mgx_mgix_decode(buf.ctypes.data_as(u8p), buf.size, C.byref(info), tb.ctypes.data_as(u8p),
                                 to.ctypes.data_as(u64p), po.ctypes.data_as(u64p), pp.ctypes.data_as(u32p))
    decode_s = time.perf_counter() - t1
    assert rc == 0, m.lib().mgx_last_error()
    keys = np.zeros(max(1, st.n_terms), dtype=np.uint64)
    goffs = np.zeros(st.n_terms + 1, dtype=np.uint64)
    gposts = np.zeros(max(1, st.n_postings), dtype=np.uint32)
    assert m.lib().mgx_index_export(gi._h, keys.ctypes.data_as(u64p), goffs.ctypes.data_as(u64p),
                                    gposts.ctypes.data_as(u32p)) == 0
    ok = (info.n_terms == st.n_terms and info.n_postings == st.n_postings and np.array_equal(po, goffs) and
          np.array_equal(pp[:info.n_postings], gposts[:st.n_postings]))
    out["mgix_export"] = {"stream_bytes": len(stream), "save_seconds": dt, "save_GB_per_s": len(stream) / dt / 1e9,
                          "decode_seconds": decode_s, "terms": int(info.n_terms), "postings": int(info.n_postings),
                          "decodes_to_device_csr": bool(ok)}
    flush()
    assert ok
    # Index::LoadFromStream into a second device index; its stream saved again must be the same bytes
    g2 = m.Index(2, 0, True, device=0)
    t2 = time.perf_counter()
    g2.load_mgix(stream)
    load_s = time.perf_counter() - t2
    again = g2.save_mgix()
    out["mgix_import"] = {"load_seconds": load_s, "documents_in_the_stream": int(g2.stats().n_docs),
                          "saved_again_is_identical": bool(again == stream)}
    flush()
    assert again == stream
    g2.close()
    del again
    del stream, buf, tb, to, po, pp, keys, goffs, gposts
    # CPU oracle beside it, on a bounded sample of the same queries
    oi = pyoracle.OracleLib(pyoracle.PORT_LIB).index(2, 0, True)
    t0 = time.perf_counter()
    oi.build_bulk(c.doc_ids, c.arena, c.offsets, os.cpu_count() or 1)
    out["cpu_index_build_s"] = round(time.perf_counter() - t0, 2)
    cpu_classes = {"search_or_3": (or_qs, lambda q: oi.search_or(q)),
                   "search_by_threshold_4_of_6": (thr_qs, lambda q: oi.search_by_threshold(q, 4)),
                   "fuzzy_d1": (fuzzy, lambda q: oi.search_fuzzy(q, 1)[0]),
                   "fuzzy_d1_verify_all": (fuzzy, lambda q: oi.search_fuzzy(q, 1, verify_text=1)[0]),
                   "synonyms_2x2": (syn, lambda q: oi.search_synonyms(q)[0]),
                   "synonyms_2x2_verify_all": (syn, lambda q: oi.search_synonyms(q, verify_text=1)[0])}
    for name, (qs, cpu_fn) in cpu_classes.items():
        sample = qs[:args.cpu_sample]
        want, cdt = timed(cpu_fn, sample, warmup=0)
        same = all(np.array_equal(g, w) for g, w in zip(results[name], want))
        out["classes"][name]["cpu_baseline"] = {
            "value": len(sample) / cdt, "unit": "queries/s", "cores": 1, "kind": "port",
            "sample": f"first {len(sample)} of the {len(qs)} queries, oracle port, one thread"}
        out["classes"][name]["parity_on_sample"] = bool(same)
        flush()
        assert same, f"{name}: GPU result sets differ from the oracle"
    print(flush(), flush=True)


if __name__ == "__main__":
    main()
