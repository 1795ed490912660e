#!/usr/bin/env python
"""oracle/gen_golden.py — TEST INFRASTRUCTURE ONLY.

Generates tests/golden/ref_tokenizer.json and tests/golden/ref_pipeline.json by RUNNING THE
REFERENCE'S OWN CODE (oracle/_ref/libmygram_ref.so = the unmodified sources under
/root/reference/src compiled with the shims in oracle/shim/, see oracle/Makefile) on seeded
inputs. The GPU box has no /root/reference, so the outputs are committed as small fixtures;
tests/test_golden_fixtures.py replays them against the oracle (CPU) and the CUDA path (gpu).

Usage (in the build container, where /root/reference exists):
    make -C oracle ref && python oracle/gen_golden.py
"""
import base64
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from pyoracle import REF_LIB, OracleLib  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CJK = [chr(0x4E00 + i) for i in range(24)] + ["㐀", "豈", "\U00020000", "\U0002b820"]
OTHER = ["あ", "い", "ア", "ー", "한", "😀", "é"]
ASCII = list("abcdefg xy01.")  # inputs are pre-normalised (lower case): Index::NormalizeText stays on the host
BAD = [b"\xff", b"\x80", b"\xe6", b"\xc0\xaf", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xe6\x9d", b"\xf0\x9f\x98", b"\x00"]
CONFIGS = [(2, 1, True), (2, 2, True), (3, 2, False), (1, 1, True), (3, 3, True), (2, 1, False), (1, 2, True)]


def rand_text(rnd, max_units, bad):
    units = []
    for _ in range(rnd.randint(0, max_units)):
        r = rnd.random()
        if bad and r < 0.1:
            units.append(rnd.choice(BAD))
        elif r < 0.45:
            units.append(rnd.choice(CJK).encode())
        elif r < 0.6:
            units.append(rnd.choice(OTHER).encode())
        else:
            units.append(rnd.choice(ASCII).encode())
    return b"".join(units)


def b64(x):
    return base64.b64encode(x).decode("ascii")


def main():
    R = OracleLib(REF_LIB)
    rnd = random.Random(20261018)

    # ---- tokenizer
    texts = [rand_text(rnd, 14, True) for _ in range(160)] + [rand_text(rnd, 700, True) for _ in range(3)]
    tok = {"_generated_by": "oracle/gen_golden.py from the reference's own sources (oracle/_ref)", "texts": [b64(t) for t in texts],
           "codepoints": [R.utf8_to_codepoints(t) for t in texts],
           "count_code_points": [R.count_code_points(t) for t in texts], "hybrid": []}
    for (a, k, c) in CONFIGS:
        tok["hybrid"].append({"a": a, "k": k, "cross": c,
                              "ngrams": [[b64(g) for g in R.ngrams("hybrid", t, a, k, c)] for t in texts]})
    tok["query"] = []
    for (a, k, c) in [(2, 0, True), (0, 0, True), (2, 1, True), (0, 2, False), (3, 0, True), (1, 0, False)]:
        tok["query"].append({"a": a, "k": k, "cross": c,
                             "ngrams": [[b64(g) for g in R.ngrams("query", t, a, k, c)] for t in texts[:80]]})
    json.dump(tok, open(os.path.join(OUT, "ref_tokenizer.json"), "w"), indent=0)

    # ---- index + pipeline
    docs = [rand_text(rnd, 18, False) for _ in range(600)]
    ids = (np.arange(len(docs), dtype=np.uint32) * 2 + 11)
    pipe = {"_generated_by": tok["_generated_by"], "docs": [b64(d) for d in docs], "ids": ids.tolist(), "cases": []}
    for (ng, kj, cross) in [(2, 0, True), (2, 1, True), (3, 2, False), (1, 1, True)]:
        idx = R.index(ng, kj, cross)
        idx.add_texts(ids, docs)
        case = {"ngram": ng, "kanji": kj, "cross": cross, "term_count": idx.term_count(),
                "total_postings": idx.total_postings(), "bm25_stats": list(idx.bm25_stats())}
        # postings of sampled n-grams (generated through the reference tokenizer)
        eff = kj if kj > 0 else ng
        grams = sorted({g for d in docs[:60] for g in R.ngrams("hybrid", d, ng, eff, cross)})
        grams = rnd.sample(grams, min(40, len(grams))) + [b"\xe9\xbe\x98\xe9\xbe\x98", b""]
        case["postings"] = [{"term": b64(g), "docs": idx.postings(g).tolist()} for g in grams]
        # set algebra
        sets = []
        for _ in range(25):
            terms = rnd.sample(grams, rnd.randint(1, 3))
            cands = sorted(rnd.sample(ids.tolist(), 30)) if rnd.random() < 0.5 else rnd.sample(ids.tolist(), 30)
            sets.append({"terms": [b64(t) for t in terms],
                         "and": idx.search_and(terms).tolist(),
                         "and_top3_rev": idx.search_and(terms, 3, True).tolist(),
                         "or": idx.search_or(terms).tolist(),
                         "not": idx.search_not(ids[::3], terms).tolist(),
                         "cands": cands, "filter": idx.filter_by_ngrams(cands, terms).tolist()})
        case["sets"] = sets
        # pipeline queries
        queries, nots = [], []
        for _ in range(60):
            t = docs[rnd.randrange(len(docs))].decode("utf-8")
            if len(t) < 3:
                continue
            terms = []
            for _ in range(rnd.randint(1, 3)):
                ln = rnd.randint(1, 4)
                st = rnd.randrange(0, max(1, len(t) - ln + 1))
                terms.append(t[st:st + ln].encode())
            queries.append(terms)
            t2 = docs[rnd.randrange(len(docs))].decode("utf-8")
            nots.append([t2[:2].encode()] if (len(t2) >= 2 and rnd.random() < 0.3) else [])
        case["queries"] = [[b64(t) for t in q] for q in queries]
        case["not_terms"] = [[b64(t) for t in q] for q in nots]
        case["runs"] = []
        for kw in (dict(score=True, descending=True, limit=10, offset=0, verify_text=0),
                   dict(score=True, descending=False, limit=5, offset=2, verify_text=0),
                   dict(score=False, descending=True, limit=20, offset=0, verify_text=1)):
            r = idx.query_batch(queries, not_terms=nots, want_sets=True, **kw)
            case["runs"].append({"params": kw, "total": r.total.tolist(), "df": r.df.tolist(),
                                 "count": r.count.tolist(),
                                 "ids": [r.ids[q, :r.count[q]].tolist() for q in range(len(queries))],
                                 "scores": [[float.hex(float(x)) for x in r.scores[q, :r.count[q]]] for q in range(len(queries))],
                                 "sets": [s.tolist() for s in r.sets]})
        pipe["cases"].append(case)
    json.dump(pipe, open(os.path.join(OUT, "ref_pipeline.json"), "w"), indent=0)
    for f in ("ref_tokenizer.json", "ref_pipeline.json"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


def gen_mgix():
    """tests/golden/ref_mgix.json: Index::SaveToStream output of the reference's own sources for two small indexes
    (one with a list above 4096 entries, i.e. a Roaring body with a bitset container), with the CSR the reference
    itself reports (PostingList::GetAll per term)."""
    import zlib
    from pyoracle import PORT_LIB
    ref, port = OracleLib(REF_LIB), OracleLib(PORT_LIB)
    rnd = random.Random(0x316)
    words = ["".join(rnd.choice("abcde") for _ in range(3)) for _ in range(20)]
    big = [("zq" + rnd.choice(words)).encode() for _ in range(4500)]
    small = [rand_text(rnd, 12, False) for _ in range(150)]
    out = {"generator": "oracle/gen_golden.py mgix", "cases": []}
    for cfg, docs, first in (((2, 0, True), big, 1), ((2, 1, True), small, 100)):
        ids = np.arange(first, first + len(docs), dtype=np.uint32)
        ri, pi = ref.index(*cfg), port.index(*cfg)
        ri.add_texts(ids, docs)
        pi.add_texts(ids, docs)
        terms = [bytes(t) for t in pi.export()[0]]  # the reference has no term enumeration besides the stream itself
        assert len(terms) == ri.term_count()
        posts = [ri.postings(t) for t in terms]
        offs = np.concatenate([[0], np.cumsum([p.size for p in posts])]).astype(np.uint64)
        flat = np.concatenate(posts).astype(np.uint32)
        out["cases"].append({"config": [cfg[0], cfg[1] if cfg[1] > 0 else cfg[0], int(cfg[2])],
                             "stream_b64": base64.b64encode(ri.save_stream()).decode(),
                             "terms_hex": [t.hex() for t in terms], "posting_offsets": offs.tolist(),
                             "n_postings": int(flat.size), "postings_crc32": zlib.crc32(flat.tobytes()),
                             "largest_list": int(max(p.size for p in posts))})
    path = os.path.join(OUT, "ref_mgix.json")
    json.dump(out, open(path, "w"), indent=0)
    print("ref_mgix.json", os.path.getsize(path), "bytes")


def gen_expanded():
    """tests/golden/ref_expanded.json: results of the reference's own ExecuteWithFuzzy / ExecuteWithSynonyms (through
    oracle/_ref) over the documents of ref_pipeline.json, for the configurations of its cases."""
    ref = OracleLib(REF_LIB)
    pipe = json.load(open(os.path.join(OUT, "ref_pipeline.json")))
    docs = [base64.b64decode(d) for d in pipe["docs"]]
    ids = np.array(pipe["ids"], dtype=np.uint32)
    rnd = random.Random(0xE5)
    texts = [d.decode("utf-8", "ignore") for d in docs]
    texts = [t for t in texts if len(t) >= 4]

    def piece(lo, hi):
        t = rnd.choice(texts)
        ln = rnd.randint(lo, min(hi, len(t)))
        st = rnd.randrange(0, len(t) - ln + 1)
        return t[st:st + ln]

    def misspell(s):
        if rnd.random() < 0.3:
            return s
        i = rnd.randrange(len(s))
        return s[:i] + rnd.choice(["x", "東", ""]) + s[i + 1:]

    out = {"generator": "oracle/gen_golden.py expanded", "cases": []}
    for case in pipe["cases"]:
        cfg = (case["ngram"], case["kanji"], case["cross"])
        idx = ref.index(*cfg)
        idx.add_texts(ids, docs)
        fuzzy, syn = [], []
        for _ in range(60):
            terms = [misspell(piece(2, 7)) for _ in range(rnd.randint(1, 2))]
            nots = [piece(1, 3)] if rnd.random() < 0.25 else []
            dist, vt = rnd.randint(0, 2), rnd.randrange(3)
            r, empty = idx.search_fuzzy(terms, dist, nots, verify_text=vt)
            fuzzy.append({"terms": [b64(t.encode()) for t in terms], "not": [b64(t.encode()) for t in nots],
                          "distance": dist, "verify_text": vt, "ids": r.tolist(), "empty_term": empty})
            groups = [[piece(1, 4) for _ in range(rnd.randint(1, 3))] for _ in range(rnd.randint(1, 3))]
            r, empty = idx.search_synonyms(groups, nots, verify_text=vt)
            syn.append({"groups": [[b64(v.encode()) for v in g] for g in groups], "not": [b64(t.encode()) for t in nots],
                        "verify_text": vt, "ids": r.tolist(), "empty_term": empty})
        out["cases"].append({"ngram": cfg[0], "kanji": cfg[1], "cross": cfg[2], "fuzzy": fuzzy, "synonyms": syn})
    path = os.path.join(OUT, "ref_expanded.json")
    json.dump(out, open(path, "w"), indent=0)
    print("ref_expanded.json", os.path.getsize(path), "bytes,",
          sum(1 for c in out["cases"] for q in c["fuzzy"] + c["synonyms"] if q["ids"]), "non-empty answers")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mgix":
        gen_mgix()
    elif len(sys.argv) > 1 and sys.argv[1] == "expanded":
        gen_expanded()
    else:
        main()
        gen_mgix()
        gen_expanded()
