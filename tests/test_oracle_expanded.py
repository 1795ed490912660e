"""Fuzzy and synonym execution paths (search_pipeline::ExecuteWithFuzzy / ExecuteWithSynonyms): the
oracle's restatement against the reference's own functions (oracle/_ref) on random corpora, every n-gram
configuration, NOT terms, verify_text modes, terms too short for an n-gram, unknown terms, empty groups."""
import random

import numpy as np
import pytest

from test_oracle_bulk import _docs

CONFIGS = [(2, 0, True), (2, 1, True), (2, 1, False), (3, 2, False), (1, 1, True), (3, 0, True)]


def expanded_cases(rnd, docs, n):
    """Random fuzzy term lists, synonym groups and NOT lists cut from the corpus (plus misspellings and strangers)."""
    def piece(lo=1, hi=7):
        t = docs[rnd.randrange(len(docs))].decode()
        if not t:
            return "ab"
        ln = rnd.randint(lo, hi)
        st = rnd.randrange(0, max(1, len(t) - ln + 1))
        return t[st:st + ln]

    def misspell(s):
        if len(s) < 2 or rnd.random() < 0.4:
            return s
        i = rnd.randrange(len(s))
        return s[:i] + rnd.choice(["x", "東", ""]) + s[i + 1:]

    cases = []
    for _ in range(n):
        fuzzy_terms = [misspell(piece(2, 8)) for _ in range(rnd.randint(1, 3))]
        groups = [[piece(1, 4) for _ in range(rnd.randint(0 if rnd.random() < 0.05 else 1, 3))]
                  for _ in range(rnd.randint(1, 3))]
        if rnd.random() < 0.1:
            groups[0].append("zzzzqq")  # unknown variant
        if rnd.random() < 0.05:
            groups[-1].append("")
        nots = [piece(1, 3) for _ in range(rnd.randint(0, 2))] if rnd.random() < 0.4 else []
        cases.append((fuzzy_terms, groups, nots, rnd.randint(0, 2)))
    return cases


@pytest.mark.parametrize("cfg", CONFIGS)
def test_fuzzy_and_synonyms_agree_with_reference_sources(oracle, reflib, cfg):
    rnd = random.Random(0xF0 + (hash(cfg) & 0xFFF))
    docs = _docs(rnd, 1200)
    ids = np.arange(1, len(docs) + 1, dtype=np.uint32)
    pi, ri = oracle.index(*cfg), reflib.index(*cfg)
    pi.add_texts(ids, docs)
    ri.add_texts(ids, docs)
    nonempty = 0
    for fuzzy_terms, groups, nots, dist in expanded_cases(rnd, docs, 150):
        a = pi.search_fuzzy(fuzzy_terms, dist, nots)
        b = ri.search_fuzzy(fuzzy_terms, dist, nots)
        assert a[1] == b[1] and np.array_equal(a[0], b[0]), (fuzzy_terms, dist, nots)
        nonempty += len(a[0]) > 0
        for vt in (0, 1, 2):
            a = pi.search_synonyms(groups, nots, verify_text=vt)
            b = ri.search_synonyms(groups, nots, verify_text=vt)
            assert a[1] == b[1] and np.array_equal(a[0], b[0]), (groups, nots, vt)
            nonempty += len(a[0]) > 0
    assert nonempty > 50
    # the documented corner cases
    a, b = pi.search_fuzzy([], 1), ri.search_fuzzy([], 1)
    assert a[1] is True and b[1] is True and len(a[0]) == 0 and len(b[0]) == 0
    a, b = pi.search_synonyms([]), ri.search_synonyms([])
    assert a[1] is True and b[1] is True and len(a[0]) == 0 and len(b[0]) == 0
    # verify_text that applies to the terms needs the edit-distance verification: the restatement declines
    assert pi.search_fuzzy(["ab"], 1, verify_text=1) is None
