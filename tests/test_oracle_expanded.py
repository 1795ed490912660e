"""Fuzzy and synonym execution paths (search_pipeline::ExecuteWithFuzzy / ExecuteWithSynonyms): the
oracle's restatement against the reference's own functions (oracle/_ref) on random corpora, every n-gram
configuration, NOT terms, verify_text modes, terms too short for an n-gram, unknown terms, empty groups."""
import random

import numpy as np
import pytest

from test_oracle_bulk import _docs

CONFIGS = [(2, 0, True), (2, 1, True), (2, 1, False), (3, 2, False), (1, 1, True), (3, 0, True)]


def spaced_docs(rnd, n):
    """The random corpus of test_oracle_bulk with the word separators ContainsFuzzyMatch knows (tab, newline, U+3000,
    U+00A0) mixed in, so the edit-distance verification meets ASCII words, CJK words and mixed ones."""
    seps = [b" ", b"\t", b"\n", "\u3000".encode(), "\u00a0".encode()]
    docs = []
    for d in _docs(rnd, n):
        parts = d.split(b" ")
        out = parts[0]
        for part in parts[1:]:
            out += rnd.choice(seps) + part
        docs.append(out)
    return docs


def expanded_cases(rnd, docs, n):
    """Random fuzzy term lists, synonym groups and NOT lists cut from the corpus (plus misspellings and strangers)."""
    def piece(lo=1, hi=7):
        t = docs[rnd.randrange(len(docs))].decode("utf-8", "ignore")
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
    docs = spaced_docs(rnd, 1200)
    ids = np.arange(1, len(docs) + 1, dtype=np.uint32)
    pi, ri = oracle.index(*cfg), reflib.index(*cfg)
    pi.add_texts(ids, docs)
    ri.add_texts(ids, docs)
    nonempty = verified_nonempty = 0
    for fuzzy_terms, groups, nots, dist in expanded_cases(rnd, docs, 150):
        for vt in (0, 1, 2):  # 1 / 2: PostFilterByFuzzyText with ContainsFuzzyMatch (utils/edit_distance.cpp)
            a = pi.search_fuzzy(fuzzy_terms, dist, nots, verify_text=vt)
            b = ri.search_fuzzy(fuzzy_terms, dist, nots, verify_text=vt)
            assert a[1] == b[1] and np.array_equal(a[0], b[0]), (fuzzy_terms, dist, nots, vt)
            nonempty += len(a[0]) > 0
            verified_nonempty += vt == 1 and len(a[0]) > 0
        for vt in (0, 1, 2):
            a = pi.search_synonyms(groups, nots, verify_text=vt)
            b = ri.search_synonyms(groups, nots, verify_text=vt)
            assert a[1] == b[1] and np.array_equal(a[0], b[0]), (groups, nots, vt)
            nonempty += len(a[0]) > 0
    assert nonempty > 50 and verified_nonempty > 5
    # the documented corner cases
    a, b = pi.search_fuzzy([], 1), ri.search_fuzzy([], 1)
    assert a[1] is True and b[1] is True and len(a[0]) == 0 and len(b[0]) == 0
    a, b = pi.search_synonyms([]), ri.search_synonyms([])
    assert a[1] is True and b[1] is True and len(a[0]) == 0 and len(b[0]) == 0


# ContainsFuzzyMatch known answers transcribed from the reference's own unit tests
# (tests/utils/edit_distance_test.cpp, line of each EXPECT in the last column)
FUZZY_KAT = [
    ("hello world", "hello", 0, True, 86), ("hello world", "hallo", 1, True, 90), ("hello world", "xxxxx", 1, False, 94),
    ("the quick brown fox", "quikc", 2, True, 99), ("", "hello", 1, False, 103), ("hello world", "", 1, True, 108),
    ("a", "", 1, True, 109), ("restaurant", "restrant", 2, True, 114), ("東京都 大阪府", "東京市", 1, True, 119),
    ("私は東京都に住む", "東京市", 1, True, 123), ("私は東西京に住む", "東京市", 1, False, 127),
    ("concatenate", "cat", 0, False, 131), ("alpha beta gamma", "zzzzz", 2, False, 135),
    ("a b c hello d e", "hello", 0, True, 139), ("hello\tworld", "hello", 0, True, 155),
    ("hello\tworld", "world", 0, True, 156), ("hello\nworld", "hello", 0, True, 160),
    ("hello\r\nworld", "world", 0, True, 161), ("one\ttwo\nthree four", "three", 0, True, 165),
    ("one\ttwo\nthree four", "threa", 1, True, 166), ("a b c", "b", 0, True, 203), ("a b c", "x", 1, True, 204),
]


def test_contains_fuzzy_match_known_answers(oracle):
    for text, term, dist, want, line in FUZZY_KAT:
        assert oracle.contains_fuzzy_match(text, term, dist) == want, (text, term, dist, f"edit_distance_test.cpp:{line}")


def test_contains_fuzzy_match_agrees_with_reference_sources(oracle, reflib):
    for text, term, dist, want, line in FUZZY_KAT:
        assert reflib.contains_fuzzy_match(text, term, dist) == want, line
    rnd = random.Random(99)
    alphabet = ["a", "b", "c", " ", "\t", "東", "京", "\u3000", "\u00a0", "é"]
    raw = [b"\xe3\x80", b"\xc2", b"\xff", b"\xe6\x9d"]
    for _ in range(20000):
        text = "".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 14))).encode()
        term = "".join(rnd.choice(alphabet[:3] + alphabet[5:7] + ["é"]) for _ in range(rnd.randint(0, 5))).encode()
        if rnd.random() < 0.15:
            cut = rnd.randrange(len(text) + 1)
            text = text[:cut] + rnd.choice(raw) + text[cut:]
        if rnd.random() < 0.05:
            term += rnd.choice(raw)
        d = rnd.randint(0, 3)
        assert oracle.contains_fuzzy_match(text, term, d) == reflib.contains_fuzzy_match(text, term, d), (text, term, d)
