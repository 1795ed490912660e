"""tests/support/corpus.py — synthetic corpora and query batches of SURVEY.md §8(d).

Thin ctypes wrapper over corpusgen.c plus the query samplers. Shared by tests/ and bench.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcorpusgen.so")
_lib = None

SEEDS = {"C1": 0xC1, "C2": 0xC2, "C3": 0xC3, "C4": 0xC4, "C5": 0xC5}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-C", HERE])
        L = C.CDLL(LIB)
        L.corpus_gen_create.restype = C.c_void_p
        L.corpus_gen_create.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_double, C.c_uint32, C.c_uint32]
        L.corpus_gen_destroy.argtypes = [C.c_void_p]
        L.corpus_gen_sizes.restype = C.c_uint64
        L.corpus_gen_sizes.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        L.corpus_gen_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        L.corpus_gen_docs.restype = C.c_uint64
        L.corpus_gen_docs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class Corpus:
    """doc_ids (uint32, ascending), arena (uint8), offsets (uint64[n+1])."""

    def __init__(self, doc_ids, arena, offsets, kind):
        self.doc_ids, self.arena, self.offsets, self.kind = doc_ids, arena, offsets, kind

    @property
    def n_docs(self):
        return int(self.doc_ids.size)

    def text(self, i):
        return self.arena[int(self.offsets[i]):int(self.offsets[i + 1])].tobytes()


def generate(kind, n_docs, seed, first_doc=0, first_id=1, alphabet=None, zipf_s=1.0, min_len=None, max_len=None,
             threads=None, arena_out=None, offsets_out=None, alloc=None):
    """kind 'cjk' (C2-style: 8192 ideographs, 16..112 cps) or 'ascii' (C1-style: 4096 words, 8..40 words).
    Document i of the call is GLOBAL document first_doc+i with id first_id+first_doc+i, so shards of one
    corpus can be generated independently."""
    k = {"cjk": 0, "ascii": 1}[kind]
    alphabet = alphabet or (8192 if k == 0 else 4096)
    min_len = min_len if min_len is not None else (16 if k == 0 else 8)
    max_len = max_len if max_len is not None else (112 if k == 0 else 40)
    threads = threads or min(32, os.cpu_count() or 1)
    L = lib()
    g = L.corpus_gen_create(k, seed, alphabet, zipf_s, min_len, max_len)
    try:
        offsets = offsets_out if offsets_out is not None else np.zeros(n_docs + 1, dtype=np.uint64)
        total = L.corpus_gen_sizes(g, first_doc, n_docs, offsets.ctypes.data, threads)
        if arena_out is not None:
            arena = arena_out
        elif alloc is not None:
            arena = alloc(max(1, total))  # e.g. a pinned-memory allocator
        else:
            arena = np.zeros(max(1, total), dtype=np.uint8)
        assert arena.size >= total
        L.corpus_gen_fill(g, first_doc, n_docs, offsets.ctypes.data, arena.ctypes.data, threads)
    finally:
        L.corpus_gen_destroy(g)
    doc_ids = (np.arange(n_docs, dtype=np.uint64) + first_id + first_doc).astype(np.uint32)
    return Corpus(doc_ids, arena[:max(1, total)], offsets, kind)


def sample_queries(corpus, n_queries, seed, n_terms=3, min_cp=2, max_cp=4):
    """C2-style queries: n_terms substrings of min_cp..max_cp code points cut from ONE random document
    (guarantees a non-empty AND). For 'ascii' corpora the terms are whole words of the document."""
    rng = np.random.default_rng(seed)
    queries = []
    while len(queries) < n_queries:
        d = int(rng.integers(0, corpus.n_docs))
        text = corpus.text(d).decode("utf-8")
        if corpus.kind == "ascii":
            words = text.split(" ")
            if len(words) < n_terms:
                continue
            pick = rng.choice(len(words), size=n_terms, replace=False)
            queries.append([words[int(i)].encode() for i in pick])
            continue
        if len(text) < max_cp:
            continue
        terms = []
        for _ in range(n_terms):
            ln = int(rng.integers(min_cp, max_cp + 1))
            st = int(rng.integers(0, len(text) - ln + 1))
            terms.append(text[st:st + ln].encode("utf-8"))
        queries.append(terms)
    return queries


def docs_by_index(kind, seed, doc_indices, alphabet=None, zipf_s=1.0, min_len=None, max_len=None):
    """Texts (list[bytes]) of arbitrary GLOBAL document indices of a corpus, without generating the corpus."""
    k = {"cjk": 0, "ascii": 1}[kind]
    alphabet = alphabet or (8192 if k == 0 else 4096)
    min_len = min_len if min_len is not None else (16 if k == 0 else 8)
    max_len = max_len if max_len is not None else (112 if k == 0 else 40)
    L = lib()
    g = L.corpus_gen_create(k, seed, alphabet, zipf_s, min_len, max_len)
    try:
        idx = np.ascontiguousarray(doc_indices, dtype=np.uint64)
        offsets = np.zeros(idx.size + 1, dtype=np.uint64)
        total = L.corpus_gen_docs(g, idx.ctypes.data, idx.size, offsets.ctypes.data, None)
        text = np.zeros(max(1, total), dtype=np.uint8)
        L.corpus_gen_docs(g, idx.ctypes.data, idx.size, offsets.ctypes.data, text.ctypes.data)
    finally:
        L.corpus_gen_destroy(g)
    raw = text.tobytes()
    return [raw[int(offsets[i]):int(offsets[i + 1])] for i in range(idx.size)]


def sample_queries_global(kind, corpus_seed, n_docs_total, n_queries, seed, n_terms=3, min_cp=2, max_cp=4, **gen_kw):
    """Same sampler as sample_queries() but over a corpus identified only by (kind, seed, size): every rank of a
    sharded run gets the identical batch without holding the whole corpus."""
    rng = np.random.default_rng(seed)
    queries = []
    while len(queries) < n_queries:
        need = n_queries - len(queries)
        picks = rng.integers(0, n_docs_total, size=need)
        texts = docs_by_index(kind, corpus_seed, picks, **gen_kw)
        for raw in texts:
            text = raw.decode("utf-8")
            if kind == "ascii":
                words = text.split(" ")
                if len(words) < n_terms:
                    continue
                pick = rng.choice(len(words), size=n_terms, replace=False)
                queries.append([words[int(i)].encode() for i in pick])
                continue
            if len(text) < max_cp:
                continue
            terms = []
            nt = n_terms if isinstance(n_terms, int) else int(rng.integers(n_terms[0], n_terms[1] + 1))
            for _ in range(nt):
                ln = int(rng.integers(min_cp, max_cp + 1))
                st = int(rng.integers(0, len(text) - ln + 1))
                terms.append(text[st:st + ln].encode("utf-8"))
            queries.append(terms)
    return queries
