"""oracle/pyoracle.py — TEST INFRASTRUCTURE ONLY.

ctypes binding for the oracle C ABI (oracle/oracle.h). The same binding drives
either library, because both export the identical entry points:

  * ``oracle/liboracle.so``            the CPU restatement (oracle.cpp), kind "port";
  * ``oracle/_ref/libmygram_ref.so``   the reference's own unmodified sources +
                                       shims (ref_driver.cpp), kind "reference".

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libmygram_ref.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


class QueryParams(C.Structure):
    _fields_ = [
        ("ngram_size", C.c_int32),
        ("kanji_ngram_size", C.c_int32),
        ("cross_boundary", C.c_int32),
        ("compute_score", C.c_int32),
        ("descending", C.c_int32),
        ("limit", C.c_uint32),
        ("offset", C.c_uint32),
        ("filter_threshold", C.c_uint32),
        ("k1", C.c_double),
        ("b", C.c_double),
        ("total_docs_override", C.c_uint64),
        ("total_len_override", C.c_uint64),
        ("verify_text", C.c_int32),
        ("reserved", C.c_int32),
    ]


def _ptr(arr, typ):
    if arr is None:
        return None
    return arr.ctypes.data_as(typ)


def pack_strings(strings):
    """list[bytes] -> (uint8 arena, uint64 offsets[n+1])"""
    offsets = np.zeros(len(strings) + 1, dtype=np.uint64)
    if strings:
        offsets[1:] = np.cumsum([len(s) for s in strings], dtype=np.uint64)
    arena = np.frombuffer(b"".join(strings), dtype=np.uint8).copy() if strings else np.zeros(0, np.uint8)
    if arena.size == 0:
        arena = np.zeros(1, np.uint8)  # keep a valid pointer
    return arena, offsets


def as_bytes(s):
    return s.encode("utf-8") if isinstance(s, str) else bytes(s)


def pack_filter_columns(n_docs, columns):
    """-> (values uint64[n_cols * n_docs], nulls uint8[...], types int32[n_cols], string table list[bytes])"""
    vals = np.zeros(len(columns) * n_docs, dtype=np.uint64)
    nulls = np.zeros(len(columns) * n_docs, dtype=np.uint8)
    types = np.asarray([c[0] for c in columns], dtype=np.int32)
    strings = []
    for ci, (typ, values) in enumerate(columns):
        for row, v in enumerate(values):
            i = ci * n_docs + row
            if v is None:
                nulls[i] = 1
            elif typ == 11:
                vals[i] = len(strings)
                strings.append(as_bytes(v))
            elif typ == 12:
                vals[i] = np.float64(v).view(np.uint64)
            else:
                vals[i] = np.int64(int(v)).view(np.uint64) if int(v) < 0 else np.uint64(int(v))
    return vals, nulls, types, strings


def pack_filter_arrays(n_docs, columns):
    """Same packing for columns that already are arrays: columns = list of (type_code, values_u64, string table or
    None); for strings (type 11) values index the table and row i refers to strings[i] of the returned table."""
    vals = np.zeros(len(columns) * n_docs, dtype=np.uint64)
    nulls = np.zeros(len(columns) * n_docs, dtype=np.uint8)
    types = np.asarray([c[0] for c in columns], dtype=np.int32)
    strings = []
    for ci, (typ, values, table) in enumerate(columns):
        v = np.ascontiguousarray(values, dtype=np.uint64)
        if typ == 11:
            base = len(strings)
            strings += [as_bytes(t) for t in table]
            v = v + np.uint64(base)
        vals[ci * n_docs:(ci + 1) * n_docs] = v
    return vals, nulls, types, strings


@dataclass
class BatchResult:
    ids: np.ndarray      # [Q, stride] uint32
    scores: np.ndarray   # [Q, stride] float64
    count: np.ndarray    # [Q] uint32
    total: np.ndarray    # [Q] uint64
    df: np.ndarray       # [n_terms] uint64
    sets: list | None    # per-query ascending result sets (if requested)


class OracleLib:
    def __init__(self, path=PORT_LIB):
        self.path = path
        self.kind = "reference" if os.path.abspath(path) == os.path.abspath(REF_LIB) else "port"
        L = self.lib = C.CDLL(path)
        L.orc_utf8_to_codepoints.restype = C.c_uint64
        L.orc_utf8_to_codepoints.argtypes = [u8p, C.c_uint64, u32p, C.c_uint64]
        L.orc_codepoints_to_utf8.restype = C.c_uint64
        L.orc_codepoints_to_utf8.argtypes = [u32p, C.c_uint64, u8p]
        L.orc_count_code_points.restype = C.c_uint64
        L.orc_count_code_points.argtypes = [u8p, C.c_uint64]
        L.orc_is_cjk_ideograph.restype = C.c_int
        L.orc_is_cjk_ideograph.argtypes = [C.c_uint32]
        L.orc_ngrams.restype = C.c_int64
        L.orc_ngrams.argtypes = [C.c_int, u8p, C.c_uint64, C.c_int, C.c_int, C.c_int, u8p, C.c_uint64, u64p, C.c_uint64]
        L.orc_index_create.restype = C.c_void_p
        L.orc_index_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_index_destroy.argtypes = [C.c_void_p]
        L.orc_index_add_document.restype = C.c_int
        L.orc_index_add_document.argtypes = [C.c_void_p, C.c_uint32, u8p, C.c_uint64]
        L.orc_index_add_batch.argtypes = [C.c_void_p, u32p, u8p, u64p, C.c_uint64, C.c_uint64]
        L.orc_index_build_bulk.restype = C.c_int
        L.orc_index_build_bulk.argtypes = [C.c_void_p, u32p, u8p, u64p, C.c_uint64, C.c_int]
        L.orc_index_remove_document.argtypes = [C.c_void_p, C.c_uint32, u8p, C.c_uint64]
        L.orc_index_update_document.argtypes = [C.c_void_p, C.c_uint32, u8p, C.c_uint64, u8p, C.c_uint64]
        L.orc_index_term_count.restype = C.c_uint64
        L.orc_index_term_count.argtypes = [C.c_void_p]
        L.orc_index_posting_size.restype = C.c_uint64
        L.orc_index_posting_size.argtypes = [C.c_void_p, u8p, C.c_uint64]
        L.orc_index_total_postings.restype = C.c_uint64
        L.orc_index_total_postings.argtypes = [C.c_void_p]
        L.orc_index_get_postings.restype = C.c_uint64
        L.orc_index_get_postings.argtypes = [C.c_void_p, u8p, C.c_uint64, u32p, C.c_uint64]
        L.orc_index_export.restype = C.c_uint64
        L.orc_index_export.argtypes = [C.c_void_p, u8p, u64p, u64p, u32p, u64p]
        L.orc_index_bm25_stats.argtypes = [C.c_void_p, u64p, u64p]
        for name in ("orc_search_or",):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, u32p, C.c_uint64]
        L.orc_search_and.restype = C.c_uint64
        L.orc_search_and.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_uint64, C.c_int, u32p, C.c_uint64]
        L.orc_search_not.restype = C.c_uint64
        L.orc_search_not.argtypes = [C.c_void_p, u32p, C.c_uint64, u8p, u64p, C.c_uint64, u32p, C.c_uint64]
        L.orc_filter_by_ngrams.restype = C.c_uint64
        L.orc_filter_by_ngrams.argtypes = [C.c_void_p, u32p, C.c_uint64, u8p, u64p, C.c_uint64, u32p, C.c_uint64]
        L.orc_search_by_threshold.restype = C.c_uint64
        L.orc_search_by_threshold.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_uint64, u32p, C.c_uint64]
        L.orc_compute_idf.restype = C.c_double
        L.orc_compute_idf.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_count_term_occurrences.restype = C.c_uint32
        L.orc_count_term_occurrences.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64]
        L.orc_score_documents.argtypes = [C.c_void_p, u32p, C.c_uint64, u8p, u64p, u64p, C.c_uint64, C.c_uint64,
                                          C.c_double, C.c_double, C.c_double, f64p]
        L.orc_sort_by_score.restype = C.c_uint64
        L.orc_sort_by_score.argtypes = [u32p, f64p, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, u32p]
        L.orc_query_batch.restype = C.c_int
        L.orc_query_batch.argtypes = [C.c_void_p, C.POINTER(QueryParams), C.c_uint64, u8p, u64p, u64p, u8p, u64p, u64p,
                                      C.c_uint64, u32p, f64p, u32p, u64p, u64p, u32p, C.c_uint64, u64p, C.c_int]
        L.orc_eval_boolean.restype = C.c_uint64
        L.orc_eval_boolean.argtypes = [C.c_void_p, i32p, i32p, C.c_uint64, u8p, u64p, u32p, C.c_uint64]
        L.orc_contains_fuzzy_match.restype = C.c_int
        L.orc_contains_fuzzy_match.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64, C.c_uint32]
        L.orc_search_fuzzy.restype = C.c_uint64
        L.orc_search_fuzzy.argtypes = [C.c_void_p, C.POINTER(QueryParams), u8p, u64p, C.c_uint64, C.c_uint32, u8p, u64p,
                                       C.c_uint64, u32p, C.c_uint64, i32p]
        L.orc_search_synonyms.restype = C.c_uint64
        L.orc_search_synonyms.argtypes = [C.c_void_p, C.POINTER(QueryParams), u8p, u64p, u64p, C.c_uint64, u8p, u64p,
                                          C.c_uint64, u32p, C.c_uint64, i32p]

    # ---- tokenizer ----
    def apply_filters(self, n_docs, first_doc_id, columns, filters, results):
        """columns: list of (type_code, values, nulls); values is a list of python values per row (None = NULL;
        str/bytes for strings, float for doubles, int/bool otherwise) -- or the tuple pack_filter_columns /
        pack_filter_arrays returned, packed once for many calls. filters: list of (col, op, literal).
        -> the ids of `results` that pass (ApplyFiltersWithBitmap, search_pipeline.cpp:1196-1237)."""
        if isinstance(columns, tuple) and len(columns) == 4 and isinstance(columns[0], np.ndarray):
            vals, nulls, types, strings = columns
            columns = list(types)
        else:
            vals, nulls, types, strings = pack_filter_columns(n_docs, columns)
        sbytes, soffs = pack_strings(strings if strings else [b""])
        fc = np.asarray([f[0] for f in filters], dtype=np.uint32)
        fo = np.asarray([f[1] for f in filters], dtype=np.uint8)
        lb, lo = pack_strings([as_bytes(f[2]) for f in filters] if filters else [b""])
        res = np.ascontiguousarray(results, dtype=np.uint32)
        out = np.zeros(max(1, res.size), dtype=np.uint32)
        fn = self.lib.orc_apply_filters
        fn.restype = C.c_uint64
        fn.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, i32p, u64p, u8p, u8p, u64p, C.c_uint32, u32p, u8p, u8p, u64p,
                       u32p, C.c_uint64, u32p]
        one32 = np.zeros(1, np.uint32)
        one8 = np.zeros(1, np.uint8)
        n = fn(n_docs, first_doc_id, len(columns), _ptr(types if types.size else np.zeros(1, np.int32), i32p),
               _ptr(vals if vals.size else np.zeros(1, np.uint64), u64p), _ptr(nulls if nulls.size else one8, u8p),
               _ptr(sbytes, u8p), _ptr(soffs, u64p), len(filters), _ptr(fc if fc.size else one32, u32p),
               _ptr(fo if fo.size else one8, u8p), _ptr(lb, u8p), _ptr(lo, u64p),
               _ptr(res if res.size else one32, u32p), res.size, _ptr(out, u32p))
        return out[:n].copy()

    def utf8_to_codepoints(self, text):
        b = as_bytes(text)
        buf = np.frombuffer(b, dtype=np.uint8).copy() if b else np.zeros(1, np.uint8)
        out = np.zeros(max(1, len(b)), dtype=np.uint32)
        n = self.lib.orc_utf8_to_codepoints(_ptr(buf, u8p), len(b), _ptr(out, u32p), out.size)
        return out[:n].tolist()

    def codepoints_to_utf8(self, cps):
        arr = np.asarray(cps, dtype=np.uint32)
        if arr.size == 0:
            return b""
        out = np.zeros(arr.size * 4, dtype=np.uint8)
        n = self.lib.orc_codepoints_to_utf8(_ptr(arr, u32p), arr.size, _ptr(out, u8p))
        return out[:n].tobytes()

    def count_code_points(self, text):
        b = as_bytes(text)
        buf = np.frombuffer(b, dtype=np.uint8).copy() if b else np.zeros(1, np.uint8)
        return int(self.lib.orc_count_code_points(_ptr(buf, u8p), len(b)))

    def is_cjk(self, cp):
        return bool(self.lib.orc_is_cjk_ideograph(cp))

    def ngrams(self, mode, text, a, k=1, cross=True):
        """mode: 'plain' | 'hybrid' | 'query' -> list[bytes]"""
        m = {"plain": 0, "hybrid": 1, "query": 2}[mode]
        b = as_bytes(text)
        buf = np.frombuffer(b, dtype=np.uint8).copy() if b else np.zeros(1, np.uint8)
        cap_n = len(b) + 1
        cap_b = len(b) * max(2, a, k) + 16
        out = np.zeros(cap_b, dtype=np.uint8)
        offs = np.zeros(cap_n + 1, dtype=np.uint64)
        n = self.lib.orc_ngrams(m, _ptr(buf, u8p), len(b), a, k, int(cross), _ptr(out, u8p), cap_b, _ptr(offs, u64p),
                                cap_n)
        assert n >= 0, "ngram buffer too small"
        raw = out.tobytes()
        return [raw[int(offs[i]):int(offs[i + 1])] for i in range(n)]

    # ---- index ----
    def contains_fuzzy_match(self, text, term, max_distance):
        t = np.frombuffer(as_bytes(text) or b"\0", dtype=np.uint8).copy()
        q = np.frombuffer(as_bytes(term) or b"\0", dtype=np.uint8).copy()
        return bool(self.lib.orc_contains_fuzzy_match(_ptr(t, u8p), len(as_bytes(text)), _ptr(q, u8p),
                                                      len(as_bytes(term)), max_distance))

    def index(self, ngram_size=2, kanji_ngram_size=0, cross_boundary=True):
        return OracleIndex(self, ngram_size, kanji_ngram_size, cross_boundary)

    # ---- bm25 ----
    def compute_idf(self, n, df):
        return float(self.lib.orc_compute_idf(n, df))

    def count_term_occurrences(self, text, term):
        t, q = as_bytes(text), as_bytes(term)
        tb = np.frombuffer(t, dtype=np.uint8).copy() if t else np.zeros(1, np.uint8)
        qb = np.frombuffer(q, dtype=np.uint8).copy() if q else np.zeros(1, np.uint8)
        return int(self.lib.orc_count_term_occurrences(_ptr(tb, u8p), len(t), _ptr(qb, u8p), len(q)))

    def sort_by_score(self, results, scores, descending=True, limit=0, offset=0):
        r = np.ascontiguousarray(results, dtype=np.uint32)
        s = np.ascontiguousarray(scores, dtype=np.float64)
        out = np.zeros(max(1, r.size), dtype=np.uint32)
        n = self.lib.orc_sort_by_score(_ptr(r, u32p), _ptr(s, f64p), r.size, int(descending), limit, offset,
                                       _ptr(out, u32p))
        return out[:n].copy()


class OracleIndex:
    def __init__(self, lib: OracleLib, ngram_size, kanji_ngram_size, cross_boundary):
        self.L = lib
        self.ngram_size = ngram_size
        self.kanji_ngram_size = kanji_ngram_size
        self.cross_boundary = cross_boundary
        self.h = lib.lib.orc_index_create(ngram_size, kanji_ngram_size, int(cross_boundary))
        self._keep = []  # arrays borrowed by build_bulk

    def close(self):
        if self.h:
            self.L.lib.orc_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_document(self, doc_id, text):
        b = as_bytes(text)
        buf = np.frombuffer(b, dtype=np.uint8).copy() if b else np.zeros(1, np.uint8)
        return self.L.lib.orc_index_add_document(self.h, doc_id, _ptr(buf, u8p), len(b))

    def add_batch(self, doc_ids, arena, offsets, batch=1000):
        doc_ids = np.ascontiguousarray(doc_ids, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.L.lib.orc_index_add_batch(self.h, _ptr(doc_ids, u32p), _ptr(arena, u8p), _ptr(offsets, u64p),
                                       doc_ids.size, batch)

    def add_texts(self, doc_ids, texts, batch=1000):
        arena, offsets = pack_strings([as_bytes(t) for t in texts])
        self.add_batch(doc_ids, arena, offsets, batch)

    def build_bulk(self, doc_ids, arena, offsets, n_threads=0):
        doc_ids = np.ascontiguousarray(doc_ids, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._keep += [doc_ids, arena, offsets]
        rc = self.L.lib.orc_index_build_bulk(self.h, _ptr(doc_ids, u32p), _ptr(arena, u8p), _ptr(offsets, u64p),
                                             doc_ids.size, n_threads)
        assert rc == 0, f"orc_index_build_bulk failed: {rc}"

    def remove_document(self, doc_id, text):
        b = as_bytes(text)
        buf = np.frombuffer(b, dtype=np.uint8).copy() if b else np.zeros(1, np.uint8)
        self.L.lib.orc_index_remove_document(self.h, doc_id, _ptr(buf, u8p), len(b))

    def update_document(self, doc_id, old, new):
        o, n = as_bytes(old), as_bytes(new)
        ob = np.frombuffer(o, dtype=np.uint8).copy() if o else np.zeros(1, np.uint8)
        nb = np.frombuffer(n, dtype=np.uint8).copy() if n else np.zeros(1, np.uint8)
        self.L.lib.orc_index_update_document(self.h, doc_id, _ptr(ob, u8p), len(o), _ptr(nb, u8p), len(n))

    def term_count(self):
        return int(self.L.lib.orc_index_term_count(self.h))

    def total_postings(self):
        return int(self.L.lib.orc_index_total_postings(self.h))

    def posting_size(self, term):
        b = as_bytes(term)
        buf = np.frombuffer(b, dtype=np.uint8).copy() if b else np.zeros(1, np.uint8)
        return int(self.L.lib.orc_index_posting_size(self.h, _ptr(buf, u8p), len(b)))

    def postings(self, term):
        b = as_bytes(term)
        buf = np.frombuffer(b, dtype=np.uint8).copy() if b else np.zeros(1, np.uint8)
        n = self.posting_size(term)
        out = np.zeros(max(1, n), dtype=np.uint32)
        n2 = self.L.lib.orc_index_get_postings(self.h, _ptr(buf, u8p), len(b), _ptr(out, u32p), out.size)
        return out[:n2].copy()

    def export(self):
        """-> (terms: list[bytes] in byte order, posting_offsets uint64[T+1], postings uint32[P])"""
        tb = np.zeros(1, dtype=np.uint64)
        T = self.L.lib.orc_index_export(self.h, None, None, None, None, _ptr(tb, u64p))
        if T == 2**64 - 1:
            raise NotImplementedError("export is not available for the reference library")
        P = self.total_postings()
        term_bytes = np.zeros(max(1, int(tb[0])), dtype=np.uint8)
        term_offs = np.zeros(T + 1, dtype=np.uint64)
        post_offs = np.zeros(T + 1, dtype=np.uint64)
        posts = np.zeros(max(1, P), dtype=np.uint32)
        self.L.lib.orc_index_export(self.h, _ptr(term_bytes, u8p), _ptr(term_offs, u64p), _ptr(post_offs, u64p),
                                    _ptr(posts, u32p), _ptr(tb, u64p))
        raw = term_bytes.tobytes()
        terms = [raw[int(term_offs[i]):int(term_offs[i + 1])] for i in range(T)]
        return terms, post_offs, posts[:P]

    def bm25_stats(self):
        a = np.zeros(1, dtype=np.uint64)
        b = np.zeros(1, dtype=np.uint64)
        self.L.lib.orc_index_bm25_stats(self.h, _ptr(a, u64p), _ptr(b, u64p))
        return int(a[0]), int(b[0])

    def _terms(self, terms):
        return pack_strings([as_bytes(t) for t in terms])

    def _call_set(self, fn, terms, *mid, cap=None):
        arena, offs = self._terms(terms)
        cap = cap or 1024
        while True:
            out = np.zeros(cap, dtype=np.uint32)
            n = fn(arena, offs, out, cap)
            if n <= cap:
                return out[:n].copy()
            cap = int(n)

    def search_and(self, terms, limit=0, reverse=False):
        f = self.L.lib.orc_search_and
        return self._call_set(lambda a, o, out, cap: f(self.h, _ptr(a, u8p), _ptr(o, u64p), len(terms), limit,
                                                       int(reverse), _ptr(out, u32p), cap), terms)

    def search_or(self, terms):
        f = self.L.lib.orc_search_or
        return self._call_set(lambda a, o, out, cap: f(self.h, _ptr(a, u8p), _ptr(o, u64p), len(terms),
                                                       _ptr(out, u32p), cap), terms)

    def search_not(self, all_docs, terms):
        f = self.L.lib.orc_search_not
        all_docs = np.ascontiguousarray(all_docs, dtype=np.uint32)
        ad = all_docs if all_docs.size else np.zeros(1, np.uint32)
        return self._call_set(lambda a, o, out, cap: f(self.h, _ptr(ad, u32p), all_docs.size, _ptr(a, u8p),
                                                       _ptr(o, u64p), len(terms), _ptr(out, u32p), cap), terms,
                              cap=max(1024, all_docs.size))

    def filter_by_ngrams(self, candidates, terms):
        f = self.L.lib.orc_filter_by_ngrams
        c = np.ascontiguousarray(candidates, dtype=np.uint32)
        cd = c if c.size else np.zeros(1, np.uint32)
        return self._call_set(lambda a, o, out, cap: f(self.h, _ptr(cd, u32p), c.size, _ptr(a, u8p), _ptr(o, u64p),
                                                       len(terms), _ptr(out, u32p), cap), terms,
                              cap=max(1024, c.size))

    def search_by_threshold(self, terms, threshold):
        f = self.L.lib.orc_search_by_threshold
        return self._call_set(lambda a, o, out, cap: f(self.h, _ptr(a, u8p), _ptr(o, u64p), len(terms), threshold,
                                                       _ptr(out, u32p), cap), terms)

    def score_documents(self, candidates, terms, dfs, total_docs, avgdl, k1=1.2, b=0.75):
        c = np.ascontiguousarray(candidates, dtype=np.uint32)
        arena, offs = self._terms(terms)
        d = np.ascontiguousarray(dfs, dtype=np.uint64)
        out = np.zeros(max(1, c.size), dtype=np.float64)
        cd = c if c.size else np.zeros(1, np.uint32)
        dd = d if d.size else np.zeros(1, np.uint64)
        self.L.lib.orc_score_documents(self.h, _ptr(cd, u32p), c.size, _ptr(arena, u8p), _ptr(offs, u64p),
                                       _ptr(dd, u64p), len(terms), total_docs, avgdl, k1, b, _ptr(out, f64p))
        return out[:c.size].copy()

    def eval_boolean(self, ops, args, terms):
        o = np.ascontiguousarray(ops, dtype=np.int32)
        a = np.ascontiguousarray(args, dtype=np.int32)
        arena, offs = self._terms(terms)
        cap = 1024
        while True:
            out = np.zeros(cap, dtype=np.uint32)
            n = self.L.lib.orc_eval_boolean(self.h, _ptr(o, i32p), _ptr(a, i32p), o.size, _ptr(arena, u8p),
                                            _ptr(offs, u64p), _ptr(out, u32p), cap)
            if n <= cap:
                return out[:n].copy()
            cap = int(n)

    # ---- MGIX stream (reference sources only: Index::SaveToStream / LoadFromStream) ----
    def save_stream(self):
        f = self.L.lib.ref_index_save_stream
        f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, u8p, C.c_uint64]
        n = f(self.h, None, 0)
        assert n > 0, "SaveToStream failed"
        out = np.zeros(n, dtype=np.uint8)
        assert f(self.h, _ptr(out, u8p), n) == n
        return out.tobytes()

    def load_stream(self, data):
        """Returns 0 on success, else the reference's ErrorCode."""
        f = self.L.lib.ref_index_load_stream
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, u8p, C.c_uint64]
        buf = np.frombuffer(bytes(data), dtype=np.uint8).copy() if len(data) else np.zeros(1, np.uint8)
        return int(f(self.h, _ptr(buf, u8p), len(data)))

    def _pipeline_params(self, raw_ngram, raw_kanji, verify_text):
        return QueryParams(self.ngram_size if raw_ngram is None else raw_ngram,
                           self.kanji_ngram_size if raw_kanji is None else raw_kanji, int(self.cross_boundary),
                           0, 1, 0, 0, 1000, 1.2, 0.75, 0, 0, verify_text, 0)

    def search_fuzzy(self, terms, max_distance, not_terms=(), raw_ngram=None, raw_kanji=None, verify_text=0):
        """ExecuteWithFuzzy over normalised terms -> (ascending ids, empty_term_detected)."""
        p = self._pipeline_params(raw_ngram, raw_kanji, verify_text)
        arena, offs = self._terms(terms)
        narena, noffs = self._terms(not_terms)
        empty = np.zeros(1, dtype=np.int32)
        cap = 1024
        while True:
            out = np.zeros(cap, dtype=np.uint32)
            n = self.L.lib.orc_search_fuzzy(self.h, C.byref(p), _ptr(arena, u8p), _ptr(offs, u64p), len(terms),
                                            max_distance, _ptr(narena, u8p), _ptr(noffs, u64p), len(not_terms),
                                            _ptr(out, u32p), cap, _ptr(empty, i32p))
            if n <= cap:
                return out[:n].copy(), bool(empty[0])
            cap = int(n)

    def search_synonyms(self, groups, not_terms=(), raw_ngram=None, raw_kanji=None, verify_text=0):
        """ExecuteWithSynonyms over expanded groups (list of lists of variants) -> (ascending ids,
        empty_term_detected)."""
        p = self._pipeline_params(raw_ngram, raw_kanji, verify_text)
        flat, gbeg = [], [0]
        for g in groups:
            flat += list(g)
            gbeg.append(len(flat))
        arena, offs = self._terms(flat)
        gbeg = np.asarray(gbeg, dtype=np.uint64)
        narena, noffs = self._terms(not_terms)
        empty = np.zeros(1, dtype=np.int32)
        cap = 1024
        while True:
            out = np.zeros(cap, dtype=np.uint32)
            n = self.L.lib.orc_search_synonyms(self.h, C.byref(p), _ptr(arena, u8p), _ptr(offs, u64p),
                                               _ptr(gbeg, u64p), len(groups), _ptr(narena, u8p), _ptr(noffs, u64p),
                                               len(not_terms), _ptr(out, u32p), cap, _ptr(empty, i32p))
            if n <= cap:
                return out[:n].copy(), bool(empty[0])
            cap = int(n)

    def query_batch(self, queries, not_terms=None, score=True, descending=True, limit=100, offset=0,
                    filter_threshold=1000, k1=1.2, b=0.75, n_threads=1, want_sets=False, stride=None,
                    total_docs_override=0, total_len_override=0, raw_ngram=None, raw_kanji=None, verify_text=0):
        """queries: list[list[bytes|str]] search terms per query; not_terms likewise (optional)."""
        flat, qbeg = [], [0]
        for q in queries:
            flat += [as_bytes(t) for t in q]
            qbeg.append(len(flat))
        arena, offs = pack_strings(flat)
        qbeg = np.asarray(qbeg, dtype=np.uint64)
        if not_terms is not None:
            nflat, nbeg = [], [0]
            for q in not_terms:
                nflat += [as_bytes(t) for t in q]
                nbeg.append(len(nflat))
            narena, noffs = pack_strings(nflat)
            nbeg = np.asarray(nbeg, dtype=np.uint64)
        else:
            narena = noffs = nbeg = None
        Q = len(queries)
        stride = stride or max(1, limit if limit else 1)
        p = QueryParams(self.ngram_size if raw_ngram is None else raw_ngram,
                        self.kanji_ngram_size if raw_kanji is None else raw_kanji, int(self.cross_boundary),
                        int(score), int(descending), limit, offset, filter_threshold, k1, b, total_docs_override,
                        total_len_override, verify_text, 0)
        ids = np.zeros((Q, stride), dtype=np.uint32)
        scores = np.zeros((Q, stride), dtype=np.float64)
        count = np.zeros(Q, dtype=np.uint32)
        total = np.zeros(Q, dtype=np.uint64)
        df = np.zeros(max(1, len(flat)), dtype=np.uint64)
        sets = None
        sets_offs = None
        sets_buf = None
        cap = 0
        if want_sets:
            sets_offs = np.zeros(Q + 1, dtype=np.uint64)
            cap = 1 << 20
            sets_buf = np.zeros(cap, dtype=np.uint32)
        while True:
            rc = self.L.lib.orc_query_batch(self.h, C.byref(p), Q, _ptr(arena, u8p), _ptr(offs, u64p), _ptr(qbeg, u64p),
                                            _ptr(narena, u8p), _ptr(noffs, u64p), _ptr(nbeg, u64p), stride,
                                            _ptr(ids, u32p), _ptr(scores, f64p), _ptr(count, u32p), _ptr(total, u64p),
                                            _ptr(df, u64p), _ptr(sets_buf, u32p), cap, _ptr(sets_offs, u64p),
                                            n_threads)
            assert rc == 0, f"orc_query_batch rc={rc}"
            if want_sets and int(sets_offs[Q]) > cap:
                cap = int(sets_offs[Q])
                sets_buf = np.zeros(cap, dtype=np.uint32)
                continue
            break
        if want_sets:
            sets = [sets_buf[int(sets_offs[q]):int(sets_offs[q + 1])].copy() for q in range(Q)]
        return BatchResult(ids, scores, count, total, df[:len(flat)], sets)
