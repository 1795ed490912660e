"""mygram-db_b200 — host-side Python mirror of the reference's index/query interface
over the C ABI of libmgx.so (include/mgx.h).

The directory name carries a hyphen (it is the name the project was given), so
import it through :func:`load` in ``mgx_loader.py`` at the repository root, or::

    import importlib.util, sys
    spec = importlib.util.spec_from_file_location("mygram_db_b200", ".../mygram-db_b200/__init__.py")

What lives here is plumbing only: ctypes signatures, numpy buffer marshalling and
three small classes whose method names and argument meaning follow the reference
(``Index`` — src/index/index.h:49-413, ``BM25Scorer`` — src/index/bm25_scorer.h:43-83,
``ResultSorter.sort_by_score`` — src/query/result_sorter.h:75). All work happens in
the CUDA kernels behind the C ABI; if libmgx.so is missing or no GPU is visible the
calls raise — there is no CPU fallback and nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MGX_LIB_PATH") or os.path.join(HERE, "libmgx.so")  # override: A/B builds of the kernels
INCLUDE = os.path.join(os.path.dirname(HERE), "include", "mgx.h")

MGX_OK = 0
MGX_ERR_UNSUPPORTED = -3
MGX_ERR_CAPACITY = -4
ERRORS = {-1: "MGX_ERR_INVALID_ARGUMENT", -2: "MGX_ERR_CUDA", -3: "MGX_ERR_UNSUPPORTED", -4: "MGX_ERR_CAPACITY",
          -5: "MGX_ERR_NO_DEVICE", -6: "MGX_ERR_FORMAT", -7: "MGX_ERR_TIMEOUT"}

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


class MgxError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERRORS.get(code, code)}: {message}")
        self.code = code


class IndexConfig(C.Structure):
    _fields_ = [("ngram_size", C.c_int32), ("kanji_ngram_size", C.c_int32), ("cross_boundary_ngrams", C.c_int32),
                ("device", C.c_int32), ("dense_threshold", C.c_double), ("max_dense_bytes", C.c_uint64),
                ("scratch_bytes", C.c_uint64), ("roaring_threshold", C.c_double)]


class IndexStats(C.Structure):
    _fields_ = [("n_docs", C.c_uint64), ("n_terms", C.c_uint64), ("n_postings", C.c_uint64),
                ("n_dense_terms", C.c_uint64), ("text_bytes", C.c_uint64), ("total_doc_length", C.c_uint64),
                ("doc_count", C.c_uint64), ("device_bytes", C.c_uint64), ("n_pair_slots", C.c_uint64),
                ("all_valid_utf8", C.c_int32), ("key_width", C.c_int32), ("last_build_ms", C.c_double)]


class IndexStatistics(C.Structure):
    _fields_ = [("total_terms", C.c_uint64), ("total_postings", C.c_uint64), ("delta_encoded_lists", C.c_uint64),
                ("roaring_bitmap_lists", C.c_uint64), ("memory_usage_bytes", C.c_uint64)]


class QueryExt(C.Structure):
    _fields_ = [("prog_ops", C.POINTER(C.c_int32)), ("prog_args", C.POINTER(C.c_int32)), ("q_prog_begin", u64p),
                ("filter_col", u32p), ("filter_op", u8p), ("filter_bytes", u8p), ("filter_offsets", u64p),
                ("q_filter_begin", u64p)]


class ExpandedQuery(C.Structure):
    _fields_ = [("ngram_size", C.c_int32), ("kanji_ngram_size", C.c_int32), ("cross_boundary", C.c_int32),
                ("verify_text", C.c_int32), ("not_bytes", u8p), ("not_offsets", u64p), ("n_not", C.c_uint64),
                ("filter_col", u32p), ("filter_op", u8p), ("filter_bytes", u8p), ("filter_offsets", u64p),
                ("n_filters", C.c_uint64)]


class MgixInfo(C.Structure):
    _fields_ = [("version", C.c_uint32), ("ngram_size", C.c_int32), ("kanji_ngram_size", C.c_int32),
                ("cross_boundary", C.c_int32), ("normalize_nfkc", C.c_int32), ("normalize_lower", C.c_int32),
                ("normalize_width", C.c_char * 32), ("n_terms", C.c_uint64), ("n_postings", C.c_uint64),
                ("term_bytes", C.c_uint64)]


class QueryParams(C.Structure):
    _fields_ = [("ngram_size", C.c_int32), ("kanji_ngram_size", C.c_int32), ("cross_boundary", C.c_int32),
                ("compute_score", C.c_int32), ("descending", C.c_int32), ("limit", C.c_uint32), ("offset", C.c_uint32),
                ("verify_text", C.c_int32), ("k1", C.c_double), ("b", C.c_double), ("total_docs", C.c_uint64),
                ("total_doc_length", C.c_uint64)]


class ShardRecordLayout(C.Structure):
    """mgx_shard_record_layout_t: byte offsets of the parts of one shard's packed top-k record (+ status block)."""
    _fields_ = [("scores_offset", C.c_uint64), ("total_offset", C.c_uint64), ("ids_offset", C.c_uint64),
                ("count_offset", C.c_uint64), ("status_offset", C.c_uint64), ("bytes", C.c_uint64)]


class BatchStats(C.Structure):
    _fields_ = [("ms_plan", C.c_double), ("ms_df_kernel", C.c_double), ("ms_and_kernel", C.c_double),
                ("ms_topk_kernel", C.c_double), ("ms_total", C.c_double), ("launches", C.c_uint64),
                ("n_df_tiles", C.c_uint64), ("n_and_tiles", C.c_uint64), ("algo_bytes_intersect", C.c_uint64),
                ("algo_bytes_score", C.c_uint64), ("algo_bytes_df", C.c_uint64), ("algo_bytes_df_lists", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("driver_entries", C.c_uint64),
                ("result_docs", C.c_uint64), ("df_candidates", C.c_uint64), ("unique_terms", C.c_uint64),
                ("ms_df_stream_kernel", C.c_double), ("df_stream_terms", C.c_uint64), ("df_stream_bytes", C.c_uint64),
                ("df_stream_hits", C.c_uint64), ("df_scanned_docs", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def build(force=False):
    """Compile libmgx.so in-tree with nvcc for sm_100a (see Makefile)."""
    if force:
        subprocess.check_call(["make", "-C", HERE, "clean"])
    subprocess.check_call(["make", "-C", HERE, "-j4"])


def lib():
    """The loaded C ABI. Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MgxError(-2, f"{LIB_PATH} is missing: run `make -C {HERE}` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.mgx_last_error.restype = C.c_char_p
    L.mgx_version.restype = C.c_char_p
    L.mgx_kernel_launch_count.restype = C.c_uint64
    L.mgx_index_create.argtypes = [C.POINTER(IndexConfig), C.POINTER(C.c_void_p)]
    L.mgx_index_destroy.argtypes = [C.c_void_p]
    L.mgx_index_build.argtypes = [C.c_void_p, u32p, u8p, u64p, C.c_uint64]
    L.mgx_index_build_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.mgx_index_get_stats.argtypes = [C.c_void_p, C.POINTER(IndexStats)]
    L.mgx_index_get_statistics.argtypes = [C.c_void_p, C.POINTER(IndexStatistics)]
    L.mgx_index_optimize.argtypes = [C.c_void_p, C.c_uint64]
    L.mgx_index_trim.argtypes = [C.c_void_p]
    L.mgx_index_clear.argtypes = [C.c_void_p]
    L.mgx_index_add_document.argtypes = [C.c_void_p, C.c_uint32, u8p, C.c_uint64, C.POINTER(C.c_int32)]
    L.mgx_index_update_document.argtypes = [C.c_void_p, C.c_uint32, u8p, C.c_uint64, u8p, C.c_uint64]
    L.mgx_index_add_document_batch.argtypes = [C.c_void_p, u32p, u8p, u64p, C.c_uint64, u64p]
    L.mgx_index_remove_document.argtypes = [C.c_void_p, C.c_uint32, u8p, C.c_uint64]
    L.mgx_index_commit.argtypes = [C.c_void_p]
    L.mgx_index_set_commit_mode.argtypes = [C.c_void_p, C.c_int]
    L.mgx_index_posting_size.argtypes = [C.c_void_p, u8p, C.c_uint64, u64p]
    L.mgx_index_get_postings.argtypes = [C.c_void_p, u8p, C.c_uint64, u32p, C.c_uint64, u64p]
    L.mgx_index_get_posting_payload.argtypes = [C.c_void_p, u8p, C.c_uint64, u32p, u32p, u32p, C.c_uint64, u64p,
                                                C.POINTER(C.c_int)]
    L.mgx_index_export.argtypes = [C.c_void_p, u64p, u64p, u32p]
    L.mgx_index_doc_lengths.argtypes = [C.c_void_p, u32p]
    L.mgx_key_to_utf8.argtypes = [C.c_uint64, C.c_int32, u8p]
    L.mgx_key_words.argtypes = [C.c_int32]
    L.mgx_wide_key_to_utf8.argtypes = [u64p, C.c_int32, u8p]
    L.mgx_index_export_terms.argtypes = [C.c_void_p, u8p, C.c_uint64, u64p, u64p]
    L.mgx_tokenize_batch.argtypes = [C.POINTER(IndexConfig), u8p, u64p, C.c_uint64, u64p, u32p, C.c_uint64, u64p]
    L.mgx_search_and.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_uint64, C.c_int32, u32p, C.c_uint64, u64p]
    L.mgx_search_or.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, u32p, C.c_uint64, u64p]
    L.mgx_search_not.argtypes = [C.c_void_p, u32p, C.c_uint64, u8p, u64p, C.c_uint64, u32p, C.c_uint64, u64p]
    L.mgx_filter_by_ngrams.argtypes = [C.c_void_p, u32p, C.c_uint64, u8p, u64p, C.c_uint64, u32p, C.c_uint64, u64p]
    L.mgx_search_by_threshold.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, C.c_uint64, u32p, C.c_uint64, u64p]
    L.mgx_eval_boolean.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_uint64, u8p, u64p,
                                   C.c_uint64, u32p, C.c_uint64, u64p]
    L.mgx_search_fuzzy.argtypes = [C.c_void_p, C.POINTER(ExpandedQuery), u8p, u64p, C.c_uint64, C.c_uint32, u32p,
                                   C.c_uint64, u64p]
    L.mgx_search_synonyms.argtypes = [C.c_void_p, C.POINTER(ExpandedQuery), u8p, u64p, u64p, C.c_uint64, u32p,
                                      C.c_uint64, u64p]
    L.mgx_search_fuzzy_batch.argtypes = [C.c_void_p, C.POINTER(ExpandedQuery), C.c_uint64, u8p, u64p, u64p, C.c_uint32,
                                         u32p, C.c_uint64, u64p]
    L.mgx_search_synonyms_batch.argtypes = [C.c_void_p, C.POINTER(ExpandedQuery), C.c_uint64, u8p, u64p, u64p, u64p,
                                            u32p, C.c_uint64, u64p]
    L.mgx_mgix_encode.argtypes = [C.POINTER(MgixInfo), u8p, u64p, u64p, u32p, C.c_double, u8p, C.c_uint64, u64p]
    L.mgx_mgix_decode.argtypes = [u8p, C.c_uint64, C.POINTER(MgixInfo), u8p, u64p, u64p, u32p]
    L.mgx_index_save_mgix.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, u8p, C.c_uint64, u64p]
    L.mgx_index_load_mgix.argtypes = [C.c_void_p, u8p, C.c_uint64]
    L.mgx_index_set_filter_column.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, u64p, u8p, C.c_uint64, u8p, u64p,
                                              C.c_uint64]
    L.mgx_query_batch_ex.argtypes = [C.c_void_p, C.POINTER(QueryParams), C.c_uint64, u8p, u64p, u64p, u8p, u64p, u64p,
                                     C.POINTER(QueryExt), C.c_uint64, u32p, f64p, u32p, u64p, u64p]
    L.mgx_query_batch.argtypes = [C.c_void_p, C.POINTER(QueryParams), C.c_uint64, u8p, u64p, u64p, u8p, u64p, u64p,
                                  C.c_uint64, u32p, f64p, u32p, u64p, u64p]
    L.mgx_batch_prepare.argtypes = [C.c_void_p, C.POINTER(QueryParams), C.c_uint64, u8p, u64p, u64p, u8p, u64p, u64p,
                                    C.c_void_p, C.POINTER(C.c_void_p)]
    L.mgx_batch_prepare_ex.argtypes = [C.c_void_p, C.POINTER(QueryParams), C.c_uint64, u8p, u64p, u64p, u8p, u64p, u64p,
                                       C.POINTER(QueryExt), C.c_void_p, C.POINTER(C.c_void_p)]
    L.mgx_batch_plan_device.argtypes = [C.c_void_p]
    L.mgx_batch_get_stats.argtypes = [C.c_void_p, C.POINTER(BatchStats)]
    L.mgx_batch_term_slots.restype = C.c_uint64
    L.mgx_batch_term_slots.argtypes = [C.c_void_p]
    L.mgx_batch_df_device.argtypes = [C.c_void_p, C.c_void_p]
    L.mgx_batch_search_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
    L.mgx_batch_destroy.argtypes = [C.c_void_p]
    L.mgx_merge_topk_device.argtypes = [C.c_int32, C.c_void_p, C.POINTER(QueryParams), C.c_uint32, C.c_uint64,
                                        C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    L.mgx_shard_record_layout.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(ShardRecordLayout)]
    L.mgx_batch_search_packed_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.mgx_merge_topk_packed_device.argtypes = [C.c_int32, C.c_void_p, C.POINTER(QueryParams), C.c_uint32, C.c_uint64,
                                               C.c_uint64, C.c_void_p, C.c_void_p]
    L.mgx_index_last_batch_stats.argtypes = [C.c_void_p, C.POINTER(BatchStats)]
    L.mgx_batch_set_streamed.argtypes = [C.c_void_p, C.c_int32]
    L.mgx_batch_overflowed.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    L.mgx_batch_reset.argtypes = [C.c_void_p]
    L.mgx_comm_unique_id.argtypes = [u8p]
    L.mgx_comm_create.argtypes = [u8p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    L.mgx_comm_destroy.argtypes = [C.c_void_p]
    L.mgx_comm_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                C.POINTER(C.c_int32)]
    L.mgx_sharded_batch_enqueue.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p]
    L.mgx_sharded_batch_finish.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
    L.mgx_share_open.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.POINTER(C.c_void_p)]
    L.mgx_share_close.argtypes = [C.c_void_p]
    L.mgx_share_close.restype = None
    L.mgx_share_publish.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32]
    L.mgx_share_import.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(QueryParams), C.c_void_p, C.c_int32,
                                   C.POINTER(C.c_void_p)]
    L.mgx_score_documents.argtypes = [C.c_void_p, u32p, C.c_uint64, u8p, u64p, u64p, C.c_uint64, C.c_uint64,
                                      C.c_double, C.c_double, C.c_double, f64p]
    L.mgx_sort_by_score.argtypes = [C.c_void_p, u32p, f64p, C.c_uint64, C.c_int32, C.c_uint32, C.c_uint32, u32p, u64p]
    _lib = L
    return L


def exported_symbols_in_header():
    """Names of every function include/mgx.h declares (for the symbol-presence test)."""
    import re
    text = open(INCLUDE).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mgx_[a-z0-9_]+)\s*\(", text)))


def _check(rc):
    if rc != MGX_OK:
        raise MgxError(rc, lib().mgx_last_error().decode("utf-8", "replace"))


def _ptr(arr, typ):
    return None if arr is None else arr.ctypes.data_as(typ)


def _bytes(s):
    return s.encode("utf-8") if isinstance(s, str) else bytes(s)


def pack_strings(strings):
    """list[bytes|str] -> (uint8 arena (>=1 byte), uint64 offsets[n+1])"""
    bs = [_bytes(s) for s in strings]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    joined = b"".join(bs)
    arena = np.frombuffer(joined, dtype=np.uint8).copy() if joined else np.zeros(1, np.uint8)
    return arena, offsets


def mgix_encode(terms, posting_offsets, postings, ngram_size=2, kanji_ngram_size=2, cross_boundary=True,
                normalize_nfkc=True, normalize_width="keep", normalize_lower=True, roaring_min_len=0.0):
    """Index::SaveToStream of an index given as CSR (terms: list[bytes]; doc ids ascending per term) -> the MGIX v4
    stream (bytes). Host-only codec."""
    info = MgixInfo(4, ngram_size, kanji_ngram_size, int(cross_boundary), int(normalize_nfkc), int(normalize_lower),
                    _bytes(normalize_width), len(terms), 0, 0)
    tb, to = pack_strings(terms)
    po = np.ascontiguousarray(posting_offsets, dtype=np.uint64)
    pp = np.ascontiguousarray(postings, dtype=np.uint32)
    if pp.size == 0:
        pp = np.zeros(1, np.uint32)
    n = C.c_uint64(0)
    rc = lib().mgx_mgix_encode(C.byref(info), _ptr(tb, u8p), _ptr(to, u64p), _ptr(po, u64p), _ptr(pp, u32p),
                               roaring_min_len, None, 0, C.byref(n))
    if rc != -4:
        _check(rc)
    out = np.zeros(n.value, dtype=np.uint8)
    _check(lib().mgx_mgix_encode(C.byref(info), _ptr(tb, u8p), _ptr(to, u64p), _ptr(po, u64p), _ptr(pp, u32p),
                                 roaring_min_len, _ptr(out, u8p), out.size, C.byref(n)))
    return out[:n.value].tobytes()


def mgix_decode(data):
    """Index::LoadFromStream of an MGIX stream -> (info dict, terms list[bytes] ascending, posting offsets,
    postings). Raises MgxError (MGX_ERR_FORMAT) for a stream the reference would reject. Host-only codec."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8).copy() if len(data) else np.zeros(1, np.uint8)
    info = MgixInfo()
    _check(lib().mgx_mgix_decode(_ptr(buf, u8p), len(data), C.byref(info), None, None, None, None))
    tb = np.zeros(max(1, info.term_bytes), dtype=np.uint8)
    to = np.zeros(info.n_terms + 1, dtype=np.uint64)
    po = np.zeros(info.n_terms + 1, dtype=np.uint64)
    pp = np.zeros(max(1, info.n_postings), dtype=np.uint32)
    _check(lib().mgx_mgix_decode(_ptr(buf, u8p), len(data), C.byref(info), _ptr(tb, u8p), _ptr(to, u64p),
                                 _ptr(po, u64p), _ptr(pp, u32p)))
    raw = tb.tobytes()
    terms = [raw[int(to[i]):int(to[i + 1])] for i in range(info.n_terms)]
    meta = {k: getattr(info, k) for k in ("version", "ngram_size", "kanji_ngram_size", "cross_boundary",
                                          "normalize_nfkc", "normalize_lower", "n_terms", "n_postings")}
    meta["normalize_width"] = info.normalize_width.decode()
    return meta, terms, po, pp[:info.n_postings].copy()


def key_to_utf8(key, width):
    """Packed key (width <= 3), or the words of a wide key (array of mgx_key_words(width) uint64), -> n-gram bytes."""
    out = np.zeros(48, dtype=np.uint8)
    if width <= 3:
        n = lib().mgx_key_to_utf8(int(key), width, _ptr(out, u8p))
    else:
        words = np.ascontiguousarray(key, dtype=np.uint64)
        n = lib().mgx_wide_key_to_utf8(_ptr(words, u64p), width, _ptr(out, u8p))
    if n < 0:
        raise MgxError(n, lib().mgx_last_error().decode())
    return out[:n].tobytes()


def tokenize_batch(texts, ngram_size=2, kanji_ngram_size=0, cross_boundary=True, device=0):
    """GenerateHybridNgrams for every text on the GPU -> list (per doc) of n-gram byte strings,
    in generation order (string_utils.cpp:452-509)."""
    arena, offsets = pack_strings(texts)
    cfg = IndexConfig(ngram_size, kanji_ngram_size, int(cross_boundary), device, 0.0, 0, 0, 0.0)
    cap = max(1, int(offsets[-1]))
    eff_kanji = kanji_ngram_size if kanji_ngram_size > 0 else ngram_size
    width = max(ngram_size, eff_kanji)
    nw = max(1, lib().mgx_key_words(width))
    keys = np.zeros(cap * nw, dtype=np.uint64)
    docs = np.zeros(cap, dtype=np.uint32)
    n = C.c_uint64(0)
    _check(lib().mgx_tokenize_batch(C.byref(cfg), _ptr(arena, u8p), _ptr(offsets, u64p), len(texts), _ptr(keys, u64p),
                                    _ptr(docs, u32p), cap, C.byref(n)))
    out = [[] for _ in texts]
    if width <= 3:
        for k, d in zip(keys[:n.value], docs[:n.value]):
            out[int(d)].append(key_to_utf8(k, width))
    else:
        for i in range(n.value):
            out[int(docs[i])].append(key_to_utf8(keys[i * nw:(i + 1) * nw], width))
    return out


@dataclass
class BatchResult:
    ids: np.ndarray     # [Q, stride] uint32
    scores: np.ndarray  # [Q, stride] float64
    count: np.ndarray   # [Q] uint32
    total: np.ndarray   # [Q] uint64
    df: np.ndarray      # [n_term_slots] uint64


def flatten_queries(queries):
    flat, begin = [], [0]
    for q in queries:
        flat += [_bytes(t) for t in q]
        begin.append(len(flat))
    arena, offsets = pack_strings(flat)
    return arena, offsets, np.asarray(begin, dtype=np.uint64), len(flat)


class Index:
    """Mirror of mygramdb::index::Index (src/index/index.h:49-413) backed by a device-resident shard."""

    def __init__(self, ngram_size=2, kanji_ngram_size=0, cross_boundary_ngrams=True, device=0, dense_threshold=0.0,
                 max_dense_bytes=0, scratch_bytes=0, roaring_threshold=0.18):
        self.ngram_size = ngram_size
        self.kanji_ngram_size = kanji_ngram_size
        self.cross_boundary_ngrams = cross_boundary_ngrams
        self.device = device
        cfg = IndexConfig(ngram_size, kanji_ngram_size, int(cross_boundary_ngrams), device, dense_threshold,
                          max_dense_bytes, scratch_bytes, roaring_threshold)
        self._h = C.c_void_p()
        _check(lib().mgx_index_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().mgx_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- build -------------------------------------------------------------------------------------
    def add_document_batch(self, doc_ids, texts):
        """Index::AddDocumentBatch (index.cpp:76-119): ADDITIVE, ids in any order; returns the number of documents
        whose text yields at least one n-gram. Journaled; folded in before the next read."""
        arena, offsets = pack_strings(texts)
        ids = np.ascontiguousarray(doc_ids, dtype=np.uint32)
        if ids.size == 0:
            ids = np.zeros(1, dtype=np.uint32)
        n = C.c_uint64(0)
        _check(lib().mgx_index_add_document_batch(self._h, _ptr(ids, u32p), _ptr(arena, u8p), _ptr(offsets, u64p),
                                                  len(texts), C.byref(n)))
        return int(n.value)

    @staticmethod
    def _text_arg(text):
        b = _bytes(text)
        buf = np.frombuffer(b, dtype=np.uint8).copy() if b else np.zeros(1, np.uint8)
        return buf, len(b)

    def add_document(self, doc_id, text):
        """Index::AddDocument (index.cpp:39-74) + the DocumentStore mirror; False if the text yields no n-gram."""
        buf, n = self._text_arg(text)
        ok = C.c_int32(0)
        _check(lib().mgx_index_add_document(self._h, doc_id, _ptr(buf, u8p), n, C.byref(ok)))
        return bool(ok.value)

    def update_document(self, doc_id, old_text, new_text):
        """Index::UpdateDocument (index.cpp:121-173)."""
        ob, on = self._text_arg(old_text)
        nb, nn = self._text_arg(new_text)
        _check(lib().mgx_index_update_document(self._h, doc_id, _ptr(ob, u8p), on, _ptr(nb, u8p), nn))

    def remove_document(self, doc_id, text):
        """Index::RemoveDocument (index.cpp:175-197)."""
        buf, n = self._text_arg(text)
        _check(lib().mgx_index_remove_document(self._h, doc_id, _ptr(buf, u8p), n))

    def get_statistics(self):
        """Index::GetStatistics (index.cpp:604-633)."""
        s = IndexStatistics()
        _check(lib().mgx_index_get_statistics(self._h, C.byref(s)))
        return s

    def optimize(self, total_docs):
        """Index::Optimize(total_docs) (index_optimization.cpp:36-120)."""
        _check(lib().mgx_index_optimize(self._h, total_docs))

    def trim(self):
        """Release the build workspace kept between rebuilds."""
        _check(lib().mgx_index_trim(self._h))

    def clear(self):
        """Index::Clear (index.cpp:635-641)."""
        _check(lib().mgx_index_clear(self._h))

    def commit(self):
        _check(lib().mgx_index_commit(self._h))

    def set_commit_mode(self, overlapped):
        """overlapped=True: reading calls neither commit journaled mutations nor wait for a commit in progress (they
        answer from the current generation); commit() publishes them (mgx_index_set_commit_mode)."""
        _check(lib().mgx_index_set_commit_mode(self._h, 1 if overlapped else 0))

    def build(self, doc_ids, arena, offsets):
        doc_ids = np.ascontiguousarray(doc_ids, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if arena.size == 0:
            arena = np.zeros(1, np.uint8)
        _check(lib().mgx_index_build(self._h, _ptr(doc_ids, u32p), _ptr(arena, u8p), _ptr(offsets, u64p),
                                     doc_ids.size))

    def build_device(self, d_doc_ids_ptr, d_text_ptr, d_offsets_ptr, n_docs):
        _check(lib().mgx_index_build_device(self._h, d_doc_ids_ptr, d_text_ptr, d_offsets_ptr, n_docs))

    def stats(self):
        s = IndexStats()
        _check(lib().mgx_index_get_stats(self._h, C.byref(s)))
        return s

    def term_count(self):
        return int(self.stats().n_terms)

    def posting_size(self, term):
        b = _bytes(term)
        buf = np.frombuffer(b, dtype=np.uint8).copy() if b else np.zeros(1, np.uint8)
        out = C.c_uint64(0)
        _check(lib().mgx_index_posting_size(self._h, _ptr(buf, u8p), len(b), C.byref(out)))
        return out.value

    count = posting_size

    def posting_payload(self, term):
        """(local docs, first-occurrence words, second-occurrence words, (offset, next, prev) bits) of one n-gram;
        see mgx_index_get_posting_payload."""
        b = _bytes(term)
        buf = np.frombuffer(b, dtype=np.uint8).copy() if b else np.zeros(1, np.uint8)
        n = C.c_uint64(0)
        layout = (C.c_int * 3)()
        _check(lib().mgx_index_get_posting_payload(self._h, _ptr(buf, u8p), len(b), None, None, None, 0, C.byref(n),
                                                   layout))
        cap = max(1, n.value)
        docs = np.zeros(cap, np.uint32)
        first = np.zeros(cap, np.uint32)
        second = np.zeros(cap, np.uint32)
        _check(lib().mgx_index_get_posting_payload(self._h, _ptr(buf, u8p), len(b), _ptr(docs, u32p),
                                                   _ptr(first, u32p), _ptr(second, u32p), cap, C.byref(n), layout))
        k = n.value
        return docs[:k], first[:k], second[:k], tuple(layout)

    def export(self):
        """-> (terms list[bytes] ascending, posting offsets uint64[T+1], postings uint32[P] global ids)"""
        s = self.stats()
        keys = np.zeros(max(1, s.n_terms), dtype=np.uint64)
        offs = np.zeros(s.n_terms + 1, dtype=np.uint64)
        posts = np.zeros(max(1, s.n_postings), dtype=np.uint32)
        _check(lib().mgx_index_export(self._h, _ptr(keys, u64p), _ptr(offs, u64p), _ptr(posts, u32p)))
        if s.key_width <= 3:
            terms = [key_to_utf8(k, s.key_width) for k in keys[:s.n_terms]]
        else:
            terms = self.export_terms()
        return terms, offs, posts[:s.n_postings]

    def export_terms(self):
        """The index's n-grams as byte strings in ascending order (any key width)."""
        s = self.stats()
        tb = np.zeros(max(1, s.n_terms * 4 * s.key_width), dtype=np.uint8)
        toff = np.zeros(s.n_terms + 1, dtype=np.uint64)
        nb = C.c_uint64(0)
        _check(lib().mgx_index_export_terms(self._h, _ptr(tb, u8p), len(tb), _ptr(toff, u64p), C.byref(nb)))
        raw = tb.tobytes()
        return [raw[int(toff[t]):int(toff[t + 1])] for t in range(s.n_terms)]

    def load_mgix(self, stream):
        """Index::LoadFromStream (index_serialization.cpp:279-613): the stream's posting lists replace the index
        content; the shard then holds no document text (see mgx_index_load_mgix)."""
        buf = np.frombuffer(bytes(stream), dtype=np.uint8)
        _check(lib().mgx_index_load_mgix(self._h, _ptr(buf, u8p), buf.size))

    def save_mgix(self, normalize_nfkc=True, normalize_width="keep", normalize_lower=True):
        """Index::SaveToStream (index_serialization.cpp:111-224) of the device index -> bytes."""
        st = self.stats()
        n = C.c_uint64(0)
        # one pass when the guess holds (np.zeros pages are only touched when written); exact size otherwise
        out = np.zeros(64 + 40 * st.n_terms + 9 * st.n_postings // 2, dtype=np.uint8)
        rc = lib().mgx_index_save_mgix(self._h, int(normalize_nfkc), _bytes(normalize_width), int(normalize_lower),
                                       _ptr(out, u8p), out.size, C.byref(n))
        if rc == -4:
            out = np.zeros(n.value, dtype=np.uint8)
            rc = lib().mgx_index_save_mgix(self._h, int(normalize_nfkc), _bytes(normalize_width),
                                           int(normalize_lower), _ptr(out, u8p), out.size, C.byref(n))
        _check(rc)
        return out[:n.value].tobytes()

    def doc_lengths(self):
        s = self.stats()
        out = np.zeros(max(1, s.n_docs), dtype=np.uint32)
        _check(lib().mgx_index_doc_lengths(self._h, _ptr(out, u32p)))
        return out[:s.n_docs]

    # -- set algebra -------------------------------------------------------------------------------
    def _set_call(self, fn, terms, pre=(), cap=4096):
        arena, offsets = pack_strings(terms)
        while True:
            out = np.zeros(cap, dtype=np.uint32)
            n = C.c_uint64(0)
            rc = fn(*pre, _ptr(arena, u8p), _ptr(offsets, u64p), len(terms), out, cap, n)
            if rc == -4:
                cap = max(int(n.value), cap * 2)
                continue
            _check(rc)
            return out[:n.value].copy()

    def search_and(self, terms, limit=0, reverse=False):
        """Index::SearchAnd (index.cpp:199-368); terms are n-gram strings."""
        L = lib()
        return self._set_call(lambda a, o, nt, out, cap, n: L.mgx_search_and(self._h, a, o, nt, limit, int(reverse),
                                                                              _ptr(out, u32p), cap, C.byref(n)), terms)

    def search_or(self, terms):
        L = lib()
        return self._set_call(lambda a, o, nt, out, cap, n: L.mgx_search_or(self._h, a, o, nt, _ptr(out, u32p), cap,
                                                                             C.byref(n)), terms)

    def search_not(self, all_docs, terms):
        L = lib()
        ad = np.ascontiguousarray(all_docs, dtype=np.uint32)
        adp = ad if ad.size else np.zeros(1, np.uint32)
        return self._set_call(lambda a, o, nt, out, cap, n: L.mgx_search_not(self._h, _ptr(adp, u32p), ad.size, a, o,
                                                                              nt, _ptr(out, u32p), cap, C.byref(n)),
                              terms, cap=max(4096, ad.size))

    def filter_by_ngrams(self, candidates, terms):
        L = lib()
        c = np.ascontiguousarray(candidates, dtype=np.uint32)
        cp = c if c.size else np.zeros(1, np.uint32)
        return self._set_call(lambda a, o, nt, out, cap, n: L.mgx_filter_by_ngrams(self._h, _ptr(cp, u32p), c.size, a,
                                                                                    o, nt, _ptr(out, u32p), cap,
                                                                                    C.byref(n)),
                              terms, cap=max(4096, c.size))

    def postings(self, term):
        return self.search_and([term])

    def search_by_threshold(self, terms, threshold):
        """Index::SearchByThreshold (index.cpp:488-578); terms are n-gram strings."""
        L = lib()
        return self._set_call(lambda a, o, nt, out, cap, n: L.mgx_search_by_threshold(self._h, a, o, nt, threshold,
                                                                                       _ptr(out, u32p), cap,
                                                                                       C.byref(n)), terms)

    def eval_boolean(self, ops, args, terms):
        """QueryNode::Evaluate (query_ast.cpp:67-161) of a postfix program: op 0 TERM(arg = term index),
        1 AND(arg = children), 2 OR(arg = children), 3 NOT. `terms` are search terms (not n-grams)."""
        L = lib()
        o = np.ascontiguousarray(ops, dtype=np.int32)
        a = np.ascontiguousarray(args, dtype=np.int32)
        op = o if o.size else np.zeros(1, np.int32)
        ap = a if a.size else np.zeros(1, np.int32)
        i32p = C.POINTER(C.c_int32)
        return self._set_call(lambda ar, of, nt, out, cap, n: L.mgx_eval_boolean(
            self._h, op.ctypes.data_as(i32p), ap.ctypes.data_as(i32p), o.size, ar, of, nt, _ptr(out, u32p), cap,
            C.byref(n)), terms)

    # -- fuzzy / synonym execution paths ---------------------------------------------------------------
    def _expanded(self, not_terms, filters, verify_text, raw_ngram, raw_kanji):
        eq = ExpandedQuery()
        eq.ngram_size = self.ngram_size if raw_ngram is None else raw_ngram
        eq.kanji_ngram_size = self.kanji_ngram_size if raw_kanji is None else raw_kanji
        eq.cross_boundary = int(self.cross_boundary_ngrams)
        eq.verify_text = verify_text
        nb, no = pack_strings(list(not_terms))
        filters = list(filters or [])
        fc = np.asarray([f[0] for f in filters] or [0], dtype=np.uint32)
        fo = np.asarray([f[1] for f in filters] or [0], dtype=np.uint8)
        lb, lo = pack_strings([_bytes(f[2]) for f in filters] or [b""])
        eq.not_bytes, eq.not_offsets, eq.n_not = _ptr(nb, u8p), _ptr(no, u64p), len(not_terms)
        eq.filter_col, eq.filter_op = _ptr(fc, u32p), _ptr(fo, u8p)
        eq.filter_bytes, eq.filter_offsets, eq.n_filters = _ptr(lb, u8p), _ptr(lo, u64p), len(filters)
        return eq, (nb, no, fc, fo, lb, lo)

    def search_fuzzy(self, terms, max_distance, not_terms=(), filters=None, verify_text=0, raw_ngram=None,
                     raw_kanji=None):
        """search_pipeline::ExecuteWithFuzzy (search_pipeline.cpp:1659-1740) over normalised terms; filters: list of
        (column id, op 0..5, literal). Returns the ascending doc ids."""
        L = lib()
        eq, keep = self._expanded(not_terms, filters, verify_text, raw_ngram, raw_kanji)
        return self._set_call(lambda a, o, nt, out, cap, n: L.mgx_search_fuzzy(
            self._h, C.byref(eq), a, o, nt, max_distance, _ptr(out, u32p), cap, C.byref(n)), list(terms))

    def search_synonyms(self, groups, not_terms=(), filters=None, verify_text=0, raw_ngram=None, raw_kanji=None):
        """search_pipeline::ExecuteWithSynonyms (search_pipeline.cpp:1580-1657) over expanded groups: a list of
        lists of variants (the normalised term and its synonyms). Returns the ascending doc ids."""
        L = lib()
        eq, keep = self._expanded(not_terms, filters, verify_text, raw_ngram, raw_kanji)
        flat, gbeg = [], [0]
        for g in groups:
            flat += list(g)
            gbeg.append(len(flat))
        gb = np.asarray(gbeg, dtype=np.uint64)
        return self._set_call(lambda a, o, nt, out, cap, n: L.mgx_search_synonyms(
            self._h, C.byref(eq), a, o, _ptr(gb, u64p), len(groups), _ptr(out, u32p), cap, C.byref(n)), flat)

    def _grouped_call(self, fn, n_queries, cap=1 << 24):
        """fn(out, cap, offsets) -> status; repeats the call once with the reported size if the capacity was too small
        (the batch runs again: callers that know their result sizes pass a fitting capacity)."""
        offs = np.zeros(n_queries + 1, dtype=np.uint64)
        for _ in range(2):
            out = np.empty(max(1, cap), dtype=np.uint32)
            rc = fn(out, cap, offs)
            if rc == MGX_ERR_CAPACITY:
                cap = int(offs[-1])
                continue
            _check(rc)
            return [out[int(offs[q]):int(offs[q + 1])].copy() for q in range(n_queries)]
        _check(rc)

    def search_fuzzy_batch(self, queries, max_distance, not_terms=(), filters=None, verify_text=0, raw_ngram=None,
                           raw_kanji=None):
        """mgx_search_fuzzy_batch: `queries` is a list of term lists; returns one ascending id array per query."""
        L = lib()
        eq, keep = self._expanded(not_terms, filters, verify_text, raw_ngram, raw_kanji)
        flat, qbeg = [], [0]
        for q in queries:
            flat += [_bytes(t) for t in q]
            qbeg.append(len(flat))
        a, o = pack_strings(flat if flat else [b""])
        qb = np.asarray(qbeg, dtype=np.uint64)
        return self._grouped_call(lambda out, cap, offs: L.mgx_search_fuzzy_batch(
            self._h, C.byref(eq), len(queries), _ptr(a, u8p), _ptr(o, u64p), _ptr(qb, u64p), max_distance,
            _ptr(out, u32p), cap, _ptr(offs, u64p)), len(queries))

    def search_synonyms_batch(self, queries, not_terms=(), filters=None, verify_text=0, raw_ngram=None, raw_kanji=None):
        """mgx_search_synonyms_batch: `queries` is a list of group lists (each group a list of variants)."""
        L = lib()
        eq, keep = self._expanded(not_terms, filters, verify_text, raw_ngram, raw_kanji)
        flat, gbeg, qbeg = [], [0], [0]
        for groups in queries:
            for g in groups:
                flat += [_bytes(v) for v in g]
                gbeg.append(len(flat))
            qbeg.append(len(gbeg) - 1)
        a, o = pack_strings(flat if flat else [b""])
        gb = np.asarray(gbeg, dtype=np.uint64)
        qb = np.asarray(qbeg, dtype=np.uint64)
        return self._grouped_call(lambda out, cap, offs: L.mgx_search_synonyms_batch(
            self._h, C.byref(eq), len(queries), _ptr(a, u8p), _ptr(o, u64p), _ptr(gb, u64p), _ptr(qb, u64p),
            _ptr(out, u32p), cap, _ptr(offs, u64p)), len(queries))

    # -- batched pipeline --------------------------------------------------------------------------
    def params(self, score=True, descending=True, limit=100, offset=0, verify_text=0, k1=1.2, b=0.75, total_docs=0,
               total_doc_length=0, raw_ngram=None, raw_kanji=None):
        return QueryParams(self.ngram_size if raw_ngram is None else raw_ngram,
                           self.kanji_ngram_size if raw_kanji is None else raw_kanji,
                           int(self.cross_boundary_ngrams), int(score), int(descending), limit, offset, verify_text,
                           k1, b, total_docs, total_doc_length)

    def set_filter_column(self, column, type_code, values):
        """Device mirror of one DocumentStore filter column. type_code = FilterValue variant index (1 bool, 2 int8,
        ... 8 int64, 9 uint64, 10 TIME, 11 string, 12 double); values: one python value per document of the index,
        None = NULL."""
        n = len(values)
        vals = np.zeros(max(1, n), dtype=np.uint64)
        nulls = np.zeros(max(1, n), dtype=np.uint8)
        strings = []
        for i, v in enumerate(values):
            if v is None:
                nulls[i] = 1
            elif type_code == 11:
                vals[i] = len(strings)
                strings.append(_bytes(v))
            elif type_code == 12:
                vals[i] = np.float64(v).view(np.uint64)
            else:
                vals[i] = np.int64(int(v)).view(np.uint64) if int(v) < 0 else np.uint64(int(v))
        sb, so = pack_strings(strings if strings else [b""])
        _check(lib().mgx_index_set_filter_column(self._h, column, type_code, _ptr(vals, u64p), _ptr(nulls, u8p), n,
                                                 _ptr(sb, u8p), _ptr(so, u64p), len(strings)))

    def set_filter_column_arrays(self, column, type_code, values_u64, nulls_u8=None, strings=None):
        """Same as set_filter_column for columns that already are arrays: values_u64[i] = the integer / bool value,
        the bits of the double, or (type 11) an index into `strings`; nulls_u8[i] != 0 marks NULL."""
        vals = np.ascontiguousarray(values_u64, dtype=np.uint64)
        nulls = None if nulls_u8 is None else np.ascontiguousarray(nulls_u8, dtype=np.uint8)
        strs = [_bytes(s) for s in (strings or [])]
        sb, so = pack_strings(strs if strs else [b""])
        _check(lib().mgx_index_set_filter_column(self._h, column, type_code, _ptr(vals, u64p), _ptr(nulls, u8p),
                                                 vals.size, _ptr(sb, u8p), _ptr(so, u64p), len(strs)))

    @staticmethod
    def build_ext(programs=None, filters=None):
        """(mgx_query_ext_t, arrays to keep alive) for per-query boolean programs / column conditions, or (None, [])."""
        ext = None
        keep = []
        if programs is not None or filters is not None:
            ext = QueryExt()
            i32p = C.POINTER(C.c_int32)
            if programs is not None:
                ops, args, pbeg = [], [], [0]
                for pr in programs:
                    if pr is not None:
                        ops += list(pr[0])
                        args += list(pr[1])
                    pbeg.append(len(ops))
                o = np.asarray(ops if ops else [0], dtype=np.int32)
                a = np.asarray(args if args else [0], dtype=np.int32)
                pb = np.asarray(pbeg, dtype=np.uint64)
                keep += [o, a, pb]
                ext.prog_ops, ext.prog_args, ext.q_prog_begin = o.ctypes.data_as(i32p), a.ctypes.data_as(i32p), _ptr(pb, u64p)
            if filters is not None:
                cols, fops, lits, fbeg = [], [], [], [0]
                for fl in filters:
                    for (c, op, lit) in (fl or []):
                        cols.append(c)
                        fops.append(op)
                        lits.append(_bytes(lit))
                    fbeg.append(len(cols))
                fc = np.asarray(cols if cols else [0], dtype=np.uint32)
                fo = np.asarray(fops if fops else [0], dtype=np.uint8)
                lb, lo = pack_strings(lits if lits else [b""])
                fb = np.asarray(fbeg, dtype=np.uint64)
                keep += [fc, fo, lb, lo, fb]
                ext.filter_col, ext.filter_op = _ptr(fc, u32p), _ptr(fo, u8p)
                ext.filter_bytes, ext.filter_offsets, ext.q_filter_begin = _ptr(lb, u8p), _ptr(lo, u64p), _ptr(fb, u64p)
        return ext, keep

    def query_batch(self, queries, not_terms=None, stride=None, programs=None, filters=None, **kw):
        """Batch of SEARCH queries (regular path of ExecuteFullPipeline + ScoreDocuments + SortByScore).
        programs: per query None or (ops, args) — a boolean postfix program over the query's own terms;
        filters: per query a list of (column id, op 0..5, literal)."""
        p = self.params(**kw)
        arena, offsets, qbeg, n_slots = flatten_queries(queries)
        if not_terms is not None:
            narena, noffsets, nbeg, _ = flatten_queries(not_terms)
        else:
            narena = noffsets = nbeg = None
        ext, keep = self.build_ext(programs, filters)
        return self.query_batch_flat(p, len(queries), arena, offsets, qbeg, narena, noffsets, nbeg, n_slots, stride,
                                     ext=ext)

    def query_batch_flat(self, p, n_queries, arena, offsets, qbeg, narena=None, noffsets=None, nbeg=None,
                         n_slots=None, stride=None, out=None, ext=None):
        stride = stride or max(1, p.limit)
        if n_slots is None:
            n_slots = int(qbeg[-1])
        if out is None:
            out = BatchResult(np.zeros((n_queries, stride), dtype=np.uint32),
                              np.zeros((n_queries, stride), dtype=np.float64), np.zeros(n_queries, dtype=np.uint32),
                              np.zeros(n_queries, dtype=np.uint64), np.zeros(max(1, n_slots), dtype=np.uint64))
        _check(lib().mgx_query_batch_ex(self._h, C.byref(p), n_queries, _ptr(arena, u8p), _ptr(offsets, u64p),
                                        _ptr(qbeg, u64p), _ptr(narena, u8p), _ptr(noffsets, u64p), _ptr(nbeg, u64p),
                                        C.byref(ext) if ext is not None else None, stride, _ptr(out.ids, u32p),
                                        _ptr(out.scores, f64p), _ptr(out.count, u32p), _ptr(out.total, u64p),
                                        _ptr(out.df, u64p)))
        out.df = out.df[:n_slots]
        return out

    def last_batch_stats(self):
        s = BatchStats()
        _check(lib().mgx_index_last_batch_stats(self._h, C.byref(s)))
        return s


class BM25Scorer:
    """Mirror of mygramdb::index::BM25Scorer (src/index/bm25_scorer.h:43-83)."""

    @staticmethod
    def score_documents(index: Index, candidates, search_terms, term_doc_freqs, total_docs, avg_doc_length, k1=1.2,
                        b=0.75):
        if len(search_terms) != len(term_doc_freqs):
            # bm25_scorer.cpp:51-55 -> ErrorCode::kInvalidArgument
            raise MgxError(-1, "BM25 search_terms and term_doc_freqs must have identical lengths")
        c = np.ascontiguousarray(candidates, dtype=np.uint32)
        arena, offsets = pack_strings(search_terms)
        dfs = np.ascontiguousarray(term_doc_freqs, dtype=np.uint64)
        dfp = dfs if dfs.size else np.zeros(1, np.uint64)
        out = np.zeros(max(1, c.size), dtype=np.float64)
        cp = c if c.size else np.zeros(1, np.uint32)
        _check(lib().mgx_score_documents(index._h, _ptr(cp, u32p), c.size, _ptr(arena, u8p), _ptr(offsets, u64p),
                                         _ptr(dfp, u64p), len(search_terms), total_docs, avg_doc_length, k1, b,
                                         _ptr(out, f64p)))
        return out[:c.size]


class ResultSorter:
    """ResultSorter::SortByScore (src/query/result_sorter.cpp:661-716)."""

    @staticmethod
    def sort_by_score(index: Index, results, scores, descending=True, limit=100, offset=0):
        r = np.ascontiguousarray(results, dtype=np.uint32)
        s = np.ascontiguousarray(scores, dtype=np.float64)
        out = np.zeros(max(1, r.size if limit == 0 else min(limit, r.size)), dtype=np.uint32)  # limit 0 = everything
        n = C.c_uint64(0)
        rp = r if r.size else np.zeros(1, np.uint32)
        sp = s if s.size else np.zeros(1, np.float64)
        _check(lib().mgx_sort_by_score(index._h, _ptr(rp, u32p), _ptr(sp, f64p), r.size, int(descending), limit, offset,
                                       _ptr(out, u32p), C.byref(n)))
        return out[:n.value].copy()
