"""CPU-side checks of the drop-in boundary: libmgx.so loads, exports every symbol include/mgx.h declares, refuses to
run without a GPU (no CPU fallback), and the host-only helpers behave. No compute calls are made here."""
import ctypes as C
import os

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(mgx):
    lib = mgx.lib()
    names = mgx.exported_symbols_in_header()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.mgx_version()


def test_header_cites_reference_interfaces():
    text = open(os.path.join(os.path.dirname(__file__), "..", "include", "mgx.h")).read()
    for cite in ("index.cpp:199-368", "index.cpp:76-119", "bm25_scorer.cpp:47-99", "result_sorter.cpp:661-716",
                 "search_pipeline.cpp:1757-2059", "string_utils.cpp:452-509"):
        assert cite in text, cite
    assert "torch" not in text.lower().replace("no torch", "")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_no_cpu_fallback_without_a_device(mgx):
    if _has_gpu():
        pytest.skip("a GPU is visible; the refusal path needs a GPU-less host")
    with pytest.raises(mgx.MgxError) as e:
        mgx.Index(2, 0, True)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)
    cfg = mgx.IndexConfig(2, 0, 1, 0, 0.0, 0, 0)
    n = C.c_uint64(0)
    text = np.frombuffer(b"abc", dtype=np.uint8).copy()
    offs = np.array([0, 3], dtype=np.uint64)
    keys = np.zeros(8, np.uint64)
    docs = np.zeros(8, np.uint32)
    rc = mgx.lib().mgx_tokenize_batch(C.byref(cfg), text.ctypes.data_as(mgx.u8p), offs.ctypes.data_as(mgx.u64p), 1,
                                      keys.ctypes.data_as(mgx.u64p), docs.ctypes.data_as(mgx.u32p), 8, C.byref(n))
    assert rc == -5


def test_expanded_and_mgix_entry_points_validate_on_the_host(mgx):
    """mgx_search_fuzzy / mgx_search_synonyms / mgx_index_save_mgix refuse null handles before touching a device;
    the MGIX codec is host-only and reports sizes, capacity errors and rejected streams through its status codes."""
    lib = mgx.lib()
    n = C.c_uint64(7)
    eq = mgx.ExpandedQuery()
    assert lib.mgx_search_fuzzy(None, C.byref(eq), None, None, 0, 1, None, 0, C.byref(n)) == -1
    assert lib.mgx_search_synonyms(None, C.byref(eq), None, None, None, 0, None, 0, C.byref(n)) == -1
    assert lib.mgx_index_save_mgix(None, 1, b"keep", 1, None, 0, C.byref(n)) == -1
    info = mgx.MgixInfo()
    assert lib.mgx_mgix_encode(None, None, None, None, None, 0.0, None, 0, C.byref(n)) == -1
    assert lib.mgx_mgix_decode(None, 0, C.byref(info), None, None, None, None) == -1
    # sizing call, then a buffer one byte short, then the exact size
    stream = mgx.mgix_encode([b"ab", b"bc"], [0, 2, 3], [1, 5, 9], 2, 2, True)
    tb, to = mgx.pack_strings([b"ab", b"bc"])
    po = np.array([0, 2, 3], dtype=np.uint64)
    pp = np.array([1, 5, 9], dtype=np.uint32)
    hdr = mgx.MgixInfo(4, 2, 2, 1, 1, 1, b"keep", 2, 0, 0)
    args = (C.byref(hdr), tb.ctypes.data_as(mgx.u8p), to.ctypes.data_as(mgx.u64p), po.ctypes.data_as(mgx.u64p),
            pp.ctypes.data_as(mgx.u32p), 0.0)
    assert lib.mgx_mgix_encode(*args, None, 0, C.byref(n)) == -4 and n.value == len(stream)
    short = np.zeros(len(stream) - 1, np.uint8)
    assert lib.mgx_mgix_encode(*args, short.ctypes.data_as(mgx.u8p), short.size, C.byref(n)) == -4
    exact = np.zeros(len(stream), np.uint8)
    assert lib.mgx_mgix_encode(*args, exact.ctypes.data_as(mgx.u8p), exact.size, C.byref(n)) == 0
    assert exact.tobytes() == stream
    # decode with capacities that are too small is a capacity error, not a partial answer
    buf = np.frombuffer(stream, dtype=np.uint8).copy()
    small = mgx.MgixInfo()
    small.n_terms, small.n_postings, small.term_bytes = 1, 3, 4
    t2, o2, q2, p2 = np.zeros(8, np.uint8), np.zeros(3, np.uint64), np.zeros(3, np.uint64), np.zeros(3, np.uint32)
    assert lib.mgx_mgix_decode(buf.ctypes.data_as(mgx.u8p), buf.size, C.byref(small), t2.ctypes.data_as(mgx.u8p),
                               o2.ctypes.data_as(mgx.u64p), q2.ctypes.data_as(mgx.u64p),
                               p2.ctypes.data_as(mgx.u32p)) == -4
    assert small.n_terms == 2 and small.n_postings == 3  # the sizes needed come back
    buf[10] ^= 1
    assert lib.mgx_mgix_decode(buf.ctypes.data_as(mgx.u8p), buf.size, C.byref(info), None, None, None, None) == -6
    assert b"kStorageCRCMismatch" in lib.mgx_last_error()


def test_argument_validation_is_host_side(mgx):
    lib = mgx.lib()
    assert lib.mgx_index_create(None, None) == -1
    assert b"null" in lib.mgx_last_error()
    out = np.zeros(16, np.uint8)
    assert lib.mgx_key_to_utf8(0, 9, out.ctypes.data_as(mgx.u8p)) == -1
    assert lib.mgx_index_get_stats(None, None) == -1
    assert lib.mgx_batch_term_slots(None) == 0
    lib.mgx_index_destroy(None)
    lib.mgx_batch_destroy(None)


def test_packed_key_order_is_utf8_byte_order(mgx, oracle):
    """Dictionary order: integer order of packed keys == bytewise order of the UTF-8 n-grams
    (the reference sorts std::string n-grams, string_utils.h:192-196)."""
    rng = np.random.default_rng(5)
    cps = [0, 1, 0x41, 0x7F, 0x80, 0x7FF, 0x800, 0x3042, 0x4E00, 0xFFFF, 0x10000, 0x20000, 0x10FFFF]
    keys, strings = [], []
    for _ in range(400):
        n = int(rng.integers(1, 4))
        seq = [int(rng.choice(cps)) for _ in range(n)]
        key = 0
        for j in range(3):
            key = (key << 21) | ((seq[j] + 1) if j < n else 0)
        keys.append(key)
        strings.append(oracle.codepoints_to_utf8(seq))
        assert mgx.key_to_utf8(key, 3) == strings[-1]
    order_k = sorted(range(len(keys)), key=lambda i: keys[i])
    order_s = sorted(range(len(keys)), key=lambda i: strings[i])
    assert [strings[i] for i in order_k] == [strings[i] for i in order_s]


def test_wide_key_order_is_utf8_byte_order(mgx, oracle):
    """Keys of 4..10 code points: mgx_key_words(width) words of three 21-bit fields, compared word by word ==
    bytewise order of the UTF-8 n-grams; mgx_wide_key_to_utf8 decodes them."""
    lib = mgx.lib()
    assert [lib.mgx_key_words(w) for w in range(1, 11)] == [1, 1, 1, 2, 2, 2, 3, 3, 3, 4]
    assert lib.mgx_key_words(0) < 0 and lib.mgx_key_words(11) < 0
    rng = np.random.default_rng(6)
    cps = [0, 1, 0x41, 0x7F, 0x80, 0x7FF, 0x800, 0x3042, 0x4E00, 0xFFFF, 0x10000, 0x20000, 0x10FFFF]
    for width in (4, 6, 7, 10):
        nw = lib.mgx_key_words(width)
        keys, strings = [], []
        for _ in range(300):
            n = int(rng.integers(1, width + 1))
            seq = [int(rng.choice(cps)) for _ in range(n)]
            words = []
            for w in range(nw):
                word = 0
                for f in range(3):
                    j = 3 * w + f
                    word = (word << 21) | ((seq[j] + 1) if j < n else 0)
                words.append(word)
            keys.append(tuple(words))
            strings.append(oracle.codepoints_to_utf8(seq))
            assert mgx.key_to_utf8(np.array(words, dtype=np.uint64), width) == strings[-1]
        order_k = sorted(range(len(keys)), key=lambda i: keys[i])
        order_s = sorted(range(len(keys)), key=lambda i: strings[i])
        assert [strings[i] for i in order_k] == [strings[i] for i in order_s]


def _build_adapter_example(tmp_path, source="adapter_example.cpp"):
    import subprocess
    root = os.path.join(os.path.dirname(__file__), "..")
    exe = str(tmp_path / source.replace(".cpp", ""))
    libdir = os.path.abspath(os.path.join(root, "mygram-db_b200"))
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-o", exe,
                           os.path.join(libdir, "adapter", source), "-L" + libdir, "-lmgx",
                           "-Wl,-rpath," + libdir])
    return exe


def test_cpp_adapter_compiles_and_links(mgx, tmp_path):
    """The C++17 adapter (reference class signatures over the C ABI) builds against libmgx.so with plain g++."""
    import subprocess
    exe = _build_adapter_example(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    if _has_gpu():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 1 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_adapter_runs_on_gpu(mgx, tmp_path):
    import subprocess
    exe = _build_adapter_example(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SearchAnd({bc,cd}) -> 2 docs" in r.stdout  # tests/index/index_search_test.cpp:393-418


def test_cpp_adapter_test_program_compiles(mgx, tmp_path):
    _build_adapter_example(tmp_path, "adapter_test.cpp")


@pytest.mark.gpu
def test_cpp_adapter_assertions_on_gpu(mgx, tmp_path):
    """adapter/adapter_test.cpp: the reference's own unit-test cases (index_search_test, index_basic_test,
    query_ast_test, bm25_scorer_test, result_sorter_test, index_serialization_test) asserted through the C++ adapter."""
    import subprocess
    exe = _build_adapter_example(tmp_path, "adapter_test.cpp")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ADAPTER TESTS OK" in r.stdout, r.stdout + r.stderr


def test_shard_record_layout_matches_the_python_protocol(mgx):
    """The packed per-shard top-k record (one all-gather, one D2H): the C layout and sharded.py's agree, parts are
    aligned for their element types, and views into a record round-trip."""
    torch = pytest.importorskip("torch")
    from mygram_db_b200 import sharded
    lib = mgx.lib()
    for q, s in ((0, 100), (1, 1), (7, 3), (4096, 100), (65536, 10)):
        lay = mgx.ShardRecordLayout()
        assert lib.mgx_shard_record_layout(q, s, C.byref(lay)) == 0
        py = sharded.record_layout(q, s)
        assert (lay.scores_offset, lay.total_offset, lay.ids_offset, lay.count_offset, lay.bytes) == \
               (py["scores"], py["total"], py["ids"], py["count"], py["bytes"])
        assert lay.scores_offset % 8 == 0 and lay.total_offset % 8 == 0
        assert lay.ids_offset % 4 == 0 and lay.count_offset % 4 == 0 and lay.bytes % 16 == 0
        assert lay.bytes >= q * s * 12 + q * 12
    assert lib.mgx_shard_record_layout(1, 1, None) == -1
    rec = torch.zeros((2, sharded.record_layout(5, 3)["bytes"]), dtype=torch.uint8)
    ids, scores, count, total = sharded.record_views(rec, 5, 3)
    assert ids.shape == (2, 5, 3) and scores.shape == (2, 5, 3) and count.shape == (2, 5) and total.shape == (2, 5)
    ids[1, 4, 2] = 77
    scores[0, 0, 0] = 1.5
    count[1, 0] = 3
    total[0, 4] = 1 << 40
    again = sharded.record_views(rec, 5, 3)
    assert int(again[0][1, 4, 2]) == 77 and float(again[1][0, 0, 0]) == 1.5
    assert int(again[2][1, 0]) == 3 and int(again[3][0, 4]) == 1 << 40


def test_share_ring_opens_and_attaches_without_a_device(mgx):
    """mgx_share_open: rank 0 creates the shared-memory ring, another handle attaches with the same geometry; a
    handle that disagrees about the geometry is refused. No compute call, so it runs without a GPU."""
    import ctypes as C
    import os
    L = mgx.lib()
    name = f"/mgx_cpu_share_{os.getpid()}".encode()
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert L.mgx_share_open(name, 2, 0, 4, 1 << 16, C.byref(a)) == 0
    assert L.mgx_share_open(name, 2, 1, 4, 1 << 16, C.byref(b)) == 0
    assert L.mgx_share_open(name, 2, 1, 8, 1 << 16, C.byref(c)) != 0 and not c.value
    assert L.mgx_share_open(name, 2, 2, 4, 1 << 16, C.byref(c)) == -1
    assert os.path.exists("/dev/shm/" + name.decode()[1:])
    L.mgx_share_close(b)
    L.mgx_share_close(a)
    assert not os.path.exists("/dev/shm/" + name.decode()[1:])
