"""Streamed batches (no host read-back between planning and the tile kernels), the overflow repeat, the C-level
sharded pipeline (mgx_sharded_batch_*) on one shard, and -- where the box has two GPUs -- the same pipeline over real
NCCL. Every path must give the answers of the synchronous form, bit for bit."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import corpus as corpus_mod

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def same_answers(a, b, limit):
    assert np.array_equal(a.count, b.count)
    assert np.array_equal(a.total, b.total)
    assert np.array_equal(a.df, b.df)
    valid = np.arange(a.ids.shape[1])[None, :] < a.count[:, None]
    assert np.array_equal(a.ids[valid], b.ids[valid])
    assert np.array_equal(a.scores[valid].view(np.uint64), b.scores[valid].view(np.uint64))


@pytest.fixture(scope="module")
def shard(mgx):
    c = corpus_mod.generate("cjk", 60000, 0xC2, alphabet=400, min_len=8, max_len=60)
    gi = mgx.Index(2, 0, True)
    gi.build(c.doc_ids, c.arena, c.offsets)
    return c, gi


@pytest.mark.parametrize("score,limit,offset", [(True, 100, 0), (True, 10, 5), (False, 50, 0), (True, 1000, 24)])
def test_streamed_equals_synchronous_form(mgx, oracle, shard, monkeypatch, score, limit, offset):
    c, gi = shard
    qs = corpus_mod.sample_queries(c, 700, 5, n_terms=3, min_cp=2, max_cp=4)
    monkeypatch.setenv("MGX_NO_STREAMED", "1")
    sync = gi.query_batch(qs, score=score, limit=limit, offset=offset)
    n_sync = gi.last_batch_stats().launches
    monkeypatch.delenv("MGX_NO_STREAMED")
    streamed = gi.query_batch(qs, score=score, limit=limit, offset=offset)
    n_streamed = gi.last_batch_stats().launches
    same_answers(streamed, sync, limit)
    assert n_streamed < n_sync, (n_streamed, n_sync)  # fused planning, no tile-map kernels
    if score and offset == 0 and limit == 100:
        oi = oracle.index(2, 0, True)
        oi.build_bulk(c.doc_ids, c.arena, c.offsets, 8)
        o = oi.query_batch(qs[:200], score=True, limit=100, n_threads=8)
        for q in range(200):
            n = int(o.count[q])
            assert int(streamed.total[q]) == int(o.total[q]) and int(streamed.count[q]) == n
            assert np.array_equal(streamed.ids[q, :n], o.ids[q, :n])
            assert np.allclose(streamed.scores[q, :n], o.scores[q, :n], rtol=1e-9, atol=0)


def test_streamed_both_df_paths(mgx, shard, monkeypatch):
    c, gi = shard
    qs = corpus_mod.sample_queries(c, 3000, 6, n_terms=2, min_cp=2, max_cp=3)
    out = {}
    for mode in ("tiles", "stream"):
        monkeypatch.setenv("MGX_DF_MODE", mode)
        out[mode] = gi.query_batch(qs, score=True, limit=100)
        st = gi.last_batch_stats()
        assert (st.df_stream_terms > 0) == (mode == "stream")
    same_answers(out["tiles"], out["stream"], 100)


def test_workspace_overflow_is_repeated_in_the_synchronous_form(mgx, shard, monkeypatch):
    c, gi = shard
    qs = corpus_mod.sample_queries(c, 500, 7, n_terms=2, min_cp=2, max_cp=3)
    want = gi.query_batch(qs, score=True, limit=100)
    monkeypatch.setenv("MGX_STREAM_TILE_CAP", "3")  # far fewer tiles than the batch has
    got = gi.query_batch(qs, score=True, limit=100)
    same_answers(got, want, 100)
    monkeypatch.delenv("MGX_STREAM_TILE_CAP")
    again = gi.query_batch(qs, score=True, limit=100)  # the stream's workspace recovers
    same_answers(again, want, 100)


def test_df_signature_filter_changes_no_answer(mgx, shard, monkeypatch):
    """The df stage rules entries out on the payload signatures before it probes a list or reads text; with
    MGX_DF_NO_SIG it runs the unfiltered form (every entry through the membership stage). Same answers, same df,
    streamed and synchronous, and far fewer candidates with the filter."""
    c, gi = shard
    qs = corpus_mod.sample_queries(c, 3000, 11, n_terms=3, min_cp=2, max_cp=4)
    monkeypatch.setenv("MGX_DF_MODE", "tiles")
    want = gi.query_batch(qs, score=True, limit=100)
    cand = gi.last_batch_stats().df_candidates
    assert cand > 0
    for envs in (("MGX_DF_NO_SIG",), ("MGX_NO_STREAMED",), ("MGX_DF_NO_SIG", "MGX_NO_STREAMED")):
        for e in envs:
            monkeypatch.setenv(e, "1")
        got = gi.query_batch(qs, score=True, limit=100)
        cand_here = gi.last_batch_stats().df_candidates
        for e in envs:
            monkeypatch.delenv(e)
        same_answers(got, want, 100)
        if "MGX_DF_NO_SIG" in envs:
            assert cand_here > 2 * cand, (cand_here, cand)


def test_df_terms_at_document_ends_and_long_terms(mgx, oracle, monkeypatch):
    """The df stage compares candidates at the recorded occurrences of the driver n-gram. Terms that end exactly at /
    would run past the end of a document (the arena holds the documents back to back), terms longer than the 12 bytes
    compared in registers, and repeated n-grams (tiny alphabet) must all agree with the oracle, with the signature
    filter on and off."""
    rnd = np.random.default_rng(17)
    alphabet = [chr(0x4E00 + i) for i in range(12)] + list("ab")
    docs = []
    for _ in range(6000):
        n = int(rnd.integers(1, 40))
        docs.append("".join(alphabet[int(i)] for i in rnd.integers(0, len(alphabet), n)).encode())
    ids = np.arange(1, len(docs) + 1, dtype=np.uint32)
    qs = []
    for _ in range(1500):
        d = docs[int(rnd.integers(0, len(docs)))].decode()
        n = int(rnd.integers(2, 9))  # up to 8 code points = 24 bytes: past the 12 compared in registers
        if len(d) < n:
            continue
        a = int(rnd.integers(0, len(d) - n + 1))
        if rnd.random() < 0.5:
            a = len(d) - n  # ends exactly where its document ends: one more character would cross into the next
        t = d[a:a + n]
        extra = alphabet[int(rnd.integers(0, len(alphabet)))]
        qs.append([t.encode(), (t + extra).encode()] if rnd.random() < 0.5 else [t.encode()])
    oi = oracle.index(2, 0, True)
    oi.add_texts(ids, docs)
    want = oi.query_batch(qs, score=True, limit=50)
    gi = mgx.Index(2, 0, True)
    gi.add_document_batch(ids, docs)
    monkeypatch.setenv("MGX_DF_MODE", "tiles")
    for env in (None, "MGX_DF_NO_SIG"):
        if env:
            monkeypatch.setenv(env, "1")
        got = gi.query_batch(qs, score=True, limit=50)
        if env:
            monkeypatch.delenv(env)
        assert np.array_equal(got.df, want.df), env
        assert np.array_equal(got.total, want.total) and np.array_equal(got.count, want.count)
        for q in range(len(qs)):  # (neighbours with scores within rounding may swap against the CPU)
            n = int(want.count[q])
            assert sorted(got.ids[q, :n]) == sorted(want.ids[q, :n]), qs[q]


def test_boolean_programs_and_filters_streamed(mgx, shard, monkeypatch):
    c, gi = shard
    rng = np.random.default_rng(3)
    n = c.n_docs
    status = (np.arange(n, dtype=np.uint64) % 3) + 1
    gi.set_filter_column_arrays(0, 8, status)
    base = corpus_mod.sample_queries(c, 400, 9, n_terms=3, min_cp=2, max_cp=3)
    queries, programs, filters = [], [], []
    for i, (a, b, cc) in enumerate(base):
        kind = i % 4
        if kind == 0:
            queries.append([a, b]); programs.append(([0, 0, 1], [0, 1, 2])); filters.append([])
        elif kind == 1:
            queries.append([a, b]); programs.append(([0, 0, 2], [0, 1, 2])); filters.append([])
        elif kind == 2:
            queries.append([a, b]); programs.append(([0, 0, 3, 1], [0, 1, 0, 2])); filters.append([])
        else:
            queries.append([a, b, cc]); programs.append(([0, 0, 2, 0, 1], [0, 1, 2, 2, 2])); filters.append([(0, 0, "1")])
    monkeypatch.setenv("MGX_NO_STREAMED", "1")
    sync = gi.query_batch(queries, programs=programs, filters=filters, score=False, limit=100)
    monkeypatch.delenv("MGX_NO_STREAMED")
    streamed = gi.query_batch(queries, programs=programs, filters=filters, score=False, limit=100)
    assert np.array_equal(sync.count, streamed.count) and np.array_equal(sync.total, streamed.total)
    valid = np.arange(100)[None, :] < sync.count[:, None]
    assert np.array_equal(sync.ids[valid], streamed.ids[valid])


def test_sharded_pipeline_single_shard(mgx, shard):
    """mgx_sharded_batch_enqueue / _finish without a communicator = one shard: same answers as mgx_query_batch,
    several batches in flight on separate streams, re-armed batch objects."""
    import torch
    import mgx_loader
    sharded = __import__("importlib").import_module("mygram_db_b200.sharded")
    c, gi = shard
    device = torch.device("cuda", 0)
    params = gi.params(score=True, limit=100, offset=0)
    comm = sharded.ShardComm(mgx, None, device)
    pipe = sharded.ShardPipeline(mgx, gi, params, 100, comm)
    streams = [torch.cuda.Stream(device=device) for _ in range(3)]
    batches, want = [], []
    for i in range(5):
        qs = corpus_mod.sample_queries(c, 300, 100 + i, n_terms=3, min_cp=2, max_cp=4)
        want.append(gi.query_batch(qs, score=True, limit=100))
        arena, offs, qbeg, _ = mgx.flatten_queries(qs)
        batches.append((arena, offs, qbeg))
    lay = sharded.record_layout(300, 100)
    for round_ in range(2):
        prepared, outs = [], []
        for i, (arena, offs, qbeg) in enumerate(batches):
            prepared.append(pipe.prepare(arena, offs, qbeg, 300, streams[i % 3]))
            outs.append(torch.empty(lay["bytes"], dtype=torch.uint8, pin_memory=True))
        for i, p in enumerate(prepared):
            pipe.enqueue(p, 0, outs[i])
        for i, p in enumerate(prepared):
            pipe.finish(p)
            if round_ == 1:  # re-arm and run the same object again
                pipe.rearm(p)
                pipe.enqueue(p, 0, outs[i])
                pipe.finish(p)
            ids, scores, count, total = [t.numpy() for t in sharded.record_views(outs[i], 300, 100)]
            w = want[i]
            assert np.array_equal(count.view(np.uint32), w.count) and np.array_equal(total.view(np.uint64), w.total)
            valid = np.arange(100)[None, :] < w.count[:, None]
            assert np.array_equal(ids.view(np.uint32)[valid], w.ids[valid])
            assert np.array_equal(scores[valid].view(np.uint64), w.scores[valid].view(np.uint64))
            pipe.release(p)
    assert pipe.repeats == 0


def test_stride_below_limit_plus_offset_is_refused(mgx, shard):
    import torch
    c, gi = shard
    params = gi.params(score=True, limit=10, offset=5)
    qs = corpus_mod.sample_queries(c, 4, 1)
    arena, offs, qbeg, _ = mgx.flatten_queries(qs)
    L = mgx.lib()
    h = C.c_void_p()
    mgx._check(L.mgx_batch_prepare(gi._h, C.byref(params), 4, mgx._ptr(arena, mgx.u8p), mgx._ptr(offs, mgx.u64p),
                                   mgx._ptr(qbeg, mgx.u64p), None, None, None, None, C.byref(h)))
    rec = torch.empty(1 << 16, dtype=torch.uint8, device="cuda")
    assert L.mgx_batch_search_packed_device(h, None, 10, C.c_void_p(rec.data_ptr())) == -1  # MGX_ERR_INVALID_ARGUMENT
    assert L.mgx_sharded_batch_enqueue(None, 0, h, 10, None) == -1
    L.mgx_batch_destroy(h)


def test_concurrent_single_calls_from_threads(mgx, oracle, shard):
    """The reference calls Index::SearchAnd from its worker pool (thread_pool.cpp:33): concurrent calls on one
    handle must be safe and give the single-threaded answers, with a writer (AddDocument + commit) in between."""
    from concurrent.futures import ThreadPoolExecutor
    c, gi = shard
    terms = []
    for d in range(0, 4000, 40):
        t = c.text(d).decode()
        terms.append([t[0:2].encode(), t[1:3].encode()])
    want = [gi.search_and(t) for t in terms]

    def one(i):
        r = gi.search_and(terms[i % len(terms)])
        return np.array_equal(r, want[i % len(terms)])

    with ThreadPoolExecutor(max_workers=16) as ex:
        assert all(ex.map(one, range(800)))
    qs = corpus_mod.sample_queries(c, 64, 77)
    base = gi.query_batch(qs, score=True, limit=20)

    def two(i):
        if i % 2 == 0:
            return np.array_equal(gi.search_and(terms[i % len(terms)]), want[i % len(terms)])
        r = gi.query_batch(qs, score=True, limit=20)
        return np.array_equal(r.total, base.total) and np.array_equal(r.count, base.count)

    with ThreadPoolExecutor(max_workers=12) as ex:
        assert all(ex.map(two, range(200)))


@pytest.mark.parametrize("config", ["c2", "c4"])
def test_two_rank_nccl_pipeline_equals_single_shard(config):
    """Real NCCL, world size 2 (skipped on a one-GPU box): the library's own exchanges must reproduce the answers
    of the unsharded index bit for bit, with several batches in flight and with a forced workspace overflow."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = os.path.join(ROOT, "tests", "support", "nccl_shard_check.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29571", script, "--config", config]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL SHARD CHECK OK" in r.stdout


def _rand_docs(rnd, n, alphabet="abcdefgh漢字東京都"):
    return [("".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 30)))).encode() for _ in range(n)]


def test_add_document_batch_is_additive_like_the_loader_loop(mgx, oracle):
    """Index::AddDocumentBatch as InitialLoader::FlushBatch calls it (initial_loader.cpp:450-512, index.cpp:76-119):
    1000-document batches one after the other, ids in any order; single-document mutations in between; the index
    must equal the oracle's incrementally maintained one after every read."""
    import random
    from test_gpu_parity import assert_same_index
    rnd = random.Random(5)
    gi = mgx.Index(2, 0, True)
    oi = oracle.index(2, 0, True)
    ids = list(range(1, 6001))
    docs = dict(zip(ids, _rand_docs(rnd, len(ids))))
    order = ids[:3000] + rnd.sample(ids[3000:], 3000)  # three ascending batches, then shuffled ones
    for b0 in range(0, len(order), 1000):
        batch = order[b0:b0 + 1000]
        n = gi.add_document_batch(batch, [docs[i] for i in batch])
        oi.add_texts(np.asarray(batch, dtype=np.uint32), [docs[i] for i in batch])
        assert n == sum(1 for i in batch if len(docs[i].decode()) >= 2)
        if b0 == 2000:  # a read in the middle of the load commits what has arrived so far
            assert_same_index(gi, oi)
            assert gi.stats().n_docs == 3000
            gi.remove_document(17, docs[17])
            oi.remove_document(17, docs[17])
            gi.update_document(18, docs[18], b"abcabc")
            oi.update_document(18, docs[18], b"abcabc")
            docs[18] = b"abcabc"
    assert_same_index(gi, oi)
    assert gi.stats().n_docs == 5999
    st = gi.stats()
    assert (st.total_doc_length, st.doc_count) == oi.bm25_stats()
    assert np.array_equal(gi.search_and([b"ab", b"bc"]), oi.search_and([b"ab", b"bc"]))


def test_filter_columns_follow_the_documents_through_mutations(mgx, oracle):
    """every journal commit used to drop the filter columns silently. They now follow the documents:
    survivors keep their values, documents the journal added or replaced are NULL until the column is set again."""
    import random
    rnd = random.Random(11)
    n = 4000
    ids = np.arange(10, 10 + 2 * n, 2, dtype=np.uint32)
    docs = _rand_docs(rnd, n, alphabet="abcd")
    gi = mgx.Index(2, 0, True)
    arena, offs = mgx.pack_strings(docs)
    gi.build(ids, arena, offs)
    status = [(i % 3) + 1 for i in range(n)]
    cat = [rnd.choice([b"x", b"y", b"z"]) for _ in range(n)]
    gi.set_filter_column(0, 8, status)
    gi.set_filter_column(1, 11, cat)
    qs = [[b"ab"], [b"bc"], [b"cd", b"da"]]
    fl = [[(0, 0, "1")], [(1, 0, "y")], [(0, 1, "2"), (1, 1, "x")]]
    before = gi.query_batch(qs, filters=fl, score=False, limit=4000)
    # mutations: remove some, update some (their column values become NULL), add new ones in gaps and at the end
    removed = set(int(i) for i in rnd.sample(list(ids), 300))
    for d in removed:
        gi.remove_document(d, b"")
    updated = set(int(i) for i in rnd.sample([int(i) for i in ids if int(i) not in removed], 300))
    new_text = {}
    for d in updated:
        new_text[d] = b"abcdabcd"
        gi.update_document(d, b"", new_text[d])
    added = {11: b"abcd", 13: b"bcda", 20001: b"cdab"}
    for d, t in added.items():
        gi.add_document(d, t)
    after = gi.query_batch(qs, filters=fl, score=False, limit=4000)
    # expectation from first principles: model the store
    rows = {}
    for i, d in enumerate(ids):
        d = int(d)
        if d in removed:
            continue
        if d in updated:
            rows[d] = (new_text[d], None, None)
        else:
            rows[d] = (docs[i], status[i], cat[i])
    for d, t in added.items():
        rows[d] = (t, None, None)

    def expect(q, f):
        out = []
        for d in sorted(rows):
            text, st_, ct = rows[d]
            if not all(t in text for t in q):
                continue
            ok = True
            for col, op, lit in f:
                v = st_ if col == 0 else ct
                lit_v = int(lit) if col == 0 else lit.encode()
                eq = v is not None and v == lit_v
                ok = ok and (eq if op == 0 else not eq)  # FilterIndex semantics: NULL is not indexed, != keeps it
            if ok:
                out.append(d)
        return out

    for qi in range(len(qs)):
        want = expect(qs[qi], fl[qi])
        assert int(after.total[qi]) == len(want), (qi, int(after.total[qi]), len(want))
        assert after.ids[qi, :len(want)].tolist() == want
    assert int(before.total[0]) > 0


def test_or_rooted_programs_are_expanded_by_driver(mgx, oracle, shard, monkeypatch):
    """A OR B, (A AND B) OR C, A OR B OR C with filters / offsets: the driver expansion (one internal query per child
    of the root, folded on the device) must equal the pass over every document, and the oracle."""
    c, gi = shard
    n = c.n_docs
    gi.set_filter_column_arrays(0, 8, (np.arange(n, dtype=np.uint64) % 3) + 1)
    base = corpus_mod.sample_queries(c, 300, 21, n_terms=3, min_cp=2, max_cp=3)
    queries, programs, filters = [], [], []
    for i, (a, b, cc) in enumerate(base):
        kind = i % 5
        if kind == 0:      # A OR B
            queries.append([a, b]); programs.append(([0, 0, 2], [0, 1, 2]))
        elif kind == 1:    # (A AND B) OR C
            queries.append([a, b, cc]); programs.append(([0, 0, 1, 0, 2], [0, 1, 2, 2, 2]))
        elif kind == 2:    # A OR B OR C
            queries.append([a, b, cc]); programs.append(([0, 0, 0, 2], [0, 1, 2, 3]))
        elif kind == 3:    # A OR (NOT B): cannot be expanded (a child that needs every document)
            queries.append([a, b]); programs.append(([0, 0, 3, 2], [0, 1, 0, 2]))
        else:              # A AND B (not OR-rooted)
            queries.append([a, b]); programs.append(([0, 0, 1], [0, 1, 2]))
        filters.append([(0, 0, "1")] if i % 3 == 0 else [])
    oi = oracle.index(2, 0, True)
    oi.build_bulk(c.doc_ids, c.arena, c.offsets, 8)
    cols = [(8, ((np.arange(n) % 3) + 1).tolist())]
    for limit, offset in ((100, 0), (7, 5), (0, 3)):
        monkeypatch.setenv("MGX_NO_OR_EXPANSION", "1")
        plain = gi.query_batch(queries, programs=programs, filters=filters, score=False, limit=limit, offset=offset,
                               stride=64 if limit == 0 else None)
        monkeypatch.delenv("MGX_NO_OR_EXPANSION")
        exp = gi.query_batch(queries, programs=programs, filters=filters, score=False, limit=limit, offset=offset,
                             stride=64 if limit == 0 else None)
        assert np.array_equal(plain.total, exp.total) and np.array_equal(plain.count, exp.count)
        valid = np.arange(plain.ids.shape[1])[None, :] < plain.count[:, None]
        assert np.array_equal(plain.ids[valid], exp.ids[valid])
    for qi in range(0, len(queries), 7):
        full = oi.eval_boolean(programs[qi][0], programs[qi][1], queries[qi])
        want = oracle.apply_filters(n, int(c.doc_ids[0]), cols, filters[qi], full) if filters[qi] else full
        got = gi.query_batch([queries[qi]], programs=[programs[qi]], filters=[filters[qi]], score=False, limit=100)
        assert int(got.total[0]) == want.size
        k = min(100, want.size)
        assert np.array_equal(got.ids[0, :k], want[:k])


def test_shared_compiled_batches_equal_local_compiles(mgx, shard):
    """mgx_share_*: a batch compiled by one shard process and imported by another must give the answers of a local
    compile, bit for bit -- also on a shard that holds invalid UTF-8 while the compiling shard did not (the streaming
    df pass the compiler chose is not exact there), after a re-arm, and a batch with column conditions is refused by
    publish and import alike. Both "ranks" live in this process (two handles on one segment)."""
    import torch
    sharded = __import__("importlib").import_module("mygram_db_b200.sharded")
    c, gi = shard
    device = torch.device("cuda", 0)
    # second shard: other documents, one of them not valid UTF-8
    c2 = corpus_mod.generate("cjk", 30000, 0xC7, alphabet=400, min_len=8, max_len=60)
    arena2 = c2.arena.copy()
    arena2[int(c2.offsets[5])] = 0xFF
    g2 = mgx.Index(2, 0, True)
    g2.build(c2.doc_ids, arena2, c2.offsets)
    assert g2.stats().all_valid_utf8 == 0 and gi.stats().all_valid_utf8 == 1
    params1 = gi.params(score=True, limit=100, offset=0)
    params2 = g2.params(score=True, limit=100, offset=0)
    comm = sharded.ShardComm(mgx, None, device)
    pub = sharded.ShardPipeline(mgx, gi, params1, 100, comm)
    imp = sharded.ShardPipeline(mgx, g2, params2, 100, comm)
    name = f"/mgx_test_share_{os.getpid()}"
    pub.open_share(name, 2, 0, n_slots=2, slot_bytes=1 << 20)
    imp.open_share(name, 2, 1, n_slots=2, slot_bytes=1 << 20)
    st = torch.cuda.Stream(device=device)
    lay = sharded.record_layout(3000, 100)
    out = torch.empty(lay["bytes"], dtype=torch.uint8, pin_memory=True)

    def answers(pipe, p):
        pipe.enqueue(p, 0, out)
        pipe.finish(p)
        return [t.numpy().copy() for t in sharded.record_views(out, 3000, 100)]

    for seq in range(5):  # more batches than slots: the ring wraps
        qs = corpus_mod.sample_queries(c, 3000, 200 + seq, n_terms=2, min_cp=2, max_cp=3)  # large: streaming df pass
        arena, offs, qbeg, _ = mgx.flatten_queries(qs)
        p0 = pub.prepare_shared(seq * 2, arena, offs, qbeg, 3000, st)         # rank 0's turn: compile + publish
        p1 = imp.prepare_shared(seq * 2, arena, offs, qbeg, 3000, st)         # rank 1 imports
        want1 = g2.query_batch(qs, score=True, limit=100)
        want0 = gi.query_batch(qs, score=True, limit=100)
        for pipe, p, w in ((pub, p0, want0), (imp, p1, want1)):
            for again in range(2):
                ids, scores, count, total = answers(pipe, p)
                assert np.array_equal(count.view(np.uint32), w.count) and np.array_equal(total.view(np.uint64), w.total)
                valid = np.arange(100)[None, :] < w.count[:, None]
                assert np.array_equal(ids.view(np.uint32)[valid], w.ids[valid])
                assert np.array_equal(scores[valid].view(np.uint64), w.scores[valid].view(np.uint64))
                pipe.rearm(p)
            pipe.release(p)
        # the odd sequence numbers are rank 1's turn: the invalid-UTF-8 shard compiles, the clean shard imports
        p1 = imp.prepare_shared(seq * 2 + 1, arena, offs, qbeg, 3000, st)
        p0 = pub.prepare_shared(seq * 2 + 1, arena, offs, qbeg, 3000, st)
        ids, scores, count, total = answers(pub, p0)
        assert np.array_equal(total.view(np.uint64), want0.total) and np.array_equal(count.view(np.uint32), want0.count)
        pub.release(p0)
        imp.release(p1)
    # a batch with a column condition cannot travel: publish and import agree, both sides compile locally
    n = gi.stats().n_docs
    gi.set_filter_column_arrays(0, 8, (np.arange(n) % 3).astype(np.int64))
    qs = corpus_mod.sample_queries(c, 8, 999)
    arena, offs, qbeg, _ = mgx.flatten_queries(qs)
    ext, keep = mgx.Index.build_ext(None, [[(0, 0, "1")]] * 8)
    L = mgx.lib()
    h = C.c_void_p()
    mgx._check(L.mgx_batch_prepare_ex(gi._h, C.byref(params1), 8, mgx._ptr(arena, mgx.u8p), mgx._ptr(offs, mgx.u64p),
                                      mgx._ptr(qbeg, mgx.u64p), None, None, None, C.byref(ext), None, C.byref(h)))
    assert L.mgx_share_publish(pub.share, 10, h, 1000) == mgx.MGX_ERR_UNSUPPORTED
    h2 = C.c_void_p()
    assert L.mgx_share_import(imp.share, 10, g2._h, C.byref(params2), None, 1000, C.byref(h2)) == mgx.MGX_ERR_UNSUPPORTED
    assert L.mgx_share_import(imp.share, 11, g2._h, C.byref(params2), None, 50, C.byref(h2)) == -7  # MGX_ERR_TIMEOUT: never published
    L.mgx_batch_destroy(h)
    imp.close_share()
    pub.close_share()
    g2.close()


def test_sort_by_score_any_offset_and_limit_zero(mgx):
    """ResultSorter::SortByScore (result_sorter.cpp:661-716) takes any offset and limit 0 = everything from offset on;
    the device version produces windows beyond the best 1024 records in runs bounded by selected keys."""
    rnd = np.random.default_rng(5)
    idx = mgx.Index(2, 0, True)
    idx.add_document_batch([1], ["ab"])
    n = 7000
    docs = rnd.permutation(50000)[:n].astype(np.uint32)
    scores = np.round(rnd.random(n) * 40) / 8.0  # many ties: the doc id decides
    for desc in (True, False):
        if desc:
            order = np.lexsort((-docs.astype(np.int64), -scores))
        else:
            order = np.lexsort((docs.astype(np.int64), scores))
        want = docs[order]
        for limit, offset in [(100, 0), (1000, 24), (1000, 25), (1000, 3000), (1000, 6500), (2500, 1000), (0, 0),
                              (0, 1500), (0, 6999), (0, 7000), (5, 7001), (1024, 1024), (1, 6999)]:
            got = mgx.ResultSorter.sort_by_score(idx, docs, scores, desc, limit, offset)
            end = n if limit == 0 else min(n, offset + limit)
            assert np.array_equal(got, want[min(offset, n):end]), (desc, limit, offset)


@pytest.mark.parametrize("streamed", [True, False])
def test_scored_batches_with_deep_offsets(mgx, oracle, shard, monkeypatch, streamed):
    """SORT _score LIMIT 1000 OFFSET n for n beyond the old 1024-record bound, and limit 0 (everything): the window
    equals the same rows of a run that returns the whole ranking, and the oracle's answer."""
    c, gi = shard
    if not streamed:
        monkeypatch.setenv("MGX_NO_STREAMED", "1")
    qs = corpus_mod.sample_queries(c, 24, 31, n_terms=1, min_cp=2, max_cp=2)  # single bigrams: long result lists
    full = gi.query_batch(qs, score=True, limit=0, offset=0, stride=60000)
    assert int(full.total.max()) > 3000
    oi = oracle.index(2, 0, True)
    oi.build_bulk(c.doc_ids, c.arena, c.offsets, 8)
    for limit, offset in [(1000, 1024), (1000, 2500), (300, 5000), (0, 1200)]:
        stride = 60000 if limit == 0 else limit
        r = gi.query_batch(qs, score=True, limit=limit, offset=offset, stride=stride)
        o = oi.query_batch(qs, score=True, limit=limit if limit else 60000, offset=offset, n_threads=8)
        for q in range(len(qs)):
            t = int(full.total[q])
            lo = min(offset, t)
            hi = t if limit == 0 else min(t, offset + limit)
            assert int(r.total[q]) == t and int(r.count[q]) == hi - lo
            assert np.array_equal(r.ids[q, :hi - lo], full.ids[q, lo:hi])
            assert np.array_equal(r.scores[q, :hi - lo].view(np.uint64), full.scores[q, lo:hi].view(np.uint64))
            assert int(o.count[q]) == hi - lo and np.array_equal(o.ids[q, :hi - lo], r.ids[q, :hi - lo])


def test_search_or_and_threshold_are_driven_by_lists(mgx, oracle, shard, monkeypatch):
    """Index::SearchOr / SearchByThreshold (index.cpp:410-448, 488-578) expanded into one list-driven query per
    sufficient list (a document in >= t of n lists is in one of any n - t + 1 of them) instead of a pass over every
    document: same answers as the single pass and as the oracle, for dense and sparse lists, unknown n-grams,
    repeated terms and more lists than the expansion takes."""
    import random
    c, gi = shard
    oi = oracle.index(2, 0, True)
    oi.build_bulk(c.doc_ids, c.arena, c.offsets, 8)
    rnd = random.Random(12)
    texts = [c.text(rnd.randrange(60000)).decode() for _ in range(400)]
    grams = [t[i:i + 2].encode() for t in texts for i in range(0, min(len(t) - 1, 6))]
    cases = []
    for _ in range(30):
        n = rnd.choice([1, 2, 3, 5, 8, 17, 24])
        terms = [rnd.choice(grams) for _ in range(n)]
        if rnd.random() < 0.3:
            terms.append("一鿿".encode())  # an n-gram no document holds
        if rnd.random() < 0.3:
            terms.append(terms[0])
        cases.append(terms)
    for terms in cases:
        want_or = oi.search_or(terms)
        uniq = len(set(terms))
        thresholds = sorted({1, 2, max(1, uniq // 2), max(1, uniq - 1), uniq})
        want_thr = [oi.search_by_threshold(terms, t) for t in thresholds]
        for pin in (False, True):
            if pin:
                monkeypatch.setenv("MGX_NO_OR_EXPANSION", "1")
            else:
                monkeypatch.delenv("MGX_NO_OR_EXPANSION", raising=False)
            assert np.array_equal(gi.search_or(terms), want_or), (pin, terms)
            for t, w in zip(thresholds, want_thr):
                assert np.array_equal(gi.search_by_threshold(terms, t), w), (pin, t, terms)
    monkeypatch.delenv("MGX_NO_OR_EXPANSION", raising=False)
