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
