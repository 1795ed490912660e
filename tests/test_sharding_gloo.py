"""N > 1 protocol on CPU: world_size 2 over the gloo backend.

The sharding protocol of mygram-db_b200/sharded.py (df all-reduce -> per-shard top-k with GLOBAL statistics ->
all-gather -> merge) is driven here with an oracle-backed stand-in for the CUDA backend, and the merged answer must
equal the single-index answer bit for bit: doc-ID-range sharding with global (N, sum dl, df) cannot change a score.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIMIT, OFFSET = 10, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleShardBackend:
    """Same interface as MgxShardBackend, computed by the CPU oracle over this rank's shard."""

    def __init__(self, torch, index, total_docs, total_len, score, descending):
        self.torch, self.index = torch, index
        self.total_docs, self.total_len = total_docs, total_len
        self.score, self.descending = score, descending
        self.stride = LIMIT + OFFSET

    def local_df(self, batch):
        r = self.index.query_batch(batch, score=True, limit=1)
        self._slots = len(r.df)
        return self.torch.from_numpy(r.df.astype(np.int64))

    def search(self, batch, df):
        # the oracle recomputes df locally; inject the GLOBAL df by scoring the shard's result sets explicitly
        r = self.index.query_batch(batch, score=False, limit=0, stride=1, want_sets=True)
        Q, S = len(batch), self.stride
        ids = np.zeros((Q, S), np.uint32)
        scores = np.zeros((Q, S), np.float64)
        count = np.zeros(Q, np.int32)
        total = np.zeros(Q, np.int64)
        dfs = df.numpy().astype(np.uint64)
        slot = 0
        avgdl = self.total_len / self.total_docs if self.total_docs else 0.0
        for q, terms in enumerate(batch):
            tdf = dfs[slot:slot + len(terms)]
            slot += len(terms)
            docs = r.sets[q]
            total[q] = len(docs)
            if self.score:
                # term order = ascending estimated size, stable (search_pipeline.cpp:2012-2014)
                est = [min([self.index.posting_size(g) for g in self.index.L.ngrams("query", t, 2, 0, True)] or [2**63])
                       for t in terms]
                order = sorted(range(len(terms)), key=lambda i: est[i])
                sc = self.index.score_documents(docs, [terms[i] for i in order], [int(tdf[i]) for i in order],
                                                self.total_docs, avgdl)
                top = self.index.L.sort_by_score(docs, sc, self.descending, S, 0)
                pos = {int(d): i for i, d in enumerate(docs)}
                count[q] = len(top)
                ids[q, :len(top)] = top
                scores[q, :len(top)] = [sc[pos[int(d)]] for d in top]
            else:
                top = docs[:S]
                count[q] = len(top)
                ids[q, :len(top)] = top
        t = self.torch
        from mygram_db_b200.sharded import record_layout, record_views
        rec = t.zeros(record_layout(Q, S)["bytes"], dtype=t.uint8)
        v_ids, v_scores, v_count, v_total = record_views(rec, Q, S)
        v_ids.copy_(t.from_numpy(ids.view(np.int32)))
        v_scores.copy_(t.from_numpy(scores))
        v_count.copy_(t.from_numpy(count))
        v_total.copy_(t.from_numpy(total))
        self._q = Q
        return rec

    def merge(self, records):
        from mygram_db_b200.sharded import merge_topk_reference, record_views
        ids_all, scores_all, count_all, total_all = record_views(records, self._q, self.stride)
        return merge_topk_reference(self.score, self.descending, LIMIT, OFFSET, ids_all.numpy().view(np.uint32),
                                    scores_all.numpy(), count_all.numpy(), total_all.numpy(), LIMIT)


def _worker(rank, world, port, score, descending, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "support")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import corpus as corpus_mod
    import mgx_loader
    import pyoracle
    mgx_loader.load()
    from mygram_db_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = sharded.TorchDist(dist)
    n_total = 6000
    lo, hi = sharded.shard_range(n_total, world, rank)
    c = corpus_mod.generate("cjk", hi - lo, 0xC2, first_doc=lo, alphabet=128, min_len=4, max_len=30)
    lib = pyoracle.OracleLib(pyoracle.PORT_LIB)
    idx = lib.index(2, 0, True)
    idx.build_bulk(c.doc_ids, c.arena, c.offsets, 2)
    tl, dc = idx.bm25_stats()
    g = torch.tensor([dc, tl], dtype=torch.int64)
    comm.all_reduce_sum(g)
    backend = OracleShardBackend(torch, idx, int(g[0]), int(g[1]), score, descending)
    qs = corpus_mod.sample_queries_global("cjk", 0xC2, n_total, 120, 5, n_terms=2, min_cp=2, max_cp=3, alphabet=128,
                                          min_len=4, max_len=30)
    ids, scores, count, total = sharded.run_sharded_batch(backend, comm, qs)
    if rank == 0:
        np.savez(os.path.join(out_dir, "merged.npz"), ids=ids, scores=scores, count=count, total=total)
    dist.destroy_process_group()


@pytest.mark.parametrize("score,descending", [(True, True), (True, False), (False, True)])
def test_two_shards_equal_one_index(tmp_path, oracle, score, descending):
    import torch.multiprocessing as mp
    import corpus as corpus_mod
    port = _free_port()
    mp.spawn(_worker, args=(2, port, score, descending, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "merged.npz")
    c = corpus_mod.generate("cjk", 6000, 0xC2, alphabet=128, min_len=4, max_len=30)
    idx = oracle.index(2, 0, True)
    idx.build_bulk(c.doc_ids, c.arena, c.offsets, 2)
    qs = corpus_mod.sample_queries_global("cjk", 0xC2, 6000, 120, 5, n_terms=2, min_cp=2, max_cp=3, alphabet=128,
                                          min_len=4, max_len=30)
    want = idx.query_batch(qs, score=score, descending=descending, limit=LIMIT, offset=OFFSET)
    assert np.array_equal(got["total"], want.total)
    assert np.array_equal(got["count"], want.count)
    for q in range(len(qs)):
        n = int(want.count[q])
        assert got["ids"][q, :n].tolist() == want.ids[q, :n].tolist(), (q, qs[q])
        if score:
            assert np.array_equal(got["scores"][q, :n], want.scores[q, :n]), (q, qs[q])
    assert int(want.total.max()) > LIMIT + OFFSET  # the merge really had to choose


def _set_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "support")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import corpus as corpus_mod
    import mgx_loader
    import pyoracle
    mgx_loader.load()
    from mygram_db_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = sharded.TorchDist(dist)
    lo, hi = sharded.shard_range(SET_DOCS, world, rank)
    c = corpus_mod.generate("cjk", hi - lo, 0xC2, first_doc=lo, alphabet=128, min_len=4, max_len=30)
    idx = pyoracle.OracleLib(pyoracle.PORT_LIB).index(2, 0, True)
    idx.build_bulk(c.doc_ids, c.arena, c.offsets, 2)
    answers = []
    for kind, arg in _set_queries():
        if kind == "fuzzy":
            local = idx.search_fuzzy(arg, 1, verify_text=1)[0]
        elif kind == "synonyms":
            local = idx.search_synonyms(arg)[0]
        else:
            local = idx.search_or(arg)
        answers.append(sharded.gather_doc_id_sets(comm, local))
    if rank == 0:
        np.savez(os.path.join(out_dir, "sets.npz"), *answers)
    dist.destroy_process_group()


SET_DOCS = 5000


def _set_queries():
    import corpus as corpus_mod
    c = corpus_mod.generate("cjk", SET_DOCS, 0xC2, alphabet=128, min_len=4, max_len=30)
    rng = np.random.default_rng(12)
    out = []
    for i in range(30):
        t = c.text(int(rng.integers(0, SET_DOCS))).decode()
        u = c.text(int(rng.integers(0, SET_DOCS))).decode()
        if i % 3 == 0:
            out.append(("fuzzy", [t[:2] + chr(0x4E00 + 200) + t[3:5]]))   # one substituted character, distance 1
        elif i % 3 == 1:
            out.append(("synonyms", [[t[:2], u[:3]], [t[1:3], u[2:4]]]))
        else:
            out.append(("or", [t[:2], u[:2], "zz"]))
    out.append(("fuzzy", ["zzzz"]))  # empty everywhere
    return out


def test_two_shards_answer_set_queries_like_one_index(tmp_path, oracle):
    """Un-scored doc-id sets (SearchOr, the fuzzy and the synonym paths) over two doc-range shards: the shard answers
    concatenated in rank order (sharded.gather_doc_id_sets: a size all-gather + an id all-gather, no merge) equal the
    answer of one index over the whole corpus."""
    import torch.multiprocessing as mp
    import corpus as corpus_mod
    port = _free_port()
    mp.spawn(_set_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "sets.npz")
    c = corpus_mod.generate("cjk", SET_DOCS, 0xC2, alphabet=128, min_len=4, max_len=30)
    idx = oracle.index(2, 0, True)
    idx.build_bulk(c.doc_ids, c.arena, c.offsets, 2)
    nonempty = 0
    for i, (kind, arg) in enumerate(_set_queries()):
        if kind == "fuzzy":
            want = idx.search_fuzzy(arg, 1, verify_text=1)[0]
        elif kind == "synonyms":
            want = idx.search_synonyms(arg)[0]
        else:
            want = idx.search_or(arg)
        assert np.array_equal(got[f"arr_{i}"], want), (kind, arg)
        nonempty += want.size > 0
    assert nonempty >= 20


def test_shard_ranges_cover_everything():
    sys.path.insert(0, ROOT)
    import mgx_loader
    mgx_loader.load()
    from mygram_db_b200.sharded import shard_range
    for n in (0, 1, 7, 10_000_000):
        for g in (1, 2, 4, 8):
            r = [shard_range(n, g, k) for k in range(g)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(g - 1))


def test_gather_doc_id_sets_single_process_and_empty_answers():
    sys.path.insert(0, ROOT)
    import mgx_loader
    mgx_loader.load()
    from mygram_db_b200 import sharded
    comm = sharded.NoDist()
    ids = np.array([3, 9, 4_000_000_000], dtype=np.uint32)  # ids above 2^31 survive the int32 transport
    assert np.array_equal(sharded.gather_doc_id_sets(comm, ids), ids)
    assert sharded.gather_doc_id_sets(comm, np.zeros(0, dtype=np.uint32)).size == 0
