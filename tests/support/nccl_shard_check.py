"""tests/support/nccl_shard_check.py — run under torchrun with >= 2 ranks, one GPU each.

Doc-id-range shards of one corpus, the library's own NCCL exchanges (mgx_sharded_batch_*), several batches in flight;
rank 0 also holds the WHOLE corpus as one shard and checks that every merged answer equals the unsharded one bit for
bit (ids in order, scores, counts, totals) and the CPU oracle's on a sample. One round forces a workspace overflow on
rank 1 only, so the collective repeat is exercised. Prints 'NCCL SHARD CHECK OK' on success."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "support")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--docs", type=int, default=200_000)
    ap.add_argument("--batch", type=int, default=512)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import corpus as corpus_mod
    import mgx_loader
    mgx = mgx_loader.load()
    sharded = __import__("importlib").import_module("mygram_db_b200.sharded")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    seed = 0xC2 if args.config == "c2" else 0xC4
    kw = dict(alphabet=600, min_len=8, max_len=60)
    lo, hi = sharded.shard_range(args.docs, world, rank)
    c = corpus_mod.generate("cjk", hi - lo, seed, first_doc=lo, **kw)
    index = mgx.Index(2, 0, True, device=local)
    index.build(c.doc_ids, c.arena, c.offsets)
    scored = args.config == "c2"
    if not scored:
        index.set_filter_column_arrays(0, 8, (np.arange(lo, hi, dtype=np.uint64) % 3) + 1)
    st = index.stats()
    g = torch.tensor([st.doc_count, st.total_doc_length], dtype=torch.int64, device=device)
    dist.all_reduce(g)
    params = index.params(score=scored, limit=100, offset=0, total_docs=int(g[0]), total_doc_length=int(g[1]))
    comm = sharded.ShardComm(mgx, dist, device, n_lanes=2)
    pipe = sharded.ShardPipeline(mgx, index, params, 100, comm)
    lay = sharded.record_layout(args.batch, 100)
    streams = [torch.cuda.Stream(device=device) for _ in range(2)]

    one = None
    if rank == 0:
        full = corpus_mod.generate("cjk", args.docs, seed, **kw)
        one = mgx.Index(2, 0, True, device=local)
        one.build(full.doc_ids, full.arena, full.offsets)
        if not scored:
            one.set_filter_column_arrays(0, 8, (np.arange(args.docs, dtype=np.uint64) % 3) + 1)

    def make_batch(i):
        base = corpus_mod.sample_queries_global("cjk", seed, args.docs, args.batch, 500 + i, n_terms=3, min_cp=2,
                                                max_cp=3, **kw)
        if scored:
            return base, None, None
        queries, programs, filters = [], [], []
        for j, (a, b, cc) in enumerate(base):
            kind = j % 4
            if kind == 0:
                queries.append([a, b]); programs.append(([0, 0, 1], [0, 1, 2])); filters.append([])
            elif kind == 1:
                queries.append([a, b]); programs.append(([0, 0, 2], [0, 1, 2])); filters.append([])
            elif kind == 2:
                queries.append([a, b]); programs.append(([0, 0, 3, 1], [0, 1, 0, 2])); filters.append([])
            else:
                queries.append([a, b, cc]); programs.append(([0, 0, 2, 0, 1], [0, 1, 2, 2, 2])); filters.append([(0, 0, "1")])
        return queries, programs, filters

    failures = 0
    for round_ in range(3):
        if round_ == 2 and rank == 1:
            os.environ["MGX_STREAM_TILE_CAP"] = "2"  # only this shard overflows: every rank must repeat the batch
        batches = [make_batch(round_ * 4 + i) for i in range(4)]
        prepared, outs, keep = [], [], []
        for i, (qs, programs, filters) in enumerate(batches):
            arena, offs, qbeg, _ = mgx.flatten_queries(qs)
            ext, k = mgx.Index.build_ext(programs, filters)
            keep.append((arena, offs, qbeg, ext, k))
            prepared.append(pipe.prepare(arena, offs, qbeg, args.batch, streams[i % 2], ext=ext))
            outs.append(torch.empty(lay["bytes"], dtype=torch.uint8, pin_memory=True))
        for i, p in enumerate(prepared):
            pipe.enqueue(p, i % 2, outs[i])
        for i, p in enumerate(prepared):
            pipe.finish(p)
            pipe.release(p)
        if round_ == 2:
            os.environ.pop("MGX_STREAM_TILE_CAP", None)
            rep = torch.tensor([pipe.repeats], dtype=torch.int64, device=device)
            dist.all_reduce(rep, op=dist.ReduceOp.MIN)
            if int(rep[0]) < 1:
                print(f"rank {rank}: the forced overflow did not repeat the batch on every rank", flush=True)
                failures += 1
        if rank == 0:
            for i, (qs, programs, filters) in enumerate(batches):
                ids, scores, count, total = [t.numpy() for t in sharded.record_views(outs[i], args.batch, 100)]
                w = one.query_batch(qs, programs=programs, filters=filters, score=scored, limit=100)
                valid = np.arange(100)[None, :] < w.count[:, None]
                ok = (np.array_equal(count.view(np.uint32), w.count) and np.array_equal(total.view(np.uint64), w.total)
                      and np.array_equal(ids.view(np.uint32)[valid], w.ids[valid]) and
                      (not scored or np.array_equal(scores[valid].view(np.uint64), w.scores[valid].view(np.uint64))))
                if not ok:
                    print(f"round {round_} batch {i}: sharded answer differs from the single-shard answer", flush=True)
                    failures += 1
                    if failures <= 3:
                        cu, tu, iu = count.view(np.uint32), total.view(np.uint64), ids.view(np.uint32)
                        for q in range(args.batch):
                            n = int(w.count[q])
                            bad = []
                            if int(cu[q]) != n:
                                bad.append(f"count {int(cu[q])} vs {n}")
                            if int(tu[q]) != int(w.total[q]):
                                bad.append(f"total {int(tu[q])} vs {int(w.total[q])}")
                            m = min(n, int(cu[q]))
                            if not np.array_equal(iu[q, :m], w.ids[q, :m]):
                                k = int(np.nonzero(iu[q, :m] != w.ids[q, :m])[0][0])
                                bad.append(f"ids differ from rank {k}: {iu[q, k:k + 4]} vs {w.ids[q, k:k + 4]}; scores "
                                           f"{scores[q, k:k + 4]} vs {w.scores[q, k:k + 4]}")
                            elif scored and not np.array_equal(scores[q, :m].view(np.uint64), w.scores[q, :m].view(np.uint64)):
                                k = int(np.nonzero(scores[q, :m] != w.scores[q, :m])[0][0])
                                bad.append(f"scores differ at rank {k}: {scores[q, k]!r} vs {w.scores[q, k]!r} "
                                           f"(rel {abs(scores[q, k] - w.scores[q, k]) / max(abs(w.scores[q, k]), 1e-300):.3e})")
                            if bad:
                                print(f"  query {q} {qs[q]}: " + "; ".join(bad), flush=True)
                                break
    if rank == 0 and scored:
        import pyoracle
        oi = pyoracle.OracleLib(pyoracle.PORT_LIB).index(2, 0, True)
        oi.build_bulk(full.doc_ids, full.arena, full.offsets, 8)
        qs = batches[0][0][:128]
        o = oi.query_batch(qs, score=True, limit=100, n_threads=8)
        ids, scores, count, total = [t.numpy() for t in sharded.record_views(outs[0], args.batch, 100)]
        for q in range(len(qs)):
            n = int(o.count[q])
            if not (int(total[q]) == int(o.total[q]) and int(count[q]) == n and
                    np.array_equal(ids.view(np.uint32)[q, :n], o.ids[q, :n]) and
                    np.allclose(scores[q, :n], o.scores[q, :n], rtol=1e-9, atol=0)):
                print(f"query {q}: sharded answer differs from the CPU oracle", flush=True)
                failures += 1
                break
    f = torch.tensor([failures], dtype=torch.int64, device=device)
    dist.all_reduce(f)
    comm.close()
    dist.destroy_process_group()
    if int(f[0]) != 0:
        sys.exit(1)
    if rank == 0:
        print("NCCL SHARD CHECK OK", flush=True)


if __name__ == "__main__":
    main()
