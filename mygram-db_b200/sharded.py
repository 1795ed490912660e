"""sharded.py — doc-ID-range sharding of the search core across the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch). The reference has no
distributed layer; this is the B200-native addition of SURVEY.md §8(e):

  * shard r owns the documents of global index range [r*ceil(N/G), (r+1)*ceil(N/G)) and builds its own
    dictionary / CSR / text arena from them -- the build needs no exchange;
  * every rank sees the whole query batch;
  * exchange 1 (tiny): all-reduce(SUM) of the per-term verified document frequencies and, once per index
    generation, of the corpus statistics (doc_count, total_doc_length) -- BM25 must use GLOBAL statistics so
    that scores are identical for every shard count;
  * exchange 2: ONE all-gather of the fixed-size per-shard top-k record (ids u32, scores f64, counts, totals packed
    in one buffer, `record_layout`), followed by the rank-based merge kernel (mgx_merge_topk_packed_device).

The protocol is written against a small backend interface so that the same code drives the CUDA library
(`MgxShardBackend`) and, in the CPU test-suite, an oracle-backed stand-in over the gloo backend.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_range(n_docs_total, world_size, rank):
    per = -(-n_docs_total // world_size)
    lo = min(n_docs_total, rank * per)
    hi = min(n_docs_total, lo + per)
    return lo, hi


class NoDist:
    """Single-process stand-in for torch.distributed."""
    world_size = 1
    rank = 0

    def all_reduce_sum(self, t):
        return t

    def all_gather(self, t):
        return t.unsqueeze(0) if hasattr(t, "unsqueeze") else t[None]

    def barrier(self):
        pass


class TorchDist:
    """torch.distributed wrapper (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, dist):
        self.dist = dist
        self.world_size = dist.get_world_size()
        self.rank = dist.get_rank()

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t

    def all_gather(self, t):
        import torch
        out = torch.empty((self.world_size,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1))
        return out

    def barrier(self):
        self.dist.barrier()


def gather_doc_id_sets(comm, local_ids, device=None):
    """Un-scored result sets of one query across the shards (the Index::Search* style calls and the fuzzy / synonym
    paths, which are per-document predicates and need no global statistics): every shard answers over its own doc-id
    range, and because the ranges are ordered, the concatenation of the shard answers in rank order IS the ascending
    global answer — no merge. Exchange: the sizes (one small all-gather), then the ids padded to the largest shard
    answer (one all-gather). `local_ids`: this shard's ascending global doc ids (numpy uint32). Returns the global
    answer (numpy uint32) on every rank."""
    import torch
    local = torch.from_numpy(np.ascontiguousarray(local_ids, dtype=np.uint32).view(np.int32).copy())
    if device is not None:
        local = local.to(device)
    n_local = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = comm.all_gather(n_local).view(-1).cpu()
    width = int(sizes.max()) if sizes.numel() else 0
    if width == 0:
        return np.zeros(0, dtype=np.uint32)
    padded = torch.zeros(width, dtype=torch.int32, device=local.device)
    padded[:local.numel()] = local
    everyone = comm.all_gather(padded).cpu().numpy().view(np.uint32)
    parts = [everyone[r, :int(sizes[r])] for r in range(everyone.shape[0])]
    out = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint32)
    assert out.size < 2 or bool(np.all(out[1:] > out[:-1])), "shard answers must come from ordered doc-id ranges"
    return out


def record_layout(n_queries, stride):
    """Byte offsets of one shard's packed top-k record: [scores f64 Q*S][total i64 Q][ids i32 Q*S][count i32 Q],
    padded to 16 bytes, then a 16-byte status block (word 0: the shard's streamed batch overflowed its workspace) --
    the same layout as mgx_shard_record_layout, include/mgx.h."""
    o_scores = 0
    o_total = n_queries * stride * 8
    o_ids = o_total + n_queries * 8
    o_count = o_ids + n_queries * stride * 4
    o_status = (o_count + n_queries * 4 + 15) & ~15
    return {"scores": o_scores, "total": o_total, "ids": o_ids, "count": o_count, "status": o_status,
            "bytes": o_status + 16}


def record_views(rec, n_queries, stride):
    """(ids, scores, count, total) views into packed records: `rec` is a uint8 tensor [..., bytes] (one record or
    the gathered [G, bytes]); the views keep the leading dimensions."""
    import torch
    lay = record_layout(n_queries, stride)
    lead = tuple(rec.shape[:-1])
    Q, S = n_queries, stride

    def part(off, n, dtype, shape):
        return rec[..., off:off + n].view(dtype).reshape(lead + shape)

    return (part(lay["ids"], Q * S * 4, torch.int32, (Q, S)), part(lay["scores"], Q * S * 8, torch.float64, (Q, S)),
            part(lay["count"], Q * 4, torch.int32, (Q,)), part(lay["total"], Q * 8, torch.int64, (Q,)))


def run_sharded_batch(backend, comm, batch):
    """One query batch over all shards. `backend` implements:
        local_df(batch)            -> int64 tensor [n_term_slots]   (this shard's verified df)
        search(batch, global_df)   -> uint8 tensor [record bytes]: this shard's packed top-k record (record_layout)
        merge(records [G, bytes])  -> (ids [Q,S] int32-as-uint32, scores [Q,S] f64, count [Q] int32, total [Q] int64)
    Exchanges: one all-reduce of the per-term df, ONE all-gather of the packed records.
    Returns the merged (ids, scores, count, total), identical on every rank."""
    df = backend.local_df(batch)
    if comm.world_size > 1:
        df = comm.all_reduce_sum(df)
    rec = backend.search(batch, df)
    if comm.world_size == 1:
        return backend.merge(rec[None])
    return backend.merge(comm.all_gather(rec))


class MgxShardBackend:
    """CUDA backend: staged C ABI (mgx_batch_*) on the current torch stream."""

    def __init__(self, mgx, index, params, stride, device):
        import torch
        self.torch = torch
        self.mgx = mgx
        self.index = index
        self.params = params
        self.stride = stride
        self.device = device
        self.L = mgx.lib()
        self.stats = []
        self.collect_stats = False

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def prepare(self, arena, offsets, qbeg, n_queries, stream=None):
        """Host compile + H2D of one batch; returns an opaque prepared batch. `stream` (a torch.cuda.Stream; default:
        the caller's current stream) is the stream every later stage of this batch must be enqueued on."""
        h = C.c_void_p()
        m = self.mgx
        st = C.c_void_p(stream.cuda_stream) if stream is not None else self._stream()
        m._check(self.L.mgx_batch_prepare(self.index._h, C.byref(self.params), n_queries, m._ptr(arena, m.u8p),
                                          m._ptr(offsets, m.u64p), m._ptr(qbeg, m.u64p), None, None, None,
                                          st, C.byref(h)))
        return {"h": h, "n_queries": n_queries, "n_slots": int(self.L.mgx_batch_term_slots(h))}

    def release(self, batch):
        if self.collect_stats:
            s = self.mgx.BatchStats()
            self.mgx._check(self.L.mgx_batch_get_stats(batch["h"], C.byref(s)))
            self.stats.append(s.as_dict())
        self.L.mgx_batch_destroy(batch["h"])

    def local_df(self, batch):
        t = self.torch
        df = t.empty(max(1, batch["n_slots"]), dtype=t.int64, device=self.device)  # every slot is written
        self.mgx._check(self.L.mgx_batch_plan_device(batch["h"]))
        self.mgx._check(self.L.mgx_batch_df_device(batch["h"], C.c_void_p(df.data_ptr())))
        return df

    def search(self, batch, df):
        t = self.torch
        Q, S = batch["n_queries"], self.stride
        # count / total are written for every query; ids / scores are valid up to count[q] (no fill kernels)
        rec = t.empty(record_layout(Q, S)["bytes"], dtype=t.uint8, device=self.device)
        self.mgx._check(self.L.mgx_batch_search_packed_device(batch["h"], C.c_void_p(df.data_ptr()), S,
                                                              C.c_void_p(rec.data_ptr())))
        self._shape = (Q, S)
        return rec

    def merge(self, records):
        """records: uint8 [G, record bytes] (contiguous, as gathered). Returns views into ONE merged record, so that
        `merged_record` is a single device-to-host copy for the caller."""
        t = self.torch
        Q, S = self._shape
        G = records.shape[0]
        out = t.empty(records.shape[1], dtype=t.uint8, device=self.device)
        self.mgx._check(self.L.mgx_merge_topk_packed_device(
            self.device.index if self.device.index is not None else 0, self._stream(), C.byref(self.params), G, Q, S,
            C.c_void_p(records.data_ptr()), C.c_void_p(out.data_ptr())))
        self.merged_record = out
        return record_views(out, Q, S)


def merge_topk_reference(params_compute_score, descending, limit, offset, ids_all, scores_all, count_all, total_all,
                         stride):
    """Plain numpy statement of mgx_merge_topk_device (SortByScore's comparator,
    result_sorter.cpp:681-686). Used by the gloo tests and to check the merge kernel."""
    G, Q, _ = ids_all.shape
    ids = np.zeros((Q, stride), dtype=np.uint32)
    scores = np.zeros((Q, stride), dtype=np.float64)
    count = np.zeros(Q, dtype=np.uint32)
    total = total_all.sum(axis=0).astype(np.uint64)
    for q in range(Q):
        recs = []
        for g in range(G):
            c = int(count_all[g, q])
            recs += [(float(scores_all[g, q, i]), int(ids_all[g, q, i]), g, i) for i in range(c)]
        if params_compute_score:
            recs.sort(key=lambda r: (-r[0], -r[1]) if descending else (r[0], r[1]))
        else:
            recs.sort(key=lambda r: (r[2], r[3]))  # concatenate shard runs in shard order
        skip = min(len(recs), offset)
        keep = recs[skip:]
        if limit:
            keep = keep[:limit]
        keep = keep[:stride]
        count[q] = len(keep)
        for i, r in enumerate(keep):
            ids[q, i] = r[1]
            scores[q, i] = r[0]
    return ids, scores, count, total


class ShardComm:
    """The library's own NCCL communicator lanes (mgx_comm_*): torch.distributed only carries the bootstrap -- rank 0
    draws one NCCL unique id per lane and broadcasts them -- the exchanges of the data path are then issued by
    libmgx.so itself on its own highest-priority streams (mgx_sharded_batch_*). world_size 1: no communicator."""

    def __init__(self, mgx, dist, device, n_lanes=2):
        import torch
        self.mgx = mgx
        self.L = mgx.lib()
        self.h = C.c_void_p()
        self.world_size = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.n_lanes = n_lanes
        if self.world_size == 1:
            return
        ids = torch.zeros(n_lanes * 128, dtype=torch.uint8)
        if self.rank == 0:
            buf = np.zeros(n_lanes * 128, dtype=np.uint8)
            for lane in range(n_lanes):
                mgx._check(self.L.mgx_comm_unique_id(buf[lane * 128:].ctypes.data_as(mgx.u8p)))
            ids = torch.from_numpy(buf)
        ids = ids.to(device)
        dist.broadcast(ids, src=0)
        host = ids.cpu().numpy().copy()
        dev_index = device.index if device.index is not None else 0
        mgx._check(self.L.mgx_comm_create(host.ctypes.data_as(mgx.u8p), n_lanes, self.world_size, self.rank, dev_index,
                                          C.byref(self.h)))

    def handle(self):
        return self.h if self.h.value else None

    def nccl_version(self):
        v = C.c_int32()
        self.L.mgx_comm_info(self.handle(), None, None, None, C.byref(v))
        return int(v.value)

    def close(self):
        if self.h.value:
            self.L.mgx_comm_destroy(self.h)
            self.h = C.c_void_p()


class ShardPipeline:
    """Whole batches through mgx_sharded_batch_enqueue / _finish: plan, df, all-reduce, search, all-gather, merge and
    the device-to-host copy of the merged record are ONE enqueue on the batch's stream (and the lane's communicator
    stream); finish waits and, if any shard's streamed batch overflowed its workspace, repeats it collectively."""

    def __init__(self, mgx, index, params, stride, comm):
        self.mgx, self.index, self.params, self.stride, self.comm = mgx, index, params, stride, comm
        self.L = mgx.lib()
        self.stats = []
        self.repeats = 0

    def prepare(self, arena, offsets, qbeg, n_queries, stream, ext=None):
        """Host compile + H2D of one batch on `stream` (a torch.cuda.Stream); ext: mgx_query_ext_t or None."""
        h = C.c_void_p()
        m = self.mgx
        m._check(self.L.mgx_batch_prepare_ex(self.index._h, C.byref(self.params), n_queries, m._ptr(arena, m.u8p),
                                             m._ptr(offsets, m.u64p), m._ptr(qbeg, m.u64p), None, None, None,
                                             C.byref(ext) if ext is not None else None,
                                             C.c_void_p(stream.cuda_stream), C.byref(h)))
        return {"h": h, "n_queries": n_queries}

    def open_share(self, name, n_ranks, rank, n_slots=8, slot_bytes=4 << 20):
        """Attach to the node's ring of compiled batches (mgx_share_open; rank 0 creates it). Afterwards
        prepare_shared() compiles a batch on ONE rank only."""
        h = C.c_void_p()
        self.mgx._check(self.L.mgx_share_open(name.encode(), n_ranks, rank, n_slots, slot_bytes, C.byref(h)))
        self.share, self.share_ranks, self.share_rank = h, n_ranks, rank

    def close_share(self):
        if getattr(self, "share", None) is not None:
            self.L.mgx_share_close(self.share)
            self.share = None

    def prepare_shared(self, seq, arena, offsets, qbeg, n_queries, stream, ext=None, timeout_ms=60000):
        """Batch number `seq` of the node's common sequence: compiled by rank seq % n_ranks and published, imported by
        the others. A batch the channel cannot carry (column conditions, larger than a slot) is compiled locally by
        every rank -- publish and import report the same status for it."""
        m = self.mgx
        if getattr(self, "share", None) is None or self.share_ranks == 1:
            return self.prepare(arena, offsets, qbeg, n_queries, stream, ext)
        if seq % self.share_ranks == self.share_rank:
            p = self.prepare(arena, offsets, qbeg, n_queries, stream, ext)
            rc = self.L.mgx_share_publish(self.share, seq, p["h"], timeout_ms)
            if rc not in (m.MGX_OK, m.MGX_ERR_UNSUPPORTED, m.MGX_ERR_CAPACITY):
                m._check(rc)
            return p
        h = C.c_void_p()
        rc = self.L.mgx_share_import(self.share, seq, self.index._h, C.byref(self.params),
                                     C.c_void_p(stream.cuda_stream), timeout_ms, C.byref(h))
        if rc in (m.MGX_ERR_UNSUPPORTED, m.MGX_ERR_CAPACITY):
            return self.prepare(arena, offsets, qbeg, n_queries, stream, ext)
        m._check(rc)
        return {"h": h, "n_queries": n_queries}

    def rearm(self, batch):
        """Back to the uploaded state (the compiled batch is copied from its pinned staging buffer again), so the
        same batch object can be enqueued once more."""
        self.mgx._check(self.L.mgx_batch_reset(batch["h"]))

    def enqueue(self, batch, lane, host_record=None):
        """host_record: pinned uint8 torch tensor of record_layout(...)['bytes'] (or None: result stays on the device)."""
        ptr = C.c_void_p(host_record.data_ptr()) if host_record is not None else None
        batch["lane"], batch["host"] = lane, ptr
        self.mgx._check(self.L.mgx_sharded_batch_enqueue(self.comm.handle(), lane, batch["h"], self.stride, ptr))

    def finish(self, batch):
        """Returns the device address of the merged record (valid until release)."""
        d = C.c_void_p()
        rep = C.c_int32()
        self.mgx._check(self.L.mgx_sharded_batch_finish(self.comm.handle(), batch["lane"], batch["h"], self.stride,
                                                        batch["host"], C.byref(d), C.byref(rep)))
        self.repeats += int(rep.value)
        return d.value

    def release(self, batch, collect_stats=False):
        if collect_stats:
            s = self.mgx.BatchStats()
            self.mgx._check(self.L.mgx_batch_get_stats(batch["h"], C.byref(s)))
            self.stats.append(s.as_dict())
        self.L.mgx_batch_destroy(batch["h"])
