#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native MygramDB search core.

Workload (BASELINE.json configs[1], SURVEY.md §8d "C2"): synthetic 10M-document CJK corpus
(8192 ideographs, Zipf(1.0), 16..112 code points per document), bigram index (ngram_size=2,
kanji_ngram_size=0 -> 2), batches of 4096 queries of 3 terms (2..4-code-point substrings of one random
document), AND + verified df + BM25 (k1=1.2, b=0.75) + top-100, `SORT _score DESC`, verify_text off.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  (CPU arm: the oracle / reference sources on host cores)

A "step" is one batch of 4096 queries through the whole hot path. Metric: queries/s.
  value  = device-resident throughput: the compiled batches are already in HBM; the timed region holds the
           planning kernels, df, (df all-reduce), intersect+score, top-k, (all-gather + merge); CUDA events on the
           launch stream, max over ranks. Every step uses a DIFFERENT batch and the index (postings + text, several
           GB) is far larger than the 126 MB L2, so no step is served from cache.
  e2e    = the same steps through the public call with HOST buffers: host query compile + H2D of the batch +
           all device work + D2H of ids/scores/counts, wall clock, max over ranks.
  N > 1  = STRONG scaling: the same 10M-document corpus is sharded by doc-id range over the ranks.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "support"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))  # only the CPU legs (cpu_baseline / --impl reference) import it

METRIC = "batched queries/sec (AND+BM25 top-k)"
UNIT = "queries/s"
K1, B = 1.2, 0.75
TOPK = 100
CORPUS_KIND, CORPUS_SEED = "cjk", 0xC2


def workload_name(args):
    return (f"C2: synthetic {args.docs}-doc CJK corpus (seed 0xC2, 8192 ideographs Zipf 1.0, 16-112 cp/doc), "
            f"ngram_size=2, batches of {args.batch} x 3-term AND + BM25 top-{TOPK}")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed regions, through NVML in a background thread
    (`nvidia-smi -lms` was measured to stall the CUDA context for 50-200 ms per query on this box; an NVML
    handle held open does not). Falls back to one-second nvidia-smi polling if pynvml is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
               ("hw_thermal_slowdown", 0x40))

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.proc = None
        self.lines = []

    def _nvml_loop(self, nv, h):
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def start(self):
        if os.environ.get("BENCH_NO_CLOCKS"):
            return
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.gpu]) if visible and visible.split(",")[self.gpu].isdigit() else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            q = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "1000", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                self.sm.append(float(f[1]))
                self.mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def stop(self):
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml" if self.proc is None else "nvidia-smi"}


def flatten(queries):
    flat, begin = [], [0]
    for q in queries:
        flat += q
        begin.append(len(flat))
    offs = np.zeros(len(flat) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(t) for t in flat], dtype=np.uint64)
    arena = np.frombuffer(b"".join(flat), dtype=np.uint8).copy()
    return arena, offs, np.asarray(begin, dtype=np.uint64)


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_index(args, doc_ids, arena, offsets, kind):
    import pyoracle
    path = pyoracle.REF_LIB if kind == "reference" else pyoracle.PORT_LIB
    lib = pyoracle.OracleLib(path)
    idx = lib.index(2, 0, True)
    t0 = time.perf_counter()
    idx.build_bulk(doc_ids, arena, offsets, os.cpu_count() or 1)
    return idx, time.perf_counter() - t0


def choose_cpu_kind(args):
    import pyoracle
    if args.ref_kind != "auto":
        return args.ref_kind
    # The reference's only build path is single-threaded (initial_loader.cpp:296-385, ~minutes per million
    # documents with the hash-map index), so above 250k documents the CPU arm uses the port, whose multi-threaded
    # bulk builder produces the identical index (tests/test_oracle_bulk.py) and whose query path restates the
    # reference's (oracle/oracle.cpp).
    if os.path.exists(pyoracle.REF_LIB) and args.docs <= 250_000:
        return "reference"
    return "port"


def cpu_run_queries(idx, queries, n_threads, budget_s):
    """Times a bounded sample: grows the sample until it costs about `budget_s` seconds (or the batch ends)."""
    n = min(len(queries), 64)
    t0 = time.perf_counter()
    res = idx.query_batch(queries[:n], score=True, descending=True, limit=TOPK, k1=K1, b=B, n_threads=n_threads)
    dt = time.perf_counter() - t0
    if dt < budget_s * 0.5 and n < len(queries):
        n2 = int(min(len(queries), max(n, n * budget_s / max(dt, 1e-3))))
        t0 = time.perf_counter()
        res = idx.query_batch(queries[:n2], score=True, descending=True, limit=TOPK, k1=K1, b=B, n_threads=n_threads)
        dt = time.perf_counter() - t0
        n = n2
    return n, dt, res


def run_reference_arm(args, rank):
    if rank != 0:
        return
    import corpus as corpus_mod
    kind = choose_cpu_kind(args)
    cores = os.cpu_count() or 1
    c = corpus_mod.generate(CORPUS_KIND, args.docs, CORPUS_SEED)
    idx, build_s = cpu_index(args, c.doc_ids, c.arena, c.offsets, kind)
    sample = max(32, min(args.batch, args.cpu_sample))
    times, done = [], 0
    for step in range(args.warmup + args.steps):
        qs = corpus_mod.sample_queries_global(CORPUS_KIND, CORPUS_SEED, args.docs, sample, 1000 + step)
        t0 = time.perf_counter()
        idx.query_batch(qs, score=True, descending=True, limit=TOPK, k1=K1, b=B, n_threads=cores)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
            done += sample
    total = sum(times)
    value = done / total
    label = ("reference sources + Roaring/abseil/spdlog shims (oracle/_ref)" if kind == "reference"
             else "oracle port of the reference path (oracle/oracle.cpp)")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32 doc ids / f64 BM25", "data": "synthetic",
        "config": {"workload": workload_name(args), "step": f"bounded sample of {sample} queries of the batch",
                   "cpu_index_build_s": round(build_s, 2)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} queries/step x {args.steps} steps, one query per thread, {label}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mgx", choices=["mgx", "reference"])
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--cpu-sample", type=int, default=256, help="queries per step of the CPU arm")
    ap.add_argument("--cpu-budget-s", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "port", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dense-threshold", type=float, default=0.0,
                    help="posting density from which a list also gets a bitmap (0 = library default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    import corpus as corpus_mod
    import mgx_loader
    mgx = mgx_loader.load()
    sharded = __import__("importlib").import_module("mygram_db_b200.sharded")

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        comm = sharded.TorchDist(dist)
    else:
        comm = sharded.NoDist()
    L = mgx.lib()

    # ---- this rank's shard of the corpus, generated straight into pinned host memory
    lo, hi = sharded.shard_range(args.docs, world, rank)
    n_local = hi - lo
    pinned = {}

    def alloc_pinned(nbytes):
        pinned["arena"] = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        return pinned["arena"].numpy()

    t0 = time.perf_counter()
    offsets_t = torch.empty(n_local + 1, dtype=torch.int64, pin_memory=True)
    c = corpus_mod.generate(CORPUS_KIND, n_local, CORPUS_SEED, first_doc=lo, alloc=alloc_pinned,
                            offsets_out=offsets_t.numpy().view(np.uint64))
    ids_t = torch.from_numpy(c.doc_ids.astype(np.int64).astype(np.uint32).view(np.int32)).pin_memory()
    gen_s = time.perf_counter() - t0

    # ---- index build: e2e (pinned host -> queryable device index) and device-resident
    index = mgx.Index(2, 0, True, device=local_rank, dense_threshold=args.dense_threshold)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    index.build(c.doc_ids, c.arena, c.offsets)
    build_e2e_s = time.perf_counter() - t0  # first build of the process: includes every device allocation
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    index.build(c.doc_ids, c.arena, c.offsets)  # rebuild from the same pinned host buffers (allocations are kept)
    build_e2e_warm_s = time.perf_counter() - t0
    st = index.stats()
    d_text = pinned["arena"][:int(c.offsets[-1])].to(device)
    d_off = offsets_t.to(device)
    d_ids = ids_t.to(device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    index.build_device(C.c_void_p(d_ids.data_ptr()), C.c_void_p(d_text.data_ptr()), C.c_void_p(d_off.data_ptr()),
                       n_local)
    build_dev_s = time.perf_counter() - t0
    st = index.stats()
    del d_text, d_off, d_ids
    build_algo_bytes = st.text_bytes + 4 * st.n_postings + 8 * st.n_terms  # SURVEY §8(d) B_build

    # ---- global corpus statistics (exchange 1, once per index generation)
    gstats = torch.tensor([st.doc_count, st.total_doc_length], dtype=torch.int64, device=device)
    if world > 1:
        comm.all_reduce_sum(gstats)
    total_docs, total_len = int(gstats[0]), int(gstats[1])
    params = index.params(score=True, descending=True, limit=TOPK, offset=0, k1=K1, b=B, total_docs=total_docs,
                          total_doc_length=total_len)
    backend = sharded.MgxShardBackend(mgx, index, params, TOPK, device)

    # ---- query batches: a different one per step, identical on every rank
    n_steps = args.warmup + args.steps
    batches = []
    for step in range(n_steps):
        qs = corpus_mod.sample_queries_global(CORPUS_KIND, CORPUS_SEED, args.docs, args.batch, 1000 + step)
        arena, offs, qbeg = flatten(qs)
        pa = torch.from_numpy(arena).pin_memory()
        po = torch.from_numpy(offs.view(np.int64)).pin_memory()
        pq = torch.from_numpy(qbeg.view(np.int64)).pin_memory()
        batches.append((qs, pa.numpy(), po.numpy().view(np.uint64), pq.numpy().view(np.uint64), (pa, po, pq)))

    clocks = ClockSampler(local_rank)

    def barrier():
        torch.cuda.synchronize()
        comm.barrier()

    # ---- value: compiled batches resident in HBM, CUDA events on the launch stream. Consecutive batches alternate
    # between two CUDA streams (as a server with two batches in flight would run them): the planning stage of a batch,
    # which ends in a small host read-back, overlaps the kernels of its predecessor instead of leaving the device idle.
    # The timed region is bracketed on the default stream: both streams start after ev0 and ev1 waits for both.
    # Only without collectives (N = 1): with NCCL kernels queued behind another batch's grid-filling kernels the two
    # ranks wait for each other (measured at N = 2: 5x slower), so sharded runs keep one batch in flight.
    in_flight = 2 if world == 1 else 1
    value_streams = [torch.cuda.Stream(device=device) for _ in range(in_flight)] if in_flight > 1 else \
        [torch.cuda.current_stream()]
    prepared = [backend.prepare(b[1], b[2], b[3], args.batch, stream=value_streams[i % in_flight])
                for i, b in enumerate(batches)]

    def run_value_step(i):
        with torch.cuda.stream(value_streams[i % in_flight]):
            return sharded.run_sharded_batch(backend, comm, prepared[i])

    for i in range(args.warmup):
        run_value_step(i)
    barrier()
    clocks.start()
    launches0 = L.mgx_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    if in_flight > 1:
        for st_ in value_streams:
            st_.wait_event(ev0)
    results = []
    for i in range(args.warmup, n_steps):
        results.append(run_value_step(i))
    if in_flight > 1:
        for st_ in value_streams:
            torch.cuda.current_stream().wait_stream(st_)
    ev1.record()
    barrier()
    gpu_launches = int(L.mgx_kernel_launch_count() - launches0)
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms[0])
    value = args.steps * args.batch / (ms_total / 1e3)
    backend.collect_stats = True
    for i, p in enumerate(prepared):
        if i < args.warmup:
            backend.collect_stats = False
        else:
            backend.collect_stats = True
        backend.release(p)
    kstats = backend.stats

    # ---- e2e: host buffers in, host buffers out, every step. The loop is software-pipelined the way a server feeds
    # a stream of batches: a compile thread prepares batch i+1 (host query compile + H2D of the compiled batch)
    # while the main thread enqueues batch i and waits for batch i-1 (the planning stage of a batch still waits for
    # its predecessor on the same stream). Every step copies its own inputs host->device and its own results
    # device->host; a step is complete when its results are in pinned host memory. Batches alternate between two
    # CUDA streams so that consecutive batches may overlap on the device.
    backend.collect_stats = False
    rec_bytes = sharded.record_layout(args.batch, TOPK)["bytes"]  # ids, scores, count, total of a batch in ONE buffer
    outs = [dict(rec=torch.empty(rec_bytes, dtype=torch.uint8, pin_memory=True), done=torch.cuda.Event())
            for _ in range(2)]
    e2e_parts = {"host_prepare_ms": 0.0, "enqueue_ms": 0.0, "wait_ms": 0.0}

    # two CUDA streams, alternating per batch: the planning stage of batch i+1 (which ends in a small host
    # read-back) overlaps the search kernels of batch i instead of leaving the device idle
    e2e_streams = [torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)]

    def e2e_prepare(b, slot):
        t_a = time.perf_counter()
        p = backend.prepare(b[1], b[2], b[3], args.batch, stream=e2e_streams[slot])  # compile + staging + H2D enqueue
        return p, 1e3 * (time.perf_counter() - t_a)

    def e2e_enqueue(p, slot, acc=None):
        t_b = time.perf_counter()
        with torch.cuda.stream(e2e_streams[slot]):
            sharded.run_sharded_batch(backend, comm, p)
            o = outs[slot]
            o["rec"].copy_(backend.merged_record, non_blocking=True)  # the merged answer: one device-to-host copy
            o["done"].record()
        if acc is not None:
            acc["enqueue_ms"] += 1e3 * (time.perf_counter() - t_b)  # plan (one size read-back), df, search, merge, D2H
        return p, slot

    def e2e_finish(pending, acc=None):
        t_a = time.perf_counter()
        p, slot = pending
        outs[slot]["done"].synchronize()
        backend.release(p)
        if acc is not None:
            acc["wait_ms"] += 1e3 * (time.perf_counter() - t_a)

    from concurrent.futures import ThreadPoolExecutor
    compiler = ThreadPoolExecutor(max_workers=1)  # ctypes releases the GIL: the compile of batch i+1 runs beside batch i

    def e2e_run(first, last, acc=None):
        if first >= last:
            return
        pending = None
        fut = compiler.submit(e2e_prepare, batches[first], first & 1)
        for i in range(first, last):
            p, prep_ms = fut.result()
            if acc is not None:
                acc["host_prepare_ms"] += prep_ms
            if i + 1 < last:
                fut = compiler.submit(e2e_prepare, batches[i + 1], (i + 1) & 1)
            cur = e2e_enqueue(p, i & 1, acc)
            if pending is not None:
                e2e_finish(pending, acc)
            pending = cur
        e2e_finish(pending, acc)

    e2e_run(0, min(args.warmup, 2))
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.warmup, n_steps, e2e_parts)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    clock_info = clocks.stop()
    e2e_value = args.steps * args.batch / float(e2e_s[0])
    h2d = int(batches[0][1].nbytes + batches[0][2].nbytes + batches[0][3].nbytes)
    d2h = int(outs[0]["rec"].numel())

    # ---- roofline of the dominant kernel (times: CUDA events around the kernel launches on the launch stream)
    peak, peak_src = measured_peaks()
    agg = {k: sum(s[k] for s in kstats) for k in kstats[0]} if kstats else {}
    kernels = {}
    if agg:
        kernels = {
            "and_tile_kernel": {"ms": agg["ms_and_kernel"], "bytes": agg["algo_bytes_intersect"] + agg["algo_bytes_score"],
                                "launches": len(kstats)},
            "df_tile_kernel": {"ms": agg["ms_df_kernel"], "bytes": agg["algo_bytes_df"] + agg["algo_bytes_df_lists"],
                               "launches": len(kstats)},
            "topk_kernel": {"ms": agg["ms_topk_kernel"], "bytes": 12 * agg["result_docs"], "launches": len(kstats)},
            "plan (lookup/term_plan/query_plan/scans)": {"ms": agg["ms_plan"], "bytes": 0, "launches": len(kstats)},
        }
    roofline = None
    if kernels:
        name = max(("and_tile_kernel", "df_tile_kernel", "topk_kernel"), key=lambda k: kernels[k]["ms"])
        kk = kernels[name]
        achieved = (kk["bytes"] / 1e9) / (kk["ms"] / 1e3) if kk["ms"] > 0 else 0.0
        # dram__bytes_read + dram__bytes_write of that kernel per launch, from the committed `ncu --set full` capture of
        # this command (profiles/r01_dram_traffic.json); null if the dominant kernel has no capture
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_dram_traffic.json")))
            if name in tj and args.docs == 10_000_000 and world == 1:
                traffic, traffic_src = tj[name]["dram_bytes_per_launch"], tj[name]["source"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kk["bytes"] / max(1, kk["launches"]),
                    "avg_launch_ms": kk["ms"] / max(1, kk["launches"]),
                    # share of the step's device time (with two batches in flight the event-timed durations of
                    # the SMALL kernels include queueing behind the other batch, so they are not summed here)
                    "step_share": (kk["ms"] / max(1, kk["launches"])) / max(1e-9, ms_total / max(1, args.steps))}

    # ---- CPU baseline beside it (rank 0, N = 1): oracle on host cores, bounded sample, plus a parity check
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind = choose_cpu_kind(args)
        idx, cpu_build_s = cpu_index(args, c.doc_ids, c.arena, c.offsets, kind)
        cores = os.cpu_count() or 1
        n, dt, res = cpu_run_queries(idx, batches[args.warmup][0], cores, args.cpu_budget_s)
        cpu_baseline = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"first {n} queries of one {args.batch}-query batch, one query per thread; "
                                  f"CPU index built in {cpu_build_s:.1f} s (not timed)"}
        g_ids, g_scores, g_count, g_total = [t.cpu().numpy() for t in results[0]]
        ok = bool(np.array_equal(g_total[:n].astype(np.uint64), res.total[:n]) and
                  np.array_equal(g_count[:n].astype(np.uint32), res.count[:n]))
        max_rel = 0.0
        for q in range(n):
            k = int(res.count[q])
            ok = ok and sorted(g_ids[q, :k].view(np.uint32).tolist()) == sorted(res.ids[q, :k].tolist())
            if k:
                max_rel = max(max_rel, float(np.max(np.abs(g_scores[q, :k] - res.scores[q, :k]) /
                                                    np.maximum(np.abs(res.scores[q, :k]), 1e-300))))
        parity = {"queries_checked": n, "doc_ids_and_totals_equal": ok, "max_rel_score_err": max_rel}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32 doc ids / f64 BM25", "data": "synthetic",
            "config": {"workload": workload_name(args), "docs_per_gpu": n_local, "sharding": f"doc-id range x{world}", "batches_in_flight": in_flight,
                       "cache_note": "a different query batch every step; index (%.1f GB resident) >> 126 MB L2" %
                                     (st.device_bytes / 1e9),
                       "terms": int(st.n_terms), "postings": int(st.n_postings), "dense_terms": int(st.n_dense_terms)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pipeline_depth": 2, "per_step_ms": {k: v / max(1, args.steps) for k, v in e2e_parts.items()}},
            "gpu_launches": gpu_launches, "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity": parity, "kernels": kernels,
            "batch_stats_per_step": {k: (v / max(1, len(kstats))) for k, v in agg.items()} if agg else None,
            "kernel_ms_note": ("two batches in flight: the event-timed durations of the small kernels (plan, top-k) "
                               "include queueing behind the other batch's grid-filling kernels") if in_flight > 1 else None,
            "kernel_ms_by_step": {k: [round(s[k], 3) for s in kstats] for k in
                                  ("ms_plan", "ms_df_kernel", "ms_and_kernel", "ms_topk_kernel", "ms_total")} if kstats else None,
            "index_build": {"docs_per_s_e2e": n_local * world / build_e2e_s,
                            "docs_per_s_e2e_rebuild": n_local * world / build_e2e_warm_s, "docs_per_s_device": n_local * world / build_dev_s,
                            "device_build_ms": st.last_build_ms, "algorithmic_bytes": int(build_algo_bytes),
                            "hbm_frac_device": (build_algo_bytes / 1e9) / max(1e-9, st.last_build_ms / 1e3) / peak,
                            "corpus_gen_s": round(gen_s, 2)},
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
