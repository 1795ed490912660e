#!/usr/bin/env python
"""bench.py — benchmarks of the B200-native MygramDB search core (BASELINE.json configs).

  python bench.py --gpus N --steps K --warmup W [--config c2|c3|c4|c5]     (N > 1: under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...                   (CPU arm: the reference path on host cores)

Configs (default c2 = BASELINE.json configs[1], the configuration the headline metric is quoted on):
  c2  10M-document CJK corpus (seed 0xC2, 8192 ideographs Zipf 1.0, 16..112 code points), bigram index, batches of
      4096 queries of 3 terms (2..4-code-point substrings of one random document): AND + verified df + BM25 + top-100.
  c3  bulk index build: tokenize + posting construction of 50M C2-style documents (seed 0xC3), doc-range sharded over
      the GPUs; metric = documents indexed per second.
  c4  mixed boolean workload over the C2-style corpus (seed 0xC4) with two filter columns: 40 % A AND B, 20 % A OR B,
      20 % A AND NOT B, 20 % (A OR B) AND C FILTER status = 1; first 100 ids + total per query.
  c5  100M documents (seed 0xC5, 8..56 code points), 65 536-query batches of 1-2 terms of 2-3 code points, top-100.

A "step" is one batch through the whole hot path (c3: one build of every shard). N > 1 is STRONG scaling: the same corpus
is sharded by doc-id range over the ranks, every rank answers every query over its shard, the library issues the two
NCCL exchanges itself (mgx_sharded_batch_*).
  value  = device-resident throughput: compiled batches already in HBM, CUDA events on the launch streams, max over
           ranks; `in_flight` batches on separate streams / communicator lanes.
  e2e    = the same steps from HOST buffers to HOST buffers (host compile + H2D + device + D2H), wall clock, max over ranks;
           at N > 1 every rank uploads the batch and rank 0 reads the merged answer back (d2h_bytes_per_step).
  parity = recorded in EVERY line: rank 0 checks the merged answer of one timed batch against the CPU oracle (ids in
           order, counts, totals, scores) and, for N > 1, against a single-shard run of the whole corpus on its own GPU
           (every query, bit for bit: the design promises identical answers for every shard count).
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
sys.path.insert(0, os.path.join(ROOT, "oracle"))  # only the CPU legs (cpu_baseline / parity / --impl reference) import it

K1, B = 1.2, 0.75
TOPK = 100

CONFIGS = {
    "c2": dict(seed=0xC2, docs=10_000_000, batch=4096, min_len=16, max_len=112, n_terms=3, min_cp=2, max_cp=4,
               metric="batched queries/sec (AND+BM25 top-k)", unit="queries/s"),
    "c3": dict(seed=0xC3, docs=50_000_000, batch=0, min_len=16, max_len=112, n_terms=3, min_cp=2, max_cp=4,
               metric="docs indexed/sec (tokenize + posting construction)", unit="docs/s"),
    "c4": dict(seed=0xC4, docs=10_000_000, batch=4096, min_len=16, max_len=112, n_terms=3, min_cp=2, max_cp=3,
               metric="batched queries/sec (mixed AND/OR/NOT + filters)", unit="queries/s"),
    "c5": dict(seed=0xC5, docs=100_000_000, batch=65536, min_len=8, max_len=56, n_terms=(1, 2), min_cp=2, max_cp=3,
               metric="batched queries/sec (AND+BM25 top-k)", unit="queries/s"),
}


def workload_name(args):
    c = CONFIGS[args.config]
    corpus = (f"synthetic {args.docs}-doc CJK corpus (seed {c['seed']:#x}, 8192 ideographs Zipf 1.0, "
              f"{c['min_len']}-{c['max_len']} cp/doc), ngram_size=2")
    if args.config == "c3":
        return f"C3: bulk index build of a {corpus}: tokenize + radix sort + segmented unique + CSR + bitmaps"
    if args.config == "c4":
        return (f"C4: {corpus} + filter columns status (int64 1..3) / category (5 strings Zipf), batches of {args.batch}: "
                "40% A AND B, 20% A OR B, 20% A AND NOT B, 20% (A OR B) AND C FILTER status = 1; first 100 ids + total")
    nt = c["n_terms"] if isinstance(c["n_terms"], int) else f"{c['n_terms'][0]}-{c['n_terms'][1]}"
    return (f"{args.config.upper()}: {corpus}, batches of {args.batch} x {nt}-term AND "
            f"({c['min_cp']}-{c['max_cp']} cp terms) + BM25 top-{TOPK}")


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


# ------------------------------------------------------------------------------------------- workloads
def gen_kw(args):
    c = CONFIGS[args.config]
    return dict(min_len=c["min_len"], max_len=c["max_len"])


def make_queries(args, step):
    """The query batch of one step, identical on every rank: (queries, programs or None, filters or None)."""
    import corpus as corpus_mod
    c = CONFIGS[args.config]
    if args.config != "c4":
        qs = corpus_mod.sample_queries_global("cjk", c["seed"], args.docs, args.batch, 1000 + step, n_terms=c["n_terms"],
                                              min_cp=c["min_cp"], max_cp=c["max_cp"], **gen_kw(args))
        return qs, None, None
    # C4: three terms per query cut from one random document (so AND shapes are non-empty), Zipf-popular by
    # construction (characters are Zipf distributed, so a random substring is a popular one more often than not)
    base = corpus_mod.sample_queries_global("cjk", c["seed"], args.docs, args.batch, 1000 + step, n_terms=3,
                                            min_cp=c["min_cp"], max_cp=c["max_cp"], **gen_kw(args))
    rng = np.random.default_rng(77 + step)
    kinds = rng.choice(4, size=args.batch, p=[0.4, 0.2, 0.2, 0.2])
    queries, programs, filters = [], [], []
    for q, kind in zip(base, kinds):
        a, b, cc = q
        if kind == 0:      # A AND B
            queries.append([a, b]); programs.append(([0, 0, 1], [0, 1, 2])); filters.append([])
        elif kind == 1:    # A OR B
            queries.append([a, b]); programs.append(([0, 0, 2], [0, 1, 2])); filters.append([])
        elif kind == 2:    # A AND NOT B
            queries.append([a, b]); programs.append(([0, 0, 3, 1], [0, 1, 0, 2])); filters.append([])
        else:              # (A OR B) AND C FILTER status = 1
            queries.append([a, b, cc]); programs.append(([0, 0, 2, 0, 1], [0, 1, 2, 2, 2])); filters.append([(0, 0, "1")])
    return queries, programs, filters


def c4_columns(n_docs, first_doc):
    """status (int64, uniform 1..3) and category (5 strings, Zipf) of documents [first_doc, first_doc + n_docs) --
    a pure function of the global document index, so shards agree."""
    idx = np.arange(first_doc, first_doc + n_docs, dtype=np.uint64)
    h = (idx * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(33)
    status = (h % np.uint64(3)) + np.uint64(1)
    u = ((idx * np.uint64(0xD6E8FEB86659FD93)) >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    w = 1.0 / np.arange(1, 6)
    cdf = np.cumsum(w / w.sum())
    category = np.searchsorted(cdf, u, side="right").clip(0, 4).astype(np.uint64)
    return status, category, [b"news", b"blog", b"wiki", b"shop", b"faq"]


# ------------------------------------------------------------------------------------------- CPU side
def cpu_index(doc_ids, arena, offsets, kind):
    import pyoracle
    path = pyoracle.REF_LIB if kind == "reference" else pyoracle.PORT_LIB
    lib = pyoracle.OracleLib(path)
    idx = lib.index(2, 0, True)
    t0 = time.perf_counter()
    idx.build_bulk(doc_ids, arena, offsets, os.cpu_count() or 1)
    return lib, idx, time.perf_counter() - t0


def cpu_answers(args, lib, idx, queries, programs, filters, columns, n_docs, first_id, n_threads):
    """The CPU restatement's answers for a list of queries. C4: QueryNode::Evaluate + ApplyFiltersWithBitmap per query
    (threads over queries; the ctypes calls release the GIL); otherwise the batched pipeline."""
    if programs is None:
        return idx.query_batch(queries, score=True, descending=True, limit=TOPK, k1=K1, b=B, n_threads=n_threads)
    from concurrent.futures import ThreadPoolExecutor

    def one(i):
        full = idx.eval_boolean(programs[i][0], programs[i][1], queries[i])
        if filters[i]:
            full = lib.apply_filters(n_docs, first_id, columns, filters[i], full)
        return full

    with ThreadPoolExecutor(max_workers=n_threads) as ex:
        sets = list(ex.map(one, range(len(queries))))
    return sets


def check_against_cpu(args, res_cpu, programs, g_ids, g_scores, g_count, g_total, n):
    """ids IN ORDER, counts, totals and scores of the first n queries."""
    ok, max_rel, bad = True, 0.0, None
    for q in range(n):
        if programs is None:
            k = int(res_cpu.count[q])
            same = (int(g_total[q]) == int(res_cpu.total[q]) and int(g_count[q]) == k and
                    np.array_equal(g_ids[q, :k], res_cpu.ids[q, :k]))
            if k:
                rel = float(np.max(np.abs(g_scores[q, :k] - res_cpu.scores[q, :k]) /
                                   np.maximum(np.abs(res_cpu.scores[q, :k]), 1e-300)))
                max_rel = max(max_rel, rel)
                same = same and rel <= 1e-5
        else:
            want = res_cpu[q]
            k = min(TOPK, want.size)
            same = int(g_total[q]) == want.size and int(g_count[q]) == k and np.array_equal(g_ids[q, :k], want[:k])
        if not same and bad is None:
            bad = q
        ok = ok and same
    return ok, max_rel, bad


def choose_cpu_kind(args):
    import pyoracle
    if args.ref_kind != "auto":
        return args.ref_kind
    # The reference's only build path is single-threaded (initial_loader.cpp:296-385, ~2 minutes per million documents
    # with the hash-map index), so above 250k documents the CPU arm runs the port, whose multi-threaded bulk builder
    # produces the identical index (tests/test_oracle_bulk.py) and whose query path restates the reference's; the
    # reference's own sources are timed beside it on a sub-corpus (calibration, below).
    if os.path.exists(pyoracle.REF_LIB) and args.docs <= 250_000:
        return "reference"
    return "port"


def calibrate_port_vs_reference(args, cores):
    """Same queries, same sub-corpus (args.calib_docs documents of the config's corpus), the reference's own sources
    (oracle/_ref, ExecuteFullPipeline per query, one query per thread) against the port: queries/s of both."""
    import corpus as corpus_mod
    import pyoracle
    if args.calib_docs <= 0 or not os.path.exists(pyoracle.REF_LIB) or args.config in ("c3", "c4"):
        return None
    c = CONFIGS[args.config]
    n = min(args.calib_docs, args.docs)
    sub = corpus_mod.generate("cjk", n, c["seed"], **gen_kw(args))
    qs = corpus_mod.sample_queries_global("cjk", c["seed"], n, max(256, 16 * cores), 4242, n_terms=c["n_terms"],
                                          min_cp=c["min_cp"], max_cp=c["max_cp"], **gen_kw(args))
    out = {"docs": n, "queries": len(qs)}
    for kind in ("reference", "port"):
        _, idx, build_s = cpu_index(sub.doc_ids, sub.arena, sub.offsets, kind)
        idx.query_batch(qs[:32], score=True, descending=True, limit=TOPK, k1=K1, b=B, n_threads=cores)
        t0 = time.perf_counter()
        idx.query_batch(qs, score=True, descending=True, limit=TOPK, k1=K1, b=B, n_threads=cores)
        out[kind + "_qps"] = len(qs) / (time.perf_counter() - t0)
        out[kind + "_build_s"] = round(build_s, 2)
        idx.close()
    out["port_over_reference"] = out["port_qps"] / out["reference_qps"]
    return out


def run_reference_arm(args, rank):
    if rank != 0:
        return
    import corpus as corpus_mod
    c = CONFIGS[args.config]
    kind = choose_cpu_kind(args)
    cores = os.cpu_count() or 1
    if args.config == "c3":
        # Index::AddDocumentBatch in 1000-document batches, the reference's single-threaded build loop
        # (initial_loader.cpp:450-512, index.cpp:76-119), on a bounded sub-sample of the corpus per step
        import pyoracle
        kind = "reference" if os.path.exists(pyoracle.REF_LIB) and args.ref_kind != "port" else "port"
        n = max(1000, args.cpu_build_docs)
        sub = corpus_mod.generate("cjk", n, c["seed"], **gen_kw(args))
        lib = pyoracle.OracleLib(pyoracle.REF_LIB if kind == "reference" else pyoracle.PORT_LIB)
        times = []
        for step in range(args.warmup + args.steps):
            idx = lib.index(2, 0, True)
            t0 = time.perf_counter()
            idx.add_batch(sub.doc_ids, sub.arena, sub.offsets, batch=1000)
            dt = time.perf_counter() - t0
            idx.close()
            if step >= args.warmup:
                times.append(dt)
        total = sum(times)
        value = n * len(times) / total
        sample = (f"Index::AddDocumentBatch over {n} documents in 1000-document batches per step (single-threaded, as the "
                  f"reference's loader), {'reference sources (oracle/_ref)' if kind == 'reference' else 'oracle port'}")
        out = base_line(args, value, total, cores=1)
        out.update({"impl": "reference", "cpu_baseline": {"value": value, "unit": c["unit"], "cores": 1, "kind": kind,
                                                          "sample": sample},
                    "e2e": {"value": value, "unit": c["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0})
        print(json.dumps(out), flush=True)
        return
    corpus = corpus_mod.generate("cjk", args.docs, c["seed"], **gen_kw(args))
    lib, idx, build_s = cpu_index(corpus.doc_ids, corpus.arena, corpus.offsets, kind)
    columns = None
    if args.config == "c4":
        import pyoracle
        status, category, names = c4_columns(args.docs, 0)
        columns = pyoracle.pack_filter_arrays(args.docs, [(8, status, None), (11, category, names)])
    # one step = a bounded sample of the batch with at least 64 queries per core, so that the step is not bound by
    # its slowest query (dynamic queue, one query per thread: the reference's only parallelism, thread_pool.cpp:33)
    sample = int(min(args.batch, max(args.cpu_sample, 64 * cores)))
    times, done = [], 0
    for step in range(args.warmup + args.steps):
        qs, programs, filters = make_queries(args, step)
        qs = qs[:sample]
        t0 = time.perf_counter()
        cpu_answers(args, lib, idx, qs, None if programs is None else programs[:sample],
                    None if filters is None else filters[:sample], columns, args.docs, 1, cores)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
            done += sample
    total = sum(times)
    value = done / total
    label = ("reference sources + Roaring/abseil/spdlog shims (oracle/_ref)" if kind == "reference"
             else "oracle port of the reference path (oracle/oracle.cpp)")
    calib = None
    try:
        calib = calibrate_port_vs_reference(args, cores) if kind == "port" else None
    except Exception as e:  # the calibration must never cost the arm its line
        calib = {"error": repr(e)}
    out = base_line(args, value, total, cores=cores)
    out.update({
        "impl": "reference",
        "cpu_baseline": {"value": value, "unit": c["unit"], "cores": cores, "kind": kind,
                         "sample": f"{sample} queries/step x {args.steps} steps (>= 64 per core), one query per thread "
                                   f"from a shared queue, {label}; CPU index built in {build_s:.1f} s (not timed)",
                         "port_vs_reference_sources": calib},
        "e2e": {"value": value, "unit": c["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    print(json.dumps(out), flush=True)


def workload_config(args):
    """`config` of BOTH arms (the driver compares them): the workload and how the timed region keeps L2 cold.
    Everything that describes one arm's run (shard sizes, batches in flight, ...) goes to `run` instead."""
    if args.config == "c3":
        note = "every pass streams the shard's (key, doc) pairs (>= 12 B x slots, GBs per shard >> 126 MB L2)"
    else:
        note = "a different query batch every step; the index (GBs per shard) is far larger than the 126 MB L2"
    return {"workload": workload_name(args), "cache_note": note}


def base_line(args, value, total_s, cores=None):
    c = CONFIGS[args.config]
    return {"metric": c["metric"], "value": value, "unit": c["unit"], "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_s / max(1, args.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 doc ids / f64 BM25" if args.config != "c3" else "u64 packed n-gram keys / u32 doc ids",
            "data": "synthetic", "config": workload_config(args)}


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mgx", choices=["mgx", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--docs", type=int, default=0, help="documents of the WHOLE corpus (0 = the config's size)")
    ap.add_argument("--batch", type=int, default=0, help="queries per batch (0 = the config's size)")
    ap.add_argument("--in-flight", type=int, default=0, help="batches in flight (0 = 6 on one GPU, 3 otherwise)")
    ap.add_argument("--min-seconds", type=float, default=1.0,
                    help="the timed K steps are repeated (same batches, in order) until the timed window is this long")
    ap.add_argument("--cpu-sample", type=int, default=256, help="minimum queries per step of the CPU arm")
    ap.add_argument("--cpu-budget-s", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--cpu-build-docs", type=int, default=50_000, help="documents per step of the c3 CPU arm")
    ap.add_argument("--calib-docs", type=int, default=250_000,
                    help="sub-corpus on which the CPU arm also times the reference's own sources (0 = off)")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "port", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-share", action="store_true",
                    help="N > 1: every rank compiles every batch (instead of one rank per batch + the shared ring)")
    ap.add_argument("--parity", default="auto", choices=["auto", "full", "gpu", "off"],
                    help="auto: CPU oracle (<= 30M docs) + single-shard GPU run (N > 1, <= 30M docs); full: both, always; "
                         "gpu: only the single-shard GPU run, at any size")
    ap.add_argument("--parity-queries", type=int, default=256)
    ap.add_argument("--dense-threshold", type=float, default=0.0,
                    help="posting density from which a list also gets a bitmap (0 = library default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    cfg = CONFIGS[args.config]
    args.docs = args.docs or cfg["docs"]
    args.batch = args.batch or cfg["batch"]

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
    L = mgx.lib()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- this rank's shard of the corpus, generated straight into pinned host memory
    lo, hi = sharded.shard_range(args.docs, world, rank)
    n_local = hi - lo
    pinned = {}

    def alloc_pinned(nbytes):
        pinned["arena"] = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        return pinned["arena"].numpy()

    t0 = time.perf_counter()
    offsets_t = torch.empty(n_local + 1, dtype=torch.int64, pin_memory=True)
    c = corpus_mod.generate("cjk", n_local, cfg["seed"], first_doc=lo, alloc=alloc_pinned,
                            offsets_out=offsets_t.numpy().view(np.uint64), **gen_kw(args))
    ids_t = torch.from_numpy(c.doc_ids.astype(np.int64).astype(np.uint32).view(np.int32)).pin_memory()
    gen_s = time.perf_counter() - t0
    peak, peak_src = measured_peaks()
    clocks = ClockSampler(local_rank)

    # ---- index build: e2e (pinned host -> queryable device index) and device-resident
    index = mgx.Index(2, 0, True, device=local_rank, dense_threshold=args.dense_threshold)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    index.build(c.doc_ids, c.arena, c.offsets)
    build_e2e_s = time.perf_counter() - t0  # first build of the process: includes every device allocation
    st = index.stats()
    d_text = pinned["arena"][:int(c.offsets[-1])].to(device)
    d_off = offsets_t.to(device)
    d_ids = ids_t.to(device)

    def build_device_once():
        index.build_device(C.c_void_p(d_ids.data_ptr()), C.c_void_p(d_text.data_ptr()), C.c_void_p(d_off.data_ptr()),
                           n_local)
        return index.stats().last_build_ms

    def build_e2e_once():
        index.build(c.doc_ids, c.arena, c.offsets)

    if args.config == "c3":
        # a step = one build of every shard from device-resident inputs (value) / from pinned host memory (e2e)
        for _ in range(args.warmup):
            build_device_once()
        barrier()
        clocks.start()
        launches0 = L.mgx_kernel_launch_count()
        dev_ms, reps = [], 0
        t_start = time.perf_counter()
        while True:
            for _ in range(args.steps):
                dev_ms.append(build_device_once())
            reps += 1
            if time.perf_counter() - t_start >= args.min_seconds or reps >= 50:
                break
        gpu_launches = int(L.mgx_kernel_launch_count() - launches0)
        ms_total = max_over_ranks(sum(dev_ms))
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            build_e2e_once()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        clock_info = clocks.stop()
        st = index.stats()
        n_steps_timed = reps * args.steps
        value = args.docs * n_steps_timed / (ms_total / 1e3)
        algo = st.text_bytes + 4 * st.n_postings + 8 * st.n_terms  # SURVEY §8(d) B_build, this shard
        build_ms = ms_total / n_steps_timed
        if rank == 0:
            out = base_line(args, value, ms_total / 1e3 / reps)
            out.update({
                "n_gpus": world,
                "run": {"docs_per_gpu": n_local, "sharding": f"doc-id range x{world}",
                        "timed_repeats": reps, "terms": int(st.n_terms), "postings": int(st.n_postings),
                        "pair_slots": int(st.n_pair_slots), "sort": os.environ.get("MGX_SORT", "one-sweep")},
                "e2e": {"value": args.docs * args.steps / e2e_s, "unit": cfg["unit"],
                        "h2d_bytes_per_step": int(c.offsets[-1]) + 8 * (n_local + 1) + 4 * n_local, "d2h_bytes_per_step": 64},
                "gpu_launches": gpu_launches, "clocks": clock_info,
                "roofline": {"bound": "hbm", "kernel": "index build (tokenize + sort + CSR, whole device pass)",
                             "achieved": (algo / 1e9) / (build_ms / 1e3), "peak": peak, "unit": "GB/s",
                             "frac": (algo / 1e9) / (build_ms / 1e3) / peak, "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": int(algo), "avg_launch_ms": build_ms,
                             "note": "B_build = text bytes + 4 B per posting + 8 B per term (SURVEY 8d); the sort passes "
                                     "are implementation traffic"},
                "cpu_baseline": None, "parity": None,
                "index_build": {"first_build_e2e_s": build_e2e_s, "corpus_gen_s": round(gen_s, 2)},
            })
            if not args.no_cpu_baseline:
                import pyoracle
                kind = "reference" if os.path.exists(pyoracle.REF_LIB) else "port"
                n = min(args.cpu_build_docs, n_local)
                libo = pyoracle.OracleLib(pyoracle.REF_LIB if kind == "reference" else pyoracle.PORT_LIB)
                oi = libo.index(2, 0, True)
                t0 = time.perf_counter()
                oi.add_batch(c.doc_ids[:n], c.arena, c.offsets[:n + 1], batch=1000)
                dt = time.perf_counter() - t0
                out["cpu_baseline"] = {"value": n / dt, "unit": cfg["unit"], "cores": 1, "kind": kind,
                                       "sample": f"Index::AddDocumentBatch over the first {n} documents of the shard in "
                                                 "1000-document batches, single-threaded as the reference's loader"}
                # parity on the sample: the device index of the same documents equals the CPU index term by term
                # (the port's export; the reference library has no CSR export -- the two are pinned to each other by
                # tests/test_oracle_bulk.py)
                gi = mgx.Index(2, 0, True, device=local_rank)
                gi.build(c.doc_ids[:n], c.arena[:int(c.offsets[n])], c.offsets[:n + 1])
                keys, goffs, gposts = gi.export()
                op = pyoracle.OracleLib(pyoracle.PORT_LIB).index(2, 0, True)
                op.build_bulk(c.doc_ids[:n], c.arena, c.offsets[:n + 1], os.cpu_count() or 1)
                oterms, ooffs, oposts = op.export()
                out["parity"] = {"docs_checked": n, "terms": int(len(oterms)),
                                 "csr_equal": bool(len(oterms) == len(keys) and [bytes(x) for x in oterms] == [bytes(x) for x in keys] and
                                                   np.array_equal(goffs, ooffs) and np.array_equal(gposts, oposts))}
                out["parity"]["ok"] = out["parity"]["csr_equal"]
                gi.close()
                oi.close()
            print(json.dumps(out), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    build_e2e_once()  # rebuild from the same pinned host buffers (allocations are kept)
    build_e2e_warm_s = time.perf_counter() - t0
    build_dev_ms = build_device_once()
    st = index.stats()
    del d_text, d_off, d_ids
    build_algo_bytes = st.text_bytes + 4 * st.n_postings + 8 * st.n_terms  # SURVEY §8(d) B_build
    columns_py = None
    if args.config == "c4":
        status, category, names = c4_columns(n_local, lo)
        index.set_filter_column_arrays(0, 8, status)
        index.set_filter_column_arrays(1, 11, category, strings=names)

    # ---- global corpus statistics (exchange 1, once per index generation)
    gstats = torch.tensor([st.doc_count, st.total_doc_length], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(gstats)
    total_docs, total_len = int(gstats[0]), int(gstats[1])
    scored = args.config != "c4"
    params = index.params(score=scored, descending=True, limit=TOPK, offset=0, k1=K1, b=B, total_docs=total_docs,
                          total_doc_length=total_len)
    # measured (gpurun_out/c2_s5o_*, c2_s5p_*): six batches in flight are 4-20 % faster than three on one GPU (the
    # fixed, latency-bound part of a batch overlaps the next batches' kernels); at N > 1 the gain of the device loop
    # (+4 %) is lost again end to end, where the compile look-ahead is shorter than the pipeline
    in_flight = args.in_flight or (6 if world == 1 else 3)
    comm = sharded.ShardComm(mgx, dist if world > 1 else None, device, n_lanes=min(4, in_flight))
    pipe = sharded.ShardPipeline(mgx, index, params, TOPK, comm)
    lay = sharded.record_layout(args.batch, TOPK)

    # ---- query batches: a different one per step, identical on every rank
    n_steps = args.warmup + args.steps
    batches = []
    for step in range(n_steps):
        qs, programs, filters = make_queries(args, step)
        arena, offs, qbeg = flatten(qs)
        pa = torch.from_numpy(arena).pin_memory()
        po = torch.from_numpy(offs.view(np.int64)).pin_memory()
        pq = torch.from_numpy(qbeg.view(np.int64)).pin_memory()
        ext, keep = mgx.Index.build_ext(programs, filters)
        batches.append(dict(qs=qs, programs=programs, filters=filters, arena=pa.numpy(),
                            offs=po.numpy().view(np.uint64), qbeg=pq.numpy().view(np.uint64), ext=ext,
                            keep=(pa, po, pq, keep)))

    streams = [torch.cuda.Stream(device=device) for _ in range(in_flight)]

    def prepare(i, slot):
        bt = batches[i]
        return pipe.prepare(bt["arena"], bt["offs"], bt["qbeg"], args.batch, streams[slot], ext=bt["ext"])

    # ---- value: compiled batches resident in HBM; `in_flight` batches on separate streams and communicator lanes.
    # Every step is ONE enqueue (plan, df, all-reduce, search, all-gather, merge) without a host synchronisation; the
    # host only waits for the batch that left the pipeline `in_flight` steps ago. When the timed window would be
    # shorter than --min-seconds the K steps are repeated: the batch objects are re-armed (their compiled form is
    # copied from pinned staging again: ~1 MB per batch, inside the timed region).
    prepared = [prepare(i, i % in_flight) for i in range(n_steps)]
    used = [False] * n_steps

    def value_enqueue(i):
        p = prepared[i]
        if used[i]:
            pipe.rearm(p)
        used[i] = True
        pipe.enqueue(p, i % in_flight % comm.n_lanes if world > 1 else 0)
        return p

    def run_pipelined(order):
        pending = []  # batch indices in flight, oldest first
        for i in order:
            while i in pending or len(pending) >= in_flight:  # a batch object is re-armed only after it has finished
                pipe.finish(prepared[pending.pop(0)])
            value_enqueue(i)
            pending.append(i)
        for j in pending:
            pipe.finish(prepared[j])

    run_pipelined(range(args.warmup))
    barrier()
    # how often the K steps have to run for the window to reach --min-seconds (estimated from one untimed pass)
    t0 = time.perf_counter()
    run_pipelined(range(args.warmup, n_steps))
    torch.cuda.synchronize()
    est = max_over_ranks(time.perf_counter() - t0)
    reps = int(max(1, min(1000, np.ceil(args.min_seconds / max(est, 1e-6)))))
    barrier()
    clocks.start()
    launches0 = L.mgx_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s_ in streams:
        s_.wait_event(ev0)
    run_pipelined([i for _ in range(reps) for i in range(args.warmup, n_steps)])
    for s_ in streams:
        torch.cuda.current_stream().wait_stream(s_)
    ev1.record()
    barrier()
    gpu_launches = int(L.mgx_kernel_launch_count() - launches0)
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    n_timed = reps * args.steps
    value = n_timed * args.batch / (ms_total / 1e3)
    repeats_value = pipe.repeats
    # the answer of the first timed batch (device record -> host) for the parity checks
    first = prepared[args.warmup]
    pipe.rearm(first)
    host_rec = torch.empty(lay["bytes"], dtype=torch.uint8, pin_memory=True)
    pipe.enqueue(first, 0, host_rec)
    pipe.finish(first)
    g_ids, g_scores, g_count, g_total = [t.numpy().copy() for t in sharded.record_views(host_rec, args.batch, TOPK)]
    g_ids = g_ids.view(np.uint32)
    for p in prepared:
        pipe.release(p)

    # ---- per-kernel times: every timed batch once more, ONE in flight, so the CUDA-event durations of the kernels
    # are not stretched by another batch sharing the device
    for i in range(args.warmup, n_steps):
        p = prepare(i, 0)
        pipe.enqueue(p, 0)
        pipe.finish(p)
        pipe.release(p, collect_stats=True)
    kstats = pipe.stats
    # B_df as SURVEY 8(d) defines it -- the text bytes of every document of SearchAnd(n-grams), the set the
    # reference's PopulateTermDocumentFrequency scans -- comes from one more batch with the signature pre-filter off
    # (MGX_DF_NO_SIG: every entry goes through the membership stage, as in round 1); the timed runs above use the
    # filter and read the text of far fewer documents (reported as text_bytes_read)
    ref_df = None
    if kstats and os.environ.get("MGX_DF_NO_SIG") is None:  # (the same decision on every rank: the batch has exchanges)
        os.environ["MGX_DF_NO_SIG"] = "1"
        try:
            n_before = len(pipe.stats)
            p = prepare(args.warmup, 0)
            pipe.enqueue(p, 0)
            pipe.finish(p)
            pipe.release(p, collect_stats=True)
            ref = pipe.stats[-1]
            ref_df = {"algo_bytes_df": ref["algo_bytes_df"], "df_candidates": ref["df_candidates"],
                      "ms_df_kernel": ref["ms_df_kernel"]}
            kstats = pipe.stats[:n_before]
        finally:
            del os.environ["MGX_DF_NO_SIG"]

    # ---- e2e: host buffers in, host buffers out, every step: compile threads prepare batches ahead (host query
    # compile + H2D of the compiled batch), the main thread enqueues them (merged record -> pinned host memory) and
    # waits for the batch that left the pipeline `in_flight` steps ago.
    from concurrent.futures import ThreadPoolExecutor
    outs = [torch.empty(lay["bytes"], dtype=torch.uint8, pin_memory=True) for _ in range(in_flight)]
    e2e_parts = {"host_prepare_ms": 0.0, "wait_for_prepared_ms": 0.0, "enqueue_ms": 0.0, "wait_ms": 0.0}
    # N > 1: the ranks of the node take turns compiling (batch number seq is compiled by rank seq % N and handed to the
    # others through the shared-memory ring of mgx_share_*), instead of every rank compiling every batch
    share_seq = [0]
    if world > 1 and not args.no_share:
        # a slot holds one compiled batch (~125 bytes per query measured on C2 / C5); the ring must fit /dev/shm
        slot_bytes = max(2 << 20, 1 << int(np.ceil(np.log2(192 * args.batch))))
        n_slots = 8
        fs = os.statvfs("/dev/shm")
        fits = [int(fs.f_bavail * fs.f_frsize >= 2 * n_slots * slot_bytes)]
        name = [f"/mgx_share_{os.environ.get('MASTER_PORT', '0')}_{os.getpid()}"]
        dist.broadcast_object_list(name, src=0)
        # rank 0 creates the ring, the others attach; if ANY rank fails, every rank falls back to compiling itself
        ok = [fits[0]]
        if rank == 0 and ok[0]:
            try:
                pipe.open_share(name[0], world, rank, n_slots=n_slots, slot_bytes=slot_bytes)
            except Exception:
                ok = [0]
        dist.broadcast_object_list(ok, src=0)
        mine = ok[0]
        if ok[0] and rank != 0:
            try:
                pipe.open_share(name[0], world, rank, n_slots=n_slots, slot_bytes=slot_bytes)
            except Exception:
                mine = 0
        flag = torch.tensor([mine], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            pipe.close_share()
            args.no_share = True
        barrier()

    # ctypes releases the GIL: batches i+1, i+2 compile beside batch i. With the shared ring a rank compiles only every
    # N-th batch and otherwise waits for another rank's publication, so more batches are prepared ahead: a worker that
    # waits for batch j must not keep this rank from starting its own batch j+1
    shared = world > 1 and not args.no_share
    ahead = min(6, world + 1) if shared else 2
    compiler = ThreadPoolExecutor(max_workers=ahead)

    def e2e_prepare(i, slot, seq):
        t_a = time.perf_counter()
        bt = batches[i]
        p = pipe.prepare_shared(seq, bt["arena"], bt["offs"], bt["qbeg"], args.batch, streams[slot], ext=bt["ext"])
        return p, 1e3 * (time.perf_counter() - t_a)

    def e2e_run(order, acc=None):
        order = list(order)
        futs, pending = {}, []
        seq0 = share_seq[0]  # the same on every rank: all ranks run the same orders
        share_seq[0] += len(order)
        for j in range(min(ahead, len(order))):
            futs[j] = compiler.submit(e2e_prepare, order[j], j % in_flight, seq0 + j)
        for j, i in enumerate(order):
            t_w = time.perf_counter()
            p, prep_ms = futs.pop(j).result()
            wait_prep = 1e3 * (time.perf_counter() - t_w)
            if j + ahead < len(order):
                futs[j + ahead] = compiler.submit(e2e_prepare, order[j + ahead], (j + ahead) % in_flight, seq0 + j + ahead)
            t_b = time.perf_counter()
            slot = j % in_flight
            # every rank uploads the step's batch; the merged answer (identical on every rank after the all-gather +
            # merge) is read back to the host by rank 0, the rank that answers the caller -- the other ranks' copies
            # stay in HBM instead of all N ranks writing the same 5 MB into host memory every step
            pipe.enqueue(p, slot % comm.n_lanes if world > 1 else 0, outs[slot] if rank == 0 else None)
            t_c = time.perf_counter()
            pending.append(p)
            if len(pending) >= in_flight:
                q = pending.pop(0)
                pipe.finish(q)
                pipe.release(q)
            if acc is not None:
                acc["host_prepare_ms"] += prep_ms
                acc["wait_for_prepared_ms"] += wait_prep
                acc["enqueue_ms"] += 1e3 * (t_c - t_b)
                acc["wait_ms"] += 1e3 * (time.perf_counter() - t_c)
        for q in pending:
            pipe.finish(q)
            pipe.release(q)

    e2e_run(range(min(args.warmup, 3)))
    barrier()
    e2e_reps = int(max(1, min(1000, np.ceil(args.min_seconds / max(est * 1.5, 1e-6)))))
    t0 = time.perf_counter()
    e2e_run([i for _ in range(e2e_reps) for i in range(args.warmup, n_steps)], e2e_parts)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clock_info = clocks.stop()
    e2e_value = e2e_reps * args.steps * args.batch / e2e_s
    h2d = int(batches[0]["arena"].nbytes + batches[0]["offs"].nbytes + batches[0]["qbeg"].nbytes)
    d2h = int(lay["bytes"])

    # ---- roofline of the dominant kernel (CUDA events around the kernel's launches, one batch in flight)
    agg = {k: sum(s[k] for s in kstats) for k in kstats[0]} if kstats else {}
    nb = max(1, len(kstats))
    kernels = {}
    roofline = None
    if agg:
        step_ms = ms_total / n_timed
        kernels = {
            "df_tile / df_units (verified df: candidate tiles)": {
                "ms": agg["ms_df_kernel"] / nb,
                "algo_bytes": ref_df["algo_bytes_df"] if ref_df else agg["algo_bytes_df"] / nb,
                "list_bytes": agg["algo_bytes_df_lists"] / nb,
                "text_bytes_read": agg["algo_bytes_df"] / nb,
                "without_signature_filter": ref_df},
            "df_stream (verified df: one pass over the text arena)": {
                "ms": agg["ms_df_stream_kernel"] / nb, "algo_bytes": agg["df_stream_bytes"] / nb},
            "and_tile (intersection + fused BM25 epilogue)": {
                "ms": agg["ms_and_kernel"] / nb, "algo_bytes": (agg["algo_bytes_intersect"] + agg["algo_bytes_score"]) / nb,
                "tiles_per_s": agg["n_and_tiles"] / max(1e-9, agg["ms_and_kernel"] / 1e3)},
            "topk (+ group pre-reduction)": {"ms": agg["ms_topk_kernel"] / nb, "algo_bytes": 12 * agg["result_docs"] / nb},
            "plan (2 launches when streamed)": {"ms": agg["ms_plan"] / nb, "algo_bytes": 0},
        }
        name = max((k for k in kernels if not k.startswith("plan")), key=lambda k: kernels[k]["ms"])
        kk = kernels[name]
        achieved = (kk["algo_bytes"] / 1e9) / (kk["ms"] / 1e3) if kk["ms"] > 0 else 0.0
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json")))
            key = name.split(" ")[0]
            if key in tj and tj[key].get("config") == args.config and tj[key].get("docs_per_gpu") == n_local:
                traffic, traffic_src = tj[key]["dram_bytes_per_launch"], tj[key]["source"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kk["algo_bytes"], "avg_launch_ms": kk["ms"],
                    "algorithmic_bytes_note": "SURVEY 8(d) per-query bytes summed over the batch on the device: for the "
                                              "df kernels B_df = text bytes of every candidate document of every scanned "
                                              "term = the documents of SearchAnd(n-grams), the set the reference's "
                                              "algorithm scans, counted by one batch with the signature pre-filter off "
                                              "(kernels[...].without_signature_filter); the timed kernel rules most of "
                                              "them out on the posting payload and reads text_bytes_read; the "
                                              "posting-list bytes of those terms are reported separately as list_bytes "
                                              "and NOT counted in `achieved`, which is therefore a rate of ALGORITHMIC "
                                              "bytes and can exceed the DRAM peak: dram_frac is the memory-system figure",
                    "step_share": kk["ms"] / max(1e-9, step_ms),
                    "dram_frac": (traffic / 1e9) / (kk["ms"] / 1e3) / peak if traffic else None}

    # ---- parity of the measured answers + the CPU baseline beside them
    cpu_baseline, parity = None, {"n_gpus": world, "checks": []}
    want_cpu = args.parity in ("auto", "full") and (args.parity == "full" or args.docs <= 30_000_000)
    want_single = world > 1 and args.parity != "off" and (args.parity in ("full", "gpu") or args.docs <= 30_000_000)
    cores = os.cpu_count() or 1
    bt = batches[args.warmup]
    if rank == 0 and (want_cpu or want_single or (world == 1 and not args.no_cpu_baseline)):
        full = c if world == 1 else corpus_mod.generate("cjk", args.docs, cfg["seed"], **gen_kw(args))
        if want_single:
            try:
                # the whole corpus as ONE shard on this GPU, same batch: the sharded answer must equal it bit for bit
                one = mgx.Index(2, 0, True, device=local_rank, dense_threshold=args.dense_threshold)
                one.build(full.doc_ids, full.arena, full.offsets)
                if args.config == "c4":
                    s_all, c_all, names = c4_columns(args.docs, 0)
                    one.set_filter_column_arrays(0, 8, s_all)
                    one.set_filter_column_arrays(1, 11, c_all, strings=names)
                p1 = one.params(score=scored, descending=True, limit=TOPK, offset=0, k1=K1, b=B)
                r1 = one.query_batch_flat(p1, args.batch, bt["arena"], bt["offs"], bt["qbeg"], ext=bt["ext"])
                valid = np.arange(TOPK)[None, :] < r1.count[:, None]
                same = (np.array_equal(r1.count, g_count.view(np.uint32)) and
                        np.array_equal(r1.total, g_total.view(np.uint64)) and np.array_equal(r1.ids[valid], g_ids[valid]) and
                        (not scored or np.array_equal(r1.scores[valid].view(np.uint64), g_scores[valid].view(np.uint64))))
                parity["checks"].append({"against": "single-shard run of the whole corpus on rank 0's GPU",
                                         "queries": args.batch, "ids_in_order_counts_totals_scores_bit_equal": bool(same)})
                one.close()
                del one
            except Exception as e:  # the measurement line must survive a failing check (reported as failed)
                parity["checks"].append({"against": "single-shard run of the whole corpus on rank 0's GPU",
                                         "queries": args.batch, "ids_in_order_counts_totals_scores_bit_equal": False,
                                         "error": repr(e)[:300]})
        if want_cpu or (world == 1 and not args.no_cpu_baseline):
            kind = choose_cpu_kind(args)
            lib, idx, cpu_build_s = cpu_index(full.doc_ids, full.arena, full.offsets, kind)
            if args.config == "c4":
                import pyoracle
                s_all, c_all, names = c4_columns(args.docs, 0)
                columns_py = pyoracle.pack_filter_arrays(args.docs, [(8, s_all, None), (11, c_all, names)])
            # a bounded sample: at least 64 queries per core and --parity-queries, at most the batch
            n = int(min(args.batch, max(args.parity_queries, 64 * cores)))
            t0 = time.perf_counter()
            res = cpu_answers(args, lib, idx, bt["qs"][:n], None if bt["programs"] is None else bt["programs"][:n],
                              None if bt["filters"] is None else bt["filters"][:n], columns_py, args.docs, 1, cores)
            dt = time.perf_counter() - t0
            if world == 1 and not args.no_cpu_baseline:
                cpu_baseline = {"value": n / dt, "unit": cfg["unit"], "cores": cores, "kind": kind,
                                "sample": f"first {n} queries of one {args.batch}-query batch (>= 64 per core), one query "
                                          f"per thread from a shared queue; CPU index built in {cpu_build_s:.1f} s (not timed)"}
            ok, max_rel, bad = check_against_cpu(args, res, bt["programs"], g_ids, g_scores, g_count.view(np.uint32),
                                                 g_total.view(np.uint64), n)
            parity["checks"].append({"against": f"CPU oracle ({kind})", "queries": n,
                                     "ids_in_order_counts_totals_equal_scores_within_1e-5": bool(ok),
                                     "max_rel_score_err": max_rel, "first_bad_query": bad})
    parity["ok"] = bool(parity["checks"]) and all(
        all(v for k, v in ch.items() if k.endswith("equal") or k.endswith("1e-5")) for ch in parity["checks"])
    if not parity["checks"]:
        parity["ok"] = None
        parity["note"] = "checks skipped for this corpus size (--parity full forces them)"

    if rank == 0:
        out = base_line(args, value, ms_total / 1e3 / reps)
        out.update({
            "n_gpus": world,
            "run": {"docs_per_gpu": n_local, "sharding": f"doc-id range x{world}",
                    "batches_in_flight": in_flight, "timed_repeats": reps, "timed_steps_total": n_timed,
                    "streamed_batches": "one enqueue per batch, no host read-back; repeated in the synchronous form "
                                        f"{repeats_value} times in the value loop (workspace overflow)",
                    "index_resident_gb": round(st.device_bytes / 1e9, 2),
                    "terms": int(st.n_terms), "postings": int(st.n_postings), "dense_terms": int(st.n_dense_terms),
                    "nccl": comm.nccl_version() if world > 1 else None},
            "e2e": {"value": e2e_value, "unit": cfg["unit"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pipeline_depth": in_flight, "timed_repeats": e2e_reps,
                    "result_readback": "rank 0 (the merged answer is identical on every rank)" if world > 1 else "rank 0",
                    "host_compile": ("every batch compiled by ONE rank of the node in turn and handed to the others "
                                     "through a shared-memory ring (mgx_share_*)") if (world > 1 and not args.no_share)
                    else "every rank compiles every batch",
                    "per_step_ms": {k: v / max(1, e2e_reps * args.steps) for k, v in e2e_parts.items()}},
            "gpu_launches": gpu_launches, "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity": parity, "kernels": kernels,
            "batch_stats_per_step": {k: (v / nb) for k, v in agg.items()} if agg else None,
            "index_build": {"docs_per_s_e2e_first": n_local * world / build_e2e_s,
                            "docs_per_s_e2e_rebuild": n_local * world / build_e2e_warm_s,
                            "docs_per_s_device": n_local * world / (build_dev_ms / 1e3), "device_build_ms": build_dev_ms,
                            "algorithmic_bytes": int(build_algo_bytes),
                            "hbm_frac_device": (build_algo_bytes / 1e9) / max(1e-9, build_dev_ms / 1e3) / peak,
                            "corpus_gen_s": round(gen_s, 2)},
        })
        print(json.dumps(out), flush=True)
    pipe.close_share()
    comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
