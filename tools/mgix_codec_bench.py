#!/usr/bin/env python
"""tools/mgix_codec_bench.py — host-only timing of the MGIX codec (no GPU needed): a synthetic CSR of 3M terms with
Pareto-sized posting lists (55M postings, twenty lists of 1.5M entries as Roaring bodies) through mgx_mgix_encode and
mgx_mgix_decode, with the CRC checked against zlib and the decoded arrays against the input.
MGX_MGIX_THREADS=1 pins the serial path."""
import sys, time, zlib, os
sys.path[:0]=[os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np, ctypes as C
import mgx_loader
m = mgx_loader.load()
rng = np.random.default_rng(1)
T = 3_000_000
sizes = np.minimum((rng.pareto(1.1, T) + 1).astype(np.int64), 2_000_000)
sizes[:20] = 1_500_000
P = int(sizes.sum())
offs = np.zeros(T+1, np.uint64); offs[1:] = np.cumsum(sizes)
gaps = rng.integers(1, 4, P, dtype=np.uint32)
posts = np.cumsum(gaps, dtype=np.uint64)
starts = posts[offs[:-1].astype(np.int64)]
posts = (posts - np.repeat(starts, sizes) + 1).astype(np.uint32)
tb = np.frombuffer(b"".join(b"%06x" % i for i in range(T)), dtype=np.uint8).copy()
to = (np.arange(T+1, dtype=np.uint64) * 6)
info = m.MgixInfo(4, 2, 2, 1, 1, 1, b"keep", T, 0, 0)
u8p,u32p,u64p = m.u8p,m.u32p,m.u64p
n = C.c_uint64(0)
out = np.zeros(8*P//2+60*T, np.uint8)
out[:] = 1
for rep in range(2):
    t0=time.perf_counter()
    rc = m.lib().mgx_mgix_encode(C.byref(info), tb.ctypes.data_as(u8p), to.ctypes.data_as(u64p), offs.ctypes.data_as(u64p), posts.ctypes.data_as(u32p), 0.0, out.ctypes.data_as(u8p), out.size, C.byref(n))
    dt=time.perf_counter()-t0
print("threads", os.environ.get("MGX_MGIX_THREADS","auto"), "terms", T, "postings", P, "encode rc", rc, round(n.value/1e6,1), "MB in", round(dt,3), "s ->", round(n.value/dt/1e9,2), "GB/s")
s = out[:n.value]
print("crc ok", zlib.crc32(s[:-4]) == int.from_bytes(s[-4:].tobytes(),'little'))
i2 = m.MgixInfo()
t0=time.perf_counter()
rc = m.lib().mgx_mgix_decode(s.ctypes.data_as(u8p), s.size, C.byref(i2), None,None,None,None)
print("decode(validate) rc", rc, round(time.perf_counter()-t0,3), "s", i2.n_terms, i2.n_postings)
tb2 = np.zeros(i2.term_bytes, np.uint8); to2=np.zeros(T+1,np.uint64); po2=np.zeros(T+1,np.uint64); pp2=np.zeros(P,np.uint32)
t0=time.perf_counter()
rc = m.lib().mgx_mgix_decode(s.ctypes.data_as(u8p), s.size, C.byref(i2), tb2.ctypes.data_as(u8p), to2.ctypes.data_as(u64p), po2.ctypes.data_as(u64p), pp2.ctypes.data_as(u32p))
print("decode(fill) rc", rc, round(time.perf_counter()-t0,3), "s", np.array_equal(pp2, posts), np.array_equal(po2, offs), np.array_equal(tb2, tb))
