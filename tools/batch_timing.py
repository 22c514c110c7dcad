"""ad-hoc: where the time of srsue_gpu_batch_submit / _wait goes for one shape (20 MHz MCS 28) from pinned host memory"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import srsue_b200 as sg
from oracle import oracle as o

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
devices = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else None
lib = sg.lib()
ocell = o.make_cell(100, 1, 1)
ocfg = o.make_cfg(ocell, sf_idx=1, cfi=1, qm=6, tbs=75376)
gen = [o.gen_subframe(ocell, ocfg, 10 + i, 30.0) for i in range(8)]
cell = sg.make_cell(100, 1, 1)
cfg = sg.make_cfg(cell, sf_idx=1, cfi=1, qm=6, tbs=75376)
sf_len = len(gen[0][1])
p = lib.srsue_gpu_host_alloc(n * sf_len * 8)
q = lib.srsue_gpu_host_alloc(n * 9422)
h_iq = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n, sf_len * 2)).view(np.complex64)
h_pl = np.ctypeslib.as_array(C.cast(q, C.POINTER(C.c_uint8)), shape=(n, 9422))
for r in range(n):
    h_iq[r] = gen[r % 8][1]
ctx = None if devices else sg.Context(0)
batch = sg.Batch(ctx, n, 0.01, 0, 4, devices=devices)
order = np.random.default_rng(1).permutation(n) if os.environ.get("PERMUTE") else np.arange(n)      # scattered rows: zero-copy gather
prep = sg.Batch.prepare([dict(cell=cell, cfg=cfg, iq=h_iq[r], payload=h_pl[r]) for r in order])
for _ in range(3):
    batch.submit_prepared(prep)
    lib.srsue_gpu_batch_wait(batch.h)
ts, tw = [], []
for _ in range(5):
    t0 = time.perf_counter()
    batch.submit_prepared(prep)
    t1 = time.perf_counter()
    lib.srsue_gpu_batch_wait(batch.h)
    t2 = time.perf_counter()
    ts.append(t1 - t0); tw.append(t2 - t1)
ok = all(np.array_equal(h_pl[r], gen[r % 8][0]) for r in range(0, n, 97))   # payload row r belongs to IQ row r in both orders
tot = np.mean(ts) + np.mean(tw)
print(dict(n=n, devices=devices, submit_ms=1e3 * np.mean(ts), wait_ms=1e3 * np.mean(tw), gbit_s=n * 75376 / tot / 1e9, ok=ok, stats=batch.stats()))
