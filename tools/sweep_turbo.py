"""BASELINE configs[3]: turbo decoder sweep -- K = 40..6144, 4/6/8 fixed iterations and CRC early stop, batches of
10^3..10^6 code blocks, int16 LLRs generated as SURVEY 8d cfg4 specifies (S = 64, Eb/N0 given), device-resident input in
the decoder layout.  Prints one JSON object per line; the CPU column is the AVX2 build of the oracle on all host threads.
usage: python tools/sweep_turbo.py > gpurun_out/turbo_sweep.jsonl"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import srsue_b200 as sg
from oracle import oracle as o

ctx = sg.Context(0)
POOL = 64


def make_pool(K, ebn0):
    out = []
    for i in range(POOL):
        rng = np.random.default_rng(40000 + i)
        c = rng.integers(0, 2, K, dtype=np.uint8)
        crcv = o.crc_bits(c[:K - 24], o.CRC24B)
        c[K - 24:] = [(crcv >> (23 - b)) & 1 for b in range(24)]
        d = o.turbo_encode(c).astype(np.float64) * 2 - 1
        sigma2 = 1.0 / (2.0 * (1.0 / 3.0) * 10.0 ** (ebn0 / 10.0))
        d = d + np.random.default_rng(5_040_000 + i).standard_normal(len(d)) * np.sqrt(sigma2)
        out.append(np.clip(np.trunc(64 * d), -2048, 2047).astype(np.int16))
    return np.stack(out)


def gpu_run(K, llrs, n_cb, iters, crc, reps=5):
    W, P, elems = ctx.tdec_geometry(K)
    d_pool = torch.zeros((POOL, elems), dtype=torch.int16, device="cuda")
    ctx.tdec_import(torch.from_numpy(llrs).cuda(), POOL, K, d_pool)
    d_tcb = d_pool[torch.arange(n_cb, device="cuda") % POOL].contiguous()
    d_bits = torch.zeros((n_cb, K // 8), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(n_cb, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ctx.tdec_decode(d_tcb, n_cb, K, iters, crc, d_bits, d_st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ctx.tdec_decode(d_tcb, n_cb, K, iters, crc, d_bits, d_st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = d_st.cpu().numpy()
    avg_it = float((st & 0xFF).mean())
    bits = np.unpackbits(d_bits[:POOL].cpu().numpy(), axis=1)
    return dict(ms=ms, cb_per_s=n_cb / ms * 1e3, mbit_per_s=n_cb * K / ms / 1e3, avg_iters=avg_it,
                int16_tops=168.0 * K * avg_it * n_cb / ms / 1e9, crc_ok_frac=float(((st >> 8) & 1).mean()) if crc else None), bits, st[:POOL]


cores = os.cpu_count() or 1
if o.have_avx2():
    o.select("avx2")
for K in (40, 104, 512, 1024, 2048, 3072, 4096, 5824, 6144):
    n_cb = 100_000 if K >= 512 else 400_000
    llrs = make_pool(K, 1.5)
    for iters, crc in ((4, 0), (6, 0), (8, 0), (8, 2)):
        r, bits, st = gpu_run(K, llrs, n_cb, iters, crc)
        row = dict(K=K, n_cb=n_cb, max_iter=iters, early_stop=bool(crc), ebn0_db=1.5, **r)
        if iters in (4, 8):
            n_cpu = max(64 * cores, 256)
            sub = llrs[np.arange(n_cpu) % POOL]
            o.tdec_mt(sub[:cores], K, cores, iters, crc)
            t0 = time.perf_counter()
            cb, ci = o.tdec_mt(sub, K, cores, iters, crc)
            dt = time.perf_counter() - t0
            row["cpu_mbit_per_s"] = n_cpu * K / dt / 1e6
            row["cpu_threads"] = cores
            row["bit_exact_vs_cpu"] = bool(np.array_equal(cb[:POOL], bits[:, :K]) and np.array_equal(ci[:POOL], st & 0xFF))
        print(json.dumps(row), flush=True)
# batch-size sweep at the largest block
llrs = make_pool(6144, 1.5)
for n_cb in (1_000, 10_000, 100_000, 1_000_000):
    r, _, _ = gpu_run(6144, llrs, n_cb, 4, 0, reps=3)
    print(json.dumps(dict(K=6144, n_cb=n_cb, max_iter=4, early_stop=False, ebn0_db=1.5, **r)), flush=True)
