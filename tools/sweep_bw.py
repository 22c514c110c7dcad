"""Device-resident chain throughput per bandwidth / grant shape (one plan each, per-stage CUDA events).
usage: python tools/sweep_bw.py > gpurun_out/bw_sweep.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import srsue_b200 as sg
from oracle import oracle as o

SHAPES = [  # prb, ports, qm, tbs, tm, snr
    (6, 1, 2, 152, 1, 10.0), (6, 1, 6, 4392, 1, 30.0), (15, 1, 4, 2216, 1, 18.0), (25, 1, 6, 11448, 1, 30.0),
    (25, 1, 6, 18336, 1, 30.0), (50, 2, 4, 6208, 2, 18.0), (50, 1, 6, 36696, 1, 30.0), (75, 1, 6, 55056, 1, 30.0),
    (100, 2, 4, 30576, 2, 15.0), (100, 1, 6, 75376, 1, 30.0)]
ctx = sg.Context(0)
for prb, ports, qm, tbs, tm, snr in SHAPES:
    B = 4096 if prb >= 50 else 16384
    ocell = o.make_cell(prb, ports, 1)
    ocfg = o.make_cfg(ocell, sf_idx=1, cfi=1, qm=qm, tbs=tbs, tm=tm)
    pool = np.stack([o.gen_subframe(ocell, ocfg, 70000 + i, snr)[1] for i in range(8)])
    cell = sg.make_cell(prb, ports, 1)
    cfg = sg.make_cfg(cell, sf_idx=1, cfi=1, qm=qm, tbs=tbs, tm=tm)
    plan = sg.PdschPlan(ctx, cell, cfg, B)
    I = plan.info
    d_iq = torch.from_numpy(pool.view(np.float32).reshape(8, -1)).cuda()[torch.arange(B, device="cuda") % 8].contiguous()
    d_sf = torch.empty((B, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.empty((B, ports * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.empty((B, 5), dtype=torch.float32, device="cuda")
    d_sb = torch.empty((B, I.sb_sf_stride), dtype=torch.int16, device="cuda")
    d_pl = torch.zeros((B, I.payload_stride), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros((B, 4), dtype=torch.int32, device="cuda")

    def step(ev=None):
        fns = [lambda: plan.ofdm_rx(B, d_iq, d_sf), lambda: plan.chest(B, d_sf, d_ce, d_meas),
               lambda: plan.pdsch_llr(B, d_sf, d_ce, d_meas, 0.01, 0, 0, d_sb), lambda: plan.pdsch_turbo(B, d_sb, 4, d_pl, d_st)]
        for i, f in enumerate(fns):
            if ev:
                ev[i].record()
            f()
        if ev:
            ev[4].record()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    reps = 5
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(reps)]
    for k in range(reps):
        step(evs[k])
    torch.cuda.synchronize()
    ms = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) / reps for i in range(4)]
    st = d_st.cpu().numpy()
    ok = int((st[:, 0] == 1).sum())
    tot = sum(ms)
    print(json.dumps(dict(prb=prb, ports=ports, qm=qm, tbs=tbs, C=I.C, K=I.Kp, batch=B, crc_ok=ok, ms=dict(fft=ms[0], chest=ms[1], demod=ms[2], turbo=ms[3]),
                          ms_total=tot, sf_per_s=B / tot * 1e3, mbit_per_s=ok * tbs / tot / 1e3,
                          avg_iters=float(st[:, 1].sum()) / (B * I.C))), flush=True)
    plan.close()
    del d_iq, d_sf, d_ce, d_sb
    torch.cuda.empty_cache()
