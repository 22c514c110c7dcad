#!/usr/bin/env python
"""bench.py -- PDSCH decoded Mbit/s @ 20 MHz, MCS 28, TM1 (BASELINE.json configs[1]) on N B200.

A "step" is one pass of the hot path (OFDM demod -> channel estimate -> equalise/demap/descramble/dematch ->
turbo decode + CRC -> transport blocks) over one batch of synthetic subframes that is already resident in
HBM.  `value` = CRC-passing transport-block bits of all ranks / max-over-ranks device time.  `e2e` is the
same metric through the host-buffer C-ABI call (H2D of the IQ and D2H of payload/status inside the timed
region).  Subframe batches are independent: ranks shard them, no collective on the data path (weak scaling).

Synthetic inputs come from the oracle's TX-side generator (oracle/lteo_tx.c), which is test infrastructure
and not on the measured path.  `--impl reference` times the CPU restatement of the srsLTE path (the
reference's own implementation, srsLTE, is not available offline -- see DESIGN.md) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = dict(prb=100, ports=1, qm=6, tbs=75376, tm=1, cfi=1, sf_idx=1, rnti=0x1234, cell_id=1, cp=0)
METRIC = "pdsch_decoded_mbit_per_s_20mhz_mcs28_tm1"


def channel_taps():
    """fixed frequency-selective channel of BASELINE configs[2]: 6 sample-spaced taps per TX port, seed 77"""
    rng = np.random.default_rng(77)
    taps = (rng.standard_normal((2, 6)) + 1j * rng.standard_normal((2, 6))) * np.array([1, .7, .5, .3, .2, .1])
    if WORKLOAD["ports"] == 4:          # --ports 4: two more independent channels (the first two stay those of configs[2])
        taps = np.concatenate([taps, (rng.standard_normal((2, 6)) + 1j * rng.standard_normal((2, 6))) * np.array([1, .7, .5, .3, .2, .1])])
    return taps / np.sqrt((abs(taps) ** 2).sum(1, keepdims=True))


def gen_pool(o, pool, snr_db, seed0):
    ocell = o.make_cell(WORKLOAD["prb"], WORKLOAD["ports"], WORKLOAD["cell_id"], cp=WORKLOAD["cp"])
    ocfg = o.make_cfg(ocell, sf_idx=WORKLOAD["sf_idx"], cfi=WORKLOAD["cfi"], rnti=WORKLOAD["rnti"], qm=WORKLOAD["qm"],
                      tbs=WORKLOAD["tbs"], tm=WORKLOAD["tm"])
    tbs, iqs = [], []
    for i in range(pool):
        # payload seed = 10000*cfg + unit index, noise seed = payload seed + 5e6 (inside gen_subframe)
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 20000 + seed0 + i, snr_db, channel_taps() if WORKLOAD["tm"] == 2 else None)
        tbs.append(tb)
        iqs.append(iq)
    return ocell, ocfg, np.stack(tbs), np.stack(iqs)


def cpu_arm(o):
    """selects the build of the CPU restatement the baselines time: AVX2 (window-parallel turbo decoder, -O3) when the
    host has it, else portable scalar C.  Returns (kind string for `sample`, restore function)."""
    if o.have_avx2():
        prev = o.select("avx2")
        return "AVX2 window-parallel int16 turbo decoder + -O3 -mavx2 front end", lambda: o.select(prev)
    return "scalar C (-O2), host without AVX2", lambda: None


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU every 200 ms while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.2)
        except Exception as e:  # noqa: BLE001
            self.reasons.add("clock sampling unavailable: %s" % type(e).__name__)

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def run_reference(args, rank, world):
    """CPU arm: the restated srsLTE path on all host threads, one subframe per thread at a time."""
    if rank != 0:
        return
    from oracle import oracle as o
    cores = os.cpu_count() or 1
    ocell, ocfg, tbs, iqs = gen_pool(o, min(args.pool, 16), args.snr, 0)
    build_kind, _ = cpu_arm(o)
    per_step = max(cores * 128, 512)
    idx = np.arange(per_step) % len(iqs)
    iq = iqs[idx]
    for _ in range(args.warmup):
        o.ue_dl_decode_mt(ocell, ocfg, iq[:cores], cores, 0.01, args.noise_mode, args.max_iter)
    t0 = time.perf_counter()
    ok_bits = 0
    for _ in range(args.steps):
        ok, payload, status = o.ue_dl_decode_mt(ocell, ocfg, iq, cores, 0.01, args.noise_mode, args.max_iter)
        ok_bits += ok * WORKLOAD["tbs"]
    dt = time.perf_counter() - t0
    val = ok_bits / dt / 1e6
    sample = "%d subframes per step x %d steps, %d threads, one subframe per thread, %s, early stop on CRC" % (
        per_step, args.steps, cores, build_kind)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mbit/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": args.label, "subframes_per_step": per_step,
                   "max_iter": args.max_iter},
        "cpu_baseline": {"value": val, "unit": "Mbit/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "subframes_per_s": val * 1e6 / WORKLOAD["tbs"],
        "note": "CPU restatement of the srsLTE path (oracle/); srsLTE itself is not installable offline",
    }))


# BASELINE configs[4]: streaming mixed-bandwidth subframe batches.  (prb, ports, qm, tbs, tm, share of the stream)
MIXED = [(6, 1, 2, 152, 1, 0.30), (15, 1, 4, 2216, 1, 0.20), (25, 1, 6, 11448, 1, 0.20), (50, 2, 4, 6208, 2, 0.15),
         (100, 1, 6, 75376, 1, 0.15)]
MIXED_SNR = {2: 10.0, 4: 18.0, 6: 30.0}


def run_harq(args, rank, local_rank, world):
    """HARQ at a BLER operating point: 20 MHz MCS 28 at an SNR where a good share of the first transmissions
    fail; every transport block owns a device-resident soft buffer (its HARQ process), failed blocks are retransmitted
    with rv 2 and combined on the device.  A step = first transmissions of the whole batch + the retransmissions of the
    failed ones, through srsue_gpu_batch_submit / _wait with host buffers.  value = goodput (bits of blocks that ended
    up decoded / time)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    import srsue_b200 as sg
    from oracle import oracle as o
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = sg.lib()
    snr = args.snr if args.snr != 30.0 else 21.4       # first-transmission BLER about 20 % with 4 iterations
    N = args.batch if args.batch != 4096 else 1024
    pool = 16
    ocell = o.make_cell(100, 1, 1)
    cell = sg.make_cell(100, 1, 1)
    cfgs, ocfgs = {}, {}
    for rv in (0, 2):
        ocfgs[rv] = o.make_cfg(ocell, sf_idx=1, cfi=1, qm=6, tbs=75376, rv=rv)
        cfgs[rv] = sg.make_cfg(cell, sf_idx=1, cfi=1, qm=6, tbs=75376, rv=rv)
    sf_len = 15 * 2048
    pl_bytes = 75376 // 8
    # pinned pools: IQ of rv 0 and rv 2 for `pool` transport blocks (same payload seed, independent noise)
    p_iq = lib.srsue_gpu_host_alloc(2 * pool * sf_len * 8)
    h_iq = np.ctypeslib.as_array(C.cast(p_iq, C.POINTER(C.c_float)), shape=(2, pool, sf_len * 2)).view(np.complex64)
    tbs = []
    for i in range(pool):
        tb0, iq0, _ = o.gen_subframe(ocell, ocfgs[0], 30000 + 1000 * rank + i, snr)
        h_iq[0, i] = iq0
        tbs.append(tb0)
        # same payload (seed), different noise: regenerate with rv 2 and a noise seed offset through the snr-preserving generator
        grid_seed = 30000 + 1000 * rank + i
        tb2, iq2, _ = o.gen_subframe(ocell, ocfgs[2], grid_seed, snr, noise_seed=grid_seed + 7_000_000)
        assert np.array_equal(tb0, tb2)
        h_iq[1, i] = iq2
    p_pl = lib.srsue_gpu_host_alloc(N * pl_bytes)
    h_pl = np.ctypeslib.as_array(C.cast(p_pl, C.POINTER(C.c_uint8)), shape=(N, pl_bytes))
    ctx = sg.Context(local_rank)
    batch = sg.Batch(ctx, N, 0.01, 0, args.max_iter)
    first = sg.Batch.prepare([dict(cell=cell, cfg=cfgs[0], iq=h_iq[0, i % pool], payload=h_pl[i], softbuffer_id=i, new_data=1)
                              for i in range(N)])

    retx_all = sg.Batch.prepare([dict(cell=cell, cfg=cfgs[2], iq=h_iq[1, i % pool], payload=h_pl[i], softbuffer_id=i, new_data=0)
                                 for i in range(N)])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        batch.submit_prepared(first)
        lib.srsue_gpu_batch_wait(batch.h)
        launches = batch.stats()["launches"]
        failed = [i for i, d in enumerate(first[0]) if d.crc_ok != 1]
        ok2 = 0
        if failed:
            arr = (sg.SfDesc * len(failed))()
            for j, i in enumerate(failed):
                arr[j] = retx_all[0][i]
            retx = (arr, None, None)
            batch.submit_prepared(retx)
            lib.srsue_gpu_batch_wait(batch.h)
            launches += batch.stats()["launches"]
            ok2 = sum(1 for d in retx[0] if d.crc_ok == 1)
        return N - len(failed), ok2, len(failed), launches

    for _ in range(args.warmup):
        r = step()
    # what ends up decoded is what was sent
    verified = all(np.array_equal(h_pl[i], tbs[i % pool]) for i, d in enumerate(first[0]) if d.crc_ok == 1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    tot = np.zeros(4)
    for _ in range(args.steps):
        tot += np.array(step(), dtype=float)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join()
    vals = torch.tensor([dt], dtype=torch.float64, device="cuda")
    sums = torch.tensor(tot.tolist(), dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    if rank == 0:
        ok1, ok2, nfail, launches = sums.tolist()
        sent = N * world * args.steps
        val = (ok1 + ok2) * 75376 / vals.item() / 1e6
        print(json.dumps({
            "metric": "pdsch_goodput_mbit_per_s_20mhz_mcs28_harq", "value": val, "unit": "Mbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": vals.item() / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": {"workload": "20MHz 100PRB TM1 64QAM MCS28 TBS75376 AWGN %gdB, rv 0 then rv 2 for failed blocks" % snr,
                       "transport_blocks_per_step_per_gpu": N, "max_iter": args.max_iter,
                       "api": "srsue_gpu_batch_submit/_wait, device-resident soft buffers, host IQ/payload buffers"},
            "bler_first_tx": nfail / sent, "bler_after_rv2": (nfail - ok2) / sent,
            "subframes_per_s": (sent + nfail) / vals.item(), "verified_bit_exact_payload": bool(verified),
            "e2e": {"value": val, "unit": "Mbit/s", "h2d_bytes_per_step": (sent + nfail) / args.steps * sf_len * 8,
                    "d2h_bytes_per_step": (sent + nfail) / args.steps * pl_bytes},
            "gpu_launches": int(launches), "clocks": sampler.summary()}))
    batch.close()
    lib.srsue_gpu_host_free(p_iq)
    lib.srsue_gpu_host_free(p_pl)
    if world > 1:
        dist.destroy_process_group()


def run_mixed(args, rank, local_rank, world, embedded=False):
    """A stream of heterogeneous batches (mixed 1.4-20 MHz bandwidths) through the batching layer with HOST buffers:
    the global stream is world x 20 batches, assigned to ranks by estimated turbo work (srsue_b200.shard), each rank
    packs its share with srsue_gpu_batch_submit / _wait.  Every number here is end to end (H2D + D2H inside)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    import srsue_b200 as sg
    from srsue_b200.shard import balance_by_work
    from oracle import oracle as o
    from srsue_b200.shard import bind_rank_to_gpu_numa
    if world > 1 and not embedded:
        bind_rank_to_gpu_numa(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1 and not embedded:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = sg.lib()
    # ---- the global stream: batches of one shape each, sizes from the shares; identical on every rank ----
    rng = np.random.default_rng(2024)
    per_rank = args.batch if args.batch != 4096 else 2048
    batches = []                       # (shape index, subframes)
    for _ in range(world):
        for j in range(20):
            m = j % len(MIXED)
            batches.append((m, max(1, int(per_rank * MIXED[m][5] / 4 * rng.uniform(0.7, 1.3)))))
    work = []
    for m, n in batches:
        s = o.cbsegm(MIXED[m][3])
        work.append(n * (s.Cp * s.Kp + s.Cm * s.Km))
    mine = balance_by_work(work, world)[rank]
    # ---- inputs of this rank: 4 distinct subframes per shape, tiled, in pinned host memory -------------
    ctx = sg.Context(local_rank)
    shapes, pinned = {}, []
    tbs_entries = {}
    for m in sorted({batches[i][0] for i in mine}):
        prb, ports, qm, tbs, tm, _ = MIXED[m]
        ocell = o.make_cell(prb, ports, 1)
        ocfg = o.make_cfg(ocell, sf_idx=1, cfi=1, qm=qm, tbs=tbs, tm=tm)
        dcis = None
        if args.blind:
            # a format 1A DCI for the full-band allocation of this shape in the first UE-specific candidate; its size goes
            # into the (synthetic, test-style) size table because these shapes are not rows of 36.213 Table 7.1.7.2.1-1
            from tests.srslte_ctypes import DciMsg, RaDlDci
            mcs = {2: 5, 4: 12, 6: 22}[qm]
            tbs_entries[(mcs if mcs < 10 else mcs - 1 if mcs < 17 else mcs - 2, prb)] = tbs
            sent = RaDlDci()
            sent.mcs_idx, sent.rv_idx, sent.alloc_type = mcs, 0, 2
            sent.type2_alloc.RB_start, sent.type2_alloc.L_crb, sent.type2_alloc.n_prb1a = 0, prb, 1
            msg = DciMsg()
            nb = lib.srslte_dci_msg_pack_pdsch(C.byref(sent), 2, C.byref(msg), prb, True)
            rk, _ = o.pdcch_regs(ocell, 1, 6)
            ss = o.pdcch_search_space(len(rk) // 9, 1, 0x1234)
            dcis = [(np.frombuffer(msg.data, np.uint8)[:nb].copy(), 0x1234, ss[0][0], ss[0][1])]
        gen = [o.gen_subframe(ocell, ocfg, 50000 + 100 * m + i + 1000 * rank, MIXED_SNR[qm], None, pcfich=args.blind, dcis=dcis) for i in range(4)]
        cell = sg.make_cell(prb, ports, 1)
        shapes[m] = dict(ocell=ocell, ocfg=ocfg, cell=cell, cfg=sg.make_cfg(cell, sf_idx=1, cfi=1, qm=qm, tbs=tbs, tm=tm),
                         tb=[g[0] for g in gen], iq=[g[1] for g in gen], tbs=tbs)
    items, truth = [], []
    total_iq = total_pl = 0
    for i in mine:
        m, n = batches[i]
        sh = shapes[m]
        sf_len = len(sh["iq"][0])
        p = lib.srsue_gpu_host_alloc(n * sf_len * 8)
        q = lib.srsue_gpu_host_alloc(n * ((sh["tbs"] + 7) // 8))
        pinned += [p, q]
        h_iq = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n, sf_len * 2)).view(np.complex64)
        h_pl = np.ctypeslib.as_array(C.cast(q, C.POINTER(C.c_uint8)), shape=(n, (sh["tbs"] + 7) // 8))
        for r in range(n):
            h_iq[r] = sh["iq"][r % 4]
            items.append(dict(cell=sh["cell"], cfg=sh["cfg"], iq=h_iq[r], payload=h_pl[r]))
            truth.append(sh["tb"][r % 4])
        total_iq += n * sf_len * 8
        total_pl += n * ((sh["tbs"] + 7) // 8)
    # arrival order: interleave the batches the way a multi-cell capture would deliver them
    perm = np.random.default_rng(7 + rank).permutation(len(items))
    items = [items[i] for i in perm]
    truth = [truth[i] for i in perm]
    batch = sg.Batch(ctx, len(items), 0.01, 0, args.max_iter)
    if args.blind:
        from tests.srslte_ctypes import install_tbs_table
        install_tbs_table(lib, tbs_entries)
        for it in items:                                  # the library is told the subframe number and the RNTI, nothing else
            it["cfg"] = sg.make_cfg(it["cell"], sf_idx=1, cfi=1, rnti=0x1234, qm=2, tbs=0)
    prepared = sg.Batch.prepare(items)
    pl_cap = max(sh["tbs"] for sh in shapes.values()) // 8

    def submit():
        if args.blind:
            batch._keep = None
            rc = lib.srsue_gpu_batch_submit_blind(batch.h, prepared[0], len(prepared[0]), 6, pl_cap)
            assert rc == 0, lib.srsue_gpu_last_error()
            batch._keep, batch._blind = prepared, True
        else:
            batch.submit_prepared(prepared)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        submit()
        res = batch.wait()
    verified = all(r["crc_ok"] == 1 and np.array_equal(r["payload"][:len(t)], t) for r, t in zip(res, truth))
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    bits = 0
    launches = 0
    for _ in range(args.steps):
        submit()
        lib.srsue_gpu_batch_wait(batch.h)
        batch._keep = None
        launches += batch.stats()["launches"]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # every step decodes the same inputs; count the CRC-passing bits of the last one outside the timed region
    bits = args.steps * sum(d.cfg.tbs for d in prepared[0] if d.crc_ok == 1)
    sampler.stop_flag = True
    sampler.join()
    vals = torch.tensor([dt], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(bits), float(len(items)) * args.steps, float(launches), float(total_iq), float(total_pl)],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    result = None
    if rank == 0:
        dt_max = vals.item()
        bits_all, sf_all, launches_all, iq_all, pl_all = sums.tolist()
        val = bits_all / dt_max / 1e6
        out = {"metric": "pdsch_decoded_mbit_per_s_mixed_bandwidth_stream", "value": val, "unit": "Mbit/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_max / args.steps * 1e3, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
               "config": {"workload": "mixed 1.4-20 MHz stream (BASELINE configs[4]): " +
                          ", ".join("%dPRB/%dport/Qm%d/TBS%d %.0f%%" % (a, b, c, d, 100 * f) for a, b, c, d, e, f in MIXED),
                          "subframes_per_step": sf_all / args.steps, "batches": len(batches), "max_iter": args.max_iter,
                          "parallelism": "batches assigned to GPUs by estimated turbo work, no collective",
                          "api": ("srsue_gpu_batch_submit_blind/_wait: PCFICH + PDCCH search + DCI -> grant in the library, host buffers"
                                  if args.blind else "srsue_gpu_batch_submit/_wait, host buffers")},
               "subframes_per_s": sf_all / dt_max, "verified_bit_exact_payload": bool(verified),
               "e2e": {"value": val, "unit": "Mbit/s", "h2d_bytes_per_step": iq_all, "d2h_bytes_per_step": pl_all},
               "gpu_launches": int(launches_all), "clocks": sampler.summary()}
        if world == 1 and not args.no_cpu_baseline and not embedded:
            cores = os.cpu_count() or 1
            build_kind, restore = cpu_arm(o)
            t0 = time.perf_counter()
            cbits = 0
            for m, sh in shapes.items():
                n = max(cores, int(16 * cores * MIXED[m][5] * 5))
                sub = np.stack([sh["iq"][i % 4] for i in range(n)])
                ok, _, _ = o.ue_dl_decode_mt(sh["ocell"], sh["ocfg"], sub, cores, 0.01, 0, args.max_iter)
                cbits += ok * sh["tbs"]
            cdt = time.perf_counter() - t0
            restore()
            out["cpu_baseline"] = {"value": cbits / cdt / 1e6, "unit": "Mbit/s", "cores": cores, "kind": "port",
                                   "sample": "same shape mix, %d threads, CPU restatement of the srsLTE path: %s" % (cores, build_kind)}
        if embedded:
            result = {"value": val, "unit": "Mbit/s", "metric": out["metric"], "subframes_per_s": out["subframes_per_s"],
                      "ms_per_step": out["ms_per_step"], "steps": args.steps, "workload": out["config"]["workload"],
                      "subframes_per_step": out["config"]["subframes_per_step"], "api": out["config"]["api"],
                      "h2d_bytes_per_step": iq_all, "d2h_bytes_per_step": pl_all, "verified_bit_exact_payload": bool(verified),
                      "note": "BASELINE configs[4]: end to end from pinned host buffers through srsue_gpu_batch_submit / _wait, "
                              "batches assigned to the ranks by estimated turbo work, no collective"}
        else:
            print(json.dumps(out))
    batch.close()
    ctx.close()
    for p in pinned:
        lib.srsue_gpu_host_free(p)
    if world > 1 and not embedded:
        dist.destroy_process_group()
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="subframes per step and per GPU (device-resident)")
    ap.add_argument("--e2e-batch", type=int, default=4096, help="subframes per step for the host-buffer measurement")
    ap.add_argument("--pool", type=int, default=512, help="distinct synthetic subframes (tiled to the batch)")
    ap.add_argument("--waterfall-snr", type=float, default=21.5, help="SNR of the waterfall leg (first-transmission BLER about 0.2)")
    ap.add_argument("--no-legs", action="store_true", help="skip the fixed-4-iteration and waterfall legs")
    ap.add_argument("--blind", action="store_true", help="--workload mixed: the caller supplies no CFI and no grant, the library decodes "
                                                         "PCFICH and PDCCH for every capture (srsue_gpu_batch_submit_blind)")
    ap.add_argument("--snr", type=float, default=30.0)
    ap.add_argument("--max-iter", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="mcs28", choices=["mcs28", "tm2", "mixed", "harq"],
                    help="mcs28: BASELINE configs[1] (the headline metric); tm2: configs[2] (2-port transmit diversity, MCS 16, "
                         "frequency-selective channel, MMSE with the estimated noise); mixed: configs[4], heterogeneous stream "
                         "through the batching layer; harq: MCS 28 at a BLER operating point with rv 2 retransmissions combined in "
                         "device-resident soft buffers")
    ap.add_argument("--multi-batch", type=int, default=4096, help="subframes per device and step of the one-process dispatcher leg (as many as the e2e leg by default)")
    ap.add_argument("--ports", type=int, default=0, choices=[0, 2, 4],
                    help="--workload tm2: 4 = a four-port cell (SFBC-FSTD, CRS of ports 2 / 3; SPEC 15c); not a BASELINE config")
    ap.add_argument("--cp", default="norm", choices=["norm", "ext"],
                    help="mcs28 / tm2: ext = extended cyclic prefix (12 symbols; SPEC 15b; MCS 28 becomes TBS 61664, the largest "
                         "that fits 11 data symbols); not a BASELINE config")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    global METRIC
    if args.workload == "tm2":
        WORKLOAD.update(ports=2, qm=4, tbs=30576, tm=2)
        METRIC = "pdsch_decoded_mbit_per_s_20mhz_mcs16_tm2"
        if args.snr == 30.0:
            args.snr = 15.0
    args.noise_mode = 1 if WORKLOAD["tm"] == 2 else 0
    args.label = ("20MHz 100PRB TM2 2-port 16QAM MCS16 TBS30576, 6-tap channel + AWGN %gdB (BASELINE configs[2])" if WORKLOAD["tm"] == 2
                  else "20MHz 100PRB TM1 64QAM MCS28 TBS75376 AWGN %gdB (BASELINE configs[1])") % args.snr
    if args.workload in ("mcs28", "tm2") and (args.ports == 4 or args.cp == "ext"):       # other cell shapes of the same path
        if args.ports == 4:
            if args.workload != "tm2":
                ap.error("--ports 4 goes with --workload tm2 (a four-port cell transmits with diversity)")
            WORKLOAD.update(ports=4)
            METRIC += "_4port"
            if args.snr == 15.0:       # ports 2 / 3 have half the pilots: the waterfall of this grant sits about 2 dB higher
                args.snr = 19.0
        if args.cp == "ext":
            WORKLOAD.update(cp=1)
            WORKLOAD.update(tbs=61664 if args.workload == "mcs28" else 22920)      # 11 data symbols: the next smaller sizes
            METRIC += "_extcp"
        args.label = "20MHz 100PRB %s %d-port %s TBS%d %s cyclic prefix, %s AWGN %gdB (cell shape beyond BASELINE's configs)" % (
            "TM2" if WORKLOAD["tm"] == 2 else "TM1", WORKLOAD["ports"], "16QAM" if WORKLOAD["qm"] == 4 else "64QAM", WORKLOAD["tbs"],
            "extended" if WORKLOAD["cp"] else "normal", "6-tap channels +" if WORKLOAD["tm"] == 2 else "", args.snr)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if args.workload == "mixed":
        run_mixed(args, rank, local_rank, world)
        return
    if args.workload == "harq":
        run_harq(args, rank, local_rank, world)
        return

    import torch
    import torch.distributed as dist
    import srsue_b200 as sg
    from oracle import oracle as o   # input generator + cpu_baseline only

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsrsue_gpu has no CPU path")
    from srsue_b200.shard import bind_rank_to_gpu_numa
    numa_cpus = bind_rank_to_gpu_numa(local_rank) if world > 1 else None      # pinned staging on the GPU's NUMA node
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # a host-side group for the one place where ranks must wait WITHOUT touching their GPU: while rank 0 drives all
    # devices through the library's own dispatcher, an NCCL barrier would park a spinning kernel on every other GPU
    cpu_group = dist.new_group(backend="gloo") if world > 1 else None

    # ---- inputs -------------------------------------------------------------------------------------
    ocell, ocfg, tbs, iqs = gen_pool(o, args.pool, args.snr, 1000 * rank)
    B = args.batch
    idx = np.arange(B) % args.pool
    ctx = sg.Context(local_rank)
    cell = sg.make_cell(WORKLOAD["prb"], WORKLOAD["ports"], WORKLOAD["cell_id"], cp=WORKLOAD["cp"])
    cfg = sg.make_cfg(cell, sf_idx=WORKLOAD["sf_idx"], cfi=WORKLOAD["cfi"], rnti=WORKLOAD["rnti"], qm=WORKLOAD["qm"],
                      tbs=WORKLOAD["tbs"], tm=WORKLOAD["tm"])
    plan = sg.PdschPlan(ctx, cell, cfg, B)
    I = plan.info
    d_pool = torch.from_numpy(iqs.view(np.float32).reshape(args.pool, -1)).cuda()
    d_iq = d_pool[torch.from_numpy(idx).cuda()].contiguous()          # [B][sf_len*2] float32, 245 760 B per subframe
    del d_pool
    d_sf = torch.empty((B, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.empty((B, WORKLOAD["ports"] * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.empty((B, 5), dtype=torch.float32, device="cuda")
    d_sb = torch.empty((B, I.sb_sf_stride), dtype=torch.int16, device="cuda")
    d_pl = torch.zeros((B, I.payload_stride), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros((B, 4), dtype=torch.int32, device="cuda")

    names = ["ofdm_fft", "chest", "equalise_demap_dematch", "turbo_crc_tb"]

    def step(ev=None):
        if ev:
            ev[0].record()
        plan.ofdm_rx(B, d_iq, d_sf)
        if ev:
            ev[1].record()
        plan.chest(B, d_sf, d_ce, d_meas)
        if ev:
            ev[2].record()
        plan.pdsch_llr(B, d_sf, d_ce, d_meas, 0.01, args.noise_mode, 0, d_sb)        # srsUE passes noise_estimate = 0.01
        if ev:
            ev[3].record()
        plan.pdsch_turbo(B, d_sb, args.max_iter, d_pl, d_st)
        if ev:
            ev[4].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    # correctness of what is being timed: every transport block passes CRC and equals what was sent
    st = d_st.cpu().numpy()
    pl = d_pl.cpu().numpy()
    verified = bool((st[:, 0] == 1).all() and np.array_equal(pl[:args.pool], tbs[idx[:args.pool]]))
    avg_iter = float(st[:, 1].sum()) / (B * I.C)

    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    launches = 0
    barrier()
    for k in range(args.steps):
        step(evs[k])
        # kernels of this library per step: FFT, channel estimate, demap / de-match, per code-block size the turbo decoder and
        # its de-interleaver, transport-block assembly
        launches += 3 + 2 * (2 if I.Cm else 1) + 1
    barrier()
    total_ms = evs[0][0].elapsed_time(evs[-1][4])
    stage_ms = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) / args.steps for i in range(4)]
    ok_bits = float((d_st[:, 0] == 1).sum().item()) * WORKLOAD["tbs"] * args.steps

    # ---- the other two operating points of SURVEY 7.3 / 8(d), same step, same buffers ------------------------
    def timed_leg(iq_dev):
        """W warm-up + K timed steps of the whole chain on iq_dev; returns ms per step, per-stage ms, status array"""
        nonlocal d_iq
        keep, d_iq = d_iq, iq_dev
        for _ in range(args.warmup):
            step()
        barrier()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
        for k in range(args.steps):
            step(ev[k])
        barrier()
        ms = ev[0][0].elapsed_time(ev[-1][4]) / args.steps
        st_ms = [sum(e[i].elapsed_time(e[i + 1]) for e in ev) / args.steps for i in range(4)]
        d_iq = keep
        return ms, st_ms, d_st.cpu().numpy()

    legs = {}
    if not args.no_legs and WORKLOAD["tm"] == 1:
        # (a) fixed iteration count: no early stop before max_iter, CRC verdicts still reported
        plan.set_min_iter(args.max_iter)
        ms, st_ms, stl = timed_leg(d_iq)
        plan.set_min_iter(1)
        legs["fixed_%d_iter" % args.max_iter] = dict(ms=ms, turbo_ms=st_ms[3], ok=int((stl[:, 0] == 1).sum()), iters=float(stl[:, 1].sum()),
                                                      snr_db=args.snr)
        launches += 0                                      # (legs are outside the headline's timed region)
        # (b) waterfall: the same grant at an SNR where a fifth of the transport blocks fail; early stop as in the headline
        wpool = min(args.pool, 256)
        _, _, wtbs, wiqs = gen_pool(o, wpool, args.waterfall_snr, 500000 + 1000 * rank)
        d_wp = torch.from_numpy(wiqs.view(np.float32).reshape(wpool, -1)).cuda()
        d_wiq = d_wp[torch.from_numpy(np.arange(B) % wpool).cuda()].contiguous()
        del d_wp
        ms, st_ms, stl = timed_leg(d_wiq)
        wpl = d_pl.cpu().numpy()
        good = stl[:wpool, 0] == 1
        legs["waterfall"] = dict(ms=ms, turbo_ms=st_ms[3], ok=int((stl[:, 0] == 1).sum()), iters=float(stl[:, 1].sum()),
                                 snr_db=args.waterfall_snr, passing_equal_sent=bool(np.array_equal(wpl[:wpool][good], wtbs[:wpool][good])))
        del d_wiq
        # leave the headline's results in the output buffers (the verification below reads them)
        step()
        barrier()

    # ---- end to end through the host-buffer call --------------------------------------------------------
    EB = min(args.e2e_batch, B)
    lib = sg.lib()
    import ctypes as C
    nbytes_iq = EB * I.sf_len * 8
    p_iq = lib.srsue_gpu_host_alloc(nbytes_iq)
    h_iq = np.ctypeslib.as_array(C.cast(p_iq, C.POINTER(C.c_float)), shape=(EB, I.sf_len * 2))
    h_iq[:] = iqs.view(np.float32).reshape(args.pool, -1)[np.arange(EB) % args.pool]
    p_pl = lib.srsue_gpu_host_alloc(EB * I.payload_stride)
    h_pl = np.ctypeslib.as_array(C.cast(p_pl, C.POINTER(C.c_uint8)), shape=(EB, I.payload_stride))
    h_st = np.zeros((EB, 4), np.int32)
    eplan = sg.PdschPlan(ctx, cell, cfg, EB)
    for _ in range(2):
        eplan.decode_batch_host(EB, h_iq, 0.01, args.noise_mode, args.max_iter, h_pl, h_st)
    barrier()
    t0 = time.perf_counter()
    e2e_bits = 0
    for _ in range(args.steps):
        eplan.decode_batch_host(EB, h_iq, 0.01, args.noise_mode, args.max_iter, h_pl, h_st)
        e2e_bits += int((h_st[:, 0] == 1).sum()) * WORKLOAD["tbs"]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_ok = bool(np.array_equal(h_pl[:args.pool], tbs[(np.arange(EB) % args.pool)[:args.pool]]))
    # the same call fed with int16 captures (the radio's wire format): half the bytes over PCIe, conversion on the FFT's loads
    pool_f = iqs.view(np.float32).reshape(args.pool, -1)
    sc16_scale = np.float32(float(np.abs(pool_f).max()) / 32000.0)
    p_q = lib.srsue_gpu_host_alloc(nbytes_iq // 2)
    h_q = np.ctypeslib.as_array(C.cast(p_q, C.POINTER(C.c_int16)), shape=(EB, I.sf_len * 2))
    h_q[:] = np.rint(pool_f / sc16_scale).astype(np.int16)[np.arange(EB) % args.pool]
    eplan.set_iq_format(True, float(sc16_scale))
    h_pl[:] = 0
    for _ in range(2):
        eplan.decode_batch_host(EB, h_q, 0.01, args.noise_mode, args.max_iter, h_pl, h_st)
    barrier()
    t0 = time.perf_counter()
    sc16_bits = 0
    for _ in range(args.steps):
        eplan.decode_batch_host(EB, h_q, 0.01, args.noise_mode, args.max_iter, h_pl, h_st)
        sc16_bits += int((h_st[:, 0] == 1).sum()) * WORKLOAD["tbs"]
    torch.cuda.synchronize()
    sc16_s = time.perf_counter() - t0
    sc16_ok = bool(np.array_equal(h_pl[:args.pool], tbs[(np.arange(EB) % args.pool)[:args.pool]]))
    sampler.stop_flag = True
    sampler.join()

    # ---- the library's own multi-GPU dispatcher: ONE process, one handle over all N GPUs ---------------------
    # (srsue_gpu_batch_create_multi: per-device contexts, streams, pinned staging and host threads inside the library;
    # rank 0 drives it while the other ranks, done with their own measurement, release their buffers and wait)
    multi = None
    if not args.no_legs and WORKLOAD["tm"] == 1:
        eplan.close()
        d_iq = d_sf = d_ce = d_sb = None
        torch.cuda.empty_cache()
        barrier()
        if rank == 0:
            nsub = min(EB, args.multi_batch) * world
            p_mpl = lib.srsue_gpu_host_alloc(nsub * I.payload_stride)
            h_mpl = np.ctypeslib.as_array(C.cast(p_mpl, C.POINTER(C.c_uint8)), shape=(nsub, I.payload_stride))
            mb = sg.Batch(None, nsub, 0.01, args.noise_mode, args.max_iter, devices=list(range(world)))
            prepared = sg.Batch.prepare([dict(cell=cell, cfg=cfg, iq=h_iq[i % EB].view(np.complex64), payload=h_mpl[i][:WORKLOAD["tbs"] // 8])
                                         for i in range(nsub)])
            for _ in range(2):
                mb.submit_prepared(prepared)
                lib.srsue_gpu_batch_wait(mb.h)
            t0 = time.perf_counter()
            mbits = 0
            t_submit = 0.0
            for _ in range(args.steps):
                ts = time.perf_counter()
                mb.submit_prepared(prepared)
                t_submit += time.perf_counter() - ts
                lib.srsue_gpu_batch_wait(mb.h)
            mdt = time.perf_counter() - t0
            # every step decodes the same inputs: count the passing blocks of the last one outside the timed region
            # (a Python loop over thousands of ctypes descriptors would be a fifth of a 5 ms step)
            mbits = args.steps * sum(1 for d in prepared[0] if d.crc_ok == 1) * WORKLOAD["tbs"]
            ok = bool(np.array_equal(h_mpl[:args.pool % EB or EB, :WORKLOAD["tbs"] // 8], tbs[(np.arange(EB) % args.pool)[:args.pool % EB or EB]]))
            multi = {"value": mbits / mdt / 1e6, "unit": "Mbit/s", "n_devices": world, "subframes_per_step": nsub,
                     "shares": [n for n, _ in mb.device_shares()], "verified_bit_exact_payload": ok,
                     "ms_per_step": mdt / args.steps * 1e3, "submit_call_ms": t_submit / args.steps * 1e3,
                     "api": "srsue_gpu_batch_create_multi + srsue_gpu_batch_submit/_wait from ONE process, host buffers (H2D and D2H inside)"}
            mb.close()
            lib.srsue_gpu_host_free(p_mpl)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)           # the other ranks wait on the host, their GPUs stay idle for rank 0

    # ---- reduce over ranks (max time, sum of units) ------------------------------------------------------
    leg_names = sorted(legs)
    vals = torch.tensor([total_ms, e2e_s, sc16_s] + [legs[n]["ms"] for n in leg_names] + [legs[n]["turbo_ms"] for n in leg_names],
                        dtype=torch.float64, device="cuda")
    sums = torch.tensor([ok_bits, float(e2e_bits), float(launches), float(sc16_bits)] + [float(legs[n]["ok"]) for n in leg_names] +
                        [legs[n]["iters"] for n in leg_names], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    total_ms_max, e2e_s_max, sc16_s_max = vals.tolist()[:3]
    ok_bits_all, e2e_bits_all, launches_all, sc16_bits_all = sums.tolist()[:4]
    nl = len(leg_names)
    leg_ms, leg_turbo_ms = vals.tolist()[3:3 + nl], vals.tolist()[3 + nl:3 + 2 * nl]
    leg_ok, leg_iters = sums.tolist()[4:4 + nl], sums.tolist()[4 + nl:4 + 2 * nl]

    # BASELINE configs[4] as a leg of the same line, so that the driver's 1/2/4/8-GPU runs carry it: the mixed-bandwidth
    # stream through the batching layer, host buffers (every rank takes part)
    mixed_leg = None
    if not args.no_legs and WORKLOAD["tm"] == 1:
        import copy
        margs = copy.copy(args)
        margs.batch, margs.steps, margs.warmup, margs.blind = 1024, 5, 3, False
        mixed_leg = run_mixed(margs, rank, local_rank, world, embedded=True)

    if rank == 0:
        value = ok_bits_all / (total_ms_max * 1e-3) / 1e6
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        sm_max = peaks.get("sm_max_mhz", 1965.0)
        # ncu figures of the shipped kernels: read from the summary tools/ncu_chain_summary.py writes (never literals here)
        ncu = None
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "chain_kernels_r02.json")))
        except Exception:  # noqa: BLE001
            pass
        # dominant kernel: the turbo decoder.  Algorithmic work per SURVEY 8d: 168 * K * iterations int16 ops
        # per code block.  Peak from the measured issue rates (tools/alu_peak.cu, profiles/alu_peak_r01.json):
        # 64 lanes/clk/SM of VIADD.16x2 on the FMA pipe (2 ops) + 64 lanes/clk/SM of VIADDMNMX.S16x2 on the ALU
        # pipe (4 ops) = 384 int16 ops/clk/SM.
        n_cb = B * I.C
        turbo_ops = 168.0 * I.Kp * avg_iter * n_cb
        alu_peak_tops = 148 * 384 * sm_max * 1e6 / 1e12
        turbo_tops = turbo_ops / (stage_ms[3] * 1e-3) / 1e12
        # algorithmic bytes per subframe: FFT 245 760 in + 134 400 out; channel estimate 134 400 in +
        # 134 400 out; K3+K4 fused: PDSCH REs 120 000 + estimates 120 000 in, soft buffer 454 584 out; turbo: soft
        # buffer in + transport block out
        np_ = WORKLOAD["ports"]
        # channel estimate: credited with what it moves (the 4 CRS symbols in, every estimate out), not with SURVEY's
        # whole-grid read (268 800 B for cfg2)
        nsym, npil = (12 if WORKLOAD["cp"] else 14), (6 if np_ == 4 else 4)        # symbols per subframe, pilot symbols read
        alg_bytes = {"ofdm_fft": I.sf_len * 8 + nsym * I.nsc * 8, "chest": npil * I.nsc * 8 + nsym * I.nsc * 8 * np_ + 20,
                     "equalise_demap_dematch": I.nof_re * 8 * (1 + min(np_, 2)) + I.C * (3 * I.Kp + 12) * 2,     # an RE is combined with two ports' estimates
                     "turbo_crc_tb": I.C * (3 * I.Kp + 12) * 2 + I.payload_stride + 4 * I.C}
        # cfg2: 380 160 / 172 820 / 694 584 / 464 058 bytes per subframe (channel estimate as moved)
        stages = []
        for i, n in enumerate(names):
            gbs = alg_bytes[n] * B / (stage_ms[i] * 1e-3) / 1e9
            stages.append({"kernel": n, "ms": stage_ms[i], "share": stage_ms[i] / sum(stage_ms), "algorithmic_gbs": gbs,
                           "frac_of_measured_hbm": gbs / hbm_peak})
        out = {
            "metric": METRIC, "value": value, "unit": "Mbit/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16", "data": "synthetic",
            "config": {"workload": args.label,
                       "subframes_per_step_per_gpu": B, "distinct_subframes": args.pool, "max_iter": args.max_iter,
                       "early_stop": "CRC24B per code block", "avg_turbo_iterations": avg_iter,
                       "l2_policy": "inputs larger than L2 (%.0f MB of IQ per step)" % (B * I.sf_len * 8 / 1e6),
                       "parallelism": "independent subframe batches per GPU, no collective",
                       "host_affinity": ("rank pinned to %d CPUs of its GPU's NUMA node" % len(numa_cpus)) if numa_cpus else "default"},
            "subframes_per_s": value * 1e6 / WORKLOAD["tbs"],
            "verified_bit_exact_payload": verified and e2e_ok,
            "e2e": {"value": e2e_bits_all / e2e_s_max / 1e6, "unit": "Mbit/s", "h2d_bytes_per_step": nbytes_iq,
                    "d2h_bytes_per_step": EB * (I.payload_stride + 16), "subframes_per_step": EB},
            "e2e_sc16": {"value": sc16_bits_all / sc16_s_max / 1e6, "unit": "Mbit/s", "h2d_bytes_per_step": nbytes_iq // 2,
                         "d2h_bytes_per_step": EB * (I.payload_stride + 16), "subframes_per_step": EB,
                         "verified_bit_exact_payload": sc16_ok,
                         "note": "same call with int16 {re, im} host captures (srsue_gpu_pdsch_plan_set_iq_format); `e2e` above "
                                 "stays on the reference's cf_t boundary"},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "alu", "kernel": "turbo_decode_crc_kernel", "achieved": turbo_tops, "peak": alu_peak_tops,
                         "unit": "Tint16op/s", "frac": turbo_tops / alu_peak_tops,
                         "traffic": (ncu["turbo"]["dram_bytes"] / ncu["batch"] * B) if ncu else None,
                         "algorithmic_ops_per_launch": turbo_ops, "launch_ms": stage_ms[3],
                         "algorithmic_bytes_per_launch": alg_bytes["turbo_crc_tb"] * B,
                         "ncu_evidence": ({"source": ncu["source"], "profiled_batch": ncu["batch"], **ncu["turbo"],
                                           "traffic_scaling": "dram bytes of the profiled launch x (bench batch / profiled batch)"}
                                          if ncu else None),
                         "peak_source": "measured VIADD.16x2 + VIADDMNMX.S16x2 issue rates at %.0f MHz (profiles/alu_peak_r01.json)" % sm_max,
                         "note": "integer-ALU bound (north star); HBM-bound front-end kernels are listed in `stages`"},
            "stages": stages,
            "clocks": sampler.summary(),
        }
        out["e2e"]["sc16"] = {k: out["e2e_sc16"][k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")}
        # what the box gives: plain cudaMemcpyAsync from pinned memory to N GPUs at once (tools/h2d_ceiling.cu)
        try:
            ceil = json.load(open(os.path.join(ROOT, "profiles", "h2d_ceiling_r02.json")))
            c = [r["gb_per_s"] for r in ceil["runs"] if r["n_gpus"] == world and r["variant"] == "pinned"]
            if c:
                sf_per_s = out["e2e"]["value"] * 1e6 / WORKLOAD["tbs"]
                out["e2e"]["h2d_gb_per_s"] = sf_per_s * I.sf_len * 8 / 1e9
                out["e2e"]["h2d_ceiling_gb_per_s"] = c[0]
                out["e2e"]["frac_of_h2d_ceiling"] = out["e2e"]["h2d_gb_per_s"] / c[0]
                out["e2e"]["ceiling_source"] = "profiles/h2d_ceiling_r02.json (all N GPUs copying concurrently from pinned host memory)"
        except Exception:  # noqa: BLE001
            pass
        if multi:
            multi["vs_e2e_of_the_rank_per_gpu_path"] = multi["value"] / out["e2e"]["value"]
            out["multi_gpu_dispatcher"] = multi
        if ncu:
            for st_ in stages:
                k = {"ofdm_fft": "fft", "chest": "chest", "equalise_demap_dematch": "demap", "turbo_crc_tb": "turbo"}[st_["kernel"]]
                if k in ncu:
                    st_["ncu"] = ncu[k]
        for i, n in enumerate(leg_names):
            nsf = B * world
            ops = 168.0 * I.Kp * leg_iters[i]
            tops = ops / world / (leg_turbo_ms[i] * 1e-3) / 1e12
            out[n] = {"value": leg_ok[i] * WORKLOAD["tbs"] / (leg_ms[i] * 1e-3) / 1e6, "unit": "Mbit/s", "ms_per_step": leg_ms[i],
                      "snr_db": legs[n]["snr_db"], "bler": 1.0 - leg_ok[i] / nsf, "avg_turbo_iterations": leg_iters[i] / (nsf * I.C),
                      "turbo_ms": leg_turbo_ms[i], "code_blocks_per_s": nsf * I.C / (leg_turbo_ms[i] * 1e-3),
                      "roofline": {"achieved": tops, "peak": alu_peak_tops, "unit": "Tint16op/s", "frac": tops / alu_peak_tops}}
            if "passing_equal_sent" in legs[n]:
                out[n]["passing_blocks_equal_sent"] = legs[n]["passing_equal_sent"]
        if mixed_leg:
            out["mixed_stream"] = mixed_leg
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            build_kind, restore = cpu_arm(o)
            n = max(cores * 64, 256)
            sub = iqs[np.arange(n) % args.pool]
            o.ue_dl_decode_mt(ocell, ocfg, sub[:cores], cores, 0.01, args.noise_mode, args.max_iter)      # warm the table caches
            reps = 8                                            # about 10 s of CPU work in total (AVX2 build)
            ok = 0
            t0 = time.perf_counter()
            for _ in range(reps):
                ok_r, cpl, _ = o.ue_dl_decode_mt(ocell, ocfg, sub, cores, 0.01, args.noise_mode, args.max_iter)
                ok += ok_r
            dt = time.perf_counter() - t0
            restore()
            same = bool(np.array_equal(cpl[:args.pool], tbs[:args.pool]))
            out["cpu_baseline"] = {"value": ok * WORKLOAD["tbs"] / dt / 1e6, "unit": "Mbit/s", "cores": cores, "kind": "port",
                                   "sample": "%d x %d subframes of the same workload, %d threads, one subframe per thread, CPU restatement "
                                             "of the srsLTE path: %s" % (reps, n, cores, build_kind),
                                   "payload_equals_gpu_payload": same}
        print(json.dumps(out), flush=True)
    eplan.close()
    plan.close()
    del h_iq, h_pl, h_q
    lib.srsue_gpu_host_free(p_iq)
    lib.srsue_gpu_host_free(p_pl)
    lib.srsue_gpu_host_free(p_q)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
