"""The batching layer (srsue_gpu_batch_*): a heterogeneous stream of subframes -- BASELINE config 5, mixed
1.4-20 MHz bandwidths, modulations, transmission modes and subframe numbers in arrival order -- must come back
exactly as the oracle decodes each subframe on its own; HARQ soft buffers stay resident between submissions
(reference behaviour: one subframe per worker, phch_recv.cc:309-369; soft buffers per process, dl_harq.cc:169-259)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MIX = [  # prb, ports, qm, tbs, tm, sf_idx, cfi, snr
    (6, 1, 2, 152, 1, 1, 1, 10.0),
    (15, 1, 4, 2216, 1, 3, 2, 16.0),
    (25, 1, 6, 11448, 1, 2, 2, 24.0),
    (50, 2, 4, 6208, 2, 4, 1, 18.0),
    (100, 1, 6, 75376, 1, 1, 1, 30.0),
    (100, 2, 4, 30576, 2, 6, 1, 15.0),
    (75, 1, 6, 55056, 1, 7, 1, 30.0),
    (25, 1, 6, 11448, 1, 0, 2, 24.0),     # subframe 0: PSS/SSS/PBCH holes
]


def _pair(sg, o, row, rv=0):
    prb, ports, qm, tbs, tm, sf, cfi, snr = row
    ocell = o.make_cell(prb, ports, 1)
    ocfg = o.make_cfg(ocell, sf_idx=sf, cfi=cfi, qm=qm, tbs=tbs, tm=tm, rv=rv)
    cell = sg.make_cell(prb, ports, 1)
    cfg = sg.make_cfg(cell, sf_idx=sf, cfi=cfi, qm=qm, tbs=tbs, tm=tm, rv=rv)
    return ocell, ocfg, cell, cfg


def test_mixed_stream_matches_oracle(gpu, oracle):
    sg, ctx = gpu
    o = oracle
    rng = np.random.default_rng(5)
    order = [int(x) for x in rng.integers(0, len(MIX), 40)]
    items, refs = [], []
    for i, m in enumerate(order):
        ocell, ocfg, cell, cfg = _pair(sg, o, MIX[m])
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 50000 + i, MIX[m][7])
        items.append(dict(cell=cell, cfg=cfg, iq=iq))
        refs.append((ocell, ocfg, iq, tb))
    b = sg.Batch(ctx, 64)
    b.submit(items)
    res = b.wait()
    st = b.stats()
    assert st["plans"] == len(set(order)) and st["launches"] >= 5 * len(set(order))
    for r, (ocell, ocfg, iq, tb) in zip(res, refs):
        rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, iq, 0.01, 0, 4)
        assert (r["crc_ok"] == 1) == (rc == 0)
        assert np.array_equal(r["payload"], pl)
        assert r["n_iter"] == avg
        assert np.allclose(r["meas"], meas, rtol=1e-4)
        assert rc == 0 and np.array_equal(pl, tb)
    # a second submission reuses the cached plans
    b.submit(items[:7])
    res2 = b.wait()
    for r, r0 in zip(res2, res[:7]):
        assert np.array_equal(r["payload"], r0["payload"]) and r["crc_ok"] == r0["crc_ok"]
    assert b.stats()["plans"] == st["plans"]
    b.close()


def test_chunking_and_contiguous_merge(gpu, oracle):
    """more subframes of one shape than the chunk capacity, IQ rows adjacent in memory (merged copies)"""
    sg, ctx = gpu
    o = oracle
    ocell, ocfg, cell, cfg = _pair(sg, o, MIX[1])
    pool = np.stack([o.gen_subframe(ocell, ocfg, 61000 + i, 16.0)[1] for i in range(6)])
    n = 2500                                   # chunk capacity is 1024
    big = np.ascontiguousarray(pool[np.arange(n) % 6])
    items = [dict(cell=cell, cfg=cfg, iq=big[i]) for i in range(n)]
    b = sg.Batch(ctx, n)
    b.submit(items)
    res = b.wait()
    ref = [o.ue_dl_decode(ocell, ocfg, pool[i], 0.01, 0, 4) for i in range(6)]
    for i, r in enumerate(res):
        rc, pl, meas, avg = ref[i % 6]
        assert (r["crc_ok"] == 1) == (rc == 0) and np.array_equal(r["payload"], pl) and r["n_iter"] == avg
    b.close()


def test_resident_harq_softbuffers(gpu, oracle):
    """two HARQ processes interleaved with untracked traffic: rv 0 fails at low SNR, rv 2 combines into the
    device-resident soft buffer in a later submission and matches the oracle carrying its own soft buffer"""
    sg, ctx = gpu
    o = oracle
    row = (25, 1, 6, 11448, 1, 2, 2, 11.0)
    other = MIX[0]
    b = sg.Batch(ctx, 16)
    sb_o = {7: None, 9: None}
    verdicts = []
    for rv in (0, 2):
        items, refs = [], []
        for sid, seed in ((7, 42), (9, 43)):
            ocell, ocfg, cell, cfg = _pair(sg, o, row, rv)
            tb, iq, _ = o.gen_subframe(ocell, ocfg, seed, row[7])
            items.append(dict(cell=cell, cfg=cfg, iq=iq, softbuffer_id=sid, new_data=1 if rv == 0 else 0))
            refs.append((sid, ocell, ocfg, iq, tb))
            oc2, og2, c2, g2 = _pair(sg, o, other)
            items.append(dict(cell=c2, cfg=g2, iq=o.gen_subframe(oc2, og2, 70 + sid, other[7])[1]))
            refs.append(None)
        b.submit(items)
        res = b.wait()
        for r, ref in zip(res, refs):
            if ref is None:
                assert r["crc_ok"] == 1
                continue
            sid, ocell, ocfg, iq, tb = ref
            if sb_o[sid] is None:
                sb_o[sid] = o.new_softbuf(o.cbsegm(ocfg.tbs).C)
            sf_o = o.ofdm_rx(row[0], iq)
            ce_o, _ = o.chest(ocell, row[5], sf_o)
            rc, pl = o.pdsch_decode(ocell, ocfg, sf_o, ce_o, 0.01, 4, softbuf=sb_o[sid])
            assert (r["crc_ok"] == 1) == (rc == 0)
            assert np.array_equal(r["payload"], pl)
            verdicts.append((rv, rc == 0))
    assert b.stats()["softbuffers"] == 2
    assert all(ok for rv, ok in verdicts if rv == 2) and not all(ok for rv, ok in verdicts if rv == 0)
    b.release_softbuffer(7)
    assert b.stats()["softbuffers"] == 1
    b.close()


def test_batch_argument_errors(gpu, oracle):
    sg, ctx = gpu
    o = oracle
    ocell, ocfg, cell, cfg = _pair(sg, o, MIX[0])
    iq = o.gen_subframe(ocell, ocfg, 1, 10.0)[1]
    b = sg.Batch(ctx, 4)
    with pytest.raises(sg.GpuError):      # same HARQ process twice in one submission
        b.submit([dict(cell=cell, cfg=cfg, iq=iq, softbuffer_id=1), dict(cell=cell, cfg=cfg, iq=iq, softbuffer_id=1)])
    with pytest.raises(sg.GpuError):      # combining needs an earlier transmission
        b.submit([dict(cell=cell, cfg=cfg, iq=iq, softbuffer_id=2, new_data=0)])
    with pytest.raises(sg.GpuError):      # more than max_subframes
        b.submit([dict(cell=cell, cfg=cfg, iq=iq)] * 5)
    b.submit([dict(cell=cell, cfg=cfg, iq=iq)])
    assert b.wait()[0]["crc_ok"] == 1
    b.close()


def test_batch_is_validated_before_anything_is_launched(gpu, oracle):
    """a descriptor that must be refused anywhere in the submission -- a combine without an earlier transmission behind
    valid buckets, a later untracked descriptor with new_data = 0 inside a bucket, a carrier offset out of range -- fails
    the whole submission before the first launch: nothing is pending, no payload buffer is written, the batch stays usable"""
    sg, ctx = gpu
    o = oracle
    ocell, ocfg, cell, cfg = _pair(sg, o, MIX[0])
    oc2, og2, c2, g2 = _pair(sg, o, MIX[1])
    iq = o.gen_subframe(ocell, ocfg, 1, 10.0)[1]
    iq2 = o.gen_subframe(oc2, og2, 2, 16.0)[1]
    b = sg.Batch(ctx, 8)
    good = [dict(cell=cell, cfg=cfg, iq=iq), dict(cell=c2, cfg=g2, iq=iq2)]
    for bad in (dict(cell=c2, cfg=g2, iq=iq2, softbuffer_id=5, new_data=0),      # no earlier transmission
                dict(cell=cell, cfg=cfg, iq=iq, new_data=0),                       # same bucket as good[0], not its first entry
                dict(cell=cell, cfg=cfg, iq=iq, cfo=1.5)):
        with pytest.raises(sg.GpuError):
            b.submit(good + [bad])
        assert b.wait() == []                                                      # nothing was accepted
        assert b.stats()["launches"] == 0
        b.submit(good)
        assert [r["crc_ok"] for r in b.wait()] == [1, 1]
    b.close()


@pytest.mark.parametrize("how", ["host_alloc", "host_register"])
def test_zero_copy_pinned_rows(gpu, oracle, how):
    """scattered subframes inside pinned regions known to the library are fetched / written by the GPU directly
    (gather / scatter kernels over UVA); results must not depend on the transfer path"""
    import ctypes as C
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    L.srsue_gpu_host_register.argtypes = [C.c_void_p, C.c_uint64]
    L.srsue_gpu_host_unregister.argtypes = [C.c_void_p]
    n = 12
    for row in (MIX[0], MIX[2]):           # 19-byte payloads (byte path) and 1431-byte payloads
        ocell, ocfg, cell, cfg = _pair(sg, o, row)
        gen = [o.gen_subframe(ocell, ocfg, 81000 + i, row[7]) for i in range(n)]
        sf_len, pb = len(gen[0][1]), (row[3] + 7) // 8
        pbp = (pb + 3) // 4 * 4 if row is MIX[2] else pb
        if how == "host_alloc":
            p, q = L.srsue_gpu_host_alloc(n * sf_len * 8), L.srsue_gpu_host_alloc(n * pbp)
            h_iq = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n, sf_len * 2)).view(np.complex64)
            h_pl = np.ctypeslib.as_array(C.cast(q, C.POINTER(C.c_uint8)), shape=(n, pbp))
        else:
            h_iq = np.zeros((n, sf_len), np.complex64)
            h_pl = np.zeros((n, pbp), np.uint8)
            assert L.srsue_gpu_host_register(h_iq.ctypes.data, h_iq.nbytes) == 0
            assert L.srsue_gpu_host_register(h_pl.ctypes.data, h_pl.nbytes) == 0
        h_pl[:] = 0xEE
        for i in range(n):
            h_iq[i] = gen[i][1]
        rev = list(range(n))[::-1]          # no two consecutive descriptors are adjacent in memory
        items = [dict(cell=cell, cfg=cfg, iq=h_iq[i], payload=h_pl[i, :pb]) for i in rev]
        b = sg.Batch(ctx, n)
        b.submit(items)
        res = b.wait()
        assert b.stats()["launches"] >= 7    # chain + gather + scatter
        for i, r in zip(rev, res):
            assert r["crc_ok"] == 1 and np.array_equal(h_pl[i, :pb], gen[i][0])
            assert (h_pl[i, pb:] == 0xEE).all()
        b.close()
        if how == "host_alloc":
            L.srsue_gpu_host_free(p)
            L.srsue_gpu_host_free(q)
        else:
            assert L.srsue_gpu_host_unregister(h_iq.ctypes.data) == 0
            assert L.srsue_gpu_host_unregister(h_pl.ctypes.data) == 0


def test_plan_cache_eviction(gpu, oracle):
    """more distinct launch shapes in one submission than the plan cache holds (16): least recently used plans are
    destroyed and rebuilt without affecting results"""
    sg, ctx = gpu
    o = oracle
    items, refs = [], []
    for sf in range(10):
        for cfi in (1, 2):
            ocell = o.make_cell(6, 1, 3)
            ocfg = o.make_cfg(ocell, sf_idx=sf, cfi=cfi, qm=2, tbs=104)
            cell = sg.make_cell(6, 1, 3)
            cfg = sg.make_cfg(cell, sf_idx=sf, cfi=cfi, qm=2, tbs=104)
            tb, iq, _ = o.gen_subframe(ocell, ocfg, 9000 + 10 * sf + cfi, 12.0)
            items.append(dict(cell=cell, cfg=cfg, iq=iq))
            refs.append(tb)
    b = sg.Batch(ctx, 64)
    for _ in range(2):                       # the second pass re-creates the evicted plans
        b.submit(items)
        res = b.wait()
        assert b.stats()["plans"] == 16
        for r, tb in zip(res, refs):
            assert r["crc_ok"] == 1 and np.array_equal(r["payload"], tb)
    b.close()


def test_mixed_stream_with_carrier_offsets(gpu, oracle):
    """Captures of different UEs carry different carrier offsets: the per-descriptor cfo is removed while each row is
    transformed (SPEC.md 14) and the result equals the oracle's rotate-then-decode, row by row; rows with cfo 0 in the
    same launch stay untouched."""
    sg, ctx = gpu
    o = oracle
    rng = np.random.default_rng(77)
    order = [int(x) for x in rng.integers(0, len(MIX), 24)]
    items, refs = [], []
    for i, m in enumerate(order):
        ocell, ocfg, cell, cfg = _pair(sg, o, MIX[m])
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 61000 + i, MIX[m][7] + 6.0)
        n = o.lib().lteo_symbol_sz(MIX[m][0])
        cfo = 0.0 if i % 3 == 0 else float(np.float32(rng.uniform(-0.4, 0.4)))
        rx = (iq.astype(np.complex128) * np.exp(2j * np.pi * cfo * np.arange(len(iq)) / n)).astype(np.complex64)
        items.append(dict(cell=cell, cfg=cfg, iq=rx, cfo=cfo))
        refs.append((ocell, ocfg, rx, tb, o.cfo_step(cfo, n)))
    b = sg.Batch(ctx, 32)
    b.submit(items)
    res = b.wait()
    for r, (ocell, ocfg, rx, tb, step) in zip(res, refs):
        x = o.cfo_correct(rx, step) if step else rx
        rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, x, 0.01, 0, 4)
        assert (r["crc_ok"] == 1) == (rc == 0) and np.array_equal(r["payload"], pl) and r["n_iter"] == avg
        assert np.allclose(r["meas"], meas, rtol=1e-4)
        assert rc == 0 and np.array_equal(pl, tb)
    items[3]["cfo"] = 1.5
    with pytest.raises(sg.GpuError):
        b.submit(items[:5])
    b.close()


def test_mixed_stream_of_int16_captures(gpu, oracle):
    """srsue_gpu_batch_set_iq_format(SC16): every descriptor points at int16 pairs (one common scale); results equal the
    oracle on the samples converted to float first, with and without a carrier offset, scattered and adjacent rows."""
    sg, ctx = gpu
    o = oracle
    rng = np.random.default_rng(99)
    order = [int(x) for x in rng.integers(0, len(MIX), 20)] + [4, 4, 4]
    scale = np.float32(4.0 / 32000.0)
    items, refs = [], []
    for i, m in enumerate(order):
        ocell, ocfg, cell, cfg = _pair(sg, o, MIX[m])
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 71000 + i, MIX[m][7] + 6.0)
        n = o.lib().lteo_symbol_sz(MIX[m][0])
        cfo = 0.0 if i % 2 == 0 else float(np.float32(rng.uniform(-0.3, 0.3)))
        rx = iq.astype(np.complex128) * np.exp(2j * np.pi * cfo * np.arange(len(iq)) / n)
        q16 = np.rint(rx.astype(np.complex64).view(np.float32) / scale).clip(-32768, 32767).astype(np.int16)
        items.append(dict(cell=cell, cfg=cfg, iq=q16, cfo=cfo))
        refs.append((ocell, ocfg, (q16.astype(np.float32) * scale).view(np.complex64), tb, o.cfo_step(cfo, n)))
    # the last three rows adjacent in host memory (one upload)
    blk = np.concatenate([it["iq"] for it in items[-3:]])
    for j, it in enumerate(items[-3:]):
        it["iq"] = blk[j * len(blk) // 3:(j + 1) * len(blk) // 3]
    b = sg.Batch(ctx, 32)
    b.set_iq_format(True, float(scale))
    b.submit(items)
    res = b.wait()
    for r, (ocell, ocfg, x, tb, step) in zip(res, refs):
        xr = o.cfo_correct(x, step) if step else x
        rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, xr, 0.01, 0, 4)
        assert (r["crc_ok"] == 1) == (rc == 0) and np.array_equal(r["payload"], pl) and r["n_iter"] == avg
        assert rc == 0 and np.array_equal(pl, tb)
    b.set_iq_format(False)
    ocell, ocfg, cell, cfg = _pair(sg, o, MIX[2])
    tb, iq, _ = o.gen_subframe(ocell, ocfg, 72000, 30.0)
    b.submit([dict(cell=cell, cfg=cfg, iq=iq)])
    assert np.array_equal(b.wait()[0]["payload"], tb)
    b.close()


def _multi_stream(sg, o, n_items, with_harq):
    items, refs = [], []
    for i in range(n_items):
        row = MIX[i % len(MIX)]
        ocell, ocfg, cell, cfg = _pair(sg, o, row)
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 8100 + i, row[7])
        it = dict(cell=cell, cfg=cfg, iq=iq)
        if with_harq and i % 3 == 0:
            it.update(softbuffer_id=1000 + i, new_data=1)
        items.append(it)
        refs.append(tb)
    return items, refs


def test_multi_gpu_handle_on_one_device(gpu, oracle):
    """srsue_gpu_batch_create_multi with a single device: the same submit / wait / stats calls, the library owns the context
    and the host thread; results as the single-device batch gives them; errors of a share surface in batch_wait"""
    sg, ctx = gpu
    o = oracle
    items, refs = _multi_stream(sg, o, 24, True)
    b = sg.Batch(None, 64, devices=[0])
    for _ in range(2):
        b.submit(items)
        res = b.wait()
        assert all(r["crc_ok"] == 1 and np.array_equal(r["payload"], t) for r, t in zip(res, refs))
    assert b.device_shares()[0][0] == len(items) and b.stats()["softbuffers"] == 8
    b.release_softbuffer(1000)
    assert b.stats()["softbuffers"] == 7
    ocell, ocfg, cell, cfg = _pair(sg, o, MIX[0])
    bad = dict(cell=cell, cfg=cfg, iq=items[0]["iq"], softbuffer_id=77, new_data=0)
    b.submit([items[0], bad])
    with pytest.raises(sg.GpuError):
        b.wait()
    b.submit(items[:4])
    assert [r["crc_ok"] for r in b.wait()] == [1, 1, 1, 1]
    b.close()


def test_multi_gpu_dispatch_pins_harq_processes(gpu, oracle):
    """two GPUs behind one handle: the submission is split by estimated turbo work, every HARQ process stays on the device
    that first saw it (rv 2 finds the soft buffer rv 0 left there), transport blocks equal the oracle's"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sg, ctx = gpu
    o = oracle
    row = (25, 1, 6, 11448, 1, 2, 2, 11.0)
    b = sg.Batch(None, 64, devices=[0, 1])
    filler, frefs = _multi_stream(sg, o, 16, False)
    sb_o, verdicts = {}, []
    for rv in (0, 2):
        items, refs = [], []
        for k, sid in enumerate((7, 9, 11, 13)):
            ocell, ocfg, cell, cfg = _pair(sg, o, row, rv)
            tb, iq, _ = o.gen_subframe(ocell, ocfg, 420 + k, row[7])
            items.append(dict(cell=cell, cfg=cfg, iq=iq, softbuffer_id=sid, new_data=1 if rv == 0 else 0))
            refs.append((sid, ocell, ocfg, iq))
            items += filler[4 * k:4 * k + 4]
            refs += [None] * 4
        b.submit(items)
        res = b.wait()
        shares = b.device_shares()
        assert len(shares) == 2 and all(n > 0 for n, _ in shares) and sum(n for n, _ in shares) == len(items)
        assert max(w for _, w in shares) < 1.6 * min(w for _, w in shares)
        fi = 0
        for r, ref in zip(res, refs):
            if ref is None:
                assert r["crc_ok"] == 1 and np.array_equal(r["payload"], frefs[fi])
                fi += 1
                continue
            sid, ocell, ocfg, iq = ref
            sb_o.setdefault(sid, o.new_softbuf(o.cbsegm(ocfg.tbs).C))
            sf_o = o.ofdm_rx(row[0], iq)
            ce_o, _ = o.chest(ocell, row[5], sf_o)
            rc, pl = o.pdsch_decode(ocell, ocfg, sf_o, ce_o, 0.01, 4, softbuf=sb_o[sid])
            assert (r["crc_ok"] == 1) == (rc == 0) and np.array_equal(r["payload"], pl)
            verdicts.append((rv, rc == 0))
    assert all(ok for rv, ok in verdicts if rv == 2) and not all(ok for rv, ok in verdicts if rv == 0)
    assert b.stats()["softbuffers"] == 4
    b.close()


def test_blind_submission_finds_every_grant(gpu, oracle):
    """srsue_gpu_batch_submit_blind: a mixed stream where the caller knows only cell, subframe number and RNTI of each capture.
    The library decodes the PCFICH, searches the PDCCH for the RNTI, turns the DCI into a grant and decodes the PDSCH --
    phch_worker.cc:254-297 for a whole batch.  CFI, grant and transport block equal what was sent and what the oracle
    decodes with the true grant; captures without a DCI for the RNTI come back empty."""
    import ctypes as C
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import DciMsg, RaDlDci, install_tbs_table
    # (prb, cfi, sf_idx, rnti, mcs, RB_start, L_crb, tbs, with_dci)
    cases = [(25, 2, 4, 0x1234, 9, 2, 10, 1544, True), (25, 1, 4, 0x1234, 9, 0, 10, 1544, True), (25, 2, 4, 0x1234, 9, 2, 10, 1544, False),
             (50, 1, 7, 0x0456, 12, 5, 25, 5736, True), (50, 3, 7, 0x0456, 3, 0, 50, 2856, True), (6, 2, 1, 0x1234, 4, 0, 6, 408, True),
             (25, 3, 5, 0xFFFF, 5, 1, 4, 296, True), (100, 1, 2, 0x2222, 20, 0, 100, 46888, True), (100, 1, 2, 0x2222, 20, 0, 100, 46888, False)]
    table = {}
    for prb, cfi, sf, rnti, mcs, start, ln, tbs, with_dci in cases:
        itbs = mcs if (mcs < 10 or rnti == 0xFFFF) else mcs - 1 if mcs < 17 else mcs - 2      # 36.213 Table 7.1.7.1-1
        table[(itbs, 3 if rnti == 0xFFFF else ln)] = tbs
    install_tbs_table(L, table)
    items, truth = [], []
    for k, (prb, cfi, sf, rnti, mcs, start, ln, tbs, with_dci) in enumerate(cases):
        si = rnti == 0xFFFF
        ocell = o.make_cell(prb, 1, 1)
        sent = RaDlDci()
        sent.mcs_idx, sent.rv_idx, sent.alloc_type = mcs, 0, 2
        sent.type2_alloc.RB_start, sent.type2_alloc.L_crb, sent.type2_alloc.n_prb1a = start, ln, 1
        m = DciMsg()
        nb = L.srslte_dci_msg_pack_pdsch(C.byref(sent), 2, C.byref(m), prb, not si)
        assert nb > 0
        rk, _ = o.pdcch_regs(ocell, cfi, 6)
        ss = o.pdcch_search_space(len(rk) // 9, sf, rnti) if not si else [(4, 0)]
        qm = 2 if mcs < 10 else 4 if mcs < 17 else 6
        prbs = list(range(start, start + ln))
        ocfg = o.make_cfg(ocell, sf_idx=sf, cfi=cfi, rnti=rnti, qm=qm, tbs=tbs, prbs=prbs)
        dcis = [(np.frombuffer(m.data, np.uint8)[:nb].copy(), rnti, ss[0][0], ss[0][1])] if with_dci else None
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 9300 + k, 26.0 if qm == 6 else 16.0, None, pcfich=True, dcis=dcis)
        cell = sg.make_cell(prb, 1, 1)
        items.append(dict(cell=cell, cfg=sg.make_cfg(cell, sf_idx=sf, cfi=1, rnti=rnti, qm=2, tbs=0), iq=iq))
        truth.append((ocell, ocfg, iq, tb, cfi, tbs, qm, prbs, with_dci))
    b = sg.Batch(ctx, 32)
    for rep in range(2):
        b.submit_blind(items, ng_x6=6)
        res = b.wait()
        for r, (ocell, ocfg, iq, tb, cfi, tbs, qm, prbs, with_dci) in zip(res, truth):
            assert r["cfi"] == cfi
            if not with_dci:
                assert r["tbs"] == 0 and r["crc_ok"] == 0
                continue
            assert r["tbs"] == tbs and r["cfg"].qm == qm and r["cfg"].rv == 0
            assert [i for i in range(ocell.nof_prb) if r["cfg"].prb_mask[i]] == prbs
            rc_o, pl_o, _, _ = o.ue_dl_decode(ocell, ocfg, iq, 0.01, 0, 4)
            assert rc_o == 0 and r["crc_ok"] == 1 and np.array_equal(r["payload"], pl_o) and np.array_equal(pl_o, tb)
    assert b.stats()["launches"] > 0
    b.close()

