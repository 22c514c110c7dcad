"""Extended cyclic prefix on the device (SPEC.md 15b): OFDM demodulation, channel estimate, equaliser / demapper / rate
de-matcher, the whole PDSCH chain, PCFICH + PDCCH search, the batching layer and the srsLTE-shaped worker sequence on
cells with cp = SRSLTE_CP_EXT -- every stage against the CPU oracle, floats bit-identical, integers bit-exact.
The cell's prefix reaches the worker at /root/reference/ue/src/phy/phch_worker.cc:74 (srslte_ue_dl_init(&ue_dl, cell))
after cell search reported it (phch_recv.cc:189)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _taps():
    rng = np.random.default_rng(78)
    taps = (rng.standard_normal((2, 5)) + 1j * rng.standard_normal((2, 5))) * np.array([1, .6, .4, .2, .1])
    return taps / np.sqrt((abs(taps) ** 2).sum(1, keepdims=True))


STAGES = {
    "bw6": dict(prb=6, ports=1, qm=2, tbs=152, tm=1, snr=10.0, taps=False, sf=1, cfi=3),
    "bw15_sf0": dict(prb=15, ports=1, qm=2, tbs=600, tm=1, snr=9.0, taps=False, sf=0, cfi=2),
    "bw25_2p_sf5": dict(prb=25, ports=2, qm=4, tbs=4008, tm=2, snr=18.0, taps=True, sf=5, cfi=2),
    "bw50": dict(prb=50, ports=1, qm=6, tbs=15264, tm=1, snr=28.0, taps=False, sf=3, cfi=1),
    "bw75_2p": dict(prb=75, ports=2, qm=4, tbs=15264, tm=2, snr=16.0, taps=True, sf=4, cfi=1),     # 1536-point FFT
    "bw100": dict(prb=100, ports=1, qm=6, tbs=61664, tm=1, snr=30.0, taps=False, sf=1, cfi=1),   # radix-16 kernel
    "bw100_2p_sf0": dict(prb=100, ports=2, qm=6, tbs=36696, tm=2, snr=30.0, taps=True, sf=0, cfi=2),
}


@pytest.mark.parametrize("name", list(STAGES))
def test_frontend_stages_match_oracle_extended_prefix(gpu, oracle, name):
    import torch
    sg, ctx = gpu
    o = oracle
    c = STAGES[name]
    taps = _taps() if c["taps"] else None
    ocell = o.make_cell(c["prb"], c["ports"], 5, cp=1)
    ocfg = o.make_cfg(ocell, sf_idx=c["sf"], cfi=c["cfi"], qm=c["qm"], tbs=c["tbs"], tm=c["tm"])
    n_sf = 3
    iq = np.stack([o.gen_subframe(ocell, ocfg, 7000 + i, c["snr"], taps)[1] for i in range(n_sf)])
    cell = sg.make_cell(c["prb"], c["ports"], 5, cp=1)
    cfg = sg.make_cfg(cell, sf_idx=c["sf"], cfi=c["cfi"], qm=c["qm"], tbs=c["tbs"], tm=c["tm"])
    plan = sg.PdschPlan(ctx, cell, cfg, n_sf)
    I = plan.info
    assert I.nof_re == len(o.pdsch_re_list(ocell, ocfg))
    nsc = I.nsc
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(n_sf, -1)).cuda()
    d_sf = torch.full((n_sf, 14 * nsc * 2), 7.0, dtype=torch.float32, device="cuda")
    d_ce = torch.full((n_sf, c["ports"] * 14 * nsc * 2), 7.0, dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((n_sf, 5), dtype=torch.float32, device="cuda")
    d_sb = torch.zeros((n_sf, I.sb_sf_stride), dtype=torch.int16, device="cuda")
    d_d = torch.zeros((n_sf, I.nof_re * 2), dtype=torch.float32, device="cuda")
    d_e = torch.zeros((n_sf, I.G), dtype=torch.int16, device="cuda")
    plan.ofdm_rx(n_sf, d_iq, d_sf)
    plan.chest(n_sf, d_sf, d_ce, d_meas)
    plan.pdsch_llr(n_sf, d_sf, d_ce, d_meas, 0.01, 0, 0, d_sb, d_d, d_e)
    torch.cuda.synchronize()
    sf_g = d_sf.cpu().numpy().view(np.complex64).reshape(n_sf, 14, nsc)
    ce_g = d_ce.cpu().numpy().view(np.complex64).reshape(n_sf, c["ports"], 14, nsc)
    # the grids keep their 14-symbol stride; rows 12 and 13 are not touched
    assert (sf_g[:, 12:] == 7.0 + 7.0j).all() and (ce_g[:, :, 12:] == 7.0 + 7.0j).all()
    meas_g = d_meas.cpu().numpy()
    dd_g = d_d.cpu().numpy().view(np.complex64)
    e_g = d_e.cpu().numpy()
    s = o.cbsegm(c["tbs"])
    for i in range(n_sf):
        sf_o = o.ofdm_rx(c["prb"], iq[i], cp=1)
        assert np.array_equal(sf_g[i, :12].reshape(-1), sf_o[:12 * nsc]), "FFT bins are expected to be bit-identical"
        ce_o, meas_o = o.chest(ocell, c["sf"], sf_o)
        assert np.array_equal(ce_g[i, :, :12], ce_o.reshape(c["ports"], 14, nsc)[:, :12]), "channel estimate"
        assert np.allclose(meas_g[i], meas_o, rtol=1e-4)
        rc, pl, dbg = o.pdsch_decode(ocell, ocfg, sf_o, ce_o, 0.01, 4, want=True)
        assert np.array_equal(dd_g[i], dbg["d"][:I.nof_re]), "equalised symbols"
        assert np.array_equal(e_g[i], dbg["e"][:I.G]), "descrambled int16 LLRs differ"
        for r in range(s.C):
            K = o.cb_len(s, r)
            t = torch.zeros(3 * K + 12, dtype=torch.int16, device="cuda")
            ctx.tdec_export(d_sb[i, r * I.sb_cb_stride:], 1, K, t)
            torch.cuda.synchronize()
            assert np.array_equal(t.cpu().numpy(), dbg["softbuf"][r, :3 * K + 12]), "soft buffer cb %d" % r
    plan.close()


CHAIN = {
    "ext_1.4MHz": dict(prb=6, ports=1, qm=2, tbs=152, tm=1, snr=10.0, taps=False, n=4, noise_mode=1, sf=1),
    "ext_20MHz_64qam": dict(prb=100, ports=1, qm=6, tbs=61664, tm=1, snr=30.0, taps=False, n=4, noise_mode=0, sf=1),
    "ext_20MHz_waterfall": dict(prb=100, ports=1, qm=6, tbs=61664, tm=1, snr=21.5, taps=False, n=5, noise_mode=0, sf=2),
    "ext_tm2_sf0": dict(prb=100, ports=2, qm=4, tbs=22920, tm=2, snr=15.0, taps=True, n=3, noise_mode=1, sf=0),
    "ext_15MHz_sf5": dict(prb=75, ports=1, qm=6, tbs=37888, tm=1, snr=30.0, taps=False, n=3, noise_mode=0, sf=5),
}


@pytest.mark.parametrize("name", list(CHAIN))
def test_chain_matches_oracle_extended_prefix(gpu, oracle, name):
    import torch
    sg, ctx = gpu
    o = oracle
    c = CHAIN[name]
    taps = _taps() if c["taps"] else None
    ocell = o.make_cell(c["prb"], c["ports"], 1, cp=1)
    ocfg = o.make_cfg(ocell, sf_idx=c["sf"], cfi=1, qm=c["qm"], tbs=c["tbs"], tm=c["tm"])
    n = c["n"]
    gen = [o.gen_subframe(ocell, ocfg, 21000 + i, c["snr"], taps) for i in range(n)]
    iq = np.stack([g[1] for g in gen])
    cell = sg.make_cell(c["prb"], c["ports"], 1, cp=1)
    cfg = sg.make_cfg(cell, sf_idx=c["sf"], cfi=1, qm=c["qm"], tbs=c["tbs"], tm=c["tm"])
    plan = sg.PdschPlan(ctx, cell, cfg, n)
    I = plan.info
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(n, -1)).cuda()
    d_pl = torch.zeros((n, I.payload_stride), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    d_meas = torch.zeros((n, 5), dtype=torch.float32, device="cuda")
    plan.decode_batch(n, d_iq, 0.01, c["noise_mode"], 4, d_pl, d_st, d_meas=d_meas)
    torch.cuda.synchronize()
    pl_g, st_g = d_pl.cpu().numpy(), d_st.cpu().numpy()
    h_pl = np.zeros((n, I.payload_stride), np.uint8)
    h_st = np.zeros((n, 4), np.int32)
    h_meas = np.zeros((n, 5), np.float32)
    plan.decode_batch_host(n, iq, 0.01, c["noise_mode"], 4, h_pl, h_st, h_meas)
    assert np.array_equal(h_pl, pl_g) and np.array_equal(h_st, st_g)
    n_ok = 0
    for i in range(n):
        rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, iq[i], 0.01, c["noise_mode"], 4)
        assert (st_g[i, 0] == 1) == (rc == 0), "CRC verdict differs (sf %d)" % i
        assert np.array_equal(pl_g[i], pl), "transport block differs from oracle (sf %d)" % i
        assert st_g[i, 2] == avg
        assert np.allclose(h_meas[i], meas, rtol=1e-4)
        if rc == 0:
            n_ok += 1
            assert np.array_equal(pl, gen[i][0])
    if "waterfall" not in name:
        assert n_ok == n
    plan.close()


@pytest.mark.parametrize("prb,ports,cid,cfi", [(6, 1, 1, 3), (6, 2, 8, 3), (25, 2, 77, 2), (100, 1, 503, 1)])
def test_pcfich_and_pdcch_match_oracle_extended_prefix(gpu, oracle, prb, ports, cid, cfi):
    """the control region of an extended-prefix cell (four symbols with CRS in the fourth at 1.4 MHz and CFI 3): CFI,
    the LLRs of every control-channel element and the blind-search verdict equal the oracle's"""
    import torch
    sg, ctx = gpu
    o = oracle
    ocell = o.make_cell(prb, ports, cid, cp=1)
    cell = sg.make_cell(prb, ports, cid, cp=1)
    sf_idx, rnti = (cid + cfi) % 10, 0x1234 + cid
    nb = sg.lib().srsue_gpu_host_dci_format_sizeof(0, prb)
    rk, _ = o.pdcch_regs(ocell, cfi, 6)
    ncce = len(rk) // 9
    ss = o.pdcch_search_space(ncce, sf_idx, rnti)
    rng = np.random.default_rng(cid)
    n = 4
    iq, sent = [], []
    for i in range(n):
        L0, n0 = ss[(3 * i + 1) % len(ss)]
        bits = rng.integers(0, 2, nb, dtype=np.uint8)
        dcis = [(bits, rnti, L0, n0)] if i != 2 else []
        ocfg = o.make_cfg(ocell, sf_idx=sf_idx, cfi=cfi, qm=2, tbs=104 if prb == 6 else 1000, tm=ports)
        iq.append(o.gen_subframe(ocell, ocfg, 4100 + i, 9.0, None, pcfich=True, dcis=dcis)[1])
        sent.append(dcis)
    iq = np.stack(iq)
    cfg = sg.make_cfg(cell, sf_idx=sf_idx, cfi=cfi, qm=2, tbs=0, tm=ports)
    plan = sg.PdschPlan(ctx, cell, cfg, n)
    I = plan.info
    n_reg, nc = plan.pdcch_info(6)
    assert n_reg == len(rk) and nc == ncce
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(n, -1)).cuda()
    d_sf = torch.zeros((n, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((n, ports * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((n, 5), dtype=torch.float32, device="cuda")
    d_llr = torch.zeros((n, 8 * n_reg), dtype=torch.int16, device="cuda")
    d_found = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    d_bits = torch.zeros((n, 64), dtype=torch.uint8, device="cuda")
    d_cfi = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_corr = torch.zeros((n, 3), dtype=torch.int32, device="cuda")
    plan.ofdm_rx(n, d_iq, d_sf)
    plan.chest(n, d_sf, d_ce, d_meas)
    plan.pcfich_decode(n, d_sf, d_ce, d_meas, 0.0, 1, d_cfi, d_corr)
    plan.pdcch_extract_llr(n, d_sf, d_ce, d_meas, 0.0, 1, d_llr)
    ncand = plan.pdcch_find_dci(n, d_llr, rnti, nb, d_found, d_bits, None)
    torch.cuda.synchronize()
    assert ncand == len(ss)
    llr_g, found, bits_g = d_llr.cpu().numpy(), d_found.cpu().numpy(), d_bits.cpu().numpy()
    for i in range(n):
        sf_o = o.ofdm_rx(prb, iq[i], cp=1)
        ce_o, meas_o = o.chest(ocell, sf_idx, sf_o)
        cfi_o, corr_o = o.pcfich_decode(ocell, sf_idx, sf_o, ce_o, meas_o[0])
        assert int(d_cfi[i]) == cfi_o == cfi and np.array_equal(d_corr[i].cpu().numpy(), corr_o)
        llr_o, _ = o.pdcch_extract_llr(ocell, sf_idx, cfi, sf_o, ce_o, meas_o[0])
        assert np.array_equal(llr_g[i], llr_o[:8 * n_reg])
        f, out, L1, n1 = o.pdcch_find_dci(llr_o, ncce, sf_idx, rnti, nb)
        assert found[i, 0] == f == (1 if sent[i] else 0)
        if f:
            assert (found[i, 1], found[i, 2]) == (L1, n1) and np.array_equal(bits_g[i, :nb], out) and np.array_equal(out, sent[i][0][0])
    plan.close()


def test_batch_layer_mixes_both_prefixes(gpu, oracle):
    """one submission with normal- and extended-prefix cells of the same bandwidth, id and grant: two plans, every
    transport block as the oracle decodes it"""
    sg, ctx = gpu
    o = oracle
    rows = [(25, 1, 4, 4008, 1, 2, 2, 20.0, 0), (25, 1, 4, 4008, 1, 2, 2, 20.0, 1), (100, 2, 4, 22920, 2, 0, 1, 16.0, 1),
            (6, 1, 2, 152, 1, 5, 3, 12.0, 1), (100, 1, 6, 61664, 1, 7, 1, 30.0, 1)]
    rng = np.random.default_rng(9)
    order = [int(x) for x in rng.integers(0, len(rows), 24)]
    items, refs = [], []
    for i, m in enumerate(order):
        prb, ports, qm, tbs, tm, sf, cfi, snr, cp = rows[m]
        ocell = o.make_cell(prb, ports, 1, cp=cp)
        ocfg = o.make_cfg(ocell, sf_idx=sf, cfi=cfi, qm=qm, tbs=tbs, tm=tm)
        cell = sg.make_cell(prb, ports, 1, cp=cp)
        cfg = sg.make_cfg(cell, sf_idx=sf, cfi=cfi, qm=qm, tbs=tbs, tm=tm)
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 52000 + i, snr)
        items.append(dict(cell=cell, cfg=cfg, iq=iq))
        refs.append((ocell, ocfg, iq, tb))
    b = sg.Batch(ctx, 32)
    b.submit(items)
    res = b.wait()
    assert b.stats()["plans"] == len(set(order))
    for r, (ocell, ocfg, iq, tb) in zip(res, refs):
        rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, iq, 0.01, 0, 4)
        assert (r["crc_ok"] == 1) == (rc == 0) and np.array_equal(r["payload"], pl) and r["n_iter"] == avg
        assert rc == 0 and np.array_equal(pl, tb)
    b.close()


def test_srslte_worker_sequence_extended_prefix(gpu, oracle):
    """phch_worker's sequence through the srsLTE-shaped symbols on a cell with cp = SRSLTE_CP_EXT: init (phch_worker.cc:74),
    FFT + estimate (:254), grant configuration (:337), PDSCH decode (:347) -- host grids of 12 symbols, oracle payload"""
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import UeDl, Cell, SoftBuffer, make_grant
    prb, qm, tbs = 25, 4, 4008
    ocell = o.make_cell(prb, 1, 1, cp=1)
    ocfg = o.make_cfg(ocell, sf_idx=3, cfi=2, qm=qm, tbs=tbs)
    tb, iq, _ = o.gen_subframe(ocell, ocfg, 123, 20.0, pcfich=True)
    q = UeDl()
    cell = Cell(nof_prb=prb, nof_ports=1, bw_idx=0, id=1, cp=1, phich_length=0, phich_resources=2)
    assert L.srslte_ue_dl_init(C.byref(q), cell) == 0
    L.srslte_ue_dl_set_rnti(C.byref(q), 0x1234)
    L.srslte_sch_set_max_noi(C.byref(q.pdsch.dl_sch), 4)
    sb = SoftBuffer()
    assert L.srslte_softbuffer_rx_init(C.byref(sb), prb) == 0
    L.srslte_softbuffer_rx_reset(C.byref(sb))
    cfi = C.c_uint32(0)
    assert L.srslte_ue_dl_decode_fft_estimate(C.byref(q), iq.ctypes.data_as(C.c_void_p), 3, C.byref(cfi)) == 0
    assert cfi.value == 2
    nsc = 12 * prb
    sf_o = o.ofdm_rx(prb, iq, cp=1)
    ce_o, meas_o = o.chest(ocell, 3, sf_o)
    sf_h = np.ctypeslib.as_array(C.cast(q.sf_symbols, C.POINTER(C.c_float)), shape=(12 * nsc * 2,)).view(np.complex64)
    ce_h = np.ctypeslib.as_array(C.cast(q.ce[0], C.POINTER(C.c_float)), shape=(12 * nsc * 2,)).view(np.complex64)
    assert np.array_equal(sf_h, sf_o[:12 * nsc]) and np.array_equal(ce_h, ce_o[0][:12 * nsc])
    grant = make_grant(prb, qm, tbs)
    assert L.srslte_ue_dl_cfg_grant(C.byref(q), C.byref(grant), cfi.value, 3, 0) == 0
    assert q.pdsch_cfg.nbits.nof_re == len(o.pdsch_re_list(ocell, ocfg)) and q.pdsch_cfg.nbits.nof_symb == 12 - 2
    payload = np.zeros(tbs // 8, np.uint8)
    ret = L.srslte_pdsch_decode_rnti(C.byref(q.pdsch), C.byref(q.pdsch_cfg), C.byref(sb), C.c_void_p(q.sf_symbols), q.ce,
                                     C.c_float(0.01), C.c_uint16(0x1234), payload.ctypes.data_as(C.c_void_p))
    rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, iq, 0.01, 0, 4)
    assert ret == 0 and rc == 0
    assert np.array_equal(payload, pl) and np.array_equal(payload, tb)
    assert L.srslte_pdsch_last_noi(C.byref(q.pdsch)) == avg
    L.srslte_softbuffer_rx_free(C.byref(sb))
    L.srslte_ue_dl_free(C.byref(q))


@pytest.mark.parametrize("prb,ports,cid", [(6, 1, 1), (25, 2, 77), (50, 1, 301), (100, 2, 0)])
def test_phich_matches_oracle_extended_prefix(gpu, oracle, prb, ports, cid):
    """spreading factor 2 (36.211 Table 6.9.1-2): two groups share a mapping unit, four sequences per group; decision
    and float metric (bit-identical) for indicators in both halves of a unit, present and absent"""
    import torch
    sg, ctx = gpu
    o = oracle
    ocell = o.make_cell(prb, ports, cid, cp=1)
    cell = sg.make_cell(prb, ports, cid, cp=1)
    sf_idx = cid % 10
    units = (6 * prb + 47) // 48
    sent = [(0, 0, 1), (0, 3, 0), (1, 2, 1), (1, 1, 0), (2 * units - 1, 0, 1), (2 * units - 2, 2, 0)]
    probe = sent + [(0, 1, None), (1, 3, None)]
    n = 3
    iq = []
    for i in range(n):
        ocfg = o.make_cfg(ocell, sf_idx=sf_idx, cfi=2, qm=2, tbs=104 if prb == 6 else 1000, tm=ports)
        iq.append(o.gen_subframe(ocell, ocfg, 8100 + i, 10.0 if i < 2 else -4.0, None, pcfich=True, phichs=sent)[1])
    iq = np.stack(iq)
    cfg = sg.make_cfg(cell, sf_idx=sf_idx, cfi=2, qm=2, tbs=0, tm=ports)
    plan = sg.PdschPlan(ctx, cell, cfg, n)
    I = plan.info
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(n, -1)).cuda()
    d_sf = torch.zeros((n, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((n, ports * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((n, 5), dtype=torch.float32, device="cuda")
    plan.ofdm_rx(n, d_iq, d_sf)
    plan.chest(n, d_sf, d_ce, d_meas)
    for g, q, ack in probe:
        d_ack = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_met = torch.zeros(n, dtype=torch.float32, device="cuda")
        plan.phich_decode(n, d_sf, d_ce, d_meas, 0.0, 1, g, q, d_ack, d_met)
        torch.cuda.synchronize()
        for i in range(n):
            sf_o = o.ofdm_rx(prb, iq[i], cp=1)
            ce_o, meas_o = o.chest(ocell, sf_idx, sf_o)
            a_o, m_o = o.phich_decode(ocell, sf_idx, sf_o, ce_o, g, q, float(meas_o[0]))
            assert int(d_ack[i]) == a_o and np.float32(d_met[i].item()) == m_o
            if i < 2 and ack is not None:
                assert a_o == ack
    # (group, sequence) of an uplink grant, both prefixes (36.213 9.1.2)
    L = sg.lib()
    for I_low, dmrs in ((0, 0), (7, 3), (prb - 1, 5)):
        g, q = C.c_int(), C.c_int()
        assert L.srsue_gpu_host_phich_index_cp(prb, 6, 1, I_low, dmrs, C.byref(g), C.byref(q)) == 0
        assert (g.value, q.value) == o.phich_index(prb, I_low, dmrs, 6, cp=1) == ((I_low + dmrs) % (2 * units), (I_low // (2 * units) + dmrs) % 4)
    with pytest.raises(Exception):
        plan.phich_decode(n, d_sf, d_ce, d_meas, 0.0, 1, 2 * units, 0, d_ack, d_met)
    with pytest.raises(Exception):
        plan.phich_decode(n, d_sf, d_ce, d_meas, 0.0, 1, 0, 4, d_ack, d_met)
    plan.close()


@pytest.mark.parametrize("prb,ports,cid", [(6, 1, 1), (6, 2, 77), (25, 2, 300), (100, 1, 503)])
def test_pbch_matches_oracle_extended_prefix(gpu, oracle, prb, ports, cid):
    """the 216-element PBCH: E = 1728 coded bits, 432 per radio frame, the circular buffer of 120 running on across the
    four frames; MIB, port count and frame position equal the oracle's and what was sent"""
    import torch
    sg, ctx = gpu
    o = oracle
    ocell = o.make_cell(prb, ports, cid, cp=1)
    cell = sg.make_cell(prb, 2, cid, cp=1)                 # the receiver estimates two ports and tries both hypotheses
    ocell_rx = o.make_cell(prb, 2, cid, cp=1)
    n = 5
    iq, mibs = [], []
    for i in range(n):
        ocfg = o.make_cfg(ocell, sf_idx=0, cfi=2, qm=2, tbs=56 if prb == 6 else 1000, tm=ports)
        mib = o.mib_pack(prb, 0, 6, 400 + i)
        iq.append(o.gen_subframe(ocell, ocfg, 8200 + i, 9.0 if i < 4 else -8.0, None, pcfich=True, mib=(mib, i % 4))[1])
        mibs.append(mib)
    iq = np.stack(iq)
    cfg = sg.make_cfg(cell, sf_idx=0, cfi=2, qm=2, tbs=0, tm=2)
    plan = sg.PdschPlan(ctx, cell, cfg, n)
    I = plan.info
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(n, -1)).cuda()
    d_sf = torch.zeros((n, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((n, 2 * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((n, 5), dtype=torch.float32, device="cuda")
    d_res = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    d_mib = torch.zeros((n, 24), dtype=torch.uint8, device="cuda")
    plan.ofdm_rx(n, d_iq, d_sf)
    plan.chest(n, d_sf, d_ce, d_meas)
    plan.pbch_decode(n, d_sf, d_ce, d_meas, 0.0, 1, d_res, d_mib)
    torch.cuda.synchronize()
    res, mib_g = d_res.cpu().numpy(), d_mib.cpu().numpy()
    for i in range(n):
        sf_o = o.ofdm_rx(prb, iq[i], cp=1)
        ce_o, meas_o = o.chest(ocell_rx, 0, sf_o)
        f, bits, p, q = o.pbch_decode(ocell_rx, sf_o, ce_o, float(meas_o[0]))
        assert res[i, 0] == f
        if f:
            assert (res[i, 1], res[i, 2]) == (p, q) and np.array_equal(mib_g[i], bits)
        if i < 4:
            assert f == 1 and p == ports and q == i % 4 and np.array_equal(bits, mibs[i])
    g = np.zeros(240, np.int32)
    assert sg.lib().srsue_gpu_host_pbch_res(C.byref(cell), g.ctypes.data_as(C.c_void_p)) == 0
    k = np.zeros(240, np.int32)
    assert o.lib().lteo_pbch_res_n(C.byref(ocell), k.ctypes.data_as(C.c_void_p)) == 216
    assert np.array_equal(g[:216], k[:216]) and (g[216:] == -1).all()
    plan.close()


def _half_frame(o, cell, first_sf, seed, snr, cfo, shift):
    out = []
    for i in range(5):
        cfg = o.make_cfg(cell, sf_idx=first_sf + i, cfi=2, qm=2, tbs=56, tm=cell.nof_ports)
        out.append(o.gen_subframe(cell, cfg, seed + i, snr, None, pcfich=True, sync=True)[1])
    x = np.concatenate(out)
    x = x * np.exp(2j * np.pi * cfo * np.arange(len(x)) / 128)
    return np.roll(x, shift).astype(np.complex64)


def test_cell_search_detects_the_cyclic_prefix(gpu, oracle):
    """srsue_gpu_cell_search_cp: with cp_mode 2 the SSS is looked for behind both prefix lengths and the better metric
    wins -- cells of both kinds, random timing and carrier offsets; every field equals the oracle's, and the prefix found
    is the one transmitted at usable SNRs.  cp_mode 0 / 1 look at one place only."""
    import torch
    sg, ctx = gpu
    o = oracle
    rng = np.random.default_rng(12)
    cases = []
    for cid in (0, 1, 77, 150, 301, 503):
        for cp in (0, 1):
            cases.append((cid, cp, 5 * int(rng.integers(0, 2)), float(rng.choice([12.0, 4.0])), float(rng.uniform(-0.3, 0.3)),
                          int(rng.integers(200, 8000))))
    cases.append((33, 1, 0, -25.0, 0.0, 300))
    bufs = np.stack([_half_frame(o, o.make_cell(6, 1 + (cid % 2), cid, cp=cp), first, 100 * cid + first, snr, cfo, shift)
                     for cid, cp, first, snr, cfo, shift in cases])
    n, ns = bufs.shape
    d_iq = torch.from_numpy(bufs.view(np.float32).reshape(n, -1)).cuda()
    for mode in (2, 0, 1):
        d_res = torch.zeros(n * C.sizeof(sg.SyncResult), dtype=torch.uint8, device="cuda")
        ctx.cell_search(d_iq, n, ns, ns, d_res, cp_mode=mode)
        torch.cuda.synchronize()
        res = (sg.SyncResult * n).from_buffer_copy(d_res.cpu().numpy().tobytes())
        for i, (cid, cp, first, snr, cfo, shift) in enumerate(cases):
            r, ref = res[i], o.pss_search(bufs[i])
            assert (r.peak_pos, r.n_id_2) == (ref["pos"], ref["n_id_2"]) and np.float32(r.peak) == ref["peak"]
            n1, sf5, corr, cp_o = o.sss_detect_cp(bufs[i], ref["pos"], ref["n_id_2"], 128, mode)
            if n1 < 0:
                assert r.valid == 0 and r.n_id_1 == -1
                continue
            assert r.valid == 1 and (r.n_id_1, r.sf5, r.cp) == (n1, sf5, cp_o) and np.float32(r.sss_corr) == corr
            if snr > 0 and mode == 2:
                assert r.cp == cp and 3 * r.n_id_1 + r.n_id_2 == cid and r.sf5 == (first == 5)


def test_init_cell_sequence_extended_prefix(gpu, oracle):
    """phch_recv::init_cell (phch_recv.cc:136-226) on an extended-prefix cell: the scan reports cp = SRSLTE_CP_EXT
    (:189), srslte_ue_mib_sync_decode with that prefix finds a subframe 0 and decodes the 216-element PBCH"""
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import Cell
    from tests.test_gpu_sync import UeSync
    cid, ports, sfn0, cfo, shift = 218, 2, 406, 0.05, 6100
    cell = o.make_cell(6, ports, cid, cp=1)
    stream = []
    for half in range(8):
        sfn = sfn0 + half // 2
        for i in range(5):
            sf = 5 * (half % 2) + i
            cfg = o.make_cfg(cell, sf_idx=sf, cfi=2, qm=2, tbs=56, tm=ports)
            mib = (o.mib_pack(50, 0, 6, sfn), sfn % 4) if sf == 0 else None
            stream.append(o.gen_subframe(cell, cfg, 1000 * half + i, 9.0, None, pcfich=True, sync=True, mib=mib)[1])
    x = np.concatenate(stream)
    x = np.roll(x * np.exp(2j * np.pi * cfo * np.arange(len(x)) / 128), shift).astype(np.complex64)
    state = {"pos": 0}
    RECV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p)

    def recv(handler, data, nsamples, ts):
        idx = (state["pos"] + np.arange(nsamples)) % len(x)
        buf = np.ascontiguousarray(x[idx])
        state["pos"] += nsamples
        C.memmove(data, buf.ctypes.data, buf.nbytes)
        return nsamples

    cb = RECV(recv)

    class Result(C.Structure):
        _fields_ = [("cell_id", C.c_uint32), ("cp", C.c_int), ("peak", C.c_float), ("mode", C.c_float), ("psr", C.c_float), ("cfo", C.c_float)]

    class CellSearch(C.Structure):
        _fields_ = [("ue_sync", UeSync), ("nof_frames_to_scan", C.c_uint32), ("detect_threshold", C.c_float), ("gpu", C.c_void_p)]

    class MibSync(C.Structure):
        _fields_ = [("ue_sync", UeSync), ("cell_id", C.c_uint32), ("gpu", C.c_void_p)]

    cs = CellSearch()
    assert L.srslte_ue_cellsearch_init(C.byref(cs), cb, None) == 0
    L.srslte_ue_cellsearch_set_nof_frames_to_scan(C.byref(cs), 6)
    L.srslte_ue_cellsearch_set_threshold.argtypes = [C.c_void_p, C.c_float]
    L.srslte_ue_cellsearch_set_threshold(C.byref(cs), 15.0)
    found = (Result * 3)()
    best = C.c_uint32(0)
    assert L.srslte_ue_cellsearch_scan(C.byref(cs), found, C.byref(best)) > 0
    assert found[best.value].cell_id == cid and found[best.value].cp == 1          # SRSLTE_CP_EXT
    L.srslte_ue_cellsearch_free(C.byref(cs))
    ms = MibSync()
    assert L.srslte_ue_mib_sync_init(C.byref(ms), found[best.value].cell_id, found[best.value].cp, cb, None) == 0
    payload = (C.c_uint8 * 24)()
    nports, off = C.c_uint32(0), C.c_uint32(0)
    assert L.srslte_ue_mib_sync_decode(C.byref(ms), 12, payload, C.byref(nports), C.byref(off)) == 1
    L.srslte_ue_mib_sync_free(C.byref(ms))
    out_cell, sfn = Cell(), C.c_uint32(0)
    L.srslte_pbch_mib_unpack(payload, C.byref(out_cell), C.byref(sfn))
    assert nports.value == ports and out_cell.nof_prb == 50 and out_cell.phich_resources == 2
    assert sfn0 <= sfn.value + off.value <= sfn0 + 3


def _random_ext_cases():
    from tests.test_gpu_chain import _random_cases
    return _random_cases(12, 20261020)


@pytest.mark.parametrize("case", _random_ext_cases(), ids=lambda c: "%dprb_%dp_qm%d_sf%d_cfi%d_n%d" % (
    c["prb"], c["ports"], c["qm"], c["sf"], c["cfi"], len(c["prbs"])))
def test_random_grants_match_oracle_extended_prefix(gpu, oracle, case):
    """whole chain on randomly drawn extended-prefix cells and grants (1 and 2 ports, scattered allocations, every subframe
    number and CFI), noise around the decoding threshold: payload, verdict, iterations and measurements equal the oracle's"""
    sg, ctx = gpu
    o = oracle
    c = case
    tm = 1 if c["ports"] == 1 else 2
    ocell = o.make_cell(c["prb"], c["ports"], c["cid"], cp=1)
    probe = o.make_cfg(ocell, sf_idx=c["sf"], cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=40, tm=tm, prbs=c["prbs"])
    nre = len(o.pdsch_re_list(ocell, probe))
    if tm == 2:
        nre -= nre % 2
    if nre * c["qm"] < 200:
        pytest.skip("allocation swallowed by the synchronisation signals")
    tbs = min(max(40, int(min(c["rate"], 0.85) * nre * c["qm"]) // 8 * 8 - 24), 61664)
    ocfg = o.make_cfg(ocell, sf_idx=c["sf"], cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=tbs, tm=tm, prbs=c["prbs"])
    cell = sg.make_cell(c["prb"], c["ports"], c["cid"], cp=1)
    cfg = sg.make_cfg(cell, sf_idx=c["sf"], cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=tbs, tm=tm, prbs=c["prbs"])
    base = 10 * np.log10(2 ** (tbs / max(nre, 1)) - 1) + 2.5
    n = 3
    iq = np.stack([o.gen_subframe(ocell, ocfg, c["seed"] + i, base + d, _taps() if (tm == 2 and i == 1) else None)[1]
                   for i, d in enumerate((6.0, 1.0, -4.0))])
    plan = sg.PdschPlan(ctx, cell, cfg, n)
    I = plan.info
    assert I.nof_re == len(o.pdsch_re_list(ocell, ocfg))
    h_pl = np.zeros((n, I.payload_stride), np.uint8)
    h_st = np.zeros((n, 4), np.int32)
    h_meas = np.zeros((n, 5), np.float32)
    plan.decode_batch_host(n, iq, 0.01, 1, 5, h_pl, h_st, h_meas)
    for i in range(n):
        rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, iq[i], 0.01, 1, 5)
        assert (h_st[i, 0] == 1) == (rc == 0), "CRC verdict differs (sf %d)" % i
        assert np.array_equal(h_pl[i], pl), "transport block differs (sf %d)" % i
        assert h_st[i, 2] == avg
        assert np.allclose(h_meas[i], meas, rtol=1e-4)
    plan.close()
