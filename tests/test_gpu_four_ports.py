"""Four-port cells on the device (SPEC.md 15c): channel estimate of ports 2 / 3, the SFBC-FSTD combiner in the PDSCH
demapper and in every control channel, the PBCH's third transmit-port hypothesis, the batching layer and the
srsLTE-shaped worker sequence -- every stage against the CPU oracle, floats bit-identical, integers bit-exact.
nof_ports reaches the worker from the MIB (/root/reference/ue/src/phy/phch_recv.cc:210) through
srslte_ue_dl_init(&ue_dl, cell) (phch_worker.cc:74)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _taps4(seed=5):
    rng = np.random.default_rng(seed)
    taps = (rng.standard_normal((4, 5)) + 1j * rng.standard_normal((4, 5))) * np.array([1, .6, .4, .2, .1])
    return taps / np.sqrt((abs(taps) ** 2).sum(1, keepdims=True))


STAGES = {
    "bw6": dict(prb=6, qm=2, tbs=104, snr=10.0, taps=False, sf=1, cfi=3, cp=0),
    "bw6_ext_sf0": dict(prb=6, qm=2, tbs=56, snr=10.0, taps=True, sf=0, cfi=2, cp=1),
    "bw25_sf5": dict(prb=25, qm=4, tbs=3240, snr=18.0, taps=True, sf=5, cfi=2, cp=0),
    "bw50_ext": dict(prb=50, qm=6, tbs=15264, snr=28.0, taps=True, sf=3, cfi=1, cp=1),
    "bw75": dict(prb=75, qm=4, tbs=12960, snr=16.0, taps=True, sf=4, cfi=1, cp=0),            # 1536-point FFT
    "bw100": dict(prb=100, qm=6, tbs=46888, snr=30.0, taps=True, sf=1, cfi=1, cp=0),          # 70 KB of estimator scratch
    "bw100_ext_sf0": dict(prb=100, qm=6, tbs=36696, snr=30.0, taps=True, sf=0, cfi=2, cp=1),
}


@pytest.mark.parametrize("name", list(STAGES))
def test_frontend_stages_match_oracle_four_ports(gpu, oracle, name):
    import torch
    sg, ctx = gpu
    o = oracle
    c = STAGES[name]
    cp, nsym = c["cp"], 12 if c["cp"] else 14
    taps = _taps4() if c["taps"] else None
    ocell = o.make_cell(c["prb"], 4, 5, cp=cp)
    ocfg = o.make_cfg(ocell, sf_idx=c["sf"], cfi=c["cfi"], qm=c["qm"], tbs=c["tbs"], tm=2)
    n_sf = 3
    iq = np.stack([o.gen_subframe(ocell, ocfg, 7000 + i, c["snr"], taps)[1] for i in range(n_sf)])
    cell = sg.make_cell(c["prb"], 4, 5, cp=cp)
    cfg = sg.make_cfg(cell, sf_idx=c["sf"], cfi=c["cfi"], qm=c["qm"], tbs=c["tbs"], tm=2)
    plan = sg.PdschPlan(ctx, cell, cfg, n_sf)
    I = plan.info
    assert I.nof_re == len(o.pdsch_re_list(ocell, ocfg))
    nsc = I.nsc
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(n_sf, -1)).cuda()
    d_sf = torch.zeros((n_sf, 14 * nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((n_sf, 4 * 14 * nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((n_sf, 5), dtype=torch.float32, device="cuda")
    d_sb = torch.zeros((n_sf, I.sb_sf_stride), dtype=torch.int16, device="cuda")
    d_d = torch.zeros((n_sf, I.nof_re * 2), dtype=torch.float32, device="cuda")
    d_e = torch.zeros((n_sf, I.G), dtype=torch.int16, device="cuda")
    plan.ofdm_rx(n_sf, d_iq, d_sf)
    plan.chest(n_sf, d_sf, d_ce, d_meas)
    plan.pdsch_llr(n_sf, d_sf, d_ce, d_meas, 0.01, 1, 0, d_sb, d_d, d_e)
    torch.cuda.synchronize()
    sf_g = d_sf.cpu().numpy().view(np.complex64).reshape(n_sf, 14, nsc)
    ce_g = d_ce.cpu().numpy().view(np.complex64).reshape(n_sf, 4, 14, nsc)
    meas_g = d_meas.cpu().numpy()
    dd_g = d_d.cpu().numpy().view(np.complex64)
    e_g = d_e.cpu().numpy()
    s = o.cbsegm(c["tbs"])
    for i in range(n_sf):
        sf_o = o.ofdm_rx(c["prb"], iq[i], cp=cp)
        assert np.array_equal(sf_g[i, :nsym].reshape(-1), sf_o[:nsym * nsc])
        ce_o, meas_o = o.chest(ocell, c["sf"], sf_o)
        ce_o = ce_o.reshape(4, 14, nsc)
        for p in range(4):
            assert np.array_equal(ce_g[i, p, :nsym], ce_o[p, :nsym]), "channel estimate of port %d" % p
        assert np.allclose(meas_g[i], meas_o, rtol=1e-4)
        rc, pl, dbg = o.pdsch_decode(ocell, ocfg, sf_o, ce_o.reshape(-1), float(meas_o[0]), 4, want=True)
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
    "p4_1.4MHz": dict(prb=6, qm=2, tbs=104, snr=10.0, n=4, noise_mode=1, sf=1, cp=0),
    "p4_20MHz_64qam": dict(prb=100, qm=6, tbs=46888, snr=30.0, n=4, noise_mode=0, sf=1, cp=0),
    "p4_20MHz_waterfall": dict(prb=100, qm=6, tbs=46888, snr=17.0, n=5, noise_mode=1, sf=2, cp=0),
    "p4_ext_sf0": dict(prb=100, qm=4, tbs=22920, snr=15.0, n=3, noise_mode=1, sf=0, cp=1),
    "p4_15MHz_sf5": dict(prb=75, qm=4, tbs=12960, snr=20.0, n=3, noise_mode=0, sf=5, cp=0),
}


@pytest.mark.parametrize("name", list(CHAIN))
def test_chain_matches_oracle_four_ports(gpu, oracle, name):
    import torch
    sg, ctx = gpu
    o = oracle
    c = CHAIN[name]
    ocell = o.make_cell(c["prb"], 4, 1, cp=c["cp"])
    ocfg = o.make_cfg(ocell, sf_idx=c["sf"], cfi=1, qm=c["qm"], tbs=c["tbs"], tm=2)
    n = c["n"]
    gen = [o.gen_subframe(ocell, ocfg, 21000 + i, c["snr"], _taps4()) for i in range(n)]
    iq = np.stack([g[1] for g in gen])
    cell = sg.make_cell(c["prb"], 4, 1, cp=c["cp"])
    cfg = sg.make_cfg(cell, sf_idx=c["sf"], cfi=1, qm=c["qm"], tbs=c["tbs"], tm=2)
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


def test_plan_rules_four_ports(gpu):
    sg, ctx = gpu
    cell = sg.make_cell(25, 4, 1)
    with pytest.raises(Exception):              # a four-port cell transmits with diversity
        sg.PdschPlan(ctx, cell, sg.make_cfg(cell, sf_idx=1, cfi=1, qm=2, tbs=1000, tm=1), 1)
    with pytest.raises(Exception):
        sg.PdschPlan(ctx, sg.make_cell(25, 3, 1), sg.make_cfg(cell, sf_idx=1, cfi=1, qm=2, tbs=1000, tm=2), 1)


@pytest.mark.parametrize("prb,cid,cfi,cp", [(6, 1, 3, 0), (6, 8, 2, 1), (25, 77, 2, 0), (50, 300, 3, 1), (100, 503, 1, 0)])
def test_pcfich_and_pdcch_match_oracle_four_ports(gpu, oracle, prb, cid, cfi, cp):
    """control region of a four-port cell (six-element REGs in symbol 1, pairs alternating between ports (0, 2) and
    (1, 3)): CFI, the LLRs of every control-channel element and the blind-search verdict equal the oracle's"""
    import torch
    sg, ctx = gpu
    o = oracle
    ocell = o.make_cell(prb, 4, cid, cp=cp)
    cell = sg.make_cell(prb, 4, cid, cp=cp)
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
        ocfg = o.make_cfg(ocell, sf_idx=sf_idx, cfi=cfi, qm=2, tbs=56 if prb == 6 else 1000, tm=2)
        iq.append(o.gen_subframe(ocell, ocfg, 4100 + i, 9.0, _taps4(cid), pcfich=True, dcis=dcis)[1])
        sent.append(dcis)
    iq = np.stack(iq)
    cfg = sg.make_cfg(cell, sf_idx=sf_idx, cfi=cfi, qm=2, tbs=0, tm=2)
    plan = sg.PdschPlan(ctx, cell, cfg, n)
    I = plan.info
    n_reg, nc = plan.pdcch_info(6)
    assert n_reg == len(rk) and nc == ncce
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(n, -1)).cuda()
    d_sf = torch.zeros((n, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((n, 4 * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
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
        sf_o = o.ofdm_rx(prb, iq[i], cp=cp)
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


@pytest.mark.parametrize("prb,cid,cp", [(6, 1, 0), (25, 77, 1), (50, 301, 0), (100, 0, 1)])
def test_phich_matches_oracle_four_ports(gpu, oracle, prb, cid, cp):
    """36.211 6.9.2: the quadruplets of a group alternate between the port pairs (0, 2) and (1, 3), starting with the
    parity of the group number; decision and float metric (bit-identical) on four independent channels"""
    import torch
    sg, ctx = gpu
    o = oracle
    ocell = o.make_cell(prb, 4, cid, cp=cp)
    cell = sg.make_cell(prb, 4, cid, cp=cp)
    sf_idx = cid % 10
    units = (6 * prb + 47) // 48
    ng = (2 if cp else 1) * units
    nseq = 4 if cp else 8
    sent = [(0, 0, 1), (0, nseq - 1, 0), (ng - 1, 2, 1), (ng - 1, 1, 0)] + ([(1, 2, 1), (1, 3, 0)] if ng > 2 else [])
    probe = sent + [(0, 1, None)]
    n = 3
    iq = []
    for i in range(n):
        ocfg = o.make_cfg(ocell, sf_idx=sf_idx, cfi=2, qm=2, tbs=56 if prb == 6 else 1000, tm=2)
        iq.append(o.gen_subframe(ocell, ocfg, 8100 + i, 10.0 if i < 2 else -4.0, _taps4(cid + 1), pcfich=True, phichs=sent)[1])
    iq = np.stack(iq)
    cfg = sg.make_cfg(cell, sf_idx=sf_idx, cfi=2, qm=2, tbs=0, tm=2)
    plan = sg.PdschPlan(ctx, cell, cfg, n)
    I = plan.info
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(n, -1)).cuda()
    d_sf = torch.zeros((n, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((n, 4 * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((n, 5), dtype=torch.float32, device="cuda")
    plan.ofdm_rx(n, d_iq, d_sf)
    plan.chest(n, d_sf, d_ce, d_meas)
    for g, q, ack in probe:
        d_ack = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_met = torch.zeros(n, dtype=torch.float32, device="cuda")
        plan.phich_decode(n, d_sf, d_ce, d_meas, 0.0, 1, g, q, d_ack, d_met)
        torch.cuda.synchronize()
        for i in range(n):
            sf_o = o.ofdm_rx(prb, iq[i], cp=cp)
            ce_o, meas_o = o.chest(ocell, sf_idx, sf_o)
            a_o, m_o = o.phich_decode(ocell, sf_idx, sf_o, ce_o, g, q, float(meas_o[0]))
            assert int(d_ack[i]) == a_o and np.float32(d_met[i].item()) == m_o
            if i < 2 and ack is not None:
                assert a_o == ack
    plan.close()


@pytest.mark.parametrize("prb,ports,cid,cp", [(6, 4, 1, 0), (6, 1, 77, 0), (25, 2, 300, 1), (25, 4, 301, 1), (100, 4, 503, 0)])
def test_pbch_three_hypotheses(gpu, oracle, prb, ports, cid, cp):
    """a receiver that estimates four ports tries 1, 2 and 4 transmit ports (CRC masks 0x0000 / 0xFFFF / 0x5555, 36.212
    5.3.1.1); cells of every kind: MIB, port count and frame position equal the oracle's and what was sent"""
    import torch
    sg, ctx = gpu
    o = oracle
    ocell = o.make_cell(prb, ports, cid, cp=cp)
    cell = sg.make_cell(prb, 4, cid, cp=cp)
    ocell_rx = o.make_cell(prb, 4, cid, cp=cp)
    n = 5
    iq, mibs = [], []
    for i in range(n):
        ocfg = o.make_cfg(ocell, sf_idx=0, cfi=2, qm=2, tbs=56 if prb == 6 else 1000, tm=1 if ports == 1 else 2)
        mib = o.mib_pack(prb, 0, 6, 400 + i)
        taps = _taps4(cid)[:ports]
        iq.append(o.gen_subframe(ocell, ocfg, 8200 + i, 11.0 if i < 4 else -8.0, taps, pcfich=True, mib=(mib, i % 4))[1])
        mibs.append(mib)
    iq = np.stack(iq)
    cfg = sg.make_cfg(cell, sf_idx=0, cfi=2, qm=2, tbs=0, tm=2)
    plan = sg.PdschPlan(ctx, cell, cfg, n)
    I = plan.info
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(n, -1)).cuda()
    d_sf = torch.zeros((n, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((n, 4 * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((n, 5), dtype=torch.float32, device="cuda")
    d_res = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    d_mib = torch.zeros((n, 24), dtype=torch.uint8, device="cuda")
    plan.ofdm_rx(n, d_iq, d_sf)
    plan.chest(n, d_sf, d_ce, d_meas)
    plan.pbch_decode(n, d_sf, d_ce, d_meas, 0.0, 1, d_res, d_mib)
    torch.cuda.synchronize()
    res, mib_g = d_res.cpu().numpy(), d_mib.cpu().numpy()
    for i in range(n):
        sf_o = o.ofdm_rx(prb, iq[i], cp=cp)
        ce_o, meas_o = o.chest(ocell_rx, 0, sf_o)
        f, bits, p, q = o.pbch_decode(ocell_rx, sf_o, ce_o, float(meas_o[0]))
        assert res[i, 0] == f
        if f:
            assert (res[i, 1], res[i, 2]) == (p, q) and np.array_equal(mib_g[i], bits)
        if i < 4:
            assert f == 1 and p == ports and q == i % 4 and np.array_equal(bits, mibs[i])
    plan.close()


def test_batch_layer_four_port_rows(gpu, oracle):
    """one submission mixing two- and four-port cells of the same bandwidth, id and grant, both prefixes: one plan per
    kind, every transport block as the oracle decodes it"""
    sg, ctx = gpu
    o = oracle
    rows = [(25, 2, 4, 3240, 2, 2, 18.0, 0), (25, 4, 4, 3240, 2, 2, 18.0, 0), (25, 4, 4, 3240, 2, 2, 18.0, 1),
            (100, 4, 6, 46888, 7, 1, 30.0, 0), (6, 4, 2, 104, 5, 3, 12.0, 0)]
    rng = np.random.default_rng(10)
    order = [int(x) for x in rng.integers(0, len(rows), 20)]
    items, refs = [], []
    for i, m in enumerate(order):
        prb, ports, qm, tbs, sf, cfi, snr, cp = rows[m]
        ocell = o.make_cell(prb, ports, 1, cp=cp)
        ocfg = o.make_cfg(ocell, sf_idx=sf, cfi=cfi, qm=qm, tbs=tbs, tm=2)
        cell = sg.make_cell(prb, ports, 1, cp=cp)
        cfg = sg.make_cfg(cell, sf_idx=sf, cfi=cfi, qm=qm, tbs=tbs, tm=2)
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 53000 + i, snr, _taps4()[:ports])
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


def test_srslte_worker_sequence_four_ports(gpu, oracle):
    """phch_worker's sequence through the srsLTE-shaped symbols on a four-port cell: init (phch_worker.cc:74), FFT +
    estimate (:254) filling ce[0..3], grant configuration (:337), PDSCH decode (:347)"""
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import UeDl, Cell, SoftBuffer, make_grant
    prb, qm, tbs = 25, 4, 3240
    ocell = o.make_cell(prb, 4, 1)
    ocfg = o.make_cfg(ocell, sf_idx=3, cfi=2, qm=qm, tbs=tbs, tm=2)
    tb, iq, _ = o.gen_subframe(ocell, ocfg, 123, 20.0, _taps4(), pcfich=True)
    q = UeDl()
    cell = Cell(nof_prb=prb, nof_ports=4, bw_idx=0, id=1, cp=0, phich_length=0, phich_resources=2)
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
    sf_o = o.ofdm_rx(prb, iq)
    ce_o, meas_o = o.chest(ocell, 3, sf_o)
    ce_o = ce_o.reshape(4, 14 * nsc)
    for p in range(4):
        ce_h = np.ctypeslib.as_array(C.cast(q.ce[p], C.POINTER(C.c_float)), shape=(14 * nsc * 2,)).view(np.complex64)
        assert np.array_equal(ce_h, ce_o[p])
    grant = make_grant(prb, qm, tbs)
    assert L.srslte_ue_dl_cfg_grant(C.byref(q), C.byref(grant), cfi.value, 3, 0) == 0
    assert q.pdsch_cfg.nbits.nof_re == len(o.pdsch_re_list(ocell, ocfg))
    payload = np.zeros(tbs // 8, np.uint8)
    ret = L.srslte_pdsch_decode_rnti(C.byref(q.pdsch), C.byref(q.pdsch_cfg), C.byref(sb), C.c_void_p(q.sf_symbols), q.ce,
                                     C.c_float(0.01), C.c_uint16(0x1234), payload.ctypes.data_as(C.c_void_p))
    rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, iq, 0.01, 0, 4)
    assert ret == 0 and rc == 0
    assert np.array_equal(payload, pl) and np.array_equal(payload, tb)
    L.srslte_softbuffer_rx_free(C.byref(sb))
    L.srslte_ue_dl_free(C.byref(q))


def test_init_cell_sequence_four_ports(gpu, oracle):
    """phch_recv::init_cell (phch_recv.cc:136-226) on a four-port cell: the MIB decoder reports nof_ports = 4 (:210)"""
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import Cell
    from tests.test_gpu_sync import UeSync
    cid, ports, sfn0, cfo, shift = 219, 4, 406, 0.05, 6100
    cell = o.make_cell(6, ports, cid)
    stream = []
    for half in range(8):
        sfn = sfn0 + half // 2
        for i in range(5):
            sf = 5 * (half % 2) + i
            cfg = o.make_cfg(cell, sf_idx=sf, cfi=2, qm=2, tbs=56, tm=2)
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

    class MibSync(C.Structure):
        _fields_ = [("ue_sync", UeSync), ("cell_id", C.c_uint32), ("gpu", C.c_void_p)]

    ms = MibSync()
    assert L.srslte_ue_mib_sync_init(C.byref(ms), cid, 0, cb, None) == 0
    payload = (C.c_uint8 * 24)()
    nports, off = C.c_uint32(0), C.c_uint32(0)
    assert L.srslte_ue_mib_sync_decode(C.byref(ms), 12, payload, C.byref(nports), C.byref(off)) == 1
    L.srslte_ue_mib_sync_free(C.byref(ms))
    out_cell, sfn = Cell(), C.c_uint32(0)
    L.srslte_pbch_mib_unpack(payload, C.byref(out_cell), C.byref(sfn))
    assert nports.value == 4 and out_cell.nof_prb == 50
    assert sfn0 <= sfn.value + off.value <= sfn0 + 3


def test_blind_batch_other_cell_shapes(gpu, oracle):
    """srsue_gpu_batch_submit_blind (phch_worker.cc:254-297 for a whole batch) on four-port and extended-prefix cells: the
    caller knows cell, subframe number and RNTI; CFI, grant and transport block equal what was sent"""
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import DciMsg, RaDlDci, install_tbs_table
    # (prb, ports, cp, cfi, sf_idx, rnti, mcs, RB_start, L_crb, tbs, with_dci)
    cases = [(25, 4, 0, 2, 4, 0x1234, 9, 2, 10, 1544, True), (25, 4, 1, 1, 4, 0x1234, 9, 0, 10, 1544, True), (25, 4, 0, 2, 4, 0x1234, 9, 2, 10, 1544, False),
             (50, 4, 0, 1, 7, 0x0456, 12, 5, 25, 5736, True), (6, 4, 0, 2, 1, 0x1234, 4, 0, 6, 408, True), (50, 2, 1, 3, 7, 0x0456, 3, 0, 50, 2856, True),
             (100, 4, 0, 1, 2, 0x2222, 20, 0, 100, 46888, True), (100, 1, 1, 1, 2, 0x2222, 20, 0, 100, 46888, True)]
    table = {}
    for prb, ports, cp, cfi, sf, rnti, mcs, start, ln, tbs, with_dci in cases:
        itbs = mcs if mcs < 10 else mcs - 1 if mcs < 17 else mcs - 2      # 36.213 Table 7.1.7.1-1
        table[(itbs, ln)] = tbs
    install_tbs_table(L, table)
    items, truth = [], []
    for k, (prb, ports, cp, cfi, sf, rnti, mcs, start, ln, tbs, with_dci) in enumerate(cases):
        ocell = o.make_cell(prb, ports, 1, cp=cp)
        tm = 1 if ports == 1 else 2
        sent = RaDlDci()
        sent.mcs_idx, sent.rv_idx, sent.alloc_type = mcs, 0, 2
        sent.type2_alloc.RB_start, sent.type2_alloc.L_crb, sent.type2_alloc.n_prb1a = start, ln, 1
        m = DciMsg()
        nb = L.srslte_dci_msg_pack_pdsch(C.byref(sent), 2, C.byref(m), prb, True)
        assert nb > 0
        rk, _ = o.pdcch_regs(ocell, cfi, 6)
        ss = o.pdcch_search_space(len(rk) // 9, sf, rnti)
        qm = 2 if mcs < 10 else 4 if mcs < 17 else 6
        prbs = list(range(start, start + ln))
        ocfg = o.make_cfg(ocell, sf_idx=sf, cfi=cfi, rnti=rnti, qm=qm, tbs=tbs, prbs=prbs, tm=tm)
        dcis = [(np.frombuffer(m.data, np.uint8)[:nb].copy(), rnti, ss[0][0], ss[0][1])] if with_dci else None
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 9400 + k, 28.0 if qm == 6 else 18.0, _taps4(k)[:ports], pcfich=True, dcis=dcis)
        cell = sg.make_cell(prb, ports, 1, cp=cp)
        items.append(dict(cell=cell, cfg=sg.make_cfg(cell, sf_idx=sf, cfi=1, rnti=rnti, qm=2, tbs=0, tm=tm), iq=iq))
        truth.append((ocell, ocfg, iq, tb, cfi, tbs, qm, prbs, with_dci))
    b = sg.Batch(ctx, 32)
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
    b.close()


def test_cpp_offline_driver_on_a_four_port_extended_prefix_capture(gpu, oracle, tmp_path):
    """driver/pdsch_offline.cc (worker = phch_worker's call sequence, batch = the batching layer) on a capture of a four-port
    cell with the extended cyclic prefix (header magic 0x53525345): transport blocks, verdicts, iterations and SNR of the oracle"""
    import os, struct, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "build", "pdsch_offline")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(root, "include"), "-o", exe,
                           os.path.join(root, "driver", "pdsch_offline.cc"), "-L" + os.path.join(root, "srsue_b200"),
                           "-lsrsue_gpu", "-Wl,-rpath," + os.path.join(root, "srsue_b200")])
    o = oracle
    for cp, magic in ((1, 0x53525345), (0, 0x53525355)):
        prb, qm, tbs, n = 50, 4, 6200, 4
        ocell = o.make_cell(prb, 4, 1, cp=cp)
        ocfg = o.make_cfg(ocell, sf_idx=1, cfi=1, qm=qm, tbs=tbs, tm=2)
        iq = np.stack([o.gen_subframe(ocell, ocfg, 700 + i, 18.0 if i != 2 else 0.0, _taps4())[1] for i in range(n)])
        inp = tmp_path / ("in%d.bin" % cp)
        with open(inp, "wb") as f:
            f.write(struct.pack("<12i", magic, prb, 4, 1, 1, 1, 0x1234, qm, tbs, 0, n, 4))
            f.write(iq.tobytes())
        ref = [o.ue_dl_decode(ocell, ocfg, iq[i], 0.01, 0, 4) for i in range(n)]
        for mode in ("worker", "batch"):
            outp = tmp_path / ("out_%s%d.bin" % (mode, cp))
            subprocess.check_call([exe, mode, str(inp), str(outp)])
            raw = open(outp, "rb").read()
            rec = 12 + tbs // 8
            assert len(raw) == n * rec
            for i in range(n):
                ack, n_iter, snr = struct.unpack_from("<iif", raw, i * rec)
                payload = np.frombuffer(raw, np.uint8, tbs // 8, i * rec + 12)
                rc, pl, meas, avg = ref[i]
                assert ack == int(rc == 0) and n_iter == avg, (mode, i)
                assert np.array_equal(payload, pl), (mode, i)
                assert abs(snr - meas[4]) <= 1e-4 * meas[4]
        assert ref[2][0] != 0 and ref[0][0] == 0      # the noisy subframe fails, the others decode


def _random_cases4(n, seed):
    """random four-port cells / grants: every bandwidth, both prefixes, all modulations, partial and scattered allocations,
    every subframe number and CFI, random cell ids and RNTIs, code rates 0.1 .. 0.85"""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        prb = int(rng.choice([6, 15, 25, 50, 75, 100]))
        nalloc = int(rng.integers(max(1, prb // 8), prb + 1))
        if rng.random() < 0.5:
            prbs = sorted(int(x) for x in rng.choice(prb, nalloc, replace=False))
        else:
            start = int(rng.integers(0, prb - nalloc + 1))
            prbs = list(range(start, start + nalloc))
        out.append(dict(prb=prb, cp=int(rng.integers(0, 2)), qm=int(rng.choice([2, 4, 6])), cid=int(rng.integers(0, 504)), sf=int(rng.integers(0, 10)),
                        cfi=int(rng.integers(1, 4)), prbs=prbs, rnti=int(rng.integers(1, 65520)), rate=float(rng.uniform(0.1, 0.85)),
                        seed=int(rng.integers(1, 1 << 30))))
    return out


@pytest.mark.parametrize("case", _random_cases4(16, 20261019), ids=lambda c: "%dprb_cp%d_qm%d_sf%d_cfi%d_n%d" % (
    c["prb"], c["cp"], c["qm"], c["sf"], c["cfi"], len(c["prbs"])))
def test_random_grants_match_oracle_four_ports(gpu, oracle, case):
    """whole chain on randomly drawn four-port cells and grants over four independent channels, noise around the decoding
    threshold so that both CRC verdicts occur: payload, verdict, iterations and measurements equal the oracle's"""
    sg, ctx = gpu
    o = oracle
    c = case
    ocell = o.make_cell(c["prb"], 4, c["cid"], cp=c["cp"])
    probe = o.make_cfg(ocell, sf_idx=c["sf"], cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=40, tm=2, prbs=c["prbs"])
    nre = len(o.pdsch_re_list(ocell, probe))
    nre -= nre % 2
    if nre * c["qm"] < 200:
        pytest.skip("allocation swallowed by the synchronisation signals")
    tbs = min(max(40, int(c["rate"] * nre * c["qm"]) // 8 * 8 - 24), 75376)
    ocfg = o.make_cfg(ocell, sf_idx=c["sf"], cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=tbs, tm=2, prbs=c["prbs"])
    cell = sg.make_cell(c["prb"], 4, c["cid"], cp=c["cp"])
    cfg = sg.make_cfg(cell, sf_idx=c["sf"], cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=tbs, tm=2, prbs=c["prbs"])
    base = 10 * np.log10(2 ** (tbs / max(nre, 1)) - 1) + 3.5
    n = 3
    iq = np.stack([o.gen_subframe(ocell, ocfg, c["seed"] + i, base + d, _taps4(c["seed"] % 1000) if i == 1 else None)[1]
                   for i, d in enumerate((7.0, 2.0, -4.0))])
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
