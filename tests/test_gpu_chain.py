"""Whole-chain parity through the C ABI: decoded transport blocks, CRC verdicts and iteration counts
against the CPU oracle, for the PDSCH configs of BASELINE.json, device-resident and host-buffer calls."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _taps():
    rng = np.random.default_rng(77)
    taps = (rng.standard_normal((2, 6)) + 1j * rng.standard_normal((2, 6))) * np.array([1, .7, .5, .3, .2, .1])
    return taps / np.sqrt((abs(taps) ** 2).sum(1, keepdims=True))


CASES = {
    "cfg1_1.4MHz_mcs0": dict(prb=6, ports=1, qm=2, tbs=152, tm=1, snr=10.0, taps=False, n=4, noise_mode=1),
    "cfg2_20MHz_mcs28": dict(prb=100, ports=1, qm=6, tbs=75376, tm=1, snr=30.0, taps=False, n=4, noise_mode=0),
    "cfg2_waterfall": dict(prb=100, ports=1, qm=6, tbs=75376, tm=1, snr=19.0, taps=False, n=6, noise_mode=0),
    "cfg3_tm2_mcs16": dict(prb=100, ports=2, qm=4, tbs=30576, tm=2, snr=15.0, taps=True, n=4, noise_mode=1),
    "filler_and_two_K": dict(prb=50, ports=1, qm=4, tbs=6208, tm=1, snr=18.0, taps=False, n=3, noise_mode=0),
    "bw75_15MHz": dict(prb=75, ports=1, qm=6, tbs=55056, tm=1, snr=30.0, taps=False, n=3, noise_mode=0),
    "low_snr_fail": dict(prb=25, ports=1, qm=6, tbs=11448, tm=1, snr=3.0, taps=False, n=3, noise_mode=0),
}


@pytest.mark.parametrize("name", list(CASES))
def test_chain_matches_oracle(gpu, oracle, name):
    import torch
    sg, ctx = gpu
    o = oracle
    c = CASES[name]
    taps = _taps() if c["taps"] else None
    ocell = o.make_cell(c["prb"], c["ports"], 1)
    ocfg = o.make_cfg(ocell, sf_idx=1, cfi=1, qm=c["qm"], tbs=c["tbs"], tm=c["tm"])
    n = c["n"]
    iq = np.stack([o.gen_subframe(ocell, ocfg, 20000 + i, c["snr"], taps)[1] for i in range(n)])
    sent = [o.gen_subframe(ocell, ocfg, 20000 + i, c["snr"], taps)[0] for i in range(n)]
    cell = sg.make_cell(c["prb"], c["ports"], 1)
    cfg = sg.make_cfg(cell, sf_idx=1, cfi=1, qm=c["qm"], tbs=c["tbs"], tm=c["tm"])
    plan = sg.PdschPlan(ctx, cell, cfg, n)
    I = plan.info
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(n, -1)).cuda()
    d_pl = torch.zeros((n, I.payload_stride), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    d_meas = torch.zeros((n, 5), dtype=torch.float32, device="cuda")
    plan.decode_batch(n, d_iq, 0.01, c["noise_mode"], 4, d_pl, d_st, d_meas=d_meas)
    torch.cuda.synchronize()
    pl_g, st_g = d_pl.cpu().numpy(), d_st.cpu().numpy()
    # host-buffer entry point (what the offline driver calls)
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
            assert np.array_equal(pl, sent[i])
    if name == "low_snr_fail":
        assert n_ok == 0
    elif name != "cfg2_waterfall":
        assert n_ok == n
    plan.close()


def test_srslte_shaped_entry_points(gpu, oracle):
    """phch_worker's call sequence through the srsLTE-compatible symbols (host buffers, batch of one)."""
    import ctypes as C
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import UeDl, Cell, Grant, SoftBuffer, make_grant
    prb, qm, tbs = 25, 4, 4968
    ocell = o.make_cell(prb, 1, 1)
    ocfg = o.make_cfg(ocell, sf_idx=1, cfi=2, qm=qm, tbs=tbs)
    tb, iq, _ = o.gen_subframe(ocell, ocfg, 321, 20.0, pcfich=True)
    q = UeDl()
    cell = Cell(nof_prb=prb, nof_ports=1, bw_idx=0, id=1, cp=0, phich_length=0, phich_resources=0)
    assert L.srslte_ue_dl_init(C.byref(q), cell) == 0
    L.srslte_ue_dl_set_rnti(C.byref(q), 0x1234)
    L.srslte_sch_set_max_noi(C.byref(q.pdsch.dl_sch), 4)
    sb = SoftBuffer()
    assert L.srslte_softbuffer_rx_init(C.byref(sb), prb) == 0
    L.srslte_softbuffer_rx_reset(C.byref(sb))
    cfi = C.c_uint32(0)
    assert L.srslte_ue_dl_decode_fft_estimate(C.byref(q), iq.ctypes.data_as(C.c_void_p), 1, C.byref(cfi)) == 0
    assert cfi.value == 2                      # decoded from the PCFICH on the device
    grant = make_grant(prb, qm, tbs)
    assert L.srslte_ue_dl_cfg_grant(C.byref(q), C.byref(grant), cfi.value, 1, 0) == 0
    assert q.pdsch_cfg.nbits.nof_re == len(o.pdsch_re_list(ocell, ocfg))
    payload = np.zeros(tbs // 8, np.uint8)
    ret = L.srslte_pdsch_decode_rnti(C.byref(q.pdsch), C.byref(q.pdsch_cfg), C.byref(sb), C.c_void_p(q.sf_symbols), q.ce,
                                     C.c_float(0.01), C.c_uint16(0x1234), payload.ctypes.data_as(C.c_void_p))
    rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, iq, 0.01, 0, 4)
    assert ret == 0 and rc == 0
    assert np.array_equal(payload, pl) and np.array_equal(payload, tb)
    assert L.srslte_pdsch_last_noi(C.byref(q.pdsch)) == avg
    L.srslte_chest_dl_get_snr.restype = C.c_float
    assert abs(L.srslte_chest_dl_get_snr(C.byref(q.chest)) - meas[4]) <= 1e-4 * meas[4]
    # host mirrors of the grid are what the oracle computes
    nsc = 12 * prb
    sf_h = np.ctypeslib.as_array(C.cast(q.sf_symbols, C.POINTER(C.c_float)), shape=(14 * nsc * 2,)).view(np.complex64)
    assert np.array_equal(sf_h, o.ofdm_rx(prb, iq))
    # srslte_tdec object
    from tests.srslte_ctypes import Tdec
    h = Tdec()
    K = 1024
    c, llr = o.gen_turbo_llrs(K, 5, ebn0_db=1.5)
    assert L.srslte_tdec_init(C.byref(h), 6144) == 0
    out = np.zeros(K // 8, np.uint8)
    assert L.srslte_tdec_run_all(C.byref(h), llr.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 4, K) == 0
    assert np.array_equal(np.unpackbits(out), o.tdec(llr, K, 4, 0)[0])
    L.srslte_tdec_free(C.byref(h))
    L.srslte_softbuffer_rx_free(C.byref(sb))
    L.srslte_ue_dl_free(C.byref(q))


@pytest.mark.parametrize("prb,ports,fmt", [(25, 1, "1A"), (50, 2, "1"), (100, 1, "1A")])
def test_phch_worker_sequence_with_pdcch_search(gpu, oracle, prb, ports, fmt):
    """work_imp's downlink sequence (phch_worker.cc:254-348) with nothing supplied by the caller but the RNTI:
    decode_fft_estimate -> CFI, pdcch_extract_llr, find_dl_dci_type -> DCI bits + CCE location, dci_msg_to_dl_grant,
    then cfg_grant + pdsch_decode_rnti on that grant."""
    import ctypes as C
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import UeDl, Cell, SoftBuffer, DciMsg, Grant, RaDlDci, install_tbs_table
    qm, tbs, cfi, sf_idx, rnti = 4, 4968 if prb == 25 else 6208, 2, 3, 0x4601
    nb = L.srsue_gpu_host_dci_format_sizeof(0 if fmt == "1A" else 1, prb)
    ocell = o.make_cell(prb, ports, 1)
    # the eNodeB side: a real DCI for a partial allocation (1A: localized type 2; 1: a type 0 RBG bitmap), MCS 12
    sent = RaDlDci()
    sent.mcs_idx, sent.harq_process, sent.rv_idx, sent.ndi = 12, 5, 0, True
    if fmt == "1A":
        sent.alloc_type = 2
        sent.type2_alloc.RB_start, sent.type2_alloc.L_crb = 2, prb - 5
        prbs = list(range(2, prb - 3))
    else:
        P = L.srslte_ra_type0_P(prb)
        nbm = -(-prb // P)
        sent.alloc_type = 0
        sent.type0_alloc.rbg_bitmask = int("".join("0" if i % 4 == 1 else "1" for i in range(nbm)), 2)
        prbs = [i for i in range(prb) if (i // P) % 4 != 1]
    install_tbs_table(L, {(11, len(prbs)): tbs})
    sent_msg = DciMsg()
    assert L.srslte_dci_msg_pack_pdsch(C.byref(sent), 2 if fmt == "1A" else 1, C.byref(sent_msg), prb, True) == nb
    ocfg = o.make_cfg(ocell, sf_idx=sf_idx, cfi=cfi, rnti=rnti, qm=qm, tbs=tbs, tm=ports, prbs=prbs)
    rk, _ = o.pdcch_regs(ocell, cfi, 6)
    ss = o.pdcch_search_space(len(rk) // 9, sf_idx, rnti)
    L0, n0 = ss[-1]                                               # the last (largest) candidate of the UE-specific space
    dci_bits = np.frombuffer(sent_msg.data, np.uint8)[:nb].copy()
    other = np.random.default_rng(1).integers(0, 2, nb, dtype=np.uint8)
    dcis = [(dci_bits, rnti, L0, n0)]
    if n0 >= 1:
        dcis.append((other, 0x0999, 1, 0))                        # somebody else's DCI in the same control region
    # an uplink grant (format 0: the size of 1A, flag 0) for the same RNTI on an earlier candidate of the search space
    nb0 = L.srsue_gpu_host_dci_format_sizeof(0, prb)
    ul_bits = np.random.default_rng(7).integers(0, 2, nb0, dtype=np.uint8)
    ul_bits[0] = 0
    taken = set(range(n0, n0 + L0)) | ({0} if n0 >= 1 else set())
    ul_at = next(((Lc, nc) for Lc, nc in ss if Lc == 2 and taken.isdisjoint(range(nc, nc + Lc))), None)
    if ul_at:
        dcis.append((ul_bits, rnti, ul_at[0], ul_at[1]))
    # HARQ indicators for two earlier uplink transmissions of this UE (phch_worker.cc:381)
    ul_tx = [(3, 1, 1), (prb // 2, 0, 0)]                          # (I_lowest, n_dmrs, ack)
    phichs = [o.phich_index(prb, Il, nd, 6) + (a,) for Il, nd, a in ul_tx]
    tb, iq, _ = o.gen_subframe(ocell, ocfg, 555, 20.0, None, pcfich=True, dcis=dcis, phichs=phichs)
    q = UeDl()
    cell = Cell(nof_prb=prb, nof_ports=ports, bw_idx=0, id=1, cp=0, phich_length=0, phich_resources=2)     # Ng = 1
    assert L.srslte_ue_dl_init(C.byref(q), cell) == 0
    L.srslte_ue_dl_set_rnti(C.byref(q), rnti)
    sb = SoftBuffer()
    assert L.srslte_softbuffer_rx_init(C.byref(sb), prb) == 0
    L.srslte_softbuffer_rx_reset(C.byref(sb))
    got_cfi = C.c_uint32(0)
    assert L.srslte_ue_dl_decode_fft_estimate(C.byref(q), iq.ctypes.data_as(C.c_void_p), sf_idx, C.byref(got_cfi)) == 0
    assert got_cfi.value == cfi
    L.srslte_ue_dl_decode_phich.restype = C.c_bool
    for Il, nd, a in ul_tx:
        assert bool(L.srslte_ue_dl_decode_phich(C.byref(q), sf_idx, Il, nd)) == bool(a)
    assert L.srslte_pdcch_extract_llr(C.byref(q.pdcch), C.c_void_p(q.sf_symbols), q.ce, C.c_float(0.0), sf_idx, got_cfi.value) == 0
    msg = DciMsg()
    assert L.srslte_ue_dl_find_dl_dci_type(C.byref(q), C.byref(msg), got_cfi.value, sf_idx, rnti, 0) == 1
    assert msg.nof_bits == nb and np.array_equal(np.frombuffer(msg.data, np.uint8)[:nb], dci_bits)
    assert n0 <= L.srslte_ue_dl_get_ncce(C.byref(q)) < n0 + L0
    # the same search in the oracle (ZF equaliser: noise 0, as phch_worker.cc:260 passes)
    sf_o = o.ofdm_rx(prb, iq)
    ce_o, _ = o.chest(ocell, sf_idx, sf_o)
    llr_o, ncce = o.pdcch_extract_llr(ocell, sf_idx, cfi, sf_o, ce_o, 0.0)
    cands = []                     # the oracle's matches in search-space order, with the shim's format-flag rule
    for Lc, nc in o.pdcch_search_space(ncce, sf_idx, rnti):
        b, r = o.pdcch_decode_candidate(llr_o[72 * nc:], Lc, nb)
        if r == rnti and (fmt != "1A" or b[0] == 1):
            cands.append((Lc, nc, b))
    assert cands and (q.last_location.L, q.last_location.ncce) == cands[0][:2] and np.array_equal(cands[0][2], dci_bits)
    # a different RNTI finds nothing
    assert L.srslte_ue_dl_find_dl_dci_type(C.byref(q), C.byref(msg), got_cfi.value, sf_idx, 0x0777, 0) == 0
    # the uplink grant of the same subframe (phch_worker.cc:426)
    if ul_at:
        ul = DciMsg()
        assert L.srslte_ue_dl_find_ul_dci(C.byref(q), C.byref(ul), got_cfi.value, sf_idx, rnti) == 1
        assert ul.nof_bits == nb0 and np.array_equal(np.frombuffer(ul.data, np.uint8)[:nb0], ul_bits)
        assert ul_at[1] <= q.last_location.ncce < ul_at[1] + ul_at[0]
    # DCI -> grant (phch_worker.cc:297) -> PDSCH
    grant, unpacked = Grant(), RaDlDci()
    assert L.srslte_dci_msg_to_dl_grant(C.byref(msg), rnti, prb, C.byref(unpacked), C.byref(grant)) == 0
    assert (grant.nof_prb, grant.Qm, grant.mcs.tbs) == (len(prbs), qm, tbs)
    assert [i for i in range(prb) if grant.prb_idx[0][i]] == prbs
    assert (unpacked.ndi, unpacked.harq_process, unpacked.rv_idx) == (True, 5, 0)
    L.srslte_ra_dl_dci_string.restype = C.c_char_p
    assert b"mcs=12 harq_pid=5" in L.srslte_ra_dl_dci_string(C.byref(unpacked))
    assert L.srslte_ue_dl_cfg_grant(C.byref(q), C.byref(grant), got_cfi.value, sf_idx, 0) == 0
    payload = np.zeros(tbs // 8, np.uint8)
    ret = L.srslte_pdsch_decode_rnti(C.byref(q.pdsch), C.byref(q.pdsch_cfg), C.byref(sb), C.c_void_p(q.sf_symbols), q.ce,
                                     C.c_float(0.01), C.c_uint16(rnti), payload.ctypes.data_as(C.c_void_p))
    assert ret == 0 and np.array_equal(payload, tb)
    L.srslte_softbuffer_rx_free(C.byref(sb))
    L.srslte_ue_dl_free(C.byref(q))


@pytest.mark.parametrize("cfi", [1, 2, 3])
def test_srslte_ue_dl_decode_wrapper_cfg1(gpu, oracle, cfi):
    """BASELINE configs[0]: a 1.4 MHz TM1 QPSK MCS 0 subframe through srslte_ue_dl_decode (the wrapper the north star
    names): PCFICH -> CFI on the device, grant from the caller, chest noise figure in the equaliser."""
    import ctypes as C
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import UeDl, Cell, make_grant
    prb, qm, tbs = 6, 2, 152
    ocell = o.make_cell(prb, 1, 1)
    q = UeDl()
    cell = Cell(nof_prb=prb, nof_ports=1, bw_idx=0, id=1, cp=0, phich_length=0, phich_resources=0)
    assert L.srslte_ue_dl_init(C.byref(q), cell) == 0
    L.srslte_ue_dl_set_rnti(C.byref(q), 0x1234)
    L.srslte_sch_set_max_noi(C.byref(q.pdsch.dl_sch), 4)
    grant = make_grant(prb, qm, tbs)
    assert L.srsue_gpu_ue_dl_set_grant(C.byref(q), C.byref(grant), 0, 0) == 0        # cfi 0: from the PCFICH
    for tti in (1, 12, 27):
        ocfg = o.make_cfg(ocell, sf_idx=tti % 10, cfi=cfi, qm=qm, tbs=tbs)
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 900 + tti, 10.0, pcfich=True)
        data = np.zeros(tbs // 8, np.uint8)
        n = L.srslte_ue_dl_decode(C.byref(q), iq.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p), tti)
        rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, iq, 0.01, 1, 4)
        assert rc == 0 and n == tbs
        assert np.array_equal(data, pl) and np.array_equal(data, tb)
    assert q.pkts_total == 3 and q.pkt_errors == 0
    L.srslte_ue_dl_free(C.byref(q))


@pytest.mark.parametrize("prb,rnti", [(25, 0x4601), (6, 0xFFFF), (50, 0x0102), (15, 0x0203)])
def test_srslte_ue_dl_decode_finds_its_own_grant(gpu, oracle, prb, rnti):
    """srslte_ue_dl_decode with no grant from the caller: CFI from the PCFICH, blind PDCCH search for the RNTI, DCI ->
    grant through the installed size table, PDSCH decode.  The SI-RNTI case takes the format 1A N_PRB^1A column."""
    import ctypes as C
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import UeDl, Cell, DciMsg, RaDlDci, install_tbs_table
    si = rnti == 0xFFFF
    cfi, tbs, mcs = (3, 296, 5) if si else (2, 2216, 9) if prb == 25 else (1, 776, 6) if prb == 50 else (2, 680, 7)
    L.srsue_gpu_host_dvrb_to_prb.restype = C.c_uint32
    ocell = o.make_cell(prb, 1, 1)
    q = UeDl()
    cell = Cell(nof_prb=prb, nof_ports=1, bw_idx=0, id=1, cp=0, phich_length=0, phich_resources=2)
    assert L.srslte_ue_dl_init(C.byref(q), cell) == 0
    L.srslte_ue_dl_set_rnti(C.byref(q), rnti)
    L.srslte_sch_set_max_noi(C.byref(q.pdsch.dl_sch), 4)
    for tti in (4, 13, 26):
        sf_idx = tti % 10
        sent = RaDlDci()
        sent.mcs_idx, sent.rv_idx = mcs, 0
        if prb == 50:
            # format 1, allocation type 1 (36.213 7.1.6.2): RBG subset 1 of P = 3, shifted, every other VRB of it
            P, p, n1 = 3, 1, 17 - 2 - 1
            sent.alloc_type, sent.type1_alloc.rbg_subset, sent.type1_alloc.shift = 1, p, True
            sent.type1_alloc.vrb_bitmask = int("10" * (n1 // 2), 2)
            subset = [i for i in range(prb) if (i // P) % P == p]
            prbs = [subset[i + len(subset) - n1] for i in range(n1) if i % 2 == 0]
            ln, fmt = len(prbs), 1
        elif prb == 15:
            # format 1A, distributed virtual resource blocks (36.211 6.2.3.2): the two slots use different PRBs
            start, ln = tti % 4, 6
            sent.alloc_type, sent.type2_alloc.mode = 2, 1
            sent.type2_alloc.RB_start, sent.type2_alloc.L_crb = start, ln
            prbs = sorted(L.srsue_gpu_host_dvrb_to_prb(prb, 0, v, 0) for v in range(start, start + ln))
            prbs1 = sorted(L.srsue_gpu_host_dvrb_to_prb(prb, 0, v, 1) for v in range(start, start + ln))
            assert prbs != prbs1
            fmt = 2
        else:
            start, ln = (1, 4) if si else (tti % 5, prb - 6)
            sent.alloc_type = 2
            sent.type2_alloc.RB_start, sent.type2_alloc.L_crb, sent.type2_alloc.n_prb1a = start, ln, 1
            prbs, fmt = list(range(start, start + ln)), 2
        if prb != 15:
            prbs1 = prbs
        install_tbs_table(L, {(mcs, 3 if si else ln): tbs})
        m = DciMsg()
        nb = L.srslte_dci_msg_pack_pdsch(C.byref(sent), fmt, C.byref(m), prb, not si)
        assert nb > 0
        rk, _ = o.pdcch_regs(ocell, cfi, 6)
        ss = o.pdcch_search_space(len(rk) // 9, sf_idx, rnti) if not si else [(4, 0)]
        ocfg = o.make_cfg(ocell, sf_idx=sf_idx, cfi=cfi, rnti=rnti, qm=2, tbs=tbs, prbs=prbs, prbs_slot1=prbs1)
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 4000 + tti, 12.0, None, pcfich=True,
                                   dcis=[(np.frombuffer(m.data, np.uint8)[:nb].copy(), rnti, ss[0][0], ss[0][1])])
        data = np.zeros(tbs // 8, np.uint8)
        n = L.srslte_ue_dl_decode(C.byref(q), iq.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p), tti)
        assert n == tbs and np.array_equal(data, tb)
        assert q.pdsch_cfg.grant.nof_prb == ln and q.pdsch_cfg.grant.mcs.idx == mcs
        assert [i for i in range(prb) if q.pdsch_cfg.grant.prb_idx[0][i]] == prbs
        assert [i for i in range(prb) if q.pdsch_cfg.grant.prb_idx[1][i]] == prbs1
        rc_o, pl_o, _, _ = o.ue_dl_decode(ocell, ocfg, iq, 0.01, 1, 4)
        assert rc_o == 0 and np.array_equal(data, pl_o)
        # a subframe without a DCI for this RNTI decodes nothing
        _, iq0, _ = o.gen_subframe(ocell, ocfg, 4100 + tti, 12.0, None, pcfich=True)
        assert L.srslte_ue_dl_decode(C.byref(q), iq0.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p), tti) == 0
    assert q.pkts_total == 3 and q.pkt_errors == 0
    L.srslte_ue_dl_free(C.byref(q))


@pytest.mark.parametrize("prb,cfo", [(25, 0.12), (75, -0.31)])
def test_worker_sequence_with_carrier_offset(gpu, oracle, prb, cfo):
    """phch_worker::set_cfo (phch_worker.cc:120) -> srsue_gpu_ue_dl_set_cfo: a 64QAM subframe received with a carrier
    offset decodes once the offset is handed to the worker, equals the oracle's rotate-then-transform bit for bit, and
    is lost without the correction."""
    import ctypes as C
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import UeDl, Cell, SoftBuffer, make_grant
    qm, tbs, cfi, sf_idx, rnti = 6, 12960 if prb == 25 else 39232, 2, 4, 0x1234
    n = o.lib().lteo_symbol_sz(prb)
    ocell = o.make_cell(prb, 1, 1)
    ocfg = o.make_cfg(ocell, sf_idx=sf_idx, cfi=cfi, rnti=rnti, qm=qm, tbs=tbs)
    tb, iq, _ = o.gen_subframe(ocell, ocfg, 77, 30.0)
    rx = (iq.astype(np.complex128) * np.exp(2j * np.pi * cfo * np.arange(len(iq)) / n)).astype(np.complex64)
    q = UeDl()
    cell = Cell(nof_prb=prb, nof_ports=1, bw_idx=0, id=1, cp=0, phich_length=0, phich_resources=2)
    assert L.srslte_ue_dl_init(C.byref(q), cell) == 0
    L.srslte_ue_dl_set_rnti(C.byref(q), rnti)
    L.srsue_gpu_ue_dl_set_cfi(C.byref(q), cfi)
    sb = SoftBuffer()
    assert L.srslte_softbuffer_rx_init(C.byref(sb), prb) == 0
    grant = make_grant(prb, qm, tbs)
    nsc = 12 * prb

    def run():
        got_cfi = C.c_uint32(0)
        L.srslte_softbuffer_rx_reset(C.byref(sb))
        assert L.srslte_ue_dl_decode_fft_estimate(C.byref(q), rx.ctypes.data_as(C.c_void_p), sf_idx, C.byref(got_cfi)) == 0
        sf = np.ctypeslib.as_array(C.cast(q.sf_symbols, C.POINTER(C.c_float)), (14 * nsc * 2,)).view(np.complex64).copy()
        assert L.srslte_ue_dl_cfg_grant(C.byref(q), C.byref(grant), cfi, sf_idx, 0) == 0
        payload = np.zeros(tbs // 8, np.uint8)
        ret = L.srslte_pdsch_decode_rnti(C.byref(q.pdsch), C.byref(q.pdsch_cfg), C.byref(sb), C.c_void_p(q.sf_symbols), q.ce,
                                         C.c_float(0.01), C.c_uint16(rnti), payload.ctypes.data_as(C.c_void_p))
        return ret, payload, sf

    ret, payload, sf = run()
    assert ret != 0 and np.array_equal(sf, o.ofdm_rx(prb, rx))                 # uncorrected: inter-carrier interference
    assert L.srsue_gpu_ue_dl_set_cfo(C.byref(q), C.c_float(cfo)) == 0
    ret, payload, sf = run()
    sf_o = o.ofdm_rx(prb, o.cfo_correct(rx, o.cfo_step(cfo, n)))
    assert np.array_equal(sf, sf_o)
    ce_o, _ = o.chest(ocell, sf_idx, sf_o)
    res_o = o.pdsch_decode(ocell, ocfg, sf_o, ce_o, 0.01, 4)
    rc_o, pl_o = res_o[0], res_o[1]
    assert ret == 0 and rc_o == 0 and np.array_equal(payload, tb) and np.array_equal(payload, pl_o)
    assert L.srsue_gpu_ue_dl_set_cfo(C.byref(q), C.c_float(0.0)) == 0
    assert run()[0] != 0
    assert L.srsue_gpu_ue_dl_set_cfo(C.byref(q), C.c_float(1.5)) != 0
    L.srslte_softbuffer_rx_free(C.byref(sb))
    L.srslte_ue_dl_free(C.byref(q))


@pytest.mark.parametrize("prb,ports,qm,tbs", [(100, 1, 6, 75376), (25, 2, 4, 6200)])
def test_batch_host_sc16_input(gpu, oracle, prb, ports, qm, tbs):
    """srsue_gpu_pdsch_plan_set_iq_format(SC16): the batched host call fed with int16 captures gives exactly what the
    oracle gives on the same samples converted to float first (srsLTE's order), and switching back to cf32 works."""
    sg, ctx = gpu
    o = oracle
    n = 6
    ocell = o.make_cell(prb, ports, 1)
    ocfg = o.make_cfg(ocell, sf_idx=2, cfi=1, qm=qm, tbs=tbs, tm=ports)
    cell = sg.make_cell(prb, ports, 1)
    cfg = sg.make_cfg(cell, sf_idx=2, cfi=1, qm=qm, tbs=tbs, tm=ports)
    sent, iqs = [], []
    for i in range(n):
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 9100 + i, 30.0)
        sent.append(tb); iqs.append(iq)
    iq = np.stack(iqs)
    peak = float(np.abs(iq.view(np.float32)).max())
    scale = np.float32(peak / 32000.0)
    q16 = np.rint(iq.view(np.float32) / scale).astype(np.int16)                      # the capture: (n, sf_len * 2) int16
    xf = (q16.astype(np.float32) * scale).view(np.complex64)
    plan = sg.PdschPlan(ctx, cell, cfg, n)
    plan.set_iq_format(True, float(scale))
    h_pl = np.zeros((n, plan.info.payload_stride), np.uint8)
    h_st = np.zeros((n, 4), np.int32)
    h_meas = np.zeros((n, 5), np.float32)
    plan.decode_batch_host(n, q16, 0.01, 1, 4, h_pl, h_st, h_meas)
    for i in range(n):
        rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, xf[i], 0.01, 1, 4)
        assert rc == 0 and h_st[i, 0] == 1
        assert np.array_equal(h_pl[i, :tbs // 8], pl) and np.array_equal(pl, sent[i])
        assert np.allclose(h_meas[i], meas, rtol=1e-4)
    # the same captures with a carrier offset each, removed through the plan's per-subframe steps
    import torch
    nfft = o.lib().lteo_symbol_sz(prb)
    cfos = [0.0, 0.2, -0.33, 0.05, 0.41, -0.12]
    rot = np.stack([(iq[i].astype(np.complex128) * np.exp(2j * np.pi * cfos[i] * np.arange(iq.shape[1]) / nfft)).astype(np.complex64)
                    for i in range(n)])
    r16 = np.rint(rot.view(np.float32) / scale).astype(np.int16)
    steps = np.array([sg.host_cfo_step(c, nfft) for c in cfos], np.int32)
    d_steps = torch.from_numpy(steps).cuda()
    plan.set_cfo(d_steps)
    h_pl[:] = 0
    plan.decode_batch_host(n, r16, 0.01, 1, 4, h_pl, h_st, h_meas)
    for i in range(n):
        xr = o.cfo_correct((r16[i].astype(np.float32) * scale).view(np.complex64), int(steps[i]))
        rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, xr, 0.01, 1, 4)
        assert rc == 0 and h_st[i, 0] == 1 and np.array_equal(h_pl[i, :tbs // 8], pl) and np.array_equal(pl, sent[i])
    plan.set_cfo(None, 0)
    plan.set_iq_format(False)
    h_pl[:] = 0
    plan.decode_batch_host(n, iq, 0.01, 1, 4, h_pl, h_st, h_meas)
    assert all(np.array_equal(h_pl[i, :tbs // 8], sent[i]) for i in range(n))
    plan.close()


def test_cpp_offline_driver_worker_and_batch(gpu, oracle, tmp_path):
    """driver/pdsch_offline.cc: the C++ host side replaying phch_worker's call sequence (mode worker) and the
    batched call (mode batch) must both reproduce the oracle's transport blocks, CRC verdicts and iterations."""
    import os, struct, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "build", "pdsch_offline")
    if not os.path.exists(exe):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(root, "include"), "-o", exe,
                               os.path.join(root, "driver", "pdsch_offline.cc"), "-L" + os.path.join(root, "srsue_b200"),
                               "-lsrsue_gpu", "-Wl,-rpath," + os.path.join(root, "srsue_b200")])
    o = oracle
    prb, qm, tbs, n = 50, 6, 21384, 5
    ocell = o.make_cell(prb, 1, 1)
    ocfg = o.make_cfg(ocell, sf_idx=1, cfi=1, qm=qm, tbs=tbs)
    iq = np.stack([o.gen_subframe(ocell, ocfg, 600 + i, 18.0 if i != 2 else 2.0)[1] for i in range(n)])
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<12i", 0x53525355, prb, 1, 1, 1, 1, 0x1234, qm, tbs, 0, n, 4))
        f.write(iq.tobytes())
    ref = [o.ue_dl_decode(ocell, ocfg, iq[i], 0.01, 0, 4) for i in range(n)]
    import torch
    modes = [("worker", []), ("batch", []), ("batch", ["--gpus", str(min(2, torch.cuda.device_count()))])]   # the last: multi-GPU handle
    for mode, extra in modes:
        outp = tmp_path / ("out_%s%s.bin" % (mode, "_multi" if extra else ""))
        subprocess.check_call([exe, mode, str(inp), str(outp)] + extra)
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


def _random_cases(n, seed):
    """random cells / grants: every bandwidth, 1-2 ports, all modulations, partial and scattered PRB allocations, every
    subframe number (0 and 5 lose PSS/SSS/PBCH REs), every CFI, random cell ids and RNTIs, code rates 0.1 .. 0.9"""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        prb = int(rng.choice([6, 15, 25, 50, 75, 100]))
        ports = int(rng.integers(1, 3))
        qm = int(rng.choice([2, 4, 6]))
        cid = int(rng.integers(0, 504))
        sf = int(rng.integers(0, 10))
        cfi = int(rng.integers(1, 4))
        nalloc = int(rng.integers(max(1, prb // 8), prb + 1))
        prbs = sorted(rng.choice(prb, nalloc, replace=False).tolist()) if rng.random() < 0.5 else list(range(int(rng.integers(0, prb - nalloc + 1)), 0))
        if not prbs:
            start = int(rng.integers(0, prb - nalloc + 1))
            prbs = list(range(start, start + nalloc))
        out.append(dict(prb=prb, ports=ports, qm=qm, cid=cid, sf=sf, cfi=cfi, prbs=prbs, rnti=int(rng.integers(1, 65520)),
                        rate=float(rng.uniform(0.1, 0.9)), seed=int(rng.integers(1, 1 << 30))))
    return out


@pytest.mark.parametrize("case", _random_cases(24, 20261018), ids=lambda c: "%dprb_%dp_qm%d_sf%d_cfi%d_n%d" % (
    c["prb"], c["ports"], c["qm"], c["sf"], c["cfi"], len(c["prbs"])))
def test_random_grants_match_oracle(gpu, oracle, case):
    """whole chain on randomly drawn cells and grants, noise around the decoding threshold of each so that both CRC
    verdicts occur: payload, verdict, iteration count and measurements equal the oracle's for every subframe"""
    import torch
    sg, ctx = gpu
    o = oracle
    c = case
    tm = 1 if c["ports"] == 1 else 2
    ocell = o.make_cell(c["prb"], c["ports"], c["cid"])
    probe = o.make_cfg(ocell, sf_idx=c["sf"], cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=40, tm=tm, prbs=c["prbs"])
    nre = len(o.pdsch_re_list(ocell, probe))
    if tm == 2:
        nre -= nre % 2
    tbs = max(40, int(c["rate"] * nre * c["qm"]) // 8 * 8 - 24)
    tbs = min(tbs, 75376)
    ocfg = o.make_cfg(ocell, sf_idx=c["sf"], cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=tbs, tm=tm, prbs=c["prbs"])
    cell = sg.make_cell(c["prb"], c["ports"], c["cid"])
    cfg = sg.make_cfg(cell, sf_idx=c["sf"], cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=tbs, tm=tm, prbs=c["prbs"])
    # Es/N0 near the Shannon limit of the spectral efficiency + margin, one subframe above and one below
    eff = tbs / max(nre, 1)
    base = 10 * np.log10(2 ** eff - 1) + 2.5
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


@pytest.mark.parametrize("kind", ["zeros", "noise_zf", "huge", "tiny"])
def test_degenerate_inputs_match_oracle(gpu, oracle, kind):
    """inputs a receiver sees in practice but a generator never makes: silence (0/0 in the zero-forcing equaliser -> NaN ->
    LLR 0), pure noise with noise_estimate 0, samples large enough to saturate every LLR, samples near the denormal
    range.  Verdicts, payload bytes, iteration counts and (NaN-aware) measurements must still equal the oracle's."""
    sg, ctx = gpu
    o = oracle
    prb, qm, tbs = 25, 4, 4968
    ocell = o.make_cell(prb, 2, 9)
    ocfg = o.make_cfg(ocell, sf_idx=4, cfi=2, qm=qm, tbs=tbs, tm=2)
    cell = sg.make_cell(prb, 2, 9)
    cfg = sg.make_cfg(cell, sf_idx=4, cfi=2, qm=qm, tbs=tbs, tm=2)
    n_samp = 15 * 512
    rng = np.random.default_rng(3)
    good = o.gen_subframe(ocell, ocfg, 77, 25.0, _taps())[1]
    if kind == "zeros":
        iq, n0 = np.zeros(n_samp, np.complex64), 0.0
    elif kind == "noise_zf":
        iq, n0 = (rng.standard_normal(n_samp) + 1j * rng.standard_normal(n_samp)).astype(np.complex64), 0.0
    elif kind == "huge":
        iq, n0 = (good * np.float32(3e18)).astype(np.complex64), 0.01
    else:
        iq, n0 = (good * np.float32(1e-30)).astype(np.complex64), 0.0
    iq2 = np.stack([iq, good])
    plan = sg.PdschPlan(ctx, cell, cfg, 2)
    I = plan.info
    h_pl = np.zeros((2, I.payload_stride), np.uint8)
    h_st = np.zeros((2, 4), np.int32)
    h_meas = np.zeros((2, 5), np.float32)
    plan.decode_batch_host(2, iq2, n0, 0, 4, h_pl, h_st, h_meas)
    for i in range(2):
        rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, iq2[i], n0, 0, 4)
        assert (h_st[i, 0] == 1) == (rc == 0) and np.array_equal(h_pl[i], pl) and h_st[i, 2] == avg
        assert np.allclose(h_meas[i], meas, rtol=1e-4, equal_nan=True)
    plan.close()


def test_four_concurrent_workers(gpu, oracle):
    """srsUE runs up to four phch_workers at once (ue/hdr/phy/phy.h:118), each with its own srslte_ue_dl_t; the shim gives
    every ue_dl its own stream, plans and decoder scratch.  Four threads decode different bandwidth-25 subframes through
    the srsLTE-shaped calls at the same time, several rounds; every result must equal the oracle's."""
    import ctypes as C
    import threading
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import UeDl, Cell, SoftBuffer, make_grant
    prb = 25
    shapes = [(2, 1384), (4, 4968), (6, 11448), (4, 2216)]          # (Qm, TBS) per worker
    rounds = 6
    work, errors = [], []
    for w, (qm, tbs) in enumerate(shapes):
        ocell = o.make_cell(prb, 1, 10 + w)
        items = []
        for r in range(rounds):
            ocfg = o.make_cfg(ocell, sf_idx=(w + r) % 10, cfi=1 + (r % 3), rnti=0x100 + w, qm=qm, tbs=tbs)
            tb, iq, _ = o.gen_subframe(ocell, ocfg, 8000 + 10 * w + r, 22.0, None, pcfich=True)
            rc, pl, meas, avg = o.ue_dl_decode(ocell, ocfg, iq, 0.01, 0, 4)
            items.append((ocfg, iq, rc, pl, avg))
        work.append((ocell, qm, tbs, items))

    def worker(w):
        try:
            ocell, qm, tbs, items = work[w]
            q = UeDl()
            cell = Cell(nof_prb=prb, nof_ports=1, bw_idx=0, id=10 + w, cp=0, phich_length=0, phich_resources=2)
            assert L.srslte_ue_dl_init(C.byref(q), cell) == 0
            L.srslte_ue_dl_set_rnti(C.byref(q), 0x100 + w)
            sb = SoftBuffer()
            assert L.srslte_softbuffer_rx_init(C.byref(sb), prb) == 0
            grant = make_grant(prb, qm, tbs)
            for _ in range(3):                                    # three passes over the items keeps all four busy
                for ocfg, iq, rc, pl, avg in items:
                    cfi = C.c_uint32(0)
                    assert L.srslte_ue_dl_decode_fft_estimate(C.byref(q), iq.ctypes.data_as(C.c_void_p), ocfg.sf_idx, C.byref(cfi)) == 0
                    assert cfi.value == ocfg.cfi
                    L.srslte_softbuffer_rx_reset(C.byref(sb))
                    assert L.srslte_ue_dl_cfg_grant(C.byref(q), C.byref(grant), cfi.value, ocfg.sf_idx, 0) == 0
                    payload = np.zeros(tbs // 8, np.uint8)
                    ret = L.srslte_pdsch_decode_rnti(C.byref(q.pdsch), C.byref(q.pdsch_cfg), C.byref(sb), C.c_void_p(q.sf_symbols), q.ce,
                                                     C.c_float(0.01), C.c_uint16(0x100 + w), payload.ctypes.data_as(C.c_void_p))
                    assert (ret == 0) == (rc == 0) and np.array_equal(payload, pl)
                    assert L.srslte_pdsch_last_noi(C.byref(q.pdsch)) == avg
            L.srslte_softbuffer_rx_free(C.byref(sb))
            L.srslte_ue_dl_free(C.byref(q))
        except Exception as e:  # noqa: BLE001
            errors.append((w, repr(e)))

    threads = [threading.Thread(target=worker, args=(w,)) for w in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
