"""Four-port cells (nof_ports from the MIB, /root/reference/ue/src/phy/phch_recv.cc:210): the oracle's additions against
independent constructions from 36.211 -- CRS of ports 2 / 3 (6.10.1.2), the precoding matrix of four-port transmit
diversity (6.3.4.3, SFBC-FSTD), the RE list, the control region (symbol 1 holds six-element REGs), the
PHICH port alternation (6.9.2) -- TX -> RX round trips, and the library's host tables against the oracle.  SPEC.md 15c.  CPU only."""
import ctypes as C

import numpy as np
import pytest


def test_crs_ports_2_and_3_equal_standard_definition(oracle):
    lib = oracle.lib()
    for cp in (0, 1):
        nslot = 6 if cp else 7
        for prb, cid in ((6, 0), (25, 77), (100, 503)):
            cell = oracle.make_cell(prb, 4, cid, cp=cp)
            for l in range(2 * nslot):
                ns, ls = l // nslot, l % nslot
                for port in (2, 3):
                    k = np.zeros(2 * prb, np.int32)
                    n = lib.lteo_crs_positions(C.byref(cell), port, l, k.ctypes.data_as(C.c_void_p))
                    if ls != 1:
                        assert n == 0
                        continue
                    v = 3 * (ns % 2) if port == 2 else 3 + 3 * (ns % 2)
                    assert n == 2 * prb and k.tolist() == [6 * m + (v + cid % 6) % 6 for m in range(2 * prb)]
            # the sequence of symbol 1 is the general r_{l, n_s}(m)
            for sf in (0, 7):
                for l in (1, nslot + 1):
                    ns, ls = 2 * sf + l // nslot, l % nslot
                    c = oracle.gold((1 << 10) * (7 * (ns + 1) + ls + 1) * (2 * cid + 1) + 2 * cid + (0 if cp else 1), 440)
                    re, im = np.zeros(2 * prb, np.int8), np.zeros(2 * prb, np.int8)
                    lib.lteo_crs_values(C.byref(cell), sf, l, re.ctypes.data_as(C.c_void_p), im.ctypes.data_as(C.c_void_p))
                    mp = np.arange(2 * prb) + 110 - prb
                    assert np.array_equal(re, 1 - 2 * c[2 * mp].astype(np.int8)) and np.array_equal(im, 1 - 2 * c[2 * mp + 1].astype(np.int8))


def test_four_port_diversity_equals_the_precoding_matrix(oracle):
    """36.211 6.3.4.3, four ports: of every four consecutive symbols x0..x3 on subcarriers k0..k3, port 0 sends
    (x0, x1, 0, 0), port 2 (-x1*, x0*, 0, 0), port 1 (0, 0, x2, x3), port 3 (0, 0, -x3*, x2*), all scaled by 1/sqrt(2)"""
    cell = oracle.make_cell(6, 4, 3)
    cfg = oracle.make_cfg(cell, sf_idx=1, cfi=3, qm=2, tbs=104, tm=2)
    tb = np.arange(13, dtype=np.uint8)
    grid = oracle.pdsch_tx_grid(cell, cfg, tb).reshape(4, 14 * 72)
    re = oracle.pdsch_re_list(cell, cfg)
    assert len(re) % 4 == 0
    two = oracle.make_cell(6, 2, 3)
    # the same bits through the two-port mapper (pinned in test_oracle.py) give the symbols x: port 0 of that grid holds
    # x_i / sqrt(2) on RE i of ITS list; rebuild x from the four-port grid instead: ports 0 / 1 carry x directly
    x = np.zeros(len(re), np.complex128)
    for i in range(0, len(re), 4):
        x[i:i + 2] = grid[0, re[i:i + 2]] * np.sqrt(2)
        x[i + 2:i + 4] = grid[1, re[i + 2:i + 4]] * np.sqrt(2)
    assert np.allclose(abs(x), 1.0)                      # QPSK symbols of unit power
    a = 1 / np.sqrt(2)
    for i in range(0, len(re), 4):
        k = re[i:i + 4]
        assert np.allclose(grid[0, k], [x[i] * a, x[i + 1] * a, 0, 0])
        assert np.allclose(grid[2, k], [-np.conj(x[i + 1]) * a, np.conj(x[i]) * a, 0, 0])
        assert np.allclose(grid[1, k], [0, 0, x[i + 2] * a, x[i + 3] * a])
        assert np.allclose(grid[3, k], [0, 0, -np.conj(x[i + 3]) * a, np.conj(x[i + 2]) * a])
    # and x are the modulation symbols of the coded bits: the two-port grid of the same cell id / grant holds them on its
    # port 0 wherever both lists coincide (symbols without CRS of ports 2 / 3)
    g2 = oracle.pdsch_tx_grid(two, oracle.make_cfg(two, sf_idx=1, cfi=3, qm=2, tbs=104, tm=2), tb).reshape(2, 14 * 72)
    re2 = oracle.pdsch_re_list(two, oracle.make_cfg(two, sf_idx=1, cfi=3, qm=2, tbs=104, tm=2))
    assert len(re2) - len(re) == 4 * 6                  # symbol 8 loses four REs per PRB (symbol 1 is control)
    n_same = int(np.argmax(re != re2[:len(re)])) if (re != re2[:len(re)]).any() else len(re)
    assert n_same == 48 + 72 + 72 + 48                  # symbols 4..7 precede the first hole (symbol 8)
    assert np.allclose(g2[0, re2[:n_same]] * np.sqrt(2), x[:n_same])


def test_pdsch_resource_elements_four_ports(oracle):
    for cp in (0, 1):
        nslot = 6 if cp else 7
        for prb, cid, sf, cfi in ((6, 3, 0, 3), (25, 77, 5, 1), (100, 500, 1, 2)):
            cell = oracle.make_cell(prb, 4, cid, cp=cp)
            cfg = oracle.make_cfg(cell, sf_idx=sf, cfi=cfi, tm=2)
            nsc, exp = 12 * prb, []
            for l in range(cfi + (1 if prb <= 10 else 0), 2 * nslot):
                ls = l % nslot
                crs = ls in (0, 1, nslot - 3)
                mid = (sf in (0, 5) and l in (nslot - 2, nslot - 1)) or (sf == 0 and nslot <= l <= nslot + 3)
                for k in range(nsc):
                    if crs and k % 3 == cid % 3:
                        continue
                    if mid and nsc // 2 - 36 <= k < nsc // 2 + 36:
                        continue
                    exp.append(l * nsc + k)
            assert oracle.pdsch_re_list(cell, cfg).tolist() == exp


def test_control_region_four_ports(oracle):
    for cp in (0, 1):
        for prb, cid in ((6, 7), (25, 100)):
            four, two = oracle.make_cell(prb, 4, cid, cp=cp), oracle.make_cell(prb, 2, cid, cp=cp)
            for cfi in (1, 2, 3):
                rk4, rl4 = oracle.pdcch_regs(four, cfi)
                rk2, rl2 = oracle.pdcch_regs(two, cfi)
                nsym = cfi + (1 if prb <= 10 else 0)
                if nsym >= 2:
                    assert (rl4 == 1).sum() == 2 * prb and (rl2 == 1).sum() == 3 * prb
                    assert set(rk4[rl4 == 1].tolist()) == set(range(0, 12 * prb, 6))
                keep4, keep2 = rl4 != 1, rl2 != 1
                assert sorted(zip(rk4[keep4].tolist(), rl4[keep4].tolist())) == sorted(zip(rk2[keep2].tolist(), rl2[keep2].tolist()))
                assert sorted(zip(rk4.tolist(), rl4.tolist())) == list(zip(rk4.tolist(), rl4.tolist()))


def _taps4(seed=5):
    rng = np.random.default_rng(seed)
    taps = (rng.standard_normal((4, 5)) + 1j * rng.standard_normal((4, 5))) * np.array([1, .6, .4, .2, .1])
    return taps / np.sqrt((abs(taps) ** 2).sum(1, keepdims=True))


@pytest.mark.parametrize("cp", [0, 1])
@pytest.mark.parametrize("prb,qm,tbs,sf", [(6, 2, 104, 0), (25, 4, 3240, 5), (75, 4, 12960, 3), (100, 6, 46888, 1)])
def test_tx_rx_round_trip_four_ports(oracle, cp, prb, qm, tbs, sf):
    """four independent channels, one per port: the transport block only comes back if every pair is combined with the
    estimates of the two ports it was sent on; PCFICH, PDCCH, PHICH and PBCH ride on the same subframe"""
    o = oracle
    cid, cfi, rnti, nb = 77, 2, 0x2345, 25
    cell = o.make_cell(prb, 4, cid, cp=cp)
    cfg = o.make_cfg(cell, sf_idx=sf, cfi=cfi, qm=qm, tbs=tbs, tm=2)
    rk, _ = o.pdcch_regs(cell, cfi, 6)
    ss = o.pdcch_search_space(len(rk) // 9, sf, rnti)
    bits = np.random.default_rng(prb).integers(0, 2, nb, dtype=np.uint8)
    units = (6 * prb + 47) // 48
    ph = [(0, 0, 1), (0, 2, 0), ((2 if cp else 1) * units - 1, 1, 1)]
    mib = o.mib_pack(prb, 0, 6, 401)
    tb, iq, s2 = o.gen_subframe(cell, cfg, 7, 26.0, _taps4(), pcfich=True, dcis=[(bits, rnti) + ss[-1]], phichs=ph,
                                mib=(mib, 1) if sf == 0 else None)
    rc, pl, meas, _ = o.ue_dl_decode(cell, cfg, iq, s2, 0, 4)
    assert rc == 0 and np.array_equal(pl, tb)
    sfo = o.ofdm_rx(prb, iq, cp=cp)
    ce, m = o.chest(cell, sf, sfo)
    assert o.pcfich_decode(cell, sf, sfo, ce, m[0])[0] == cfi
    llr, nc = o.pdcch_extract_llr(cell, sf, cfi, sfo, ce, m[0])
    f, out, L1, n1 = o.pdcch_find_dci(llr, nc, sf, rnti, nb)
    assert f == 1 and np.array_equal(out, bits)            # (a lower aggregation level at the same CCE decodes too)
    assert [o.phich_decode(cell, sf, sfo, ce, g, q, float(m[0]))[0] for g, q, _ in ph] == [a for _, _, a in ph]
    if sf == 0:
        f, got, ports, off = o.pbch_decode(cell, sfo, ce, float(m[0]))
        assert f == 1 and ports == 4 and off == 1 and np.array_equal(got, mib)
    # read as a two-port cell the same samples do not decode
    two = o.make_cell(prb, 2, cid, cp=cp)
    assert o.ue_dl_decode(two, o.make_cfg(two, sf_idx=sf, cfi=cfi, qm=qm, tbs=tbs, tm=2), iq, s2, 0, 4)[0] != 0


def test_phich_port_alternation_four_ports(oracle):
    """36.211 6.9.2: quadruplet i of group g goes out on ports (0, 2) when i + g (extended prefix: i + g / 2) is even and on
    (1, 3) otherwise"""
    o = oracle
    for cp in (0, 1):
        cell = o.make_cell(25, 4, 9, cp=cp)
        for g in range(4):
            grid = np.zeros((4, 14, 300), np.complex128)
            o.lib().lteo_phich_tx(C.byref(cell), 3, 6, g, 1, 1, grid.ctypes.data_as(C.c_void_p))
            k = o.phich_res(cell, g)
            for i in range(3):
                par = (i + (g // 2 if cp else g)) % 2
                on = abs(grid[:, 0, k[4 * i:4 * i + 4]]).sum(1) > 0
                assert on.tolist() == ([True, False, True, False] if par == 0 else [False, True, False, True])


def test_host_tables_four_ports_agree_with_oracle(oracle):
    import srsue_b200 as sg
    lib = sg.lib()
    for cp in (0, 1):
        for prb, sf, cfi in ((6, 0, 3), (25, 5, 2), (100, 1, 1), (50, 9, 3)):
            cell, ocell = sg.make_cell(prb, 4, 7, cp=cp), oracle.make_cell(prb, 4, 7, cp=cp)
            cfg, ocfg = sg.make_cfg(cell, sf_idx=sf, cfi=cfi, tm=2), oracle.make_cfg(ocell, sf_idx=sf, cfi=cfi, tm=2)
            re = np.zeros(14 * 12 * prb, np.int32)
            n = lib.srsue_gpu_host_pdsch_re(C.byref(cell), C.byref(cfg), re.ctypes.data_as(C.c_void_p))
            ref = oracle.pdsch_re_list(ocell, ocfg)
            assert n == len(ref) and np.array_equal(re[:n], ref)
            for c in (1, 2, 3):
                rk, rl = oracle.pdcch_regs(ocell, c, 6)
                re4 = np.zeros(4 * 12 * prb, np.int32)
                m = lib.srsue_gpu_host_pdcch_regs(C.byref(cell), c, 6, re4.ctypes.data_as(C.c_void_p))
                assert m == len(rk)
                lo = oracle.lib()
                exp = []
                for k0, l in zip(rk.tolist(), rl.tolist()):
                    k4 = np.zeros(4, np.int32)
                    lo.lteo_reg_res(C.byref(ocell), k0, l, k4.ctypes.data_as(C.c_void_p))
                    exp += [l * 12 * prb + int(k) for k in k4]
                assert re4[:4 * m].tolist() == exp


def test_host_tables_four_ports_all_shifts(oracle):
    """the PDSCH RE list and the control-region REGs of the library against the oracle for every CRS frequency shift
    (cell id mod 6), every bandwidth, both prefixes and the subframes with synchronisation signals / PBCH"""
    import srsue_b200 as sg
    lib = sg.lib()
    for cp in (0, 1):
        for prb in (6, 15, 25, 50, 75, 100):
            for cid in (0, 1, 2, 3, 4, 5, 503):
                for sf, cfi in ((0, 1), (5, 2), (9, 3)):
                    cell, ocell = sg.make_cell(prb, 4, cid, cp=cp), oracle.make_cell(prb, 4, cid, cp=cp)
                    cfg, ocfg = sg.make_cfg(cell, sf_idx=sf, cfi=cfi, tm=2), oracle.make_cfg(ocell, sf_idx=sf, cfi=cfi, tm=2)
                    re = np.zeros(14 * 12 * prb, np.int32)
                    n = lib.srsue_gpu_host_pdsch_re(C.byref(cell), C.byref(cfg), re.ctypes.data_as(C.c_void_p))
                    ref = oracle.pdsch_re_list(ocell, ocfg)
                    assert n == len(ref) and np.array_equal(re[:n], ref), (cp, prb, cid, sf, cfi)
                    rk, rl = oracle.pdcch_regs(ocell, cfi, 6)
                    re4 = np.zeros(4 * 12 * prb, np.int32)
                    assert lib.srsue_gpu_host_pdcch_regs(C.byref(cell), cfi, 6, re4.ctypes.data_as(C.c_void_p)) == len(rk)
