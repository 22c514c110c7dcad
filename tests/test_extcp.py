"""Extended cyclic prefix (cell search reports it, /root/reference/ue/src/phy/phch_recv.cc:189; the cell the worker is
initialised with carries it, phch_worker.cc:74): the oracle's geometry against independent constructions written from
36.211 (6 symbols per slot, 512-sample prefixes, CRS in symbols 0 and 3 of a slot, N_CP = 0 in the CRS c_init,
PSS/SSS/PBCH positions, the four-symbol control region with CRS in symbol 3), TX -> RX round trips, and the library's
host-side tables against the oracle.  SPEC.md section 15b.  CPU only."""
import ctypes as C

import numpy as np
import pytest


def test_ofdm_extended_prefix_against_numpy_fft(oracle):
    rng = np.random.default_rng(15)
    for prb, n in ((6, 128), (15, 256), (25, 512), (50, 1024), (75, 1536), (100, 2048)):
        iq = (rng.standard_normal(15 * n) + 1j * rng.standard_normal(15 * n)).astype(np.complex64)
        sf = oracle.ofdm_rx(prb, iq, cp=1).reshape(14, -1)
        nsc, cp = 12 * prb, n // 4                          # 512 samples at 30.72 Msps
        assert 12 * (n + cp) == 15 * n                      # six symbols fill a 0.5 ms slot exactly
        for l in range(12):
            pos = l * (n + cp) + cp
            X = np.fft.fft(iq[pos:pos + n].astype(np.complex128)) / np.sqrt(n)
            ref = np.concatenate([X[n - nsc // 2:], X[1:nsc // 2 + 1]])
            assert np.max(np.abs(sf[l] - ref)) / np.sqrt(np.mean(abs(ref) ** 2)) < 1e-5
        assert not sf[12:].any()                            # rows 12 and 13 of the 14-row grid stay untouched
        # modulator and demodulator are inverses
        grid = np.zeros((14, nsc), np.complex128)
        grid[:12] = rng.standard_normal((12, nsc)) + 1j * rng.standard_normal((12, nsc))
        back = oracle.ofdm_rx(prb, oracle.ofdm_tx(prb, grid, cp=1).astype(np.complex64), cp=1).reshape(14, -1)
        assert np.max(np.abs(back[:12] - grid[:12])) < 1e-4


def test_crs_extended_prefix_equals_standard_definition(oracle):
    """36.211 6.10.1: c_init = 2^10 (7 (n_s + 1) + l + 1)(2 N_ID + 1) + 2 N_ID + N_CP with N_CP = 0; ports 0 / 1 in
    symbols 0 and N_symb - 3 = 3 of each slot with v = 0 / 3 (port 0) and 3 / 0 (port 1)."""
    lib = oracle.lib()
    for prb, cid in ((6, 0), (25, 77), (100, 503)):
        for ports in (1, 2):
            cell = oracle.make_cell(prb, ports, cid, cp=1)
            for sf in (0, 3, 9):
                for l in range(14):
                    ns, ls = 2 * sf + l // 6, l % 6
                    k = np.zeros(2 * prb, np.int32)
                    if l >= 12 or ls not in (0, 3):
                        if l < 12:
                            assert lib.lteo_crs_positions(C.byref(cell), 0, l, k.ctypes.data_as(C.c_void_p)) == 0
                        continue
                    c_init = (1 << 10) * (7 * (ns + 1) + ls + 1) * (2 * cid + 1) + 2 * cid
                    c = oracle.gold(c_init, 440)
                    re, im = np.zeros(2 * prb, np.int8), np.zeros(2 * prb, np.int8)
                    lib.lteo_crs_values(C.byref(cell), sf, l, re.ctypes.data_as(C.c_void_p), im.ctypes.data_as(C.c_void_p))
                    mp = np.arange(2 * prb) + 110 - prb
                    assert np.array_equal(re, 1 - 2 * c[2 * mp].astype(np.int8)) and np.array_equal(im, 1 - 2 * c[2 * mp + 1].astype(np.int8))
                    for port in range(ports):
                        n = lib.lteo_crs_positions(C.byref(cell), port, l, k.ctypes.data_as(C.c_void_p))
                        v = (0 if ls == 0 else 3) if port == 0 else (3 if ls == 0 else 0)
                        assert n == 2 * prb and k.tolist() == [6 * m + (v + cid % 6) % 6 for m in range(2 * prb)]


def _re_list_from_the_text(prb, ports, cid, sf, cfi, prbs):
    """36.211 6.3.5 for the extended prefix: symbols cfi' .. 11, allocated PRBs ascending, subcarriers ascending, minus
    CRS (symbols 0 / 3 of a slot; both ports' positions when two are configured), minus the central 72 subcarriers of
    SSS / PSS (symbols 4 / 5 of slot 0 in subframes 0 and 5) and PBCH (symbols 0..3 of slot 1 in subframe 0)."""
    nsc, out = 12 * prb, []
    first = cfi + (1 if prb <= 10 else 0)
    for l in range(first, 12):
        ls = l % 6
        res = set()
        if ls in (0, 3):
            v0 = 0 if ls == 0 else 3
            res.add((v0 + cid % 6) % 6)
            if ports == 2:
                res.add((3 - v0 + cid % 6) % 6)
        mid = (sf in (0, 5) and l in (4, 5)) or (sf == 0 and 6 <= l <= 9)
        for p in prbs:
            for k in range(12 * p, 12 * p + 12):
                if k % 6 in res or (mid and nsc // 2 - 36 <= k < nsc // 2 + 36):
                    continue
                out.append(l * nsc + k)
    return out


def test_pdsch_resource_elements_extended_prefix(oracle):
    c100 = oracle.make_cell(100, 1, 1, cp=1)
    # ten data symbols; CRS of port 0 in symbols 3, 6 and 9
    assert len(oracle.pdsch_re_list(c100, oracle.make_cfg(c100, sf_idx=1, cfi=2))) == 10 * 1200 - 3 * 200
    for prb, ports, cid, sf, cfi, prbs in ((6, 1, 3, 0, 3, range(6)), (6, 2, 10, 5, 1, (1, 4)), (25, 2, 77, 0, 2, range(3, 20)),
                                          (100, 1, 503, 5, 1, range(100)), (50, 2, 5, 4, 3, range(0, 50, 2))):
        cell = oracle.make_cell(prb, ports, cid, cp=1)
        cfg = oracle.make_cfg(cell, sf_idx=sf, cfi=cfi, prbs=list(prbs), tm=ports)
        assert oracle.pdsch_re_list(cell, cfg).tolist() == _re_list_from_the_text(prb, ports, cid, sf, cfi, list(prbs))


def test_control_region_extended_prefix(oracle):
    """36.211 6.2.4 with the extended prefix: the fourth control symbol (<= 10 PRB, CFI 3) carries CRS, so it holds two
    REGs of six REs per PRB like symbol 0; the PHICH mapping units and the PCFICH do not move."""
    for cid in (0, 7, 100):
        ext, norm = oracle.make_cell(6, 2, cid, cp=1), oracle.make_cell(6, 2, cid)
        for cfi in (1, 2, 3):
            rk_e, rl_e = oracle.pdcch_regs(ext, cfi)
            rk_n, rl_n = oracle.pdcch_regs(norm, cfi)
            assert (rl_e == 3).sum() == (12 if cfi == 3 else 0) and (rl_n == 3).sum() == (18 if cfi == 3 else 0)
            assert sorted(zip(rk_e[rl_e < 3].tolist(), rl_e[rl_e < 3].tolist())) == sorted(zip(rk_n[rl_n < 3].tolist(), rl_n[rl_n < 3].tolist()))
            # mapping order: k' ascending, then l'
            assert sorted(zip(rk_e.tolist(), rl_e.tolist())) == list(zip(rk_e.tolist(), rl_e.tolist()))
            if cfi == 3:
                assert set(rk_e[rl_e == 3].tolist()) == set(range(0, 72, 6))
    big = oracle.make_cell(50, 2, 9, cp=1)
    assert np.array_equal(oracle.pdcch_regs(big, 3)[0], oracle.pdcch_regs(oracle.make_cell(50, 2, 9), 3)[0])


@pytest.mark.parametrize("prb,ports,qm,tbs,tm,snr,sf", [
    (6, 1, 2, 152, 1, 10.0, 1), (6, 1, 2, 104, 1, 16.0, 0), (25, 2, 4, 4008, 2, 18.0, 5), (50, 1, 6, 15264, 1, 28.0, 0),
    (75, 2, 4, 15264, 2, 16.0, 3), (100, 1, 6, 46888, 1, 30.0, 1), (100, 2, 6, 36696, 2, 30.0, 9)])
def test_tx_rx_round_trip_extended_prefix(oracle, prb, ports, qm, tbs, tm, snr, sf):
    cell = oracle.make_cell(prb, ports, 11, cp=1)
    cfg = oracle.make_cfg(cell, sf_idx=sf, cfi=2, qm=qm, tbs=tbs, tm=tm)
    rng = np.random.default_rng(3)
    taps = (rng.standard_normal((2, 4)) + 1j * rng.standard_normal((2, 4))) * np.array([1, .5, .3, .1])
    taps /= np.sqrt((abs(taps) ** 2).sum(1, keepdims=True))
    tb, iq, s2 = oracle.gen_subframe(cell, cfg, 4242, snr, taps if qm < 6 else None)
    rc, pl, meas, _ = oracle.ue_dl_decode(cell, cfg, iq, s2, 0, 4)
    assert rc == 0 and np.array_equal(pl, tb)
    assert abs(10 * np.log10(meas[4]) - snr) < 4.0          # the estimator's SNR follows the channel's
    # the same samples read with the wrong prefix do not decode
    wrong = oracle.make_cell(prb, ports, 11, cp=0)
    wcfg = oracle.make_cfg(wrong, sf_idx=sf, cfi=2, qm=qm, tbs=tbs, tm=tm)
    if len(oracle.pdsch_re_list(wrong, wcfg)) * qm > tbs + 24:
        assert oracle.ue_dl_decode(wrong, wcfg, iq, s2, 0, 4)[0] != 0


def test_host_tables_extended_prefix_agree_with_oracle(oracle):
    import srsue_b200 as sg
    lib = sg.lib()
    for prb, ports, sf, cfi in ((6, 1, 1, 3), (6, 2, 0, 1), (100, 1, 1, 1), (100, 2, 5, 2), (25, 2, 0, 3), (50, 1, 9, 2), (75, 2, 4, 1)):
        cell, ocell = sg.make_cell(prb, ports, 7, cp=1), oracle.make_cell(prb, ports, 7, cp=1)
        cfg, ocfg = sg.make_cfg(cell, sf_idx=sf, cfi=cfi), oracle.make_cfg(ocell, sf_idx=sf, cfi=cfi)
        re = np.zeros(14 * 12 * prb, np.int32)
        n = lib.srsue_gpu_host_pdsch_re(C.byref(cell), C.byref(cfg), re.ctypes.data_as(C.c_void_p))
        ref = oracle.pdsch_re_list(ocell, ocfg)
        assert n == len(ref) and np.array_equal(re[:n], ref)
        for c in (1, 2, 3):
            for ng in (1, 6, 12):
                rk, rl = oracle.pdcch_regs(ocell, c, ng)
                re4 = np.zeros(4 * 12 * prb, np.int32)
                m = lib.srsue_gpu_host_pdcch_regs(C.byref(cell), c, ng, re4.ctypes.data_as(C.c_void_p))
                assert m == len(rk)
                exp = []
                for k0, l in zip(rk.tolist(), rl.tolist()):
                    ks = [k0 + j for j in range(6) if (k0 + j) % 3 != 7 % 3] if l in (0, 3) and (l == 0 or prb <= 10) else [k0 + j for j in range(4)]
                    exp += [l * 12 * prb + k for k in ks]
                assert re4[:4 * m].tolist() == exp
