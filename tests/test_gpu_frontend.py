"""K1-K4 parity: OFDM demodulation, channel estimation, equaliser, demapper, descrambler and rate
de-matcher on the GPU (through the C ABI) against the CPU oracle on the same synthetic IQ.

North-star bars: float intermediates within 1e-4 of signal RMS (the op-order contract of SPEC.md
actually makes them identical, which is reported); int16 LLRs and the soft buffer bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _cfgs(o):
    rng = np.random.default_rng(77)
    taps = (rng.standard_normal((2, 6)) + 1j * rng.standard_normal((2, 6))) * np.array([1, .7, .5, .3, .2, .1])
    taps /= np.sqrt((abs(taps) ** 2).sum(1, keepdims=True))
    return {
        "cfg1": dict(prb=6, ports=1, qm=2, tbs=152, tm=1, snr=10.0, taps=None, sf=1),
        "cfg2": dict(prb=100, ports=1, qm=6, tbs=75376, tm=1, snr=30.0, taps=None, sf=1),
        "cfg3": dict(prb=100, ports=2, qm=4, tbs=30576, tm=2, snr=15.0, taps=taps, sf=1),
        "bw25_sf5": dict(prb=25, ports=1, qm=4, tbs=4968, tm=1, snr=20.0, taps=None, sf=5),
        "bw50_2p": dict(prb=50, ports=2, qm=6, tbs=21384, tm=2, snr=28.0, taps=taps, sf=0),
        "bw15": dict(prb=15, ports=1, qm=2, tbs=1008, tm=1, snr=8.0, taps=None, sf=3),
        "bw75": dict(prb=75, ports=1, qm=6, tbs=55056, tm=1, snr=30.0, taps=None, sf=2),          # 1536-point FFT
        "bw75_2p": dict(prb=75, ports=2, qm=4, tbs=22920, tm=2, snr=16.0, taps=taps, sf=5),
    }


def _rel(a, b):
    rms = np.sqrt(np.mean(np.abs(b) ** 2))
    return np.max(np.abs(a - b)) / rms


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "bw25_sf5", "bw50_2p", "bw15", "bw75", "bw75_2p"])
def test_frontend_stages_match_oracle(gpu, oracle, name):
    import torch
    sg, ctx = gpu
    o = oracle
    c = _cfgs(o)[name]
    ocell = o.make_cell(c["prb"], c["ports"], 1)
    ocfg = o.make_cfg(ocell, sf_idx=c["sf"], cfi=1, qm=c["qm"], tbs=c["tbs"], tm=c["tm"])
    n_sf = 3
    iqs, tbs_sent = [], []
    for i in range(n_sf):
        tb, iq, _ = o.gen_subframe(ocell, ocfg, 1000 + i, c["snr"], c["taps"])
        iqs.append(iq); tbs_sent.append(tb)
    iq = np.stack(iqs)
    cell = sg.make_cell(c["prb"], c["ports"], 1)
    cfg = sg.make_cfg(cell, sf_idx=c["sf"], cfi=1, qm=c["qm"], tbs=c["tbs"], tm=c["tm"])
    plan = sg.PdschPlan(ctx, cell, cfg, n_sf)
    I = plan.info
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(n_sf, -1)).cuda()
    d_sf = torch.zeros((n_sf, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((n_sf, c["ports"] * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((n_sf, 5), dtype=torch.float32, device="cuda")
    d_sb = torch.zeros((n_sf, I.sb_sf_stride), dtype=torch.int16, device="cuda")
    d_d = torch.zeros((n_sf, I.nof_re * 2), dtype=torch.float32, device="cuda")
    d_e = torch.zeros((n_sf, I.G), dtype=torch.int16, device="cuda")
    plan.ofdm_rx(n_sf, d_iq, d_sf)
    plan.chest(n_sf, d_sf, d_ce, d_meas)
    plan.pdsch_llr(n_sf, d_sf, d_ce, d_meas, 0.01, 0, 0, d_sb, d_d, d_e)
    torch.cuda.synchronize()
    sf_g = d_sf.cpu().numpy().view(np.complex64)
    ce_g = d_ce.cpu().numpy().view(np.complex64).reshape(n_sf, c["ports"], -1)
    meas_g = d_meas.cpu().numpy()
    dd_g = d_d.cpu().numpy().view(np.complex64)
    e_g = d_e.cpu().numpy()
    s = o.cbsegm(c["tbs"])
    exact = {}
    for i in range(n_sf):
        sf_o = o.ofdm_rx(c["prb"], iq[i])
        # independent anchor for the oracle itself: numpy FFT
        assert _rel(sf_g[i], sf_o) <= TOL
        ce_o, meas_o = o.chest(ocell, c["sf"], sf_o)
        assert _rel(ce_g[i], ce_o) <= TOL
        assert np.allclose(meas_g[i], meas_o, rtol=1e-4)
        rc, pl, dbg = o.pdsch_decode(ocell, ocfg, sf_o, ce_o, 0.01, 4, want=True)
        assert _rel(dd_g[i], dbg["d"][:I.nof_re]) <= TOL
        exact.setdefault("sf", []).append(np.array_equal(sf_g[i], sf_o))
        exact.setdefault("ce", []).append(np.array_equal(ce_g[i], ce_o))
        exact.setdefault("d", []).append(np.array_equal(dd_g[i], dbg["d"][:I.nof_re]))
        assert np.array_equal(e_g[i], dbg["e"][:I.G]), "descrambled int16 LLRs differ"
        # soft buffer: device layout -> srsLTE order, per code block
        for r in range(s.C):
            K = o.cb_len(s, r)
            t = torch.zeros(3 * K + 12, dtype=torch.int16, device="cuda")
            ctx.tdec_export(d_sb[i, r * I.sb_cb_stride:], 1, K, t)
            torch.cuda.synchronize()
            assert np.array_equal(t.cpu().numpy(), dbg["softbuf"][r, :3 * K + 12]), "soft buffer cb %d" % r
    print(name, "bit-identical floats:", {k: all(v) for k, v in exact.items()})
    assert all(all(v) for v in exact.values()), "float intermediates are expected to be bit-identical"
    plan.close()


@pytest.mark.gpu
@pytest.mark.parametrize("prb", [6, 15, 25, 50, 75, 100])
def test_ofdm_rx_with_cfo_correction_matches_oracle(gpu, oracle, prb):
    """SPEC.md 14: the rotation srsLTE's synchroniser applies to the samples (phch_recv.cc:322), fused into the first
    FFT pass; per-subframe steps (one of them 0 = identity) and one step for all subframes."""
    import torch
    sg, ctx = gpu
    o = oracle
    n = o.lib().lteo_symbol_sz(prb)
    rng = np.random.default_rng(prb)
    n_sf = 4
    iq = (rng.standard_normal((n_sf, 15 * n)) + 1j * rng.standard_normal((n_sf, 15 * n))).astype(np.complex64)
    cfos = [0.43, -0.27, 0.0, 0.0031]
    steps = np.array([sg.host_cfo_step(c, n) for c in cfos], np.int32)
    assert steps.tolist() == [o.cfo_step(c, n) for c in cfos] and steps[2] == 0
    cell = sg.make_cell(prb, 1, 1)
    plan = sg.PdschPlan(ctx, cell, sg.make_cfg(cell, sf_idx=1, cfi=1, qm=2, tbs=152), n_sf)
    d_iq = torch.from_numpy(iq.view(np.float32)).cuda()
    d_sf = torch.zeros((n_sf, 14 * 12 * prb * 2), dtype=torch.float32, device="cuda")
    plan.ofdm_rx(n_sf, d_iq, d_sf, d_cfo_steps=torch.from_numpy(steps).cuda())
    torch.cuda.synchronize()
    got = d_sf.cpu().numpy().view(np.complex64)
    for i in range(n_sf):
        assert np.array_equal(got[i], o.ofdm_rx(prb, o.cfo_correct(iq[i], int(steps[i])))), "subframe %d" % i
    assert np.array_equal(got[2], o.ofdm_rx(prb, iq[2]))
    plan.ofdm_rx(n_sf, d_iq, d_sf, cfo_step=int(steps[1]))
    torch.cuda.synchronize()
    got = d_sf.cpu().numpy().view(np.complex64)
    for i in range(n_sf):
        assert np.array_equal(got[i], o.ofdm_rx(prb, o.cfo_correct(iq[i], int(steps[1]))))
    plan.close()


@pytest.mark.parametrize("prb", [6, 15, 25, 50, 75, 100])
def test_ofdm_rx_sc16_matches_oracle(gpu, oracle, prb):
    """int16 samples (the radio's wire format): the conversion float(v) * scale on the FFT's loads equals converting first"""
    import torch
    sg, ctx = gpu
    o = oracle
    n = o.lib().lteo_symbol_sz(prb)
    rng = np.random.default_rng(200 + prb)
    n_sf = 3
    q = rng.integers(-32768, 32768, (n_sf, 15 * n, 2), dtype=np.int16)
    q[0, :4] = [[-32768, 32767], [0, 0], [1, -1], [32767, -32768]]
    cell = sg.make_cell(prb, 1, 1)
    plan = sg.PdschPlan(ctx, cell, sg.make_cfg(cell, sf_idx=1, cfi=1, qm=2, tbs=152), n_sf)
    d_sf = torch.zeros((n_sf, 14 * 12 * prb * 2), dtype=torch.float32, device="cuda")
    for scale in (1.0 / 32768.0, 3.1e-4):
        x = (q.astype(np.float32) * np.float32(scale)).view(np.complex64).reshape(n_sf, -1)
        plan.ofdm_rx_sc16(n_sf, torch.from_numpy(q).cuda(), scale, d_sf)
        torch.cuda.synchronize()
        got = d_sf.cpu().numpy().view(np.complex64)
        for i in range(n_sf):
            assert np.array_equal(got[i], o.ofdm_rx(prb, x[i])), "subframe %d scale %g" % (i, scale)
        # a live radio's case: int16 samples with a carrier offset -- convert, then rotate (SPEC.md 14), both on the loads
        steps = np.array([sg.host_cfo_step(c, n) for c in (0.31, 0.0, -0.08)], np.int32)
        plan.ofdm_rx_sc16(n_sf, torch.from_numpy(q).cuda(), scale, d_sf, d_cfo_steps=torch.from_numpy(steps).cuda())
        torch.cuda.synchronize()
        got = d_sf.cpu().numpy().view(np.complex64)
        for i in range(n_sf):
            assert np.array_equal(got[i], o.ofdm_rx(prb, o.cfo_correct(x[i], int(steps[i])))), "cfo subframe %d" % i
    plan.close()


def test_ofdm_oracle_vs_numpy_fft(oracle):
    """anchor: the oracle's OFDM demodulator against numpy's FFT (runs without a GPU too)"""
    o = oracle
    rng = np.random.default_rng(5)
    for prb, n in ((6, 128), (25, 512), (75, 1536), (100, 2048)):
        iq = (rng.standard_normal(15 * n) + 1j * rng.standard_normal(15 * n)).astype(np.complex64)
        sf = o.ofdm_rx(prb, iq).reshape(14, -1)
        pos = 0
        for l in range(14):
            pos += (160 if l % 7 == 0 else 144) * n // 2048
            X = np.fft.fft(iq[pos:pos + n].astype(np.complex128)) / np.sqrt(n)
            pos += n
            nsc = 12 * prb
            ref = np.concatenate([X[n - nsc // 2:], X[1:nsc // 2 + 1]])
            assert np.max(np.abs(sf[l] - ref)) / np.sqrt(np.mean(abs(ref) ** 2)) < 1e-5


def test_harq_accumulate_matches_oracle(gpu, oracle):
    """rv 0 then rv 2 of the same transport block into one soft buffer (dl_harq.cc:191-259 behaviour)"""
    import torch
    sg, ctx = gpu
    o = oracle
    prb, qm, tbs = 25, 6, 11448     # rate > 1 on the first transmission alone would fail; keep decodable
    ocell = o.make_cell(prb, 1, 1)
    cell = sg.make_cell(prb, 1, 1)
    sb_o = None
    d_sb = None
    for rv, seed in ((0, 42), (2, 42)):
        ocfg = o.make_cfg(ocell, sf_idx=2, cfi=2, qm=qm, tbs=tbs, rv=rv)
        cfg = sg.make_cfg(cell, sf_idx=2, cfi=2, qm=qm, tbs=tbs, rv=rv)
        tb, iq, _ = o.gen_subframe(ocell, ocfg, seed, 11.0)
        plan = sg.PdschPlan(ctx, cell, cfg, 1)
        I = plan.info
        if d_sb is None:
            d_sb = torch.zeros((1, I.sb_sf_stride), dtype=torch.int16, device="cuda")
            sb_o = o.new_softbuf(I.C)
        d_iq = torch.from_numpy(iq.view(np.float32).reshape(1, -1)).cuda()
        d_pl = torch.zeros((1, I.payload_stride), dtype=torch.uint8, device="cuda")
        d_st = torch.zeros((1, 4), dtype=torch.int32, device="cuda")
        plan.decode_batch(1, d_iq, 0.01, 0, 4, d_pl, d_st, d_softbuf=d_sb, accumulate=1 if rv else 0)
        torch.cuda.synchronize()
        sf_o = o.ofdm_rx(prb, iq)
        ce_o, _ = o.chest(ocell, 2, sf_o)
        rc, pl = o.pdsch_decode(ocell, ocfg, sf_o, ce_o, 0.01, 4, softbuf=sb_o)
        st = d_st.cpu().numpy()[0]
        assert (st[0] == 1) == (rc == 0)
        assert np.array_equal(d_pl.cpu().numpy()[0], pl)
        s = o.cbsegm(tbs)
        for r in range(s.C):
            K = o.cb_len(s, r)
            t = torch.zeros(3 * K + 12, dtype=torch.int16, device="cuda")
            ctx.tdec_export(d_sb[0, r * I.sb_cb_stride:], 1, K, t)
            torch.cuda.synchronize()
            assert np.array_equal(t.cpu().numpy(), sb_o[r, :3 * K + 12])
        plan.close()


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "bw50_2p"])
def test_fused_channel_interpolation_is_bit_identical(gpu, oracle, name):
    """The whole-chain call never materialises the estimate grid: the demapper interpolates the smoothed pilots
    per resource element.  Soft buffers must equal the unfused stage path bit for bit."""
    import torch
    sg, ctx = gpu
    o = oracle
    c = _cfgs(o)[name]
    ocell = o.make_cell(c["prb"], c["ports"], 1)
    ocfg = o.make_cfg(ocell, sf_idx=c["sf"], cfi=1, qm=c["qm"], tbs=c["tbs"], tm=c["tm"])
    n_sf = 2
    iq = np.stack([o.gen_subframe(ocell, ocfg, 3000 + i, c["snr"], c["taps"])[1] for i in range(n_sf)])
    cell = sg.make_cell(c["prb"], c["ports"], 1)
    cfg = sg.make_cfg(cell, sf_idx=c["sf"], cfi=1, qm=c["qm"], tbs=c["tbs"], tm=c["tm"])
    plan = sg.PdschPlan(ctx, cell, cfg, n_sf)
    I = plan.info
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(n_sf, -1)).cuda()
    d_sf = torch.zeros((n_sf, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((n_sf, c["ports"] * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_pil = torch.zeros((n_sf, c["ports"] * 4 * 2 * c["prb"] * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((n_sf, 5), dtype=torch.float32, device="cuda")
    d_meas2 = torch.zeros((n_sf, 5), dtype=torch.float32, device="cuda")
    d_sb = torch.zeros((n_sf, I.sb_sf_stride), dtype=torch.int16, device="cuda")
    d_sb2 = torch.full((n_sf, I.sb_sf_stride), 77, dtype=torch.int16, device="cuda")
    plan.ofdm_rx(n_sf, d_iq, d_sf)
    plan.chest(n_sf, d_sf, d_ce, d_meas)
    plan.pdsch_llr(n_sf, d_sf, d_ce, d_meas, 0.01, 1, 0, d_sb)
    plan.chest_pilots(n_sf, d_sf, d_pil, d_meas2)
    plan.pdsch_llr_fused(n_sf, d_sf, d_pil, d_meas2, 0.01, 1, 0, d_sb2)
    torch.cuda.synchronize()
    assert torch.equal(d_meas, d_meas2)
    assert torch.equal(d_sb, d_sb2)
    plan.close()


@pytest.mark.parametrize("prb,ports,cid", [(6, 1, 1), (15, 2, 77), (25, 1, 301), (50, 2, 5), (75, 2, 8), (100, 1, 503), (100, 2, 0)])
def test_pcfich_matches_oracle(gpu, oracle, prb, ports, cid):
    """srsue_gpu_pcfich_decode: CFI and the three integer correlations equal the oracle's for every subframe of a batch
    with mixed CFIs, at an SNR where decisions are reliable and at one where they are not."""
    import torch
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    ocell = o.make_cell(prb, ports, cid)
    cell = sg.make_cell(prb, ports, cid)
    sf_idx = cid % 10
    tbs = 152 if prb == 6 else 1000
    cfis = [1, 2, 3, 2, 1, 3] if prb > 10 else [1, 2, 3, 1, 2, 3]
    for snr, noise_mode in ((8.0, 1), (-8.0, 0)):
        iq = []
        for i, cfi in enumerate(cfis):
            ocfg = o.make_cfg(ocell, sf_idx=sf_idx, cfi=cfi, qm=2, tbs=tbs, tm=ports)
            iq.append(o.gen_subframe(ocell, ocfg, 3000 + i, snr, None, pcfich=True)[1])
        iq = np.stack(iq)
        n = len(cfis)
        cfg = sg.make_cfg(cell, sf_idx=sf_idx, cfi=1, qm=2, tbs=0, tm=ports)      # front-end-only plan
        plan = sg.PdschPlan(ctx, cell, cfg, n)
        I = plan.info
        d_iq = torch.from_numpy(iq.view(np.float32).reshape(n, -1)).cuda()
        d_sf = torch.zeros((n, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
        d_ce = torch.zeros((n, ports * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
        d_meas = torch.zeros((n, 5), dtype=torch.float32, device="cuda")
        d_cfi = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_corr = torch.zeros((n, 3), dtype=torch.int32, device="cuda")
        plan.ofdm_rx(n, d_iq, d_sf)
        plan.chest(n, d_sf, d_ce, d_meas)
        plan.pcfich_decode(n, d_sf, d_ce, d_meas, 0.01, noise_mode, d_cfi, d_corr)
        torch.cuda.synchronize()
        got, corr = d_cfi.cpu().numpy(), d_corr.cpu().numpy()
        for i, cfi in enumerate(cfis):
            sf_o = o.ofdm_rx(prb, iq[i])
            ce_o, meas_o = o.chest(ocell, sf_idx, sf_o)
            ref, rcorr = o.pcfich_decode(ocell, sf_idx, sf_o, ce_o, meas_o[0] if noise_mode else 0.01)
            assert got[i] == ref and np.array_equal(corr[i], rcorr)
            if snr > 0:
                assert ref == cfi
        plan.close()


@pytest.mark.parametrize("prb,ports,cid,cfi", [(6, 1, 1, 3), (15, 2, 77, 2), (25, 1, 301, 2), (50, 2, 5, 1), (75, 1, 8, 3), (100, 1, 503, 1),
                                               (100, 2, 0, 3)])
def test_pdcch_matches_oracle(gpu, oracle, prb, ports, cid, cfi):
    """PDCCH on the device: the LLRs of every control-channel element, the RNTI every search-space candidate decodes to
    (rate de-matching + tail-biting Viterbi + CRC16) and the blind-search verdict equal the oracle's, with DCIs present
    at a reliable SNR and in a subframe that is mostly noise."""
    import torch
    sg, ctx = gpu
    o = oracle
    ocell = o.make_cell(prb, ports, cid)
    cell = sg.make_cell(prb, ports, cid)
    sf_idx, rnti = (cid + cfi) % 10, 0x1234 + cid
    nb = sg.lib().srsue_gpu_host_dci_format_sizeof(0, prb)
    rk, _ = o.pdcch_regs(ocell, cfi, 6)
    ncce = len(rk) // 9
    ss = o.pdcch_search_space(ncce, sf_idx, rnti)
    rng = np.random.default_rng(cid)
    n = 4
    for snr in (8.0, -3.0):
        iq, sent = [], []
        for i in range(n):
            L0, n0 = ss[(3 * i + 1) % len(ss)]
            bits = rng.integers(0, 2, nb, dtype=np.uint8)
            dcis = [(bits, rnti, L0, n0)] if i != 2 else []         # subframe 2 carries no DCI for this UE
            ocfg = o.make_cfg(ocell, sf_idx=sf_idx, cfi=cfi, qm=2, tbs=152 if prb == 6 else 1000, tm=ports)
            iq.append(o.gen_subframe(ocell, ocfg, 4000 + i, snr, None, pcfich=True, dcis=dcis)[1])
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
        d_rem = torch.zeros((n, 24), dtype=torch.uint16, device="cuda")
        plan.ofdm_rx(n, d_iq, d_sf)
        plan.chest(n, d_sf, d_ce, d_meas)
        plan.pdcch_extract_llr(n, d_sf, d_ce, d_meas, 0.0, 1, d_llr)
        ncand = plan.pdcch_find_dci(n, d_llr, rnti, nb, d_found, d_bits, d_rem)
        torch.cuda.synchronize()
        assert ncand == len(ss)
        llr_g, found, bits_g = d_llr.cpu().numpy(), d_found.cpu().numpy(), d_bits.cpu().numpy()
        rem = d_rem.cpu().numpy().reshape(-1)[:n * ncand].reshape(n, ncand)
        for i in range(n):
            sf_o = o.ofdm_rx(prb, iq[i])
            ce_o, meas_o = o.chest(ocell, sf_idx, sf_o)
            llr_o, _ = o.pdcch_extract_llr(ocell, sf_idx, cfi, sf_o, ce_o, meas_o[0])
            assert np.array_equal(llr_g[i], llr_o[:8 * n_reg])
            for c, (L0, n0) in enumerate(ss):
                assert rem[i, c] == o.pdcch_decode_candidate(llr_o[72 * n0:], L0, nb)[1]
            f, out, L1, n1 = o.pdcch_find_dci(llr_o, ncce, sf_idx, rnti, nb)
            assert found[i, 0] == f
            if f:
                assert (found[i, 1], found[i, 2]) == (L1, n1) and np.array_equal(bits_g[i, :nb], out)
            if snr > 0:
                assert f == (1 if sent[i] else 0)
                if f:
                    assert np.array_equal(out, sent[i][0][0])
        plan.close()


@pytest.mark.parametrize("prb,ports,cid", [(6, 1, 1), (25, 2, 77), (50, 1, 301), (75, 2, 8), (100, 1, 503), (100, 2, 0)])
def test_phich_matches_oracle(gpu, oracle, prb, ports, cid):
    """HARQ indicators: decision and float metric (bit-identical) for several indicators of one subframe batch, including
    two sequences sharing a group, at a reliable SNR and in noise"""
    import torch
    sg, ctx = gpu
    o = oracle
    ocell = o.make_cell(prb, ports, cid)
    cell = sg.make_cell(prb, ports, cid)
    sf_idx = cid % 10
    groups = (6 * prb + 47) // 48
    rng = np.random.default_rng(cid + prb)
    for snr in (10.0, -10.0):
        n = 5
        ph = [(0, 0, 1), (0, 5, 0), (groups - 1, 3, 1)]
        acks = [[int(rng.integers(0, 2)) for _ in ph] for _ in range(n)]
        iq = []
        for i in range(n):
            ocfg = o.make_cfg(ocell, sf_idx=sf_idx, cfi=2, qm=2, tbs=152 if prb == 6 else 1000, tm=ports)
            iq.append(o.gen_subframe(ocell, ocfg, 6000 + i, snr, None, pcfich=True,
                                     phichs=[(g, q, a) for (g, q, _), a in zip(ph, acks[i])])[1])
        iq = np.stack(iq)
        cfg = sg.make_cfg(cell, sf_idx=sf_idx, cfi=2, qm=2, tbs=0, tm=ports)
        plan = sg.PdschPlan(ctx, cell, cfg, n)
        I = plan.info
        d_iq = torch.from_numpy(iq.view(np.float32).reshape(n, -1)).cuda()
        d_sf = torch.zeros((n, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
        d_ce = torch.zeros((n, ports * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
        d_meas = torch.zeros((n, 5), dtype=torch.float32, device="cuda")
        d_ack = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_met = torch.zeros(n, dtype=torch.float32, device="cuda")
        plan.ofdm_rx(n, d_iq, d_sf)
        plan.chest(n, d_sf, d_ce, d_meas)
        for j, (g, q, _) in enumerate(ph):
            plan.phich_decode(n, d_sf, d_ce, d_meas, 0.0, 1, g, q, d_ack, d_met)
            torch.cuda.synchronize()
            ack_g, met_g = d_ack.cpu().numpy(), d_met.cpu().numpy()
            for i in range(n):
                sf_o = o.ofdm_rx(prb, iq[i])
                ce_o, meas_o = o.chest(ocell, sf_idx, sf_o)
                a, m = o.phich_decode(ocell, sf_idx, sf_o, ce_o, g, q, meas_o[0])
                assert ack_g[i] == a and met_g[i] == m
                if snr > 0:
                    assert a == acks[i][j]
        plan.close()


@pytest.mark.parametrize("prb,ports,cid", [(6, 1, 1), (6, 2, 2), (25, 2, 77), (50, 1, 301), (75, 2, 8), (100, 1, 503), (100, 2, 0)])
def test_pbch_mib_matches_oracle(gpu, oracle, prb, ports, cid):
    """blind MIB decode from subframes 0 (port hypotheses x frame positions): verdict, MIB bits, detected number of ports
    and frame offset equal the oracle's; at a reliable SNR they equal what was sent"""
    import torch
    sg, ctx = gpu
    o = oracle
    ocell = o.make_cell(prb, ports, cid)
    cell = sg.make_cell(prb, ports, cid)
    for snr in (6.0, -9.0):
        sfns = [0, 5, 1022, 1023, 640]
        n = len(sfns)
        mibs = [o.mib_pack(prb, 0, 6, s) for s in sfns]
        iq = []
        for i, s in enumerate(sfns):
            ocfg = o.make_cfg(ocell, sf_idx=0, cfi=2, qm=2, tbs=104, tm=ports, prbs=[0] if prb > 6 else None)
            iq.append(o.gen_subframe(ocell, ocfg, 7000 + i, snr, None, pcfich=True, mib=(mibs[i], s % 4))[1])
        iq = np.stack(iq)
        cfg = sg.make_cfg(cell, sf_idx=0, cfi=2, qm=2, tbs=0, tm=ports)
        plan = sg.PdschPlan(ctx, cell, cfg, n)
        I = plan.info
        d_iq = torch.from_numpy(iq.view(np.float32).reshape(n, -1)).cuda()
        d_sf = torch.zeros((n, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
        d_ce = torch.zeros((n, ports * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
        d_meas = torch.zeros((n, 5), dtype=torch.float32, device="cuda")
        d_res = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
        d_mib = torch.zeros((n, 24), dtype=torch.uint8, device="cuda")
        plan.ofdm_rx(n, d_iq, d_sf)
        plan.chest(n, d_sf, d_ce, d_meas)
        plan.pbch_decode(n, d_sf, d_ce, d_meas, 0.0, 1, d_res, d_mib)
        torch.cuda.synchronize()
        res, mib_g = d_res.cpu().numpy(), d_mib.cpu().numpy()
        for i, s in enumerate(sfns):
            sf_o = o.ofdm_rx(prb, iq[i])
            ce_o, meas_o = o.chest(ocell, 0, sf_o)
            f, bits, p, q = o.pbch_decode(ocell, sf_o, ce_o, meas_o[0])
            assert res[i, 0] == f
            if f:
                assert (res[i, 1], res[i, 2]) == (p, q) and np.array_equal(mib_g[i], bits)
            if snr > 0:
                assert f == 1 and p == ports and q == s % 4 and np.array_equal(bits, mibs[i])
        plan.close()


def test_srslte_ue_mib_decode_shim(gpu, oracle):
    """phch_recv's MIB step (phch_recv.cc:246-253): srslte_ue_mib_decode on the samples of a subframe 0, then
    srslte_pbch_mib_unpack -> bandwidth, PHICH configuration, SFN"""
    import ctypes as C
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import UeMib, Cell
    prb, ports, cid, sfn = 50, 2, 123, 777
    ocell = o.make_cell(prb, ports, cid)
    ocfg = o.make_cfg(ocell, sf_idx=0, cfi=1, qm=2, tbs=104, tm=2, prbs=[3])
    mib = o.mib_pack(prb, 0, 3, sfn)                     # Ng = 1/2
    iq = o.gen_subframe(ocell, ocfg, 99, 8.0, None, pcfich=True, mib=(mib, sfn % 4))[1]
    q = UeMib()
    cell = Cell(nof_prb=prb, nof_ports=1, bw_idx=0, id=cid, cp=0, phich_length=0, phich_resources=0)   # ports unknown yet
    assert L.srslte_ue_mib_init(C.byref(q), cell) == 0
    L.srslte_pbch_decode_reset(C.byref(q))
    payload = (C.c_uint8 * 24)()
    nports, off = C.c_uint32(0), C.c_uint32(0)
    assert L.srslte_ue_mib_decode(C.byref(q), iq.ctypes.data_as(C.c_void_p), payload, C.byref(nports), C.byref(off)) == 1
    assert nports.value == 2 and off.value == sfn % 4 and np.array_equal(np.frombuffer(payload, np.uint8), mib)
    out_cell, out_sfn = Cell(), C.c_uint32(0)
    L.srslte_pbch_mib_unpack(payload, C.byref(out_cell), C.byref(out_sfn))
    assert out_cell.nof_prb == prb and out_cell.phich_resources == 1 and out_cell.phich_length == 0
    assert out_sfn.value + off.value == sfn              # the MIB carries SFN / 4, the offset the two LSBs
    packed = (C.c_uint8 * 24)()
    out_cell.phich_resources = 1
    L.srslte_pbch_mib_pack(C.byref(out_cell), sfn, packed)
    assert np.array_equal(np.frombuffer(packed, np.uint8), mib)
    # noise only: nothing found
    noise = (np.random.default_rng(1).standard_normal((len(iq), 2)) @ np.array([1, 1j])).astype(np.complex64)
    assert L.srslte_ue_mib_decode(C.byref(q), noise.ctypes.data_as(C.c_void_p), payload, C.byref(nports), C.byref(off)) == 0
    L.srslte_ue_mib_free(C.byref(q))


def test_pbch_tables_survive_pdcch_on_the_same_plan(gpu, oracle):
    """PBCH, then PDCCH, then PBCH again on ONE plan (the natural order on subframe 0): building the PDCCH tables must not
    release the PBCH tables -- the second MIB decode equals the first and that of a fresh plan"""
    import torch
    sg, ctx = gpu
    o = oracle
    prb, ports, cid = 25, 2, 77
    ocell = o.make_cell(prb, ports, cid)
    cell = sg.make_cell(prb, ports, cid)
    sfns = [0, 5, 1022]
    n = len(sfns)
    mibs = [o.mib_pack(prb, 0, 6, s) for s in sfns]
    ocfg = o.make_cfg(ocell, sf_idx=0, cfi=2, qm=2, tbs=104, tm=ports, prbs=[0])
    iq = np.stack([o.gen_subframe(ocell, ocfg, 7100 + i, 6.0, None, pcfich=True, mib=(mibs[i], s % 4))[1] for i, s in enumerate(sfns)])
    cfg = sg.make_cfg(cell, sf_idx=0, cfi=2, qm=2, tbs=0, tm=ports)

    def run(plan, with_pdcch):
        I = plan.info
        d_iq = torch.from_numpy(iq.view(np.float32).reshape(n, -1)).cuda()
        d_sf = torch.zeros((n, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
        d_ce = torch.zeros((n, ports * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
        d_meas = torch.zeros((n, 5), dtype=torch.float32, device="cuda")
        plan.ofdm_rx(n, d_iq, d_sf)
        plan.chest(n, d_sf, d_ce, d_meas)
        out = []
        for rep in range(2):
            d_res = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
            d_mib = torch.zeros((n, 24), dtype=torch.uint8, device="cuda")
            plan.pbch_decode(n, d_sf, d_ce, d_meas, 0.0, 1, d_res, d_mib)
            torch.cuda.synchronize()
            out.append((d_res.cpu().numpy().copy(), d_mib.cpu().numpy().copy()))
            if with_pdcch and rep == 0:
                n_reg = plan.pdcch_info(1)[0]
                d_llr = torch.zeros((n, 8 * n_reg), dtype=torch.int16, device="cuda")
                plan.pdcch_extract_llr(n, d_sf, d_ce, d_meas, 0.0, 1, d_llr)
                torch.cuda.synchronize()
        return out

    p1 = sg.PdschPlan(ctx, cell, cfg, n)
    first, second = run(p1, True)
    p1.close()
    p2 = sg.PdschPlan(ctx, cell, cfg, n)
    fresh, _ = run(p2, False)
    p2.close()
    for res, mib in (first, second, fresh):
        assert np.all(res[:, 0] == 1) and np.array_equal(res[:, 1], [ports] * n) and np.array_equal(res[:, 2], [s % 4 for s in sfns])
        assert np.array_equal(mib, np.stack(mibs))

