"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must reproduce them on
the CPU, the CUDA kernels must reproduce them on the GPU.  Parity remains unpinned against srsLTE itself."""
import hashlib
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _taps():
    rng = np.random.default_rng(77)
    t = (rng.standard_normal((2, 6)) + 1j * rng.standard_normal((2, 6))) * np.array([1, .7, .5, .3, .2, .1])
    return t / np.sqrt((abs(t) ** 2).sum(1, keepdims=True))


def test_oracle_reproduces_turbo_golden(oracle):
    tv = np.load(os.path.join(G, "turbo.npz"))
    for K in (40, 512, 1056, 5824):
        for i, llr in enumerate(tv["llr_%d" % K]):
            b, it, ok, _ = oracle.tdec(llr, K, 4, 0)
            assert np.array_equal(np.packbits(b), tv["bits4_%d" % K][i])
            b2, it2, ok2, _ = oracle.tdec(llr, K, 6, 2)
            assert (it2, ok2, int(np.packbits(b2).sum())) == tuple(tv["crcrun_%d" % K][i])


def test_oracle_reproduces_pdsch_golden(oracle):
    pv = np.load(os.path.join(G, "pdsch.npz"))
    cell = oracle.make_cell(6, 1, 1)
    cfg = oracle.make_cfg(cell, sf_idx=1, cfi=1, qm=2, tbs=152)
    iq = pv["cfg1_iq"]
    sf = oracle.ofdm_rx(6, iq)
    ce, meas = oracle.chest(cell, 1, sf)
    rc, pl, dbg = oracle.pdsch_decode(cell, cfg, sf, ce, 0.01, 4, want=True)
    assert np.array_equal(sf, pv["cfg1_sf"]) and np.array_equal(ce, pv["cfg1_ce"]) and np.array_equal(meas, pv["cfg1_meas"])
    assert np.array_equal(dbg["e"][:1656], pv["cfg1_e"])
    assert np.array_equal(dbg["softbuf"][0, :3 * 176 + 12], pv["cfg1_softbuf"])
    assert np.array_equal(pl, pv["cfg1_payload"]) and np.array_equal(pl, pv["cfg1_tb"]) and rc == int(pv["cfg1_rc"][0])
    # the generator itself is part of what the digests freeze
    tb, iq2, _ = oracle.gen_subframe(cell, cfg, 1, 10.0)
    assert np.array_equal(iq2, iq)


@pytest.mark.parametrize("idx,name,prb,ports,qm,tbs,tm,snr", [(0, "cfg2", 100, 1, 6, 75376, 1, 30.0), (1, "cfg3", 100, 2, 4, 30576, 2, 15.0)])
def test_oracle_reproduces_digests(oracle, idx, name, prb, ports, qm, tbs, tm, snr):
    pv = np.load(os.path.join(G, "pdsch.npz"))
    want = dict(kv.split("=") for kv in str(pv["digests"][idx]).split()[1:])
    cell = oracle.make_cell(prb, ports, 1)
    cfg = oracle.make_cfg(cell, sf_idx=1, cfi=1, qm=qm, tbs=tbs, tm=tm)
    tb, iq, _ = oracle.gen_subframe(cell, cfg, 20000, snr, _taps() if ports == 2 else None)
    sf = oracle.ofdm_rx(prb, iq)
    ce, meas = oracle.chest(cell, 1, sf)
    rc, pl, dbg = oracle.pdsch_decode(cell, cfg, sf, ce, 0.01, 4, want=True)
    s = oracle.cbsegm(tbs)
    assert sha(iq) == want["iq"] and sha(sf) == want["sf"] and sha(ce) == want["ce"]
    assert sha(dbg["softbuf"][:s.C, :3 * s.Kp + 12]) == want["sb"] and sha(pl) == want["payload"]
    assert ",".join(map(str, dbg["iters"])) == want["iters"] and rc == int(want["rc"])


CTRL = dict(prb=25, ports=2, cid=77, cfi=2, sf_idx=3, rnti=0x4601, nb=25)


def test_oracle_reproduces_control_golden(oracle):
    o = oracle
    cv = np.load(os.path.join(G, "control.npz"))
    c = CTRL
    cell = o.make_cell(c["prb"], c["ports"], c["cid"])
    sf = o.ofdm_rx(c["prb"], cv["iq"])
    ce, meas = o.chest(cell, c["sf_idx"], sf)
    cfi, corr = o.pcfich_decode(cell, c["sf_idx"], sf, ce, meas[0])
    assert cfi == int(cv["cfi"][0]) == c["cfi"] and np.array_equal(corr, cv["corr"])
    llr, nc = o.pdcch_extract_llr(cell, c["sf_idx"], cfi, sf, ce, meas[0])
    assert np.array_equal(llr[:len(cv["llr"])], cv["llr"])
    ss = o.pdcch_search_space(nc, c["sf_idx"], c["rnti"])
    assert np.array_equal(np.array(ss, np.int32), cv["cand"])
    assert [o.pdcch_decode_candidate(llr[72 * n:], L, c["nb"])[1] for L, n in ss] == cv["rem"].tolist()
    f, out, L1, n1 = o.pdcch_find_dci(llr, nc, c["sf_idx"], c["rnti"], c["nb"])
    assert [f, L1, n1] == cv["found"].tolist() and f == 1 and np.array_equal(out, cv["bits"]) and np.array_equal(out, cv["sent"])


def test_oracle_reproduces_cfo_golden(oracle):
    o = oracle
    v = np.load(os.path.join(G, "cfo.npz"))
    assert [o.cfo_step(0.37, 128), o.cfo_step(-0.081, 128)] == v["steps"].tolist()
    assert np.array_equal(o.cfo_table()[::64], v["tab"])
    for i, st in enumerate(v["steps"]):
        y = o.cfo_correct(v["x"], int(st))
        assert np.array_equal(y, v["y"][i]) and np.array_equal(o.ofdm_rx(6, y), v["sf"][i])
    # independent anchor: the rotation against a double-precision complex exponential (table resolution 2 pi / 4096)
    n = np.arange(1920)
    ref = v["x"].astype(np.complex128) * np.exp(-2j * np.pi * 0.37 * n / 128)
    assert np.max(np.abs(v["y"][0] - ref) / np.abs(v["x"])) < 2 * np.pi / 4096 + 1e-6


@pytest.mark.gpu
def test_gpu_reproduces_cfo_golden(gpu):
    import torch
    sg, ctx = gpu
    v = np.load(os.path.join(G, "cfo.npz"))
    cell = sg.make_cell(6, 1, 1)
    plan = sg.PdschPlan(ctx, cell, sg.make_cfg(cell, sf_idx=1, cfi=1, qm=2, tbs=152), 2)
    d_iq = torch.from_numpy(np.stack([v["x"], v["x"]]).view(np.float32)).cuda()
    d_sf = torch.zeros((2, 14 * 72 * 2), dtype=torch.float32, device="cuda")
    assert [sg.host_cfo_step(0.37, 128), sg.host_cfo_step(-0.081, 128)] == v["steps"].tolist()
    plan.ofdm_rx(2, d_iq, d_sf, d_cfo_steps=torch.from_numpy(v["steps"]).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(d_sf.cpu().numpy().view(np.complex64), v["sf"].reshape(2, -1))
    plan.close()


@pytest.mark.gpu
def test_gpu_reproduces_control_golden(gpu):
    import torch
    sg, ctx = gpu
    cv = np.load(os.path.join(G, "control.npz"))
    c = CTRL
    cell = sg.make_cell(c["prb"], c["ports"], c["cid"])
    cfg = sg.make_cfg(cell, sf_idx=c["sf_idx"], cfi=c["cfi"], qm=2, tbs=0, tm=2)
    plan = sg.PdschPlan(ctx, cell, cfg, 1)
    I = plan.info
    n_reg, ncce = plan.pdcch_info(6)
    d_iq = torch.from_numpy(cv["iq"].view(np.float32).reshape(1, -1)).cuda()
    d_sf = torch.zeros((1, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((1, 2 * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((1, 5), dtype=torch.float32, device="cuda")
    d_cfi = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_corr = torch.zeros((1, 3), dtype=torch.int32, device="cuda")
    d_llr = torch.zeros((1, 8 * n_reg), dtype=torch.int16, device="cuda")
    d_found = torch.zeros((1, 4), dtype=torch.int32, device="cuda")
    d_bits = torch.zeros((1, 64), dtype=torch.uint8, device="cuda")
    d_rem = torch.zeros((1, 24), dtype=torch.uint16, device="cuda")
    plan.ofdm_rx(1, d_iq, d_sf)
    plan.chest(1, d_sf, d_ce, d_meas)
    plan.pcfich_decode(1, d_sf, d_ce, d_meas, 0.0, 1, d_cfi, d_corr)
    plan.pdcch_extract_llr(1, d_sf, d_ce, d_meas, 0.0, 1, d_llr)
    ncand = plan.pdcch_find_dci(1, d_llr, c["rnti"], c["nb"], d_found, d_bits, d_rem)
    torch.cuda.synchronize()
    assert int(d_cfi.cpu()[0]) == int(cv["cfi"][0]) and np.array_equal(d_corr.cpu().numpy()[0], cv["corr"])
    assert np.array_equal(d_llr.cpu().numpy()[0], cv["llr"])
    assert ncand == len(cv["cand"]) and d_rem.cpu().numpy()[0, :ncand].astype(np.int32).tolist() == cv["rem"].tolist()
    found = d_found.cpu().numpy()[0]
    assert found[:3].tolist() == cv["found"].tolist() and np.array_equal(d_bits.cpu().numpy()[0, :c["nb"]], cv["bits"])
    plan.close()


@pytest.mark.gpu
def test_gpu_reproduces_turbo_golden(gpu):
    import torch
    sg, ctx = gpu
    tv = np.load(os.path.join(G, "turbo.npz"))
    for K in (40, 512, 1056, 5824):
        llrs = tv["llr_%d" % K]
        n = len(llrs)
        d_in = torch.from_numpy(llrs).cuda()
        d_bits = torch.zeros((n, K // 8), dtype=torch.uint8, device="cuda")
        d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
        ctx.tdec_run_all(d_in, n, K, 4, 0, d_bits, d_st)
        torch.cuda.synchronize()
        assert np.array_equal(d_bits.cpu().numpy(), tv["bits4_%d" % K])
        ctx.tdec_run_all(d_in, n, K, 6, 2, d_bits, d_st)
        torch.cuda.synchronize()
        st = d_st.cpu().numpy()
        got = np.stack([st & 0xFF, (st >> 8) & 1, d_bits.cpu().numpy().astype(np.int64).sum(1)], 1)
        assert np.array_equal(got, tv["crcrun_%d" % K])


@pytest.mark.gpu
def test_gpu_reproduces_pdsch_golden(gpu):
    import torch
    sg, ctx = gpu
    pv = np.load(os.path.join(G, "pdsch.npz"))
    cell = sg.make_cell(6, 1, 1)
    cfg = sg.make_cfg(cell, sf_idx=1, cfi=1, qm=2, tbs=152)
    plan = sg.PdschPlan(ctx, cell, cfg, 1)
    I = plan.info
    iq = pv["cfg1_iq"]
    d_iq = torch.from_numpy(iq.view(np.float32).reshape(1, -1)).cuda()
    d_sf = torch.zeros((1, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((1, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((1, 5), dtype=torch.float32, device="cuda")
    d_sb = torch.zeros((1, I.sb_sf_stride), dtype=torch.int16, device="cuda")
    d_e = torch.zeros((1, I.G), dtype=torch.int16, device="cuda")
    d_pl = torch.zeros((1, I.payload_stride), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros((1, 4), dtype=torch.int32, device="cuda")
    plan.ofdm_rx(1, d_iq, d_sf)
    plan.chest(1, d_sf, d_ce, d_meas)
    plan.pdsch_llr(1, d_sf, d_ce, d_meas, 0.01, 0, 0, d_sb, None, d_e)
    plan.pdsch_turbo(1, d_sb, 4, d_pl, d_st)
    t = torch.zeros(3 * 176 + 12, dtype=torch.int16, device="cuda")
    ctx.tdec_export(d_sb, 1, 176, t)
    torch.cuda.synchronize()
    assert np.array_equal(d_sf.cpu().numpy().view(np.complex64)[0], pv["cfg1_sf"])
    assert np.array_equal(d_ce.cpu().numpy().view(np.complex64), pv["cfg1_ce"])
    assert np.allclose(d_meas.cpu().numpy()[0], pv["cfg1_meas"], rtol=1e-5)
    assert np.array_equal(d_e.cpu().numpy()[0], pv["cfg1_e"])
    assert np.array_equal(t.cpu().numpy(), pv["cfg1_softbuf"])
    assert np.array_equal(d_pl.cpu().numpy()[0], pv["cfg1_payload"]) and d_st.cpu().numpy()[0, 0] == 1
    plan.close()


EXT = dict(prb=6, ports=2, cid=77, cfi=2, rnti=0x1234, qm=2, tbs=56, phich=((0, 1), (1, 2), (1, 3)))


def test_oracle_reproduces_extended_prefix_golden(oracle):
    """tests/golden/extcp.npz (make_golden.py extcp): one subframe 0 of an extended-prefix cell through every stage"""
    o = oracle
    v = np.load(os.path.join(G, "extcp.npz"))
    c = EXT
    cell = o.make_cell(c["prb"], c["ports"], c["cid"], cp=1)
    cfg = o.make_cfg(cell, sf_idx=0, cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=c["tbs"], tm=2)
    sf = o.ofdm_rx(c["prb"], v["iq"], cp=1)
    ce, meas = o.chest(cell, 0, sf)
    assert np.array_equal(sf[:12 * 72], v["sf"]) and np.array_equal(ce[:, :12 * 72], v["ce"]) and np.array_equal(meas, v["meas"])
    cfi, corr = o.pcfich_decode(cell, 0, sf, ce, meas[0])
    assert cfi == int(v["cfi"][0]) == c["cfi"] and np.array_equal(corr, v["corr"])
    for i, (g, q) in enumerate(c["phich"]):
        a, m = o.phich_decode(cell, 0, sf, ce, g, q, float(meas[0]))
        assert a == v["phich_ack"][i] and m == v["phich_metric"][i]
    f, bits, ports, off = o.pbch_decode(cell, sf, ce, float(meas[0]))
    assert [f, ports, off] == v["pbch"].tolist() and np.array_equal(bits, v["mib"]) and np.array_equal(bits, v["mib_sent"])
    rc, pl, dbg = o.pdsch_decode(cell, cfg, sf, ce, float(meas[0]), 4, want=True)
    assert rc == int(v["rc"][0]) == 0 and np.array_equal(pl, v["payload"]) and np.array_equal(pl, v["tb"])
    assert np.array_equal(dbg["e"][:len(v["e"])], v["e"])
    x = np.concatenate([v["iq"][-300:], v["iq"]])
    pk = o.pss_search(x)
    n1, sf5, scorr, cp = o.sss_detect_cp(x, pk["pos"], pk["n_id_2"], 128, 2)
    assert [pk["pos"], pk["n_id_2"], n1, sf5, cp] == v["sync"].tolist() and cp == 1 and 3 * n1 + pk["n_id_2"] == c["cid"]
    assert pk["peak"] == v["sync_f"][0] and scorr == v["sync_f"][1]


@pytest.mark.gpu
def test_gpu_reproduces_extended_prefix_golden(gpu):
    import ctypes as C
    import torch
    sg, ctx = gpu
    v = np.load(os.path.join(G, "extcp.npz"))
    c = EXT
    cell = sg.make_cell(c["prb"], c["ports"], c["cid"], cp=1)
    cfg = sg.make_cfg(cell, sf_idx=0, cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=c["tbs"], tm=2)
    plan = sg.PdschPlan(ctx, cell, cfg, 1)
    I = plan.info
    d_iq = torch.from_numpy(v["iq"].view(np.float32).reshape(1, -1)).cuda()
    d_sf = torch.zeros((1, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((1, 2 * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((1, 5), dtype=torch.float32, device="cuda")
    d_cfi = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_corr = torch.zeros((1, 3), dtype=torch.int32, device="cuda")
    d_sb = torch.zeros((1, I.sb_sf_stride), dtype=torch.int16, device="cuda")
    d_e = torch.zeros((1, I.G), dtype=torch.int16, device="cuda")
    d_res = torch.zeros((1, 4), dtype=torch.int32, device="cuda")
    d_mib = torch.zeros((1, 24), dtype=torch.uint8, device="cuda")
    d_pl = torch.zeros((1, I.payload_stride), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros((1, 4), dtype=torch.int32, device="cuda")
    plan.ofdm_rx(1, d_iq, d_sf)
    plan.chest(1, d_sf, d_ce, d_meas)
    plan.pcfich_decode(1, d_sf, d_ce, d_meas, 0.0, 1, d_cfi, d_corr)
    plan.pbch_decode(1, d_sf, d_ce, d_meas, 0.0, 1, d_res, d_mib)
    plan.pdsch_llr(1, d_sf, d_ce, d_meas, 0.0, 1, 0, d_sb, None, d_e)
    plan.decode_batch(1, d_iq, 0.0, 1, 4, d_pl, d_st)
    torch.cuda.synchronize()
    sf_g = d_sf.cpu().numpy().view(np.complex64)[0]
    ce_g = d_ce.cpu().numpy().view(np.complex64).reshape(2, -1)
    assert np.array_equal(sf_g[:12 * 72], v["sf"]) and np.array_equal(ce_g[:, :12 * 72], v["ce"])
    assert np.allclose(d_meas.cpu().numpy()[0], v["meas"], rtol=1e-5)
    assert int(d_cfi.cpu()[0]) == int(v["cfi"][0]) and np.array_equal(d_corr.cpu().numpy()[0], v["corr"])
    assert d_res.cpu().numpy()[0, :3].tolist() == v["pbch"].tolist() and np.array_equal(d_mib.cpu().numpy()[0], v["mib"])
    assert np.array_equal(d_e.cpu().numpy()[0, :len(v["e"])], v["e"])
    assert int(d_st.cpu()[0, 0]) == 1 and np.array_equal(d_pl.cpu().numpy()[0, :len(v["payload"])], v["payload"])
    for i, (g, q) in enumerate(c["phich"]):
        d_ack = torch.zeros(1, dtype=torch.int32, device="cuda")
        d_met = torch.zeros(1, dtype=torch.float32, device="cuda")
        plan.phich_decode(1, d_sf, d_ce, d_meas, 0.0, 1, g, q, d_ack, d_met)
        torch.cuda.synchronize()
        assert int(d_ack[0]) == v["phich_ack"][i] and np.float32(d_met[0].item()) == v["phich_metric"][i]
    x = np.concatenate([v["iq"][-300:], v["iq"]])
    d_x = torch.from_numpy(x.view(np.float32).reshape(1, -1)).cuda()
    d_sr = torch.zeros(C.sizeof(sg.SyncResult), dtype=torch.uint8, device="cuda")
    ctx.cell_search(d_x, 1, len(x), len(x), d_sr, cp_mode=2)
    torch.cuda.synchronize()
    r = sg.SyncResult.from_buffer_copy(d_sr.cpu().numpy().tobytes())
    assert [r.peak_pos, r.n_id_2, r.n_id_1, r.sf5, r.cp] == v["sync"].tolist()
    assert np.float32(r.peak) == v["sync_f"][0] and np.float32(r.sss_corr) == v["sync_f"][1]
    plan.close()


P4 = dict(prb=6, cid=101, cfi=2, rnti=0x2345, qm=4, tbs=208, phich=((0, 1), (0, 5)), nb=21)


def test_oracle_reproduces_four_port_golden(oracle):
    """tests/golden/fourports.npz (make_golden.py fourports): one subframe 0 of a four-port cell through every stage"""
    o = oracle
    v = np.load(os.path.join(G, "fourports.npz"))
    c = P4
    cell = o.make_cell(c["prb"], 4, c["cid"])
    cfg = o.make_cfg(cell, sf_idx=0, cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=c["tbs"], tm=2)
    sf = o.ofdm_rx(c["prb"], v["iq"])
    ce, meas = o.chest(cell, 0, sf)
    assert np.array_equal(sf, v["sf"]) and np.array_equal(ce, v["ce"]) and np.array_equal(meas, v["meas"])
    cfi, corr = o.pcfich_decode(cell, 0, sf, ce, meas[0])
    assert cfi == int(v["cfi"][0]) == c["cfi"] and np.array_equal(corr, v["corr"])
    for i, (g, q) in enumerate(c["phich"]):
        a, m = o.phich_decode(cell, 0, sf, ce, g, q, float(meas[0]))
        assert a == v["phich_ack"][i] and m == v["phich_metric"][i]
    assert v["phich_ack"].tolist() == [1, 0]
    f, bits, ports, off = o.pbch_decode(cell, sf, ce, float(meas[0]))
    assert [f, ports, off] == v["pbch"].tolist() and ports == 4 and np.array_equal(bits, v["mib"]) and np.array_equal(bits, v["mib_sent"])
    llr, nc = o.pdcch_extract_llr(cell, 0, cfi, sf, ce, meas[0])
    assert np.array_equal(llr[:len(v["llr"])], v["llr"])
    fd, out, L1, n1 = o.pdcch_find_dci(llr, nc, 0, c["rnti"], c["nb"])
    assert [fd, L1, n1] == v["dci"].tolist() and fd == 1 and np.array_equal(out, v["dci_bits"]) and np.array_equal(out, v["dci_sent"])
    rc, pl, dbg = o.pdsch_decode(cell, cfg, sf, ce, float(meas[0]), 4, want=True)
    assert rc == int(v["rc"][0]) == 0 and np.array_equal(pl, v["payload"]) and np.array_equal(pl, v["tb"])
    assert np.array_equal(dbg["e"][:len(v["e"])], v["e"])


@pytest.mark.gpu
def test_gpu_reproduces_four_port_golden(gpu):
    import torch
    sg, ctx = gpu
    v = np.load(os.path.join(G, "fourports.npz"))
    c = P4
    cell = sg.make_cell(c["prb"], 4, c["cid"])
    cfg = sg.make_cfg(cell, sf_idx=0, cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=c["tbs"], tm=2)
    plan = sg.PdschPlan(ctx, cell, cfg, 1)
    I = plan.info
    n_reg, nc = plan.pdcch_info(6)
    d_iq = torch.from_numpy(v["iq"].view(np.float32).reshape(1, -1)).cuda()
    d_sf = torch.zeros((1, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_ce = torch.zeros((1, 4 * 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
    d_meas = torch.zeros((1, 5), dtype=torch.float32, device="cuda")
    d_cfi = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_corr = torch.zeros((1, 3), dtype=torch.int32, device="cuda")
    d_sb = torch.zeros((1, I.sb_sf_stride), dtype=torch.int16, device="cuda")
    d_e = torch.zeros((1, I.G), dtype=torch.int16, device="cuda")
    d_res = torch.zeros((1, 4), dtype=torch.int32, device="cuda")
    d_mib = torch.zeros((1, 24), dtype=torch.uint8, device="cuda")
    d_pl = torch.zeros((1, I.payload_stride), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros((1, 4), dtype=torch.int32, device="cuda")
    d_llr = torch.zeros((1, 8 * n_reg), dtype=torch.int16, device="cuda")
    d_found = torch.zeros((1, 4), dtype=torch.int32, device="cuda")
    d_bits = torch.zeros((1, 64), dtype=torch.uint8, device="cuda")
    plan.ofdm_rx(1, d_iq, d_sf)
    plan.chest(1, d_sf, d_ce, d_meas)
    plan.pcfich_decode(1, d_sf, d_ce, d_meas, 0.0, 1, d_cfi, d_corr)
    plan.pbch_decode(1, d_sf, d_ce, d_meas, 0.0, 1, d_res, d_mib)
    plan.pdcch_extract_llr(1, d_sf, d_ce, d_meas, 0.0, 1, d_llr)
    plan.pdcch_find_dci(1, d_llr, c["rnti"], c["nb"], d_found, d_bits, None)
    plan.pdsch_llr(1, d_sf, d_ce, d_meas, 0.0, 1, 0, d_sb, None, d_e)
    plan.decode_batch(1, d_iq, 0.0, 1, 4, d_pl, d_st)
    torch.cuda.synchronize()
    assert np.array_equal(d_sf.cpu().numpy().view(np.complex64)[0], v["sf"])
    assert np.array_equal(d_ce.cpu().numpy().view(np.complex64).reshape(v["ce"].shape), v["ce"])
    assert np.allclose(d_meas.cpu().numpy()[0], v["meas"], rtol=1e-5)
    assert int(d_cfi.cpu()[0]) == int(v["cfi"][0]) and np.array_equal(d_corr.cpu().numpy()[0], v["corr"])
    assert d_res.cpu().numpy()[0, :3].tolist() == v["pbch"].tolist() and np.array_equal(d_mib.cpu().numpy()[0], v["mib"])
    assert np.array_equal(d_llr.cpu().numpy()[0], v["llr"])
    assert d_found.cpu().numpy()[0, :3].tolist() == v["dci"].tolist() and np.array_equal(d_bits.cpu().numpy()[0, :c["nb"]], v["dci_bits"])
    assert np.array_equal(d_e.cpu().numpy()[0, :len(v["e"])], v["e"])
    assert int(d_st.cpu()[0, 0]) == 1 and np.array_equal(d_pl.cpu().numpy()[0, :len(v["payload"])], v["payload"])
    for i, (g, q) in enumerate(c["phich"]):
        d_ack = torch.zeros(1, dtype=torch.int32, device="cuda")
        d_met = torch.zeros(1, dtype=torch.float32, device="cuda")
        plan.phich_decode(1, d_sf, d_ce, d_meas, 0.0, 1, g, q, d_ack, d_met)
        torch.cuda.synchronize()
        assert int(d_ack[0]) == v["phich_ack"][i] and np.float32(d_met[0].item()) == v["phich_metric"][i]
    plan.close()
