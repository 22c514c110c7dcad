"""Cell search (srsue_gpu_cell_search): PSS position / N_id_2 / peak power, SSS cell group and subframe 0-or-5, CFO, against
the oracle on synthetic 1.92 Msps half frames (a 6-PRB cell is exactly that rate) with random timing offsets, carrier
offsets and noise (reference: srslte_ue_cellsearch_scan, phch_recv.cc:146-177)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class Agc(C.Structure):
    """srslte_agc_t of include/srsue_gpu/srslte_compat.h"""
    _fields_ = [("gain", C.c_double), ("set_gain_callback", C.c_void_p), ("handler", C.c_void_p), ("target", C.c_float),
                ("bandwidth", C.c_float), ("max_gain", C.c_float), ("last_power", C.c_float), ("period", C.c_uint32),
                ("count", C.c_uint32), ("nof_updates", C.c_uint32)]


class UeSync(C.Structure):
    """srslte_ue_sync_t of include/srsue_gpu/srslte_compat.h"""
    _fields_ = [("agc", Agc), ("threshold", C.c_float), ("em_alpha", C.c_float), ("gpu", C.c_void_p)]


def _half_frame(o, cell, first_sf, seed, snr, cfo, shift):
    out = []
    for i in range(5):
        cfg = o.make_cfg(cell, sf_idx=first_sf + i, cfi=2, qm=2, tbs=104, tm=cell.nof_ports)
        out.append(o.gen_subframe(cell, cfg, seed + i, snr, None, pcfich=True, sync=True)[1])
    x = np.concatenate(out)
    x = x * np.exp(2j * np.pi * cfo * np.arange(len(x)) / 128)
    return np.roll(x, shift).astype(np.complex64)


def test_cell_search_matches_oracle(gpu, oracle):
    import torch
    sg, ctx = gpu
    o = oracle
    rng = np.random.default_rng(11)
    cases = []
    for cid in (0, 1, 2, 77, 150, 301, 503, 250):
        for first in (0, 5):
            cases.append((cid, first, float(rng.choice([12.0, 3.0, -2.0])), float(rng.uniform(-0.3, 0.3)), int(rng.integers(0, 8000))))
    cases.append((33, 0, -25.0, 0.0, 100))          # noise only: whatever is found must still equal the oracle
    bufs = np.stack([_half_frame(o, o.make_cell(6, 1 + (cid % 2), cid), first, 100 * cid + first, snr, cfo, shift)
                     for cid, first, snr, cfo, shift in cases])
    n, ns = bufs.shape
    d_iq = torch.from_numpy(bufs.view(np.float32).reshape(n, -1)).cuda()
    d_res = torch.zeros(n * C.sizeof(sg.SyncResult), dtype=torch.uint8, device="cuda")
    ctx.cell_search(d_iq, n, ns, ns, d_res)
    torch.cuda.synchronize()
    res = (sg.SyncResult * n).from_buffer_copy(d_res.cpu().numpy().tobytes())
    for i, (cid, first, snr, cfo, shift) in enumerate(cases):
        r, ref = res[i], o.pss_search(bufs[i])
        assert (r.peak_pos, r.n_id_2) == (ref["pos"], ref["n_id_2"])
        assert np.float32(r.peak) == ref["peak"]                                     # bit-identical correlation power
        assert abs(r.mean_power - ref["mean_power"]) <= 1e-5 * ref["mean_power"]
        assert abs(r.cfo - ref["cfo"]) <= 1e-5
        if ref["pos"] >= 137:
            n1, sf5, corr = o.sss_detect(bufs[i], ref["pos"], ref["n_id_2"])
            assert r.valid == 1 and (r.n_id_1, r.sf5) == (n1, sf5) and np.float32(r.sss_corr) == corr
        else:
            assert r.valid == 0 and r.n_id_1 == -1
        if snr > 0:
            assert r.peak_pos == (832 + shift) % 9600 and 3 * r.n_id_1 + r.n_id_2 == cid and r.sf5 == (first == 5)
            assert abs(r.cfo - cfo) < 0.05
            assert r.peak / r.mean_power > 20


def test_forced_root_and_cellsearch_shim(gpu, oracle):
    """srslte_ue_cellsearch_scan / _scan_N_id_2 as phch_recv::init_cell uses them (phch_recv.cc:146-177): the scan pulls
    5 ms frames through the radio callback, searches them in one batch and reports the cell most frames agree on"""
    import torch
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    cid, cfo = 302, 0.11
    cell = o.make_cell(6, 1, cid)
    frames = [_half_frame(o, cell, 5 * (i % 2), 500 + 10 * i, 8.0, cfo, 2000) for i in range(6)]
    # forced root through the batch API: same peak as the free search when the forced root is the right one
    bufs = np.stack(frames)
    n, ns = bufs.shape
    d_iq = torch.from_numpy(bufs.view(np.float32).reshape(n, -1)).cuda()
    d_res = torch.zeros(n * C.sizeof(sg.SyncResult), dtype=torch.uint8, device="cuda")
    ctx.cell_search(d_iq, n, ns, ns, d_res, force_n_id_2=cid % 3)
    torch.cuda.synchronize()
    res = (sg.SyncResult * n).from_buffer_copy(d_res.cpu().numpy().tobytes())
    for i in range(n):
        assert res[i].n_id_2 == cid % 3 and 3 * res[i].n_id_1 + res[i].n_id_2 == cid and res[i].sf5 == i % 2
    ctx.cell_search(d_iq, n, ns, ns, d_res, force_n_id_2=(cid + 1) % 3)
    torch.cuda.synchronize()
    wrong = (sg.SyncResult * n).from_buffer_copy(d_res.cpu().numpy().tobytes())
    assert all(w.peak < 0.2 * r.peak for w, r in zip(wrong, res))

    class Result(C.Structure):
        _fields_ = [("cell_id", C.c_uint32), ("cp", C.c_int), ("peak", C.c_float), ("mode", C.c_float), ("psr", C.c_float), ("cfo", C.c_float)]

    class CellSearch(C.Structure):
        _fields_ = [("ue_sync", UeSync), ("nof_frames_to_scan", C.c_uint32), ("detect_threshold", C.c_float), ("gpu", C.c_void_p)]

    state = {"i": 0}
    RECV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p)

    def recv(handler, data, nsamples, ts):
        f = frames[state["i"] % len(frames)]
        state["i"] += 1
        assert nsamples == len(f)
        C.memmove(data, f.ctypes.data, f.nbytes)
        return nsamples

    cb = RECV(recv)
    cs = CellSearch()
    assert L.srslte_ue_cellsearch_init(C.byref(cs), cb, None) == 0
    assert L.srslte_ue_cellsearch_set_nof_frames_to_scan(C.byref(cs), 6) == 0
    L.srslte_ue_cellsearch_set_threshold.argtypes = [C.c_void_p, C.c_float]
    L.srslte_ue_cellsearch_set_threshold(C.byref(cs), 5.0)
    found = (Result * 3)()
    best = C.c_uint32(9)
    assert L.srslte_ue_cellsearch_scan(C.byref(cs), found, C.byref(best)) == 6
    assert best.value == cid % 3 and found[best.value].cell_id == cid and found[best.value].cp == 0
    assert found[best.value].mode == 1.0 and abs(found[best.value].cfo - cfo * 15000) < 600 and found[best.value].psr > 20
    one = Result()
    assert L.srslte_ue_cellsearch_scan_N_id_2(C.byref(cs), cid % 3, C.byref(one)) == 6 and one.cell_id == cid
    # a wrong root only produces noise peaks: max / mean of ~9500 exponential variates is about ln(9500) = 9
    L.srslte_ue_cellsearch_set_threshold(C.byref(cs), 15.0)
    assert L.srslte_ue_cellsearch_scan_N_id_2(C.byref(cs), (cid + 1) % 3, C.byref(one)) == 0
    assert L.srslte_ue_cellsearch_scan_N_id_2(C.byref(cs), cid % 3, C.byref(one)) == 6
    L.srslte_ue_cellsearch_free(C.byref(cs))


def test_init_cell_sequence_cellsearch_then_mib(gpu, oracle):
    """phch_recv::init_cell (phch_recv.cc:136-226) end to end on a synthetic air interface: cell search at 1.92 Msps,
    then srslte_ue_mib_sync_decode finds a subframe 0 and decodes the MIB -> cell id, ports, bandwidth, PHICH, SFN"""
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import Cell
    cid, ports, sfn0, cfo, shift = 217, 2, 406, 0.05, 6100
    # the central 1.92 MHz of a cell look the same whatever its bandwidth: generate a 6-PRB view whose MIB announces 50 PRB
    cell = o.make_cell(6, ports, cid)
    stream = []
    for half in range(8):                                   # 40 ms of air interface
        sfn = sfn0 + half // 2
        for i in range(5):
            sf = 5 * (half % 2) + i
            cfg = o.make_cfg(cell, sf_idx=sf, cfi=2, qm=2, tbs=104, tm=ports)
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
    assert found[best.value].cell_id == cid
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
    assert sfn0 <= sfn.value + off.value <= sfn0 + 3          # one of the four frames of the synthetic stream
    packed = (C.c_uint8 * 3)()
    L.srslte_bit_pack_vector(payload, packed, 24)
    assert bytes(packed) == np.packbits(np.frombuffer(payload, np.uint8)).tobytes()
    assert L.srslte_sampling_freq_hz(50) == 15360000 and L.srslte_tti_interval(3, 10238) == 5


@pytest.mark.parametrize("prb,nfft", [(15, 256), (25, 512), (50, 1024), (100, 2048)])
def test_pss_sss_at_the_cells_own_rate(gpu, oracle, prb, nfft):
    """tracking-style search at the cell's sampling rate: one root, a window around the expected position, nfft-sample
    replica and nfft-point transforms; position, power, cell group and subframe equal the oracle's"""
    import torch
    sg, ctx = gpu
    o = oracle
    cid = 100 + prb
    cell = o.make_cell(prb, 1, cid)
    bufs = []
    for first in (0, 5):
        sfs = []
        for sf in (first, first + 1):
            cfg = o.make_cfg(cell, sf_idx=sf, cfi=2, qm=2, tbs=1000, tm=1)
            sfs.append(o.gen_subframe(cell, cfg, 40 + sf, 7.0, None, pcfich=True, sync=True)[1])
        x = np.concatenate(sfs)
        bufs.append((x * np.exp(2j * np.pi * 0.07 * np.arange(len(x)) / nfft)).astype(np.complex64))
    exp = 832 * nfft // 128
    ns = exp + 3 * nfft
    win = np.stack([b[:ns] for b in bufs])
    d_iq = torch.from_numpy(win.view(np.float32).reshape(2, -1)).cuda()
    d_res = torch.zeros(2 * C.sizeof(sg.SyncResult), dtype=torch.uint8, device="cuda")
    ctx.cell_search(d_iq, 2, ns, ns, d_res, force_n_id_2=cid % 3, first_pos=exp - 40, nfft=nfft)
    torch.cuda.synchronize()
    res = (sg.SyncResult * 2).from_buffer_copy(d_res.cpu().numpy().tobytes())
    for i in range(2):
        ref = o.pss_search(win[i], nfft, cid % 3, exp - 40)
        n1, sf5, corr = o.sss_detect(win[i], ref["pos"], ref["n_id_2"], nfft)
        r = res[i]
        assert (r.peak_pos, r.n_id_2, r.n_id_1, r.sf5) == (ref["pos"], ref["n_id_2"], n1, sf5)
        assert np.float32(r.peak) == ref["peak"] and np.float32(r.sss_corr) == corr and abs(r.cfo - ref["cfo"]) <= 1e-5
        # at an oversampled rate the correlation peak is several samples wide: noise may move it by a sample
        assert abs(r.peak_pos - exp) <= 2 and 3 * r.n_id_1 + r.n_id_2 == cid and r.sf5 == i and abs(r.cfo - 0.07) < 0.06


@pytest.mark.parametrize("cp", [0, 1])
def test_ue_sync_finds_tracks_and_follows_timing_slips(gpu, oracle, cp):
    """(both cyclic prefixes: the SSS check of the tracker looks where the cell's prefix puts it)
    srslte_ue_sync_zerocopy (phch_recv.cc:322) on a continuous 5 MHz stream that starts at an arbitrary sample, loses
    two samples at one point and repeats one at another (sampling-clock drift): 0 while searching, then exactly one aligned
    subframe per call with the right subframe index, re-aligned within a few subframes after each slip"""
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import Cell
    prb, nfft, cid = 25, 512, 183
    sf_len = 15 * nfft
    cell = o.make_cell(prb, 1, cid, cp=cp)
    sfs = []
    for sf in range(10):
        cfg = o.make_cfg(cell, sf_idx=sf, cfi=2, qm=2, tbs=1000, tm=1)
        sfs.append(o.gen_subframe(cell, cfg, 300 + sf, 10.0, None, pcfich=True, sync=True)[1])
    frame = np.concatenate(sfs)
    S = np.tile(frame, 9)                                    # 90 subframes
    starts = np.arange(90) * sf_len                          # true start of every subframe in S
    # two samples vanish inside subframe 42, one sample is repeated inside subframe 65
    p1, p2 = 42 * sf_len + 1000, 65 * sf_len + 2000
    M = np.concatenate([S[:p1], S[p1 + 2:p2], S[p2 - 1:]])
    mstart = starts.copy()
    mstart[43:] -= 2
    mstart[66:] += 1
    state = {"pos": 3333}                                    # the radio starts somewhere inside subframe 0
    RECV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p)

    def recv(handler, data, nsamples, ts):
        a = state["pos"]
        assert a + nsamples <= len(M)
        C.memmove(data, M[a:a + nsamples].ctypes.data, nsamples * 8)
        state["pos"] = a + nsamples
        return nsamples

    cb = RECV(recv)
    q = UeSync()
    c = Cell(nof_prb=prb, nof_ports=1, bw_idx=0, id=cid, cp=cp, phich_length=0, phich_resources=2)
    assert L.srslte_ue_sync_init(C.byref(q), c, cb, None) == 0
    buf = np.zeros(sf_len, np.complex64)
    delivered, zeros = [], 0
    while state["pos"] + 6 * sf_len < len(M):
        rc = L.srslte_ue_sync_zerocopy(C.byref(q), buf.ctypes.data_as(C.c_void_p))
        assert rc >= 0
        if rc == 0:
            zeros += 1
            assert not delivered, "lost synchronisation"
            continue
        end = state["pos"]                                    # the buffer ends where the stream position is now
        delivered.append((end - sf_len, L.srslte_ue_sync_get_sfidx(C.byref(q)), buf.copy()))
    assert 1 <= zeros <= 3 and len(delivered) > 60
    aligned = 0
    for a, sfidx, data in delivered:
        j = int(np.argmin(np.abs(mstart - a)))
        assert sfidx == j % 10
        if 43 <= j <= 46 or 66 <= j <= 71:                    # re-aligning after a slip (next PSS check + one subframe)
            continue
        assert a == mstart[j], (j, a - mstart[j])
        aligned += 1
    assert aligned > 50
    L.srslte_ue_sync_get_cfo.restype = C.c_float
    L.srslte_ue_sync_get_sfo.restype = C.c_float
    assert abs(L.srslte_ue_sync_get_cfo(C.byref(q))) < 600.0
    L.srslte_ue_sync_free(C.byref(q))


def test_cpp_driver_acquire_mode(gpu, oracle, tmp_path):
    """driver/pdsch_offline.cc acquire: phch_recv's start-up (cell search, MIB search, subframe synchronisation with the
    frame number from the MIB) in C++ over a capture file that wraps around"""
    import os, struct, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "build", "pdsch_offline")
    if not os.path.exists(exe):
        pytest.skip("driver not built (run __graft_entry__.build())")
    o = oracle
    cid, ports, sfn0 = 344, 1, 100
    cell = o.make_cell(6, ports, cid)
    stream = []
    for fr in range(4):                                   # 40 ms: one BCH period, so the wrap-around is seamless
        for sf in range(10):
            cfg = o.make_cfg(cell, sf_idx=sf, cfi=2, qm=2, tbs=104, tm=ports)
            mib = (o.mib_pack(6, 0, 6, sfn0 + fr), (sfn0 + fr) % 4) if sf == 0 else None
            stream.append(o.gen_subframe(cell, cfg, 100 * fr + sf, 10.0, None, pcfich=True, sync=True, mib=mib)[1])
    x = np.roll(np.concatenate(stream), 4567).astype(np.complex64)
    fin, fout = str(tmp_path / "cap.bin"), str(tmp_path / "out.txt")
    with open(fin, "wb") as f:
        f.write(struct.pack("2i", 0x53525355, len(x)))
        f.write(x.tobytes())
    r = subprocess.run([exe, "acquire", fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = open(fout).read().splitlines()
    first = dict(zip(lines[0].split()[::2], lines[0].split()[1::2]))
    assert (int(first["cell_id"]), int(first["ports"]), int(first["prb"]), int(first["phich_ng"])) == (cid, ports, 6, 2)
    ttis = [int(l.split()[1]) for l in lines if l.startswith("tti")]
    assert len(ttis) >= 3 and all(10 * sfn0 <= t <= 10 * (sfn0 + 3) and t % 10 == 0 for t in ttis)
    last = dict(zip(lines[-1].split()[::2], lines[-1].split()[1::2]))
    assert int(last["delivered"]) == 40 and int(last["sf_errors"]) == 0 and int(last["mib_decoded"]) == len(ttis)


def test_ue_sync_agc_closes_the_loop(gpu, oracle):
    """srslte_ue_sync_start_agc (phch_recv.cc:111): a radio whose samples scale with the gain it was last told; the loop
    settles the mean sample power at the target within a few blocks, keeps the synchroniser locked, honours the period
    set at phch_recv.cc:302, and srslte_agc_get_gain reports the radio's answer."""
    sg, ctx = gpu
    o = oracle
    L = sg.lib()
    from tests.srslte_ctypes import Cell
    prb, nfft, cid = 6, 128, 77
    sf_len = 15 * nfft
    cell = o.make_cell(prb, 1, cid)
    sfs = []
    for sf in range(10):
        cfg = o.make_cfg(cell, sf_idx=sf, cfi=3, qm=2, tbs=104, tm=1)
        sfs.append(o.gen_subframe(cell, cfg, 800 + sf, 15.0, None, pcfich=True, sync=True)[1])
    S = np.tile(np.concatenate(sfs), 8)
    p_nat = float(np.mean(np.abs(S) ** 2))
    radio = {"pos": 777, "gain": 0.0, "calls": []}
    g_ref = 30.0                                              # the stream has its natural power at 30 dB of gain
    RECV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p)
    GAIN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_double)

    def recv(handler, data, nsamples, ts):
        a = radio["pos"]
        x = (S[a:a + nsamples] * np.float32(10.0 ** ((radio["gain"] - g_ref) / 20.0))).astype(np.complex64)
        C.memmove(data, x.ctypes.data, nsamples * 8)
        radio["pos"] = a + nsamples
        return nsamples

    def set_gain(handler, gain):
        radio["gain"] = round(gain * 2.0) / 2.0               # the radio has 0.5 dB steps
        radio["calls"].append(radio["gain"])
        return radio["gain"]

    cb, gcb = RECV(recv), GAIN(set_gain)
    q = UeSync()
    c = Cell(nof_prb=prb, nof_ports=1, bw_idx=0, id=cid, cp=0, phich_length=0, phich_resources=2)
    assert L.srslte_ue_sync_init(C.byref(q), c, cb, None) == 0
    assert L.srslte_ue_sync_start_agc(C.byref(q), gcb, C.c_float(55.0)) == 0          # far too hot: 25 dB above natural
    L.srslte_agc_get_gain.restype = C.c_float
    buf = np.zeros(sf_len, np.complex64)
    got, powers = 0, []
    while radio["pos"] + 12 * sf_len < len(S):
        rc = L.srslte_ue_sync_zerocopy(C.byref(q), buf.ctypes.data_as(C.c_void_p))
        assert rc >= 0
        if rc == 1:
            got += 1
            powers.append(float(np.mean(np.abs(buf) ** 2)))
            if got == 20:
                n_before = len(radio["calls"])
                L.srslte_ue_sync_set_agc_period(C.byref(q), 20)
    want = g_ref + 10.0 * np.log10(0.1 / p_nat)
    assert got > 40 and abs(radio["gain"] - want) <= 1.0 and abs(L.srslte_agc_get_gain(C.byref(q.agc)) - radio["gain"]) < 1e-6
    # subframes differ in content (sync signals, empty control symbols), so single subframes sit a few dB around the target
    assert abs(10 * np.log10(np.mean(powers[10:]) / 0.1)) < 2.0 and all(abs(10 * np.log10(p / 0.1)) < 6.0 for p in powers[10:])
    assert len(radio["calls"]) - n_before <= (got - 20) // 20 + 1 and q.agc.nof_updates == len(radio["calls"])
    L.srslte_ue_sync_free(C.byref(q))
