"""ctypes binding of the CPU oracle (oracle/liblteoracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/lte_oracle.h.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by the
product package srsue_b200.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblteoracle.so")
_SO_AVX2 = os.path.join(_HERE, "_build", "liblteoracle_avx2.so")   # same sources, -O3 -mavx2 + window-parallel turbo

MAX_K = 6144
SB_STRIDE = 3 * MAX_K + 12
CRC24A, CRC24B, CRC16 = 0x1864CFB, 0x1800063, 0x11021
TD_C, TD_E, TD_INF = 511, 2047, 10000


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h", ".inc"))]
    if (not force and os.path.exists(_SO) and os.path.exists(_SO_AVX2)
            and all(min(os.path.getmtime(_SO), os.path.getmtime(_SO_AVX2)) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-B", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


class Cell(C.Structure):
    # cp: 0 = normal cyclic prefix, 1 = extended (12 symbols per subframe; grids keep their 14-row stride)
    _fields_ = [("nof_prb", C.c_int), ("nof_ports", C.c_int), ("cell_id", C.c_int), ("cp", C.c_int)]


class PdschCfg(C.Structure):
    _fields_ = [("sf_idx", C.c_int), ("cfi", C.c_int), ("rnti", C.c_int), ("qm", C.c_int),
                ("tbs", C.c_int), ("rv", C.c_int), ("tm", C.c_int), ("nof_prb_alloc", C.c_int),
                ("prb_mask", C.c_uint8 * 110)]


class CbSegm(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("tbs", "B", "C", "Kp", "Km", "Cp", "Cm", "F")]


_libs = {}
_kind = "portable"


def have_avx2():
    try:
        with open("/proc/cpuinfo") as f:
            return " avx2 " in f.read().replace("\n", " ")
    except OSError:
        return False


def select(kind):
    """choose the build every function of this module calls: "portable" (-O2 scalar C, the checker) or "avx2"
    (bit-identical, vectorised; the CPU baseline bench.py times).  Returns the previous choice."""
    global _kind
    assert kind in ("portable", "avx2")
    if kind == "avx2" and not have_avx2():
        raise RuntimeError("this host has no AVX2")
    prev, _kind = _kind, kind
    return prev


def lib():
    if _kind not in _libs:
        so = _SO if _kind == "portable" else _SO_AVX2
        build()
        try:
            L = C.CDLL(so)
        except OSError:
            build(force=True)
            L = C.CDLL(so)
        L.lteo_crc_bits.restype = C.c_uint32
        _libs[_kind] = L
    return _libs[_kind]


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def make_cell(nof_prb, nof_ports=1, cell_id=1, cp=0):
    return Cell(nof_prb, nof_ports, cell_id, cp)


def make_cfg(cell, sf_idx=1, cfi=1, rnti=0x1234, qm=2, tbs=152, rv=0, tm=1, prbs=None, prbs_slot1=None):
    """prbs: allocated PRBs (all when None); prbs_slot1: the PRBs of the second slot when they differ from the first
    (distributed virtual resource blocks) -- prb_mask then holds 1 (both slots), 2 (slot 0 only) or 4 (slot 1 only)"""
    cfg = PdschCfg()
    cfg.sf_idx, cfg.cfi, cfg.rnti, cfg.qm, cfg.tbs, cfg.rv, cfg.tm = sf_idx, cfi, rnti, qm, tbs, rv, tm
    s0 = set(range(cell.nof_prb) if prbs is None else prbs)
    s1 = s0 if prbs_slot1 is None else set(prbs_slot1)
    for p in s0 | s1:
        cfg.prb_mask[p] = 1 if (p in s0 and p in s1) else 2 if p in s0 else 4
    cfg.nof_prb_alloc = len(s0)
    return cfg


# ---- small helpers -------------------------------------------------------------------------
def qpp_Ks():
    L = lib()
    return [L.lteo_qpp_K(i) for i in range(L.lteo_qpp_table_size())]


def qpp_params(K):
    f1, f2 = C.c_int(), C.c_int()
    if lib().lteo_qpp_params(K, C.byref(f1), C.byref(f2)):
        raise ValueError(K)
    return f1.value, f2.value


def qpp_perm(K):
    pi = np.zeros(K, np.uint16)
    lib().lteo_qpp_perm(K, _p(pi))
    return pi


def window_len(K):
    return lib().lteo_window_len(K)


def crc_bits(bits, poly, order=24):
    bits = np.ascontiguousarray(bits, np.uint8)
    return lib().lteo_crc_bits(_p(bits), len(bits), C.c_uint32(poly), order)


def gold(c_init, n):
    c = np.zeros(n, np.uint8)
    lib().lteo_gold(C.c_uint32(c_init), n, _p(c))
    return c


def cbsegm(tbs):
    s = CbSegm()
    if lib().lteo_cbsegm(tbs, C.byref(s)):
        raise ValueError(tbs)
    return s


def cb_len(s, r):
    return lib().lteo_cb_len(C.byref(s), r)


def cb_E(s, G, qm, nl, r):
    return lib().lteo_cb_E(C.byref(s), G, qm, nl, r)


def pdsch_re_list(cell, cfg):
    nsc = 12 * cell.nof_prb
    re = np.zeros(14 * nsc, np.int32)
    n = lib().lteo_pdsch_re_list(C.byref(cell), C.byref(cfg), _p(re))
    return re[:n].copy()


def fft_twiddles(n):
    tw = np.zeros(n // 2, np.complex64)
    lib().lteo_fft_twiddles(n, _p(tw))
    return tw


def rm_sequence(K, F, rv):
    seq = np.zeros(3 * (K + 4), np.int32)
    n = lib().lteo_rm_sequence(K, F, rv, _p(seq))
    return seq[:n].copy()


# ---- TX ----------------------------------------------------------------------------------
def turbo_encode(c):
    c = np.ascontiguousarray(c, np.uint8)
    d = np.zeros(3 * (len(c) + 4), np.uint8)
    lib().lteo_turbo_encode(_p(c), len(c), _p(d))
    return d


def ulsch_encode(tbs, qm, nof_prb, tb_bytes, rv=0, rnti=0x1234, sf_idx=0, cell_id=1, n_symb=12):
    """UL-SCH coding (no control information) + PUSCH scrambling -> G = 12 nof_prb n_symb qm bits (one per byte)."""
    tb = np.ascontiguousarray(tb_bytes, dtype=np.uint8)
    out = np.zeros(12 * nof_prb * n_symb * qm, np.uint8)
    G = lib().lteo_ulsch_encode(int(tbs), int(qm), int(nof_prb), int(n_symb), int(rv), int(rnti), int(sf_idx), int(cell_id), _p(tb), _p(out))
    if G < 0:
        raise ValueError("lteo_ulsch_encode failed (%d)" % G)
    return out


def pdsch_tx_grid(cell, cfg, tb_bytes):
    nsc = 12 * cell.nof_prb
    grid = np.zeros((cell.nof_ports, 14, nsc), np.complex128)
    tb = np.ascontiguousarray(tb_bytes, np.uint8)
    rc = lib().lteo_pdsch_tx_grid(C.byref(cell), C.byref(cfg), _p(tb), _p(grid))
    if rc:
        raise RuntimeError("pdsch_tx_grid rc=%d" % rc)
    return grid


def ofdm_tx(nof_prb, grid, cp=0):
    n = lib().lteo_symbol_sz(nof_prb)
    grid = np.ascontiguousarray(grid, np.complex128)
    iq = np.zeros(15 * n, np.complex128)
    lib().lteo_ofdm_tx_cp(nof_prb, cp, _p(grid), _p(iq))
    return iq


# ---- RX ----------------------------------------------------------------------------------
def fft(x):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.zeros_like(x)
    lib().lteo_fft(_p(x), _p(out), len(x))
    return out


def ofdm_rx(nof_prb, iq, cp=0):
    iq = np.ascontiguousarray(iq, np.complex64)
    sf = np.zeros(14 * 12 * nof_prb, np.complex64)
    lib().lteo_ofdm_rx_cp(nof_prb, cp, _p(iq), _p(sf))
    return sf


def chest(cell, sf_idx, sf):
    sf = np.ascontiguousarray(sf, np.complex64)
    ce = np.zeros((cell.nof_ports, 14 * 12 * cell.nof_prb), np.complex64)
    meas = np.zeros(5, np.float32)
    lib().lteo_chest(C.byref(cell), sf_idx, _p(sf), _p(ce), _p(meas))
    return ce, meas


def equalize(cell, cfg, sf, ce, n0):
    sf = np.ascontiguousarray(sf, np.complex64)
    ce = np.ascontiguousarray(ce, np.complex64)
    d = np.zeros(14 * 12 * cell.nof_prb, np.complex64)
    n = C.c_int()
    lib().lteo_equalize(C.byref(cell), C.byref(cfg), _p(sf), _p(ce), C.c_float(n0), _p(d), C.byref(n))
    return d[:n.value].copy()


def demod(d, qm):
    d = np.ascontiguousarray(d, np.complex64)
    llr = np.zeros(len(d) * qm, np.int16)
    lib().lteo_demod(_p(d), len(d), qm, _p(llr))
    return llr


def descramble(llr, c_init):
    llr = np.ascontiguousarray(llr, np.int16).copy()
    lib().lteo_descramble(_p(llr), len(llr), C.c_uint32(c_init))
    return llr


def rm_rx(e, K, F, rv, w=None):
    e = np.ascontiguousarray(e, np.int16)
    if w is None:
        w = np.zeros(3 * K + 12, np.int16)
    lib().lteo_rm_rx(_p(e), len(e), K, F, rv, _p(w))
    return w


def tdec(inp, K, max_iter=4, crc_type=0, window=0):
    inp = np.ascontiguousarray(inp, np.int16)
    assert len(inp) >= 3 * K + 12
    bits = np.zeros(K, np.uint8)
    ok = C.c_int()
    la = np.zeros(K, np.int16)
    it = lib().lteo_tdec_dbg(_p(inp), K, max_iter, crc_type, _p(bits), C.byref(ok), _p(la), window)
    return bits, it, ok.value, la


def new_softbuf(ncb=13):
    return np.zeros((ncb, SB_STRIDE), np.int16)


def pdsch_decode(cell, cfg, sf, ce, noise_est, max_iter=4, softbuf=None, want=False):
    s = cbsegm(cfg.tbs)
    if softbuf is None:
        softbuf = new_softbuf(s.C)
    sf = np.ascontiguousarray(sf, np.complex64)
    ce = np.ascontiguousarray(ce, np.complex64)
    payload = np.zeros((cfg.tbs + 7) // 8, np.uint8)
    nsc = 12 * cell.nof_prb
    d = np.zeros(14 * nsc, np.complex64)
    e = np.zeros(14 * nsc * cfg.qm, np.int16)
    iters = np.zeros(s.C, np.int32)
    crc = np.zeros(s.C, np.int32)
    rc = lib().lteo_pdsch_decode(C.byref(cell), C.byref(cfg), _p(sf), _p(ce), C.c_float(noise_est), max_iter,
                                 _p(softbuf), _p(payload), _p(d), _p(e), _p(iters), _p(crc))
    if want:
        return rc, payload, dict(d=d, e=e, iters=iters, crc=crc, softbuf=softbuf)
    return rc, payload


def ue_dl_decode(cell, cfg, iq, noise_est=0.01, noise_mode=0, max_iter=4, softbuf=None):
    s = cbsegm(cfg.tbs)
    if softbuf is None:
        softbuf = new_softbuf(s.C)
    iq = np.ascontiguousarray(iq, np.complex64)
    payload = np.zeros((cfg.tbs + 7) // 8, np.uint8)
    meas = np.zeros(5, np.float32)
    avg = C.c_int()
    rc = lib().lteo_ue_dl_decode(C.byref(cell), C.byref(cfg), _p(iq), C.c_float(noise_est), noise_mode, max_iter,
                                 _p(softbuf), _p(payload), _p(meas), C.byref(avg))
    return rc, payload, meas, avg.value


def ue_dl_decode_mt(cell, cfg, iq, nthreads, noise_est=0.01, noise_mode=0, max_iter=4):
    """batch of subframes on `nthreads` host threads (bench.py cpu_baseline / reference arm)"""
    iq = np.ascontiguousarray(iq, np.complex64)
    n = iq.shape[0]
    payload = np.zeros((n, (cfg.tbs + 7) // 8), np.uint8)
    status = np.zeros((n, 2), np.int32)
    ok = lib().lteo_ue_dl_decode_mt(C.byref(cell), C.byref(cfg), _p(iq), n, C.c_float(noise_est), noise_mode, max_iter,
                                    nthreads, _p(payload), _p(status))
    return ok, payload, status


def tdec_mt(llrs, K, nthreads, max_iter=4, crc_type=0):
    llrs = np.ascontiguousarray(llrs, np.int16)
    n = llrs.shape[0]
    bits = np.zeros((n, K), np.uint8)
    iters = np.zeros(n, np.int32)
    lib().lteo_tdec_mt(_p(llrs), n, K, max_iter, crc_type, nthreads, _p(bits), _p(iters))
    return bits, iters


# ---- synthetic subframes (SURVEY.md 8d seeds) ----------------------------------------------
def pcfich_re(cell):
    k = np.zeros(16, np.int32)
    lib().lteo_pcfich_re(C.byref(cell), _p(k))
    return k


def pcfich_decode(cell, sf_idx, sf, ce, noise_est=0.01):
    sf = np.ascontiguousarray(sf, np.complex64)
    ce = np.ascontiguousarray(ce, np.complex64)
    corr = np.zeros(3, np.int32)
    cfi = lib().lteo_pcfich_decode(C.byref(cell), sf_idx, _p(sf), _p(ce), C.c_float(noise_est), _p(corr))
    return cfi, corr


class DciTx(C.Structure):
    _fields_ = [("bits", C.c_void_p), ("nof_bits", C.c_int), ("rnti", C.c_uint16), ("L", C.c_int), ("ncce", C.c_int)]


def pdcch_regs(cell, cfi, ng_x6=6):
    rk = np.zeros(12 * cell.nof_prb, np.int32)
    rl = np.zeros(12 * cell.nof_prb, np.int32)
    n = lib().lteo_pdcch_regs(C.byref(cell), cfi, ng_x6, _p(rk), _p(rl))
    return rk[:n], rl[:n]


def pdcch_quad_perm(n_quad, cell_id):
    src = np.zeros(n_quad, np.int32)
    lib().lteo_pdcch_quad_perm(n_quad, cell_id, _p(src))
    return src


def pdcch_search_space(nof_cce, sf_idx, rnti, common=False):
    cl, cn = np.zeros(32, np.int32), np.zeros(32, np.int32)
    n = lib().lteo_pdcch_search_space(nof_cce, sf_idx, C.c_uint16(rnti), int(common), _p(cl), _p(cn))
    return list(zip(cl[:n].tolist(), cn[:n].tolist()))


def dci_encode(bits, rnti, E):
    bits = np.ascontiguousarray(bits, np.uint8)
    e = np.zeros(E, np.uint8)
    lib().lteo_dci_encode(_p(bits), len(bits), C.c_uint16(rnti), E, _p(e))
    return e


def pdcch_tx(cell, sf_idx, cfi, dcis, grid, ng_x6=6):
    """dcis: list of (bits uint8 array, rnti, L, ncce); adds them to grid [ports][14][nsc] complex128"""
    keep = [np.ascontiguousarray(d[0], np.uint8) for d in dcis]
    arr = (DciTx * max(len(dcis), 1))()
    for a, k, d in zip(arr, keep, dcis):
        a.bits, a.nof_bits, a.rnti, a.L, a.ncce = k.ctypes.data, len(k), d[1], d[2], d[3]
    rc = lib().lteo_pdcch_tx(C.byref(cell), sf_idx, cfi, ng_x6, len(dcis), arr, _p(grid))
    if rc < 0:
        raise RuntimeError("pdcch_tx: DCI does not fit the control region")
    return rc


def pdcch_extract_llr(cell, sf_idx, cfi, sf, ce, noise_est=0.01, ng_x6=6):
    sf = np.ascontiguousarray(sf, np.complex64)
    ce = np.ascontiguousarray(ce, np.complex64)
    llr = np.zeros(8 * 12 * cell.nof_prb, np.int16)
    ncce = lib().lteo_pdcch_extract_llr(C.byref(cell), sf_idx, cfi, ng_x6, _p(sf), _p(ce), C.c_float(noise_est), _p(llr))
    return llr, ncce


def pdcch_decode_candidate(llr, L, nof_bits):
    llr = np.ascontiguousarray(llr, np.int16)
    out = np.zeros(nof_bits, np.uint8)
    lib().lteo_pdcch_decode_candidate.restype = C.c_uint16
    r = lib().lteo_pdcch_decode_candidate(_p(llr), L, nof_bits, _p(out))
    return out, r


def pdcch_find_dci(llr, nof_cce, sf_idx, rnti, nof_bits, common=False):
    llr = np.ascontiguousarray(llr, np.int16)
    out = np.zeros(nof_bits, np.uint8)
    L, n = C.c_int(), C.c_int()
    found = lib().lteo_pdcch_find_dci(_p(llr), nof_cce, sf_idx, C.c_uint16(rnti), int(common), nof_bits, _p(out), C.byref(L), C.byref(n))
    return found, out, L.value, n.value


def phich_index(nof_prb, I_lowest, n_dmrs, ng_x6=6, cp=0):
    g, q = C.c_int(), C.c_int()
    lib().lteo_phich_index_cp(nof_prb, ng_x6, cp, I_lowest, n_dmrs, C.byref(g), C.byref(q))
    return g.value, q.value


def phich_res(cell, n_group, ng_x6=6):
    k = np.zeros(12, np.int32)
    lib().lteo_phich_res(C.byref(cell), ng_x6, n_group, _p(k))
    return k


def phich_decode(cell, sf_idx, sf, ce, n_group, n_seq, noise_est=0.0, ng_x6=6):
    sf = np.ascontiguousarray(sf, np.complex64)
    ce = np.ascontiguousarray(ce, np.complex64)
    m = C.c_float()
    ack = lib().lteo_phich_decode(C.byref(cell), sf_idx, ng_x6, _p(sf), _p(ce), C.c_float(noise_est), n_group, n_seq, C.byref(m))
    return ack, np.float32(m.value)


def pss_search(x, nfft=128, force_n_id_2=-1, first_pos=0):
    x = np.ascontiguousarray(x, np.complex64)
    pos, nid2 = C.c_int(), C.c_int()
    cfo, mp = C.c_float(), C.c_float()
    lib().lteo_pss_search_n.restype = C.c_float
    peak = lib().lteo_pss_search_n(_p(x), len(x), nfft, force_n_id_2, first_pos, C.byref(pos), C.byref(nid2), C.byref(cfo), C.byref(mp))
    return dict(peak=np.float32(peak), pos=pos.value, n_id_2=nid2.value, cfo=np.float32(cfo.value), mean_power=np.float32(mp.value))


def cfo_step(cfo, nfft):
    lib().lteo_cfo_step.restype = C.c_int32
    return int(lib().lteo_cfo_step(C.c_float(cfo), nfft))


def cfo_correct(x, step):
    x = np.ascontiguousarray(x, np.complex64)
    y = np.zeros_like(x)
    lib().lteo_cfo_correct(_p(x), _p(y), len(x), C.c_int32(step))
    return y


def cfo_table():
    t = np.zeros(4096, np.complex64)
    lib().lteo_cfo_table(_p(t))
    return t


def sss_detect(x, peak_pos, n_id_2, nfft=128):
    x = np.ascontiguousarray(x, np.complex64)
    sf5 = C.c_int()
    corr = C.c_float()
    n1 = lib().lteo_sss_detect_n(_p(x), peak_pos, n_id_2, nfft, C.byref(sf5), C.byref(corr))
    return n1, sf5.value, np.float32(corr.value)


def sss_detect_cp(x, peak_pos, n_id_2, nfft=128, cp_mode=2):
    """cp_mode 0 / 1: normal / extended prefix only, 2: both; returns (n_id_1, sf5, corr, cp)"""
    x = np.ascontiguousarray(x, np.complex64)
    sf5, cp = C.c_int(), C.c_int()
    corr = C.c_float()
    n1 = lib().lteo_sss_detect_cp(_p(x), peak_pos, n_id_2, nfft, cp_mode, C.byref(sf5), C.byref(corr), C.byref(cp))
    return n1, sf5.value, np.float32(corr.value), cp.value


def mib_pack(nof_prb, phich_ext, ng_x6, sfn):
    """the 24 MIB bits (36.331): dl-Bandwidth, phich-Duration, phich-Resource, the 8 MSBs of the SFN, 10 spare zeros"""
    bw = [6, 15, 25, 50, 75, 100].index(nof_prb)
    ng = [1, 3, 6, 12].index(ng_x6)
    v = (bw << 21) | (int(phich_ext) << 20) | (ng << 18) | (((sfn >> 2) & 0xFF) << 10)
    return np.array([(v >> (23 - i)) & 1 for i in range(24)], np.uint8)


def pbch_decode(cell, sf, ce, noise_est=0.0):
    sf = np.ascontiguousarray(sf, np.complex64)
    ce = np.ascontiguousarray(ce, np.complex64)
    bits = np.zeros(24, np.uint8)
    p, q = C.c_int(), C.c_int()
    f = lib().lteo_pbch_decode(C.byref(cell), _p(sf), _p(ce), C.c_float(noise_est), _p(bits), C.byref(p), C.byref(q))
    return f, bits, p.value, q.value


def pbch_llr(cell, hyp_ports, sf, ce, noise_est=0.0):
    sf = np.ascontiguousarray(sf, np.complex64)
    ce = np.ascontiguousarray(ce, np.complex64)
    llr = np.zeros(480, np.int16)
    lib().lteo_pbch_llr(C.byref(cell), hyp_ports, _p(sf), _p(ce), C.c_float(noise_est), _p(llr))
    return llr


def gen_subframe(cell, cfg, seed, snr_db=30.0, taps=None, pcfich=False, dcis=None, ng_x6=6, phichs=None, mib=None, sync=False,
                 noise_seed=None):
    """One synthetic DL subframe: returns (tb_bytes, iq complex64 of 15*N_FFT samples, sigma2).

    Payload RNG: numpy default_rng(seed); noise RNG: default_rng(seed + 5_000_000).  `taps` is an
    optional [ports][ntaps] complex array of channel taps (sample-spaced, shorter than the CP)."""
    rng = np.random.default_rng(seed)
    tb = rng.integers(0, 256, (cfg.tbs + 7) // 8, dtype=np.uint8)
    grid = pdsch_tx_grid(cell, cfg, tb)
    if pcfich:          # control format indicator of this subframe in symbol 0 (off by default: older fixtures)
        lib().lteo_pcfich_tx(C.byref(cell), cfg.sf_idx, cfg.cfi, _p(grid))
    if dcis:            # PDCCHs of this subframe: list of (bits, rnti, L, ncce)
        pdcch_tx(cell, cfg.sf_idx, cfg.cfi, dcis, grid, ng_x6)
    if sync and cfg.sf_idx in (0, 5):     # PSS + SSS of the cell
        lib().lteo_sync_tx(C.byref(cell), cfg.sf_idx, _p(grid))
    if mib is not None:     # (24 MIB bits, radio frame number mod 4): PBCH of a subframe 0
        assert cfg.sf_idx == 0
        mb = np.ascontiguousarray(mib[0], np.uint8)
        lib().lteo_pbch_tx(C.byref(cell), _p(mb), int(mib[1]), _p(grid))
    for (g_, q_, ack_) in (phichs or []):     # HARQ indicators: (n_group, n_seq, ack)
        lib().lteo_phich_tx(C.byref(cell), cfg.sf_idx, ng_x6, g_, q_, ack_, _p(grid))
    n = lib().lteo_symbol_sz(cell.nof_prb)
    nsc = 12 * cell.nof_prb
    rx = np.zeros((14, nsc), np.complex128)
    k = np.arange(nsc)
    bins = np.where(k < nsc // 2, n - nsc // 2 + k, k - nsc // 2 + 1)
    for p in range(cell.nof_ports):
        if taps is None:
            h = np.ones(nsc, np.complex128)
        else:
            t = np.asarray(taps[p], np.complex128)
            h = (t[None, :] * np.exp(-2j * np.pi * np.outer(bins, np.arange(len(t))) / n)).sum(1)
        rx += grid[p] * h[None, :]
    iq = ofdm_tx(cell.nof_prb, rx, cell.cp)
    # unitary (I)FFT on both sides: a unit-power RE stays unit power and the per-RE noise variance
    # equals the per-sample one, so snr_db is Es/N0 per resource element
    sigma2 = 10.0 ** (-snr_db / 10.0)
    nrng = np.random.default_rng(seed + 5_000_000 if noise_seed is None else noise_seed)   # retransmissions: same payload, new noise
    noise = (nrng.standard_normal(len(iq)) + 1j * nrng.standard_normal(len(iq))) * np.sqrt(sigma2 / 2)
    return tb, (iq + noise).astype(np.complex64), sigma2


def gen_turbo_llrs(K, seed, ebn0_db=None, scale=64.0):
    """cfg4 input: K random info bits -> turbo encoder -> LLR int16 [3K+12] (srsLTE triples order)."""
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 2, K, dtype=np.uint8)
    d = turbo_encode(c).astype(np.float64)
    s = 2.0 * d - 1.0                                 # bit 1 -> +1 (positive LLR <=> bit 1)
    if ebn0_db is not None:
        sigma2 = 1.0 / (2.0 * (1.0 / 3.0) * 10.0 ** (ebn0_db / 10.0))
        nrng = np.random.default_rng(seed + 5_000_000)
        s = s + nrng.standard_normal(len(s)) * np.sqrt(sigma2)
    llr = np.clip(np.trunc(scale * s), -2048, 2047).astype(np.int16)
    return c, llr
