"""Host-side resource-allocation logic (srsue_b200/csrc/ra.cc) that stands in for srsLTE's srslte_dci_msg_to_dl_grant
(ue/src/phy/phch_worker.cc:297): DCI 1A / 1 payload <-> fields <-> grant.  No GPU needed; the checks are written from
36.212 5.3.3.1 and 36.213 7.1.6 / 7.1.7 independently of the C++ (different formulation of allocation type 1)."""
import ctypes as C

import numpy as np
import pytest

from tests.srslte_ctypes import DciMsg, Grant, RaDlDci, install_tbs_table

BWS = [6, 15, 25, 50, 75, 100]
FMT1, FMT1A = 1, 2


@pytest.fixture(scope="module")
def L():
    import srsue_b200 as sg
    lib = sg.lib()
    lib.srslte_ra_type2_to_riv.restype = C.c_uint32
    lib.srslte_ra_type2_n_rb.restype = C.c_uint32
    lib.srslte_ra_type0_P.restype = C.c_uint32
    return lib


def _P(n):
    return 1 if n <= 10 else 2 if n <= 26 else 3 if n <= 63 else 4


def _prbs(g):
    a = [i for i in range(110) if g.prb_idx[0][i]]
    assert a == [i for i in range(110) if g.prb_idx[1][i]] and len(a) == g.nof_prb
    return a


@pytest.mark.parametrize("n", BWS)
def test_riv_round_trip_and_width(L, n):
    seen = set()
    for s in range(n):
        for l in range(1, n - s + 1):
            riv = L.srslte_ra_type2_to_riv(l, s, n)
            assert riv < (1 << L.srslte_ra_type2_n_rb(n)) and riv not in seen
            seen.add(riv)
            lo, so = C.c_uint32(), C.c_uint32()
            L.srslte_ra_type2_from_riv(riv, C.byref(lo), C.byref(so), n, n)
            assert (lo.value, so.value) == (l, s)
    assert len(seen) == n * (n + 1) // 2 and L.srslte_ra_type0_P(n) == _P(n)


def test_mcs_table(L):
    L.srslte_ra_tbs_idx_from_mcs.restype = C.c_int
    exp_mod = [1] * 10 + [2] * 7 + [3] * 12 + [1, 2, 3]
    exp_itbs = list(range(10)) + list(range(9, 16)) + list(range(15, 27)) + [-1] * 3
    for m in range(32):
        assert L.srslte_ra_mod_from_mcs(m) == exp_mod[m] and L.srslte_ra_tbs_idx_from_mcs(m) == exp_itbs[m]


@pytest.mark.parametrize("n", BWS)
def test_format1a_crnti_round_trip(L, n):
    t = install_tbs_table(L, {(26, 100): 75376, (26, 6): 4392})
    rng = np.random.default_rng(n)
    for _ in range(60):
        d = RaDlDci()
        d.alloc_type = 2
        d.type2_alloc.RB_start = int(rng.integers(0, n))
        d.type2_alloc.L_crb = int(rng.integers(1, n - d.type2_alloc.RB_start + 1))
        d.mcs_idx, d.harq_process, d.rv_idx = int(rng.integers(0, 29)), int(rng.integers(0, 8)), int(rng.integers(0, 4))
        d.ndi, d.tpc = bool(rng.integers(0, 2)), int(rng.integers(0, 4))
        msg = DciMsg()
        nb = L.srslte_dci_msg_pack_pdsch(C.byref(d), FMT1A, C.byref(msg), n, True)
        assert nb == L.srsue_gpu_host_dci_format_sizeof(0, n) == msg.nof_bits and msg.data[0] == 1
        u, g = RaDlDci(), Grant()
        assert L.srslte_dci_msg_to_dl_grant(C.byref(msg), 0x4601, n, C.byref(u), C.byref(g)) == 0
        assert (u.type2_alloc.RB_start, u.type2_alloc.L_crb, u.mcs_idx, u.harq_process, u.rv_idx, u.ndi, u.tpc) == \
               (d.type2_alloc.RB_start, d.type2_alloc.L_crb, d.mcs_idx, d.harq_process, d.rv_idx, d.ndi, d.tpc)
        assert _prbs(g) == list(range(d.type2_alloc.RB_start, d.type2_alloc.RB_start + d.type2_alloc.L_crb))
        itbs = d.mcs_idx if d.mcs_idx < 10 else d.mcs_idx - 1 if d.mcs_idx < 17 else d.mcs_idx - 2
        assert g.mcs.tbs == t[itbs, g.nof_prb - 1] and g.mcs.idx == d.mcs_idx
        assert g.Qm == (2 if d.mcs_idx < 10 else 4 if d.mcs_idx < 17 else 6)
    # a format 0 message (flag 0) is not a downlink assignment
    msg.data[0] = 0
    assert L.srslte_dci_msg_to_dl_grant(C.byref(msg), 0x4601, n, C.byref(u), C.byref(g)) != 0


def test_format1a_si_rnti_uses_nprb1a_column(L):
    t = install_tbs_table(L, {})
    for n in (25, 50):
        for tpc_lsb in (0, 1):
            d = RaDlDci()
            d.alloc_type, d.mcs_idx = 2, 7
            d.type2_alloc.RB_start, d.type2_alloc.L_crb, d.type2_alloc.n_prb1a = 2, 4, tpc_lsb
            msg, u, g = DciMsg(), RaDlDci(), Grant()
            assert L.srslte_dci_msg_pack_pdsch(C.byref(d), FMT1A, C.byref(msg), n, False) > 0
            assert L.srslte_dci_msg_to_dl_grant(C.byref(msg), 0xFFFF, n, C.byref(u), C.byref(g)) == 0
            assert g.Qm == 2 and g.mcs.tbs == t[7, 2 + tpc_lsb - 1] and _prbs(g) == [2, 3, 4, 5]
            assert u.type2_alloc.n_prb1a == tpc_lsb


@pytest.mark.parametrize("n", BWS)
def test_format1_type0_and_type1(L, n):
    t = install_tbs_table(L, {})
    P = _P(n)
    nbm = -(-n // P)
    rng = np.random.default_rng(100 + n)
    for k in range(80):
        d = RaDlDci()
        d.mcs_idx, d.harq_process, d.rv_idx, d.ndi = int(rng.integers(0, 29)), int(rng.integers(0, 8)), int(rng.integers(0, 4)), bool(k & 1)
        if n <= 10 or k % 2 == 0:
            d.alloc_type = 0
            d.type0_alloc.rbg_bitmask = int(rng.integers(1, 1 << nbm))
            exp = [i for i in range(n) if (d.type0_alloc.rbg_bitmask >> (nbm - 1 - i // P)) & 1]
        else:
            d.alloc_type = 1
            nsub = int(np.ceil(np.log2(P)))
            n1 = nbm - nsub - 1
            p = int(rng.integers(0, P))
            d.type1_alloc.rbg_subset, d.type1_alloc.shift = p, bool(rng.integers(0, 2))
            d.type1_alloc.vrb_bitmask = int(rng.integers(1, 1 << n1))
            subset = [i for i in range(n) if (i // P) % P == p]            # the PRBs of RBG subset p, ascending
            off = len(subset) - n1 if d.type1_alloc.shift else 0
            exp = [subset[i + off] for i in range(n1) if (d.type1_alloc.vrb_bitmask >> (n1 - 1 - i)) & 1]
        msg, u, g = DciMsg(), RaDlDci(), Grant()
        nb = L.srslte_dci_msg_pack_pdsch(C.byref(d), FMT1, C.byref(msg), n, True)
        assert nb == L.srsue_gpu_host_dci_format_sizeof(1, n)
        assert L.srslte_dci_msg_to_dl_grant(C.byref(msg), 0x0100, n, C.byref(u), C.byref(g)) == 0
        assert _prbs(g) == exp and (u.alloc_type, u.mcs_idx, u.harq_process, u.rv_idx, u.ndi) == \
               (d.alloc_type, d.mcs_idx, d.harq_process, d.rv_idx, d.ndi)
        itbs = d.mcs_idx if d.mcs_idx < 10 else d.mcs_idx - 1 if d.mcs_idx < 17 else d.mcs_idx - 2
        assert g.mcs.tbs == t[itbs, len(exp) - 1]


def test_retransmission_mcs_and_distributed(L):
    install_tbs_table(L, {})
    d = RaDlDci()
    d.alloc_type, d.mcs_idx = 2, 30
    d.type2_alloc.RB_start, d.type2_alloc.L_crb = 0, 10
    msg, u, g = DciMsg(), RaDlDci(), Grant()
    assert L.srslte_dci_msg_pack_pdsch(C.byref(d), FMT1A, C.byref(msg), 25, True) > 0
    assert L.srslte_dci_msg_to_dl_grant(C.byref(msg), 0x4601, 25, C.byref(u), C.byref(g)) == 0
    assert g.mcs.tbs == 0 and g.Qm == 4                                   # size of the first transmission: MAC's to fill
    msg.data[1] = 1                                                       # distributed virtual resource blocks
    assert L.srslte_dci_msg_to_dl_grant(C.byref(msg), 0x4601, 25, C.byref(u), C.byref(g)) != 0
    assert L.srsue_gpu_ra_set_tbs_table(None, 27, 110) != 0 and L.srslte_ra_tbs_from_idx(27, 1) < 0


def test_no_table_fails_loudly():
    """A fresh process has no size table: conversions that need one return an error and say why."""
    import subprocess, sys
    code = ("import ctypes as C, srsue_b200 as sg; L = sg.lib(); "
            "assert L.srsue_gpu_ra_have_tbs_table() == 0; assert L.srslte_ra_tbs_from_idx(3, 10) < 0")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "no transport-block-size table installed" in r.stderr


def test_cqi_helpers(L):
    """phch_worker.cc:504-527: CQI index from the estimator's SNR, report timing (36.213 Table 7.2.2-1A), UCI bits"""
    L.srslte_cqi_from_snr.restype = C.c_uint8
    L.srslte_cqi_send.restype = C.c_bool
    cq = [L.srslte_cqi_from_snr(C.c_float(s)) for s in np.arange(-10, 30, 0.25)]
    assert cq[0] == 0 and cq[-1] == 15 and all(b - a in (0, 1) for a, b in zip(cq, cq[1:])) and set(cq) == set(range(16))
    for idx, period, off in ((0, 2, 0), (1, 2, 1), (2, 5, 0), (6, 5, 4), (7, 10, 0), (16, 10, 9), (17, 20, 0), (36, 20, 19),
                             (37, 40, 0), (76, 40, 39), (77, 80, 0), (157, 160, 0), (316, 160, 159), (318, 32, 0), (349, 32, 31),
                             (350, 64, 0), (414, 128, 0), (541, 128, 127)):
        hits = [t for t in range(10240) if L.srslte_cqi_send(idx, t)]
        assert hits == list(range(off, 10240, period)), idx
    assert not any(L.srslte_cqi_send(i, t) for i in (317, 542, 1023) for t in range(200))

    class Cqi(C.Structure):
        _fields_ = [("type", C.c_int), ("wideband_cqi", C.c_uint8), ("subband_cqi", C.c_uint8), ("subband_label", C.c_uint8)]
    buf = (C.c_uint8 * 64)()
    v = Cqi(0, 11, 0, 0)
    assert L.srslte_cqi_value_pack(C.byref(v), buf) == 4 and list(buf[:4]) == [1, 0, 1, 1]
    v = Cqi(1, 0, 6, 1)
    assert L.srslte_cqi_value_pack(C.byref(v), buf) == 5 and list(buf[:5]) == [0, 1, 1, 0, 1]
