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


def test_retransmission_mcs(L):
    install_tbs_table(L, {})
    d = RaDlDci()
    d.alloc_type, d.mcs_idx = 2, 30
    d.type2_alloc.RB_start, d.type2_alloc.L_crb = 0, 10
    msg, u, g = DciMsg(), RaDlDci(), Grant()
    assert L.srslte_dci_msg_pack_pdsch(C.byref(d), FMT1A, C.byref(msg), 25, True) > 0
    assert L.srslte_dci_msg_to_dl_grant(C.byref(msg), 0x4601, 25, C.byref(u), C.byref(g)) == 0
    assert g.mcs.tbs == 0 and g.Qm == 4                                   # size of the first transmission: MAC's to fill
    assert L.srsue_gpu_ra_set_tbs_table(None, 27, 110) != 0 and L.srslte_ra_tbs_from_idx(27, 1) < 0


def test_no_table_fails_loudly():
    """A fresh process has only the built-in columns: a size outside them returns an error and says why."""
    import subprocess, sys
    code = ("import ctypes as C, srsue_b200 as sg; L = sg.lib(); "
            "assert L.srsue_gpu_ra_have_tbs_table() == 0; assert L.srslte_ra_tbs_from_idx(3, 11) < 0; "
            "assert L.srslte_ra_tbs_from_idx(3, 10) == 568")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "no transport-block-size table" in r.stderr


def test_builtin_tbs_columns_have_the_structure_of_the_standard_table():
    """36.213 Table 7.1.7.2.1-1 cannot be read offline; the thirteen columns the library carries (tbs_table.inc, written
    from memory) are checked against every structural property of the real table: multiples of 8, strictly increasing in
    I_TBS, non-decreasing in N_PRB, and -- the selective one -- TBS + 24 segments into code blocks with zero filler bits
    (36.212 5.1.2).  Anchors: the sizes BASELINE.md and the reference's configs name."""
    import subprocess, sys
    code = r"""
import ctypes as C, json, srsue_b200 as sg
L = sg.lib()
cols = (C.c_int32 * 32)()
n = L.srsue_gpu_ra_builtin_tbs_columns(cols, 32)
seg = (C.c_int * 8)()
out = {}
for c in list(cols)[:n]:
    col = []
    for i in range(27):
        t = L.srslte_ra_tbs_from_idx(i, c)
        assert L.srsue_gpu_host_cbsegm(t, seg) == 0
        col.append((t, seg[7]))
    out[c] = col
print(json.dumps(out))
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)      # fresh process: nothing installed
    assert r.returncode == 0, r.stderr
    import json
    tab = {int(k): v for k, v in json.loads(r.stdout).items()}
    assert sorted(tab) == [1, 2, 3, 4, 5, 6, 10, 15, 25, 50, 75, 100, 110]
    for n_prb, col in tab.items():
        sizes = [t for t, _ in col]
        assert all(t % 8 == 0 and t >= 16 for t in sizes), n_prb
        assert all(b > a for a, b in zip(sizes, sizes[1:])), n_prb
        assert all(F == 0 for _, F in col), (n_prb, [t for t, F in col if F])          # no filler bits, ever
    ns = sorted(tab)
    for i in range(27):
        row = [tab[n][i][0] for n in ns]
        assert all(b >= a for a, b in zip(row, row[1:])), i
    assert tab[100][26][0] == 75376 and tab[75][26][0] == 55056 and tab[100][15][0] == 30576 and tab[6][0][0] == 152
    assert tab[25][20][0] == 11448 and tab[110][26][0] == 75376 and tab[1][0][0] == 16 and tab[1][26][0] == 712


def test_cqi_helpers(L):
    """phch_worker.cc:504-527: CQI index from the estimator's SNR, report timing (36.213 Table 7.2.2-1A), UCI bits"""
    L.srslte_cqi_from_snr.restype = C.c_uint8
    L.srslte_cqi_send.restype = C.c_bool
    cq = [L.srslte_cqi_from_snr(C.c_float(s)) for s in np.arange(-10, 30, 0.25)]
    assert cq[0] == 0 and cq[-1] == 15 and all(b - a in (0, 1) for a, b in zip(cq, cq[1:])) and set(cq) == set(range(16))
    # the mapping itself (srsLTE's thresholds as recalled in ra.cc): number of table entries below the SNR
    for snr, want in ((1.9, 0), (2.0, 1), (4.0, 1), (4.1, 2), (10.0, 4), (10.01, 5), (19.9, 9), (20.0, 10), (29.0, 14), (29.1, 15)):
        assert L.srslte_cqi_from_snr(C.c_float(snr)) == want, snr
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


def _dvrb(L, n, gap2):
    L.srsue_gpu_host_dvrb_to_prb.restype = C.c_uint32
    L.srsue_gpu_host_n_vrb_dl.restype = C.c_uint32
    nv = L.srsue_gpu_host_n_vrb_dl(n, gap2)
    return nv, [[L.srsue_gpu_host_dvrb_to_prb(n, gap2, v, s) for v in range(nv)] for s in (0, 1)]


def test_distributed_vrb_mapping_properties(L):
    """36.211 6.2.3.2: every distributed VRB lands on a distinct PRB inside the carrier in each slot, the two slots of a
    VRB are one gap apart (gap 1) / half an interleaving unit apart (gap 2), consecutive VRBs are spread out, and the
    1.4 MHz case equals the mapping worked out by hand from the interleaver (2 rows x 4 columns, 2 nulls)."""
    gaps1 = {6: 3, 7: 4, 10: 5, 11: 4, 12: 8, 19: 8, 20: 12, 25: 12, 26: 12, 27: 18, 44: 18, 45: 27, 49: 27, 50: 27, 63: 27,
             64: 32, 75: 32, 79: 32, 80: 48, 100: 48, 110: 48}
    for n in range(6, 111):
        for gap2 in ((0, 1) if n >= 50 else (0,)):
            nv, (s0, s1) = _dvrb(L, n, gap2)
            assert nv > 0 and nv % 2 == 0 and nv <= n
            assert len(set(s0)) == nv and len(set(s1)) == nv and max(s0 + s1) < n
            if not gap2:
                g = gaps1.get(n)
                assert nv == 2 * min(g, n - g) if g else True
                assert all(abs(a - b) == (g if g else abs(s0[0] - s1[0])) for a, b in zip(s0, s1))
            else:
                g = 9 if n <= 63 else 16
                assert nv == (n // (2 * g)) * 2 * g and all(abs(a - b) == g for a, b in zip(s0, s1))
            assert L.srsue_gpu_host_dvrb_to_prb(n, gap2, nv, 0) == 0xFFFFFFFF
    assert _dvrb(L, 6, 0)[1][0] == [0, 2, 3, 5, 1, 4] and _dvrb(L, 6, 0)[1][1] == [3, 5, 0, 2, 4, 1]
    assert _dvrb(L, 25, 0)[1][0][:8] == [0, 6, 12, 18, 1, 7, 13, 19]
    assert L.srsue_gpu_host_n_vrb_dl(25, 1) == 0                              # no second gap below 50 PRB


@pytest.mark.parametrize("n", BWS)
def test_format1a_distributed_round_trip(L, n):
    install_tbs_table(L, {})
    rng = np.random.default_rng(300 + n)
    for k in range(60):
        gap2 = int(n >= 50 and k % 2)
        nv, (s0, s1) = _dvrb(L, n, gap2)
        d = RaDlDci()
        d.alloc_type, d.mcs_idx, d.harq_process, d.rv_idx, d.ndi = 2, int(rng.integers(0, 29)), int(rng.integers(0, 8)), 1, True
        d.type2_alloc.mode, d.type2_alloc.n_gap = 1, gap2
        d.type2_alloc.RB_start = int(rng.integers(0, nv))
        lmax = nv - d.type2_alloc.RB_start
        d.type2_alloc.L_crb = int(rng.integers(1, lmax + 1))
        msg, u, g = DciMsg(), RaDlDci(), Grant()
        nb = L.srslte_dci_msg_pack_pdsch(C.byref(d), FMT1A, C.byref(msg), n, True)
        if nb < 0:
            # with the gap flag taking the field's MSB (>= 50 PRB) not every (start, length) is expressible
            assert n >= 50
            continue
        assert L.srslte_dci_msg_to_dl_grant(C.byref(msg), 0x4601, n, C.byref(u), C.byref(g)) == 0
        assert (u.type2_alloc.mode, u.type2_alloc.n_gap, u.type2_alloc.RB_start, u.type2_alloc.L_crb) == \
               (1, gap2, d.type2_alloc.RB_start, d.type2_alloc.L_crb)
        vr = range(d.type2_alloc.RB_start, d.type2_alloc.RB_start + d.type2_alloc.L_crb)
        assert [i for i in range(110) if g.prb_idx[0][i]] == sorted(s0[v] for v in vr)
        assert [i for i in range(110) if g.prb_idx[1][i]] == sorted(s1[v] for v in vr)
        assert g.nof_prb == d.type2_alloc.L_crb


def test_per_slot_allocation_tables_match_oracle(L):
    """the PDSCH resource-element list for a grant whose slots use different PRBs: library table == oracle's"""
    import srsue_b200 as sg
    from oracle import oracle as o
    for prb, ports, sf, cfi in ((25, 1, 3, 2), (50, 2, 0, 1), (100, 1, 5, 3), (6, 1, 1, 3)):
        nv, (s0, s1) = _dvrb(L, prb, 0)
        vr = range(1, min(nv, 9))
        a, b = [s0[v] for v in vr], [s1[v] for v in vr]
        cell, ocell = sg.make_cell(prb, ports, 9), o.make_cell(prb, ports, 9)
        cfg = sg.make_cfg(cell, sf_idx=sf, cfi=cfi, prbs=a, prbs_slot1=b)
        ocfg = o.make_cfg(ocell, sf_idx=sf, cfi=cfi, prbs=a, prbs_slot1=b)
        assert bytes(cfg.prb_mask) == bytes(ocfg.prb_mask) and set(bytes(cfg.prb_mask)) <= {0, 1, 2, 4}
        re = np.zeros(14 * 12 * prb, np.int32)
        n = L.srsue_gpu_host_pdsch_re(C.byref(cell), C.byref(cfg), re.ctypes.data_as(C.c_void_p))
        ref = o.pdsch_re_list(ocell, ocfg)
        assert n == len(ref) and np.array_equal(re[:n], ref)
        nsc = 12 * prb
        for idx in re[:n]:
            l, k = divmod(int(idx), nsc)
            assert (k // 12) in (a if l < 7 else b)


@pytest.mark.parametrize("n", BWS)
def test_random_dci_bits_never_escape_the_carrier(L, n):
    """blind decoding does produce false positives: any bit pattern must either be refused or give a grant whose PRBs lie
    inside the carrier, with the same number of PRBs in both slots and a size from the table"""
    t = install_tbs_table(L, {})
    rng = np.random.default_rng(4000 + n)
    accepted = 0
    for fmt, size_fmt in ((FMT1A, 0), (FMT1, 1)):
        nb = L.srsue_gpu_host_dci_format_sizeof(size_fmt, n)
        for k in range(1500):
            msg, u, g = DciMsg(), RaDlDci(), Grant()
            bits = rng.integers(0, 2, nb, dtype=np.uint8)
            C.memmove(msg.data, bits.ctypes.data, nb)
            msg.nof_bits, msg.format = nb, fmt
            rnti = 0x4601 if k % 3 else 0xFFFF
            if L.srslte_dci_msg_to_dl_grant(C.byref(msg), rnti, n, C.byref(u), C.byref(g)) != 0:
                continue
            accepted += 1
            s0 = [i for i in range(110) if g.prb_idx[0][i]]
            s1 = [i for i in range(110) if g.prb_idx[1][i]]
            assert 1 <= len(s0) == len(s1) == g.nof_prb and max(s0 + s1) < n
            assert g.Qm in (2, 4, 6) and (g.mcs.tbs == 0 or g.mcs.tbs in t)
    assert accepted > 500


def test_distributed_vrb_closed_form_equals_block_interleaver(L):
    """36.211 6.2.3.2 describes the mapping twice: as closed-form equations (what ra.cc implements) and as a block
    interleaver -- VRB numbers written row by row into N_row x 4, N_null nulls in the last N_null/2 rows of the 2nd and
    4th column, read column by column ignoring nulls; the j-th number read lands on position j of its interleaving unit,
    the second slot is shifted by half a unit, and the upper half of a unit is lifted by N_gap - unit/2."""
    def P_of(n):
        return 1 if n <= 10 else 2 if n <= 26 else 3 if n <= 63 else 4
    for n in range(6, 111):
        for gap2 in ((0, 1) if n >= 50 else (0,)):
            nv, (s0, s1) = _dvrb(L, n, gap2)
            gap = abs(s0[0] - s1[0])                                  # checked against Table 6.2.3.2-1 in the test above
            unit = 2 * gap if gap2 else nv
            P = P_of(n)
            n_row = -(-unit // (4 * P)) * P
            n_null = 4 * n_row - unit
            cells, v = {}, 0
            for r in range(n_row):
                for c in range(4):
                    if c in (1, 3) and r >= n_row - n_null // 2:
                        continue                                      # a null
                    cells[(r, c)] = v
                    v += 1
            assert v == unit
            order = [cells[(r, c)] for c in range(4) for r in range(n_row) if (r, c) in cells]
            pos = {vrb: j for j, vrb in enumerate(order)}
            for vrb in range(nv):
                u, w = divmod(vrb, unit)
                for slot, got in ((0, s0[vrb]), (1, s1[vrb])):
                    t = pos[w] if slot == 0 else (pos[w] + unit // 2) % unit
                    t += unit * u
                    want = t if t < unit // 2 else t + gap - unit // 2
                    assert got == want, (n, gap2, vrb, slot)
