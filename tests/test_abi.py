"""The C-ABI library loads on a CPU-only machine, exports every symbol include/*.h declares, its host-side
bookkeeping agrees with the oracle, and it fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include", "srsue_gpu")


def _declared(header):
    txt = open(os.path.join(INC, header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b((?:srsue_gpu|srslte)_[a-z0-9_]+)\s*\(", txt))
    return {n for n in names if not n.endswith("_t")}


@pytest.fixture(scope="module")
def lib():
    import srsue_b200 as sg
    return sg.lib()


@pytest.mark.parametrize("header", ["srsue_gpu.h", "srslte_compat.h"])
def test_every_declared_symbol_is_exported(lib, header):
    names = _declared(header)
    assert len(names) >= 20
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, "declared but not exported: %s" % missing


def test_reference_call_sites_are_covered(lib):
    # the srsLTE symbols of the DL path that srsUE calls (phch_worker.cc:74,88,104,127,254,337,
    # 347-348,359-360; dl_harq.cc:169,174,232; proc_ra.cc:60,254) plus the two the north star names
    for n in ["srslte_ue_dl_init", "srslte_ue_dl_free", "srslte_ue_dl_set_rnti", "srslte_ue_dl_decode_fft_estimate",
              "srslte_ue_dl_cfg_grant", "srslte_pdsch_decode_rnti", "srslte_sch_set_max_noi", "srslte_pdsch_last_noi",
              "srslte_chest_dl_get_snr", "srslte_chest_dl_get_rsrp", "srslte_chest_dl_get_rssi", "srslte_chest_dl_get_rsrq",
              "srslte_chest_dl_get_noise_estimate", "srslte_softbuffer_rx_init", "srslte_softbuffer_rx_reset",
              "srslte_softbuffer_rx_reset_tbs", "srslte_softbuffer_rx_free", "srslte_ue_dl_decode", "srslte_ue_dl_decode_rnti",
              "srslte_tdec_init", "srslte_tdec_free", "srslte_tdec_reset", "srslte_tdec_iteration", "srslte_tdec_decision",
              "srslte_tdec_decision_byte", "srslte_tdec_run_all", "srslte_vec_malloc", "srslte_symbol_sz",
              "srslte_pdcch_extract_llr", "srslte_ue_dl_find_dl_dci_type", "srslte_ue_dl_find_ul_dci", "srslte_ue_dl_get_ncce", "srslte_ue_dl_decode_phich",
              "srslte_ue_mib_init", "srslte_ue_mib_free", "srslte_ue_mib_decode", "srslte_pbch_decode_reset", "srslte_pbch_mib_unpack",
              "srslte_pbch_mib_pack", "srslte_ue_cellsearch_init", "srslte_ue_cellsearch_free", "srslte_ue_cellsearch_scan",
              "srslte_ue_cellsearch_scan_N_id_2", "srslte_ue_cellsearch_set_nof_frames_to_scan", "srslte_ue_cellsearch_set_threshold",
              "srslte_ue_sync_start_agc", "srslte_agc_get_gain", "srslte_ue_mib_sync_init", "srslte_ue_mib_sync_decode",
              "srslte_ue_mib_sync_free", "srslte_bit_pack_vector", "srslte_bit_unpack_vector", "srslte_bit_pack", "srslte_bit_unpack",
              "srslte_cp_string", "srslte_cell_fprint", "srslte_sampling_freq_hz", "srslte_timestamp_copy", "srslte_timestamp_add",
              "srslte_tti_interval", "srslte_ue_sync_init", "srslte_ue_sync_free", "srslte_ue_sync_zerocopy", "srslte_ue_sync_get_buffer",
              "srslte_ue_sync_get_sfidx", "srslte_ue_sync_get_cfo", "srslte_ue_sync_get_sfo", "srslte_ue_sync_set_cfo",
              "srslte_ue_sync_set_agc_period", "srslte_ue_sync_decode_sss_on_track", "srslte_ue_sync_get_last_timestamp",
              "srslte_sync_set_threshold", "srslte_sync_set_em_alpha",
              "srslte_dci_msg_to_dl_grant", "srslte_ra_dl_dci_string", "srslte_cqi_from_snr", "srslte_cqi_send",
              "srslte_cqi_value_pack", "srsue_gpu_ue_dl_set_cfo", "srsue_gpu_ra_set_tbs_table"]:
        assert hasattr(lib, n), n


def test_headers_are_plain_c_and_cxx(tmp_path):
    """the boundary is a C ABI: both headers must parse as C99 (srsLTE is C) and as C++11 (srsUE is C++)"""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "srsue_gpu/srslte_compat.h"\n'
                   'int main(void) { srslte_ra_dl_dci_t d; srslte_cqi_value_t v; srsue_gpu_sf_desc_t s; (void)d; (void)v; (void)s; return 0; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I" + inc, str(src)],
                ["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-I" + inc, "-x", "c++", str(src)]):
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr


def test_host_tables_agree_with_oracle(lib, oracle):
    import srsue_b200 as sg
    out = (C.c_int * 8)()
    for tbs in (152, 1008, 4968, 6208, 30576, 75376):
        assert lib.srsue_gpu_host_cbsegm(tbs, out) == 0
        s = oracle.cbsegm(tbs)
        assert list(out) == [s.tbs, s.B, s.C, s.Kp, s.Km, s.Cp, s.Cm, s.F]
    for K in oracle.qpp_Ks():
        W, P, n = sg.Context.tdec_geometry(K)
        assert W == oracle.window_len(K) and P == K // W
        pi = np.zeros(K, np.uint16)
        assert lib.srsue_gpu_host_qpp(K, pi.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(pi, oracle.qpp_perm(K))
    for K, F, rv in ((176, 0, 0), (1056, 24, 1), (3136, 56, 2), (5824, 0, 0), (6144, 0, 3)):
        seq = np.zeros(3 * (K + 4), np.int32)
        n = lib.srsue_gpu_host_rm_sequence(K, F, rv, seq.ctypes.data_as(C.c_void_p))
        ref = oracle.rm_sequence(K, F, rv)
        assert n == len(ref) and np.array_equal(seq[:n], ref)
    for c_init in (1, 0x48D0201, 0x7FFFFFFF):
        c = np.zeros(777, np.uint8)
        assert lib.srsue_gpu_host_gold(C.c_uint32(c_init), 777, c.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(c, oracle.gold(c_init, 777))
    for prb, ports, sf, cfi in ((6, 1, 1, 1), (100, 1, 1, 1), (100, 2, 1, 1), (100, 1, 0, 2), (25, 2, 5, 3), (50, 1, 9, 2)):
        cell, ocell = sg.make_cell(prb, ports, 7), oracle.make_cell(prb, ports, 7)
        cfg, ocfg = sg.make_cfg(cell, sf_idx=sf, cfi=cfi), oracle.make_cfg(ocell, sf_idx=sf, cfi=cfi)
        re = np.zeros(14 * 12 * prb, np.int32)
        n = lib.srsue_gpu_host_pdsch_re(C.byref(cell), C.byref(cfg), re.ctypes.data_as(C.c_void_p))
        ref = oracle.pdsch_re_list(ocell, ocfg)
        assert n == len(ref) and np.array_equal(re[:n], ref)


def test_invalid_arguments_are_rejected(lib):
    assert lib.srsue_gpu_tdec_geometry(41, None, None, None) == -2
    out = (C.c_int * 8)()
    assert lib.srsue_gpu_host_cbsegm(0, out) == -2
    assert lib.srsue_gpu_ctx_create(None, 0) == -2


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.srsue_gpu_ctx_create(C.byref(h), 0)
    assert rc == -1 and not h.value
    assert b"no CPU fallback" in lib.srsue_gpu_last_error()
    # the srsLTE-shaped entry points fail the same way
    from tests.srslte_ctypes import UeDl, Cell, Tdec
    q = UeDl()
    assert lib.srslte_ue_dl_init(C.byref(q), Cell(nof_prb=6, nof_ports=1, bw_idx=0, id=1, cp=0, phich_length=0, phich_resources=0)) == -1
    t = Tdec()
    assert lib.srslte_tdec_init(C.byref(t), 6144) == -1


def test_host_pcfich_mapping_matches_oracle():
    """the PCFICH resource-element map (36.211 6.7.4) of the library equals the oracle's for every bandwidth and a
    spread of cell ids, and avoids the CRS positions of ports 0 and 1"""
    import ctypes as C
    import numpy as np
    import srsue_b200 as sg
    from oracle import oracle as o
    L = sg.lib()
    for prb in (6, 15, 25, 50, 75, 100):
        for cid in (0, 1, 2, 5, 77, 150, 301, 503):
            cell = sg.make_cell(prb, 2, cid)
            k = np.zeros(16, np.int32)
            assert L.srsue_gpu_host_pcfich_re(C.byref(cell), k.ctypes.data_as(C.c_void_p)) == 0
            ref = o.pcfich_re(o.make_cell(prb, 2, cid))
            assert np.array_equal(k, ref)
            assert len(set(k.tolist())) == 16 and all(x % 3 != cid % 3 for x in k)


def test_host_pdcch_tables_match_oracle():
    """control-region REG map (PCFICH / PHICH groups excluded), quadruplet interleaver + cyclic shift, search spaces and
    DCI sizes of the library against the oracle / the published format sizes"""
    import ctypes as C
    import numpy as np
    import srsue_b200 as sg
    from oracle import oracle as o
    L = sg.lib()
    for prb in (6, 15, 25, 50, 75, 100):
        for cid in (0, 1, 77, 301, 503):
            ocell = o.make_cell(prb, 2, cid)
            cell = sg.make_cell(prb, 2, cid)
            for cfi in (1, 2, 3):
                for ng in (1, 3, 6, 12):
                    rk, rl = o.pdcch_regs(ocell, cfi, ng)
                    re4 = np.zeros(4 * 12 * prb, np.int32)
                    n = L.srsue_gpu_host_pdcch_regs(C.byref(cell), cfi, ng, re4.ctypes.data_as(C.c_void_p))
                    assert n == len(rk)
                    exp = []
                    for k0, l in zip(rk.tolist(), rl.tolist()):
                        ks = [k0 + j for j in range(6) if (k0 + j) % 3 != cid % 3] if l == 0 else [k0 + j for j in range(4)]
                        exp += [l * 12 * prb + k for k in ks]
                    assert re4[:4 * n].tolist() == exp
                src = np.zeros(n, np.int32)
                assert L.srsue_gpu_host_pdcch_quad_perm(n, cid, src.ctypes.data_as(C.c_void_p)) == 0
                assert np.array_equal(src, o.pdcch_quad_perm(n, cid))
                for common in (0, 1):
                    for sf in (0, 4, 9):
                        cl, cn = np.zeros(32, np.int32), np.zeros(32, np.int32)
                        m = L.srsue_gpu_host_pdcch_search_space(n // 9, sf, 0x1234 + cid, common, cl.ctypes.data_as(C.c_void_p),
                                                                cn.ctypes.data_as(C.c_void_p))
                        assert list(zip(cl[:m].tolist(), cn[:m].tolist())) == o.pdcch_search_space(n // 9, sf, 0x1234 + cid, bool(common))
    # 36.212 5.3.3.1 sizes (FDD): format 1A / format 1 at 1.4, 3, 5, 10, 15, 20 MHz
    assert [L.srsue_gpu_host_dci_format_sizeof(0, p) for p in (6, 15, 25, 50, 75, 100)] == [21, 22, 25, 27, 27, 28]
    assert [L.srsue_gpu_host_dci_format_sizeof(1, p) for p in (6, 15, 25, 50, 75, 100)] == [19, 23, 27, 31, 33, 39]


def test_host_phich_tables_match_oracle():
    import ctypes as C
    import numpy as np
    import srsue_b200 as sg
    from oracle import oracle as o
    L = sg.lib()
    for prb in (6, 15, 25, 50, 75, 100):
        for cid in (0, 1, 77, 301, 503):
            cell = sg.make_cell(prb, 1, cid)
            ocell = o.make_cell(prb, 1, cid)
            for ng in (1, 3, 6, 12):
                groups = (ng * prb + 47) // 48
                for g in range(groups):
                    k = np.zeros(12, np.int32)
                    assert L.srsue_gpu_host_phich_res(C.byref(cell), g, k.ctypes.data_as(C.c_void_p)) == 0
                    assert np.array_equal(k, o.phich_res(ocell, g, ng))
                for I_lowest, n_dmrs in ((0, 0), (3, 1), (prb - 1, 7), (prb // 2, 4)):
                    a, b = C.c_int(), C.c_int()
                    assert L.srsue_gpu_host_phich_index(prb, ng, I_lowest, n_dmrs, C.byref(a), C.byref(b)) == 0
                    assert (a.value, b.value) == o.phich_index(prb, I_lowest, n_dmrs, ng)


def test_pdcch_quadruplet_interleaver_equals_block_description():
    """36.211 6.8.5 / 36.212 5.1.4.2.1 as a matrix: <NULL>s in front up to a multiple of 32, written row by row, the columns
    permuted with the pattern of Table 5.1.4-2, read column by column without the NULLs, then a cyclic shift by the cell
    id -- an independent construction of what srsue_gpu_host_pdcch_quad_perm (and the oracle) compute in closed form."""
    import ctypes as C
    import numpy as np
    import srsue_b200 as sg
    L = sg.lib()
    P = [1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31, 0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30]
    for m in (18, 41, 59, 87, 144, 207, 360, 789):
        for cid in (0, 1, 77, 503):
            rows = -(-m // 32)
            y = [None] * (rows * 32 - m) + list(range(m))
            mat = [y[r * 32:(r + 1) * 32] for r in range(rows)]
            perm = [mat[r][P[c]] for c in range(32) for r in range(rows)]
            w = [v for v in perm if v is not None]
            want = [w[(i + cid) % m] for i in range(m)]
            src = np.zeros(m, np.int32)
            assert L.srsue_gpu_host_pdcch_quad_perm(m, cid, src.ctypes.data_as(C.c_void_p)) == 0
            assert src.tolist() == want, (m, cid)


def test_control_mappings_equal_standard_formulas():
    """36.211 literally: PCFICH quadruplet i sits in the REG that starts at k = k_bar + floor(i N_RB / 2) * 6 with
    k_bar = 6 (N_ID mod 2 N_RB) (6.7.4); PHICH group m', quadruplet i sits in the free REG number
    (N_ID + m' + floor(i n_0 / 3)) mod n_0 of symbol 0, n_0 = REGs not taken by the PCFICH (6.9.3, normal duration);
    the PDCCH search space is L {(Y_k + m) mod floor(N_CCE / L)} + i with Y_k = 39827 Y_{k-1} mod 65537, Y_{-1} = RNTI (36.213 9.1.1)."""
    import ctypes as C
    import numpy as np
    import srsue_b200 as sg
    L = sg.lib()
    for prb in (6, 15, 25, 50, 75, 100):
        nsc = 12 * prb
        for cid in (0, 3, 77, 150, 503):
            cell = sg.make_cell(prb, 2, cid)
            k = np.zeros(16, np.int32)
            assert L.srsue_gpu_host_pcfich_re(C.byref(cell), k.ctypes.data_as(C.c_void_p)) == 0
            kbar = 6 * (cid % (2 * prb))
            want, pc_regs = [], []
            for i in range(4):
                k0 = (kbar + (i * prb // 2) * 6) % nsc
                pc_regs.append(k0 // 6)
                want += [k0 + j for j in range(6) if (k0 + j) % 3 != cid % 3]
            assert k.tolist() == want
            free = [r for r in range(2 * prb) if r not in pc_regs]
            for g in range(-(-(12 * prb) // 48)):                         # ceil(Ng (N_RB / 8)) groups at the largest Ng = 2
                k12 = np.zeros(12, np.int32)
                assert L.srsue_gpu_host_phich_res(C.byref(cell), g, k12.ctypes.data_as(C.c_void_p)) == 0
                exp = []
                for i in range(3):
                    k0 = 6 * free[(cid + g + (i * len(free)) // 3) % len(free)]
                    exp += [k0 + j for j in range(6) if (k0 + j) % 3 != cid % 3]
                assert k12.tolist() == exp, (prb, cid, g)
    for ncce in (6, 12, 21, 41, 87):
        for rnti in (0x0001, 0x4601, 0xFFF3):
            for sf in range(10):
                y = rnti
                for _ in range(sf + 1):
                    y = (39827 * y) % 65537
                for common in (0, 1):
                    exp = []
                    for Lc, M in (((4, 4), (8, 2)) if common else ((1, 6), (2, 6), (4, 2), (8, 2))):
                        if ncce // Lc == 0:
                            continue
                        for m in range(M):
                            exp.append((Lc, Lc * (((0 if common else y) + m) % (ncce // Lc))))
                    cl, cn = np.zeros(32, np.int32), np.zeros(32, np.int32)
                    n = L.srsue_gpu_host_pdcch_search_space(ncce, sf, rnti, common, cl.ctypes.data_as(C.c_void_p), cn.ctypes.data_as(C.c_void_p))
                    got = list(zip(cl[:n].tolist(), cn[:n].tolist()))
                    assert sorted(set(got)) == sorted(set(exp)), (ncce, rnti, sf, common)
