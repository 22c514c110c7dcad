"""ctypes mirrors of include/srsue_gpu/srslte_compat.h for the tests (the C++ callers use the header)."""
import ctypes as C


class Cell(C.Structure):
    _fields_ = [("nof_prb", C.c_uint32), ("nof_ports", C.c_uint32), ("bw_idx", C.c_uint32), ("id", C.c_uint32),
                ("cp", C.c_int), ("phich_length", C.c_int), ("phich_resources", C.c_int)]


class Mcs(C.Structure):
    _fields_ = [("mod", C.c_int), ("tbs", C.c_int), ("idx", C.c_uint32)]


class Grant(C.Structure):
    _fields_ = [("prb_idx", (C.c_bool * 110) * 2), ("nof_prb", C.c_uint32), ("Qm", C.c_uint32), ("mcs", Mcs)]


class CbSegm(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("F", "C", "K1", "K2", "C1", "C2", "tbs")]


class NBits(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("lstart", "nof_symb", "nof_bits", "nof_re")]


class PdschCfg(C.Structure):
    _fields_ = [("cb_segm", CbSegm), ("grant", Grant), ("nbits", NBits), ("rv", C.c_uint32), ("sf_idx", C.c_uint32)]


class SoftBuffer(C.Structure):
    _fields_ = [("max_cb", C.c_uint32), ("buffer_f", C.POINTER(C.POINTER(C.c_int16))), ("gpu_shadow", C.c_void_p)]


class Sch(C.Structure):
    _fields_ = [("max_iterations", C.c_uint32), ("nof_iterations", C.c_uint32)]


class Pdsch(C.Structure):
    _fields_ = [("cell", Cell), ("rnti", C.c_uint16), ("dl_sch", Sch), ("gpu", C.c_void_p)]


class Chest(C.Structure):
    _fields_ = [("cell", Cell), ("noise_estimate", C.c_float), ("rsrp", C.c_float), ("rssi", C.c_float),
                ("rsrq", C.c_float), ("gpu", C.c_void_p)]


class Pdcch(C.Structure):
    _fields_ = [("gpu", C.c_void_p)]


class DciMsg(C.Structure):
    _fields_ = [("data", C.c_uint8 * 128), ("nof_bits", C.c_uint32), ("format", C.c_int)]


class DciLocation(C.Structure):
    _fields_ = [("L", C.c_uint32), ("ncce", C.c_uint32)]


class UeDl(C.Structure):
    _fields_ = [("pdcch", Pdcch), ("pdsch", Pdsch), ("chest", Chest), ("pdsch_cfg", PdschCfg), ("softbuffer", SoftBuffer),
                ("cell", Cell), ("sf_symbols", C.c_void_p), ("ce", C.c_void_p * 4), ("current_rnti", C.c_uint16),
                ("last_location", DciLocation), ("last_n_cce", C.c_uint32), ("pkt_errors", C.c_uint64),
                ("pkts_total", C.c_uint64), ("nof_detected", C.c_uint64), ("gpu", C.c_void_p)]


class Tdec(C.Structure):
    _fields_ = [("max_long_cb", C.c_uint32), ("n_iter", C.c_uint32), ("gpu", C.c_void_p)]


def make_grant(nof_prb, qm, tbs, prbs=None):
    g = Grant()
    for p in (range(nof_prb) if prbs is None else prbs):
        g.prb_idx[0][p] = True
        g.prb_idx[1][p] = True
    g.nof_prb = nof_prb if prbs is None else len(list(prbs))
    g.Qm = qm
    g.mcs.mod = {2: 1, 4: 2, 6: 3}[qm]
    g.mcs.tbs = tbs
    g.mcs.idx = 0
    return g


class RaType0(C.Structure):
    _fields_ = [("rbg_bitmask", C.c_uint32)]


class RaType1(C.Structure):
    _fields_ = [("vrb_bitmask", C.c_uint32), ("rbg_subset", C.c_uint32), ("shift", C.c_bool)]


class RaType2(C.Structure):
    _fields_ = [("riv", C.c_uint32), ("L_crb", C.c_uint32), ("RB_start", C.c_uint32), ("n_prb1a", C.c_int),
                ("n_gap", C.c_int), ("mode", C.c_int)]


class RaDlDci(C.Structure):
    _fields_ = [("alloc_type", C.c_int), ("type0_alloc", RaType0), ("type1_alloc", RaType1), ("type2_alloc", RaType2),
                ("mcs_idx", C.c_uint32), ("harq_process", C.c_uint32), ("rv_idx", C.c_int), ("ndi", C.c_bool),
                ("dci_is_1a", C.c_bool), ("tpc", C.c_uint32)]


def install_tbs_table(L, entries):
    """Install a SYNTHETIC 27 x 110 transport-block-size table (a smooth byte-aligned stand-in for 36.213 Table
    7.1.7.2.1-1, which this tree does not carry) with the given {(I_TBS, N_PRB): size} cells overlaid."""
    import numpy as np
    t = np.zeros((27, 110), np.int32)
    for i in range(27):
        for n in range(1, 111):
            t[i, n - 1] = 8 * ((n * (16 + 28 * i)) // 8)
    for (i, n), v in entries.items():
        t[i, n - 1] = v
    assert L.srsue_gpu_ra_set_tbs_table(t.ctypes.data_as(C.c_void_p), 27, 110) == 0
    return t


class UeMib(C.Structure):
    _fields_ = [("pbch", C.c_void_p), ("cell", Cell), ("gpu", C.c_void_p)]
