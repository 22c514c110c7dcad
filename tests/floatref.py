"""ctypes loader of tests/float_ref/float_ref.c (the independent float receiver back end of the fixed-point anchor) and the
numpy glue around it: float rate de-matching and descrambling from the (independently checked) read order and Gold
sequence, code-block segmentation and CRC checks from 36.212 5.1.1-5.1.2.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "float_ref", "float_ref.c")
        out_dir = os.path.join(HERE, "_build")
        os.makedirs(out_dir, exist_ok=True)
        so = os.path.join(out_dir, "libfloatref.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src, "-lm"])
        _lib = C.CDLL(so)
        _lib.fr_turbo_decode.restype = C.c_int
    return _lib


def turbo_decode(llr, K, f1, f2, iters=4):
    """llr: float array [3K+12] in srsLTE decoder-input order -> K hard bits (full-length float max-log-MAP)"""
    llr = np.ascontiguousarray(llr, np.float64)
    bits = np.zeros(K, np.uint8)
    rc = lib().fr_turbo_decode(llr.ctypes.data_as(C.c_void_p), K, f1, f2, iters, bits.ctypes.data_as(C.c_void_p), None)
    assert rc == 0
    return bits


def demap(d, qm, gain=1.0):
    """exact max-log LLRs (positive = bit 1) of complex symbols d for the 36.211 7.1 mappers"""
    d = np.ascontiguousarray(d, np.complex128)
    llr = np.zeros(len(d) * qm, np.float64)
    lib().fr_demap(d.ctypes.data_as(C.c_void_p), len(d), qm, C.c_double(gain), llr.ctypes.data_as(C.c_void_p))
    return llr


def rate_dematch(e, seq, K, F, w=None):
    """float soft buffer in decoder-input order: received LLRs e accumulate at seq[i mod N]; filler bits are known zeros"""
    if w is None:
        w = np.zeros(3 * K + 12, np.float64)
    np.add.at(w, seq[np.arange(len(e)) % len(seq)], e)
    for k in range(F):
        w[3 * k] = w[3 * k + 1] = -1e4
    return w


def segmentation(tbs, Ks):
    """36.212 5.1.2: (C, K+, K-, C+, C-, F) for a transport block of tbs bits (CRC24A added here)"""
    B, Z = tbs + 24, 6144
    if B <= Z:
        Cn, Bp = 1, B
    else:
        Cn = -(-B // (Z - 24))
        Bp = B + 24 * Cn
    Kp = min(k for k in Ks if Cn * k >= Bp)
    if Cn == 1:
        return 1, Kp, 0, 1, 0, Kp - Bp
    Km = max(k for k in Ks if k < Kp)
    Cm = (Cn * Kp - Bp) // (Kp - Km)
    return Cn, Kp, Km, Cn - Cm, Cm, (Cn - Cm) * Kp + Cm * Km - Bp
