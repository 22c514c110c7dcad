"""What the receiver's frozen fixed-point choices cost (oracle/SPEC.md 5-7: int16 demapper scales, the +-511 soft-buffer
clamp, parallel windows with next-iteration initialisation), measured against an INDEPENDENT receiver back end:
tests/float_ref/float_ref.c, double precision, full-length max-log-MAP written from 36.212 5.1.3.2 and an exact max-log
demapper from 36.211 7.1.  This is the parity evidence available offline beyond the transmit-side pins: the reference's own
decoder (srsLTE) runs full-length recursions, so the gap below bounds how far the restated receiver can be from it.

The BLER curves are in profiles/bler_r02.json (tests/bler_sweep.py, about 7 minutes); the tests here check the float
reference itself, a small live comparison, and the committed table."""
import json
import os

import numpy as np
import pytest

import floatref as fr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_float_reference_decodes_noiseless_blocks_and_hard_demaps(oracle):
    o = oracle
    for K in (40, 104, 512, 1056, 6144):
        f1, f2 = o.qpp_params(K)
        c = np.random.default_rng(K).integers(0, 2, K, dtype=np.uint8)
        d = o.turbo_encode(c)
        assert np.array_equal(fr.turbo_decode(2.0 * d - 1.0, K, f1, f2, 1), c), K
    # exact max-log demapper: the sign of every LLR is the transmitted bit (36.211 7.1 tables written out here)
    rng = np.random.default_rng(5)
    for qm, norm, lev in ((2, np.sqrt(2), lambda b: 1 - 2 * b[0]), (4, np.sqrt(10), lambda b: (1 - 2 * b[0]) * (2 - (1 - 2 * b[1]))),
                          (6, np.sqrt(42), lambda b: (1 - 2 * b[0]) * (4 - (1 - 2 * b[1]) * (2 - (1 - 2 * b[2]))))):
        bits = rng.integers(0, 2, (200, qm))
        sym = np.array([lev(b[0::2]) + 1j * lev(b[1::2]) for b in bits]) / norm
        llr = fr.demap(sym + 0.01 * (rng.standard_normal(200) + 1j * rng.standard_normal(200)), qm)
        assert np.array_equal((llr > 0).astype(int).reshape(200, qm), bits), qm


def test_windowed_int16_decoder_tracks_the_float_reference_live(oracle):
    """short blocks, a few hundred of them at one waterfall point: both decoders see the same noise"""
    o = oracle
    K, n, snr = 40, 300, -3.5
    f1, f2 = o.qpp_params(K)
    sigma = np.sqrt(1.0 / (2.0 * 10.0 ** (snr / 10.0)))
    ei = ef = 0
    for i in range(n):
        c = np.random.default_rng(77_000 + i).integers(0, 2, K, dtype=np.uint8)
        r = 2.0 * o.turbo_encode(c) - 1.0 + sigma * np.random.default_rng(78_000 + i).standard_normal(3 * K + 12)
        w16 = np.clip(np.trunc(64.0 * r), -2048, 2047).astype(np.int16)
        ei += int(not np.array_equal(o.tdec(w16, K, 4, 0)[0], c))
        ef += int(not np.array_equal(fr.turbo_decode(r, K, f1, f2, 4), c))
    assert 0.05 < ef / n < 0.3 and abs(ei - ef) <= 0.04 * n, (ei, ef)


def test_committed_bler_table_bounds_the_cost_of_the_fixed_point_choices():
    tab = json.load(open(os.path.join(ROOT, "profiles", "bler_r02.json")))
    sc = tab["scenarios"]
    assert set(sc) >= {"turbo_K6144_rate_1_3", "turbo_K5824_rate_0_84", "turbo_K40_rate_1_3", "chain_20MHz_MCS28", "chain_20MHz_MCS28_rv0_rv2"}
    for name, s in sc.items():
        rows = s["rows"]
        assert len(rows) >= 7 and all(r["blocks"] >= 160 for r in rows), name
        # the float reference is never worse than the fixed-point receiver beyond sampling noise, and both fall with SNR
        assert all(r["bler_float"] <= r["bler_int16"] + 0.03 for r in rows), name
        assert rows[0]["bler_int16"] > 0.5 and rows[-1]["bler_int16"] < 0.02, name
        assert s["gap_db"] is not None and s["gap_db"] > -0.05, name
    # SNR penalty at BLER 0.1 of windows + clamps + int16 scaling (SPEC 7.7 quotes these numbers)
    assert sc["turbo_K6144_rate_1_3"]["gap_db"] <= 0.3
    assert sc["turbo_K5824_rate_0_84"]["gap_db"] <= 0.3
    assert sc["turbo_K40_rate_1_3"]["gap_db"] <= 0.1
    assert sc["chain_20MHz_MCS28_rv0_rv2"]["gap_db"] <= 0.3
    # the whole MCS 28 chain (rate 0.84 windows: 0.25 dB, plus the demapper's piecewise-linear int16 LLRs clipped at 511):
    # measured 0.33 dB with 160 subframes per point (+-0.05 dB) -- slightly over the 0.3 dB the turbo code alone keeps
    assert sc["chain_20MHz_MCS28"]["gap_db"] <= 0.4
