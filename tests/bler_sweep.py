#!/usr/bin/env python
"""BLER of the receiver under test (oracle/: int16, +-511 soft buffer, parallel windows with NII, SPEC 5-7) against the
independent float receiver of tests/float_ref (full-length double-precision max-log-MAP, exact max-log demapper) on
IDENTICAL noisy inputs.  Writes profiles/bler_r02.json; tests/test_fixed_point_anchor.py asserts the gap at BLER 0.1.

    python tests/bler_sweep.py [--blocks 400] [--subframes 200]

Scenarios: turbo code alone at K = 6144 rate 1/3, K = 5824 rate 0.84 (the MCS 28 code-block shape,
punctured by the rate matcher), K = 40; the whole PDSCH chain 20 MHz MCS 28 around its waterfall; rv 0 + rv 2 combining."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import floatref as fr  # noqa: E402
from oracle import oracle as o  # noqa: E402


def crossing(snrs, bler, target=0.1):
    """SNR at which BLER crosses `target`, interpolating log10(BLER) linearly between the bracketing points"""
    pts = [(s, max(b, 1e-4)) for s, b in zip(snrs, bler)]
    for (s0, b0), (s1, b1) in zip(pts, pts[1:]):
        if b0 >= target >= b1 and b0 != b1:
            return s0 + (s1 - s0) * (np.log10(b0) - np.log10(target)) / (np.log10(b0) - np.log10(b1))
    return None


def turbo_scenario(K, E, snrs, n, scale, seed0, iters=4):
    """BPSK over AWGN, Es/N0 per coded bit; E coded bits are sent (E = 3K + 12: mother code; less: punctured through the
    rate matcher of 36.212 5.1.4.1, rv 0).  int16 path: trunc(scale * r) clamped like the demapper's output, oracle
    de-matching (+-511 soft buffer) and windowed decoder; float path: the same r through float_ref."""
    f1, f2 = o.qpp_params(K)
    seq = o.rm_sequence(K, 0, 0)
    full = E >= 3 * K + 12
    rows = []
    for snr in snrs:
        sigma = np.sqrt(1.0 / (2.0 * 10.0 ** (snr / 10.0)))
        err_i = err_f = 0
        for i in range(n):
            rng = np.random.default_rng(seed0 + i)
            c = rng.integers(0, 2, K, dtype=np.uint8)
            d = o.turbo_encode(c)                                   # 3K+12, decoder-input order
            idx = np.arange(3 * K + 12) if full else seq[np.arange(E) % len(seq)]
            r = (2.0 * d[idx] - 1.0) + sigma * np.random.default_rng(seed0 + 7_000_000 + i).standard_normal(len(idx))
            e16 = np.clip(np.trunc(scale * r), -32767, 32767).astype(np.int16)
            if full:
                w16 = np.clip(e16, -2048, 2047).astype(np.int16)    # SURVEY 8d cfg4 input; the decoder clamps to +-511 itself
                wf = r.copy()
            else:
                w16 = o.rm_rx(e16, K, 0, 0)
                wf = fr.rate_dematch(r, seq, K, 0)
            bi = o.tdec(w16, K, iters, 0)[0]
            bf = fr.turbo_decode(wf, K, f1, f2, iters)
            err_i += int(not np.array_equal(bi, c))
            err_f += int(not np.array_equal(bf, c))
        rows.append(dict(snr_db=float(snr), blocks=n, bler_int16=err_i / n, bler_float=err_f / n))
        print("  K=%d E=%d %.2f dB: int16 %.4f float %.4f" % (K, E, snr, err_i / n, err_f / n), flush=True)
    return rows


def chain_scenario(snrs, n, seed0, combine_rv2=False, iters=4):
    """20 MHz, MCS 28 (TBS 75376, 13 code blocks of K = 5824), AWGN.  Both receivers start from the oracle's equalised
    symbols (float front end, checked against numpy's FFT elsewhere); from there the int16 path is lteo_pdsch_decode, the
    float path demaps exactly, descrambles, de-matches in float and runs the full-length decoder on every code block."""
    Ks = o.qpp_Ks()
    cell = o.make_cell(100, 1, 1)
    cfgs = {rv: o.make_cfg(cell, sf_idx=1, cfi=1, rnti=0x1234, qm=6, tbs=75376, tm=1, rv=rv) for rv in (0, 2)}
    Cn, Kp, Km, Cp, Cm, F = fr.segmentation(75376, Ks)
    assert (Cn, Kp, Cm, F) == (13, 5824, 0, 0)
    f1, f2 = o.qpp_params(Kp)
    G = len(o.pdsch_re_list(cell, cfgs[0])) * 6
    Gp = G // 6
    Es = [6 * (Gp // Cn) if r <= Cn - (Gp % Cn) - 1 else 6 * -(-Gp // Cn) for r in range(Cn)]
    c_init = (0x1234 << 14) + (1 << 9) + 1
    scr = o.gold(c_init, G).astype(np.float64)
    rows = []
    for snr in snrs:
        err_i = err_f = 0
        for i in range(n):
            soft_i = o.new_softbuf(Cn)
            soft_f = [None] * Cn
            ok_i = ok_f = False
            for t, rv in enumerate((0, 2) if combine_rv2 else (0,)):
                tb, iq, _ = o.gen_subframe(cell, cfgs[rv], seed0 + i, snr, noise_seed=seed0 + i + 5_000_000 + 3_000_000 * t)
                sf = o.ofdm_rx(100, iq)
                ce, _ = o.chest(cell, 1, sf)
                rc, pl, dbg = o.pdsch_decode(cell, cfgs[rv], sf, ce, 0.01, iters, softbuf=soft_i, want=True)
                ok_i = rc == 0 and np.array_equal(pl, tb)
                llr = fr.demap(dbg["d"][:Gp].astype(np.complex128), 6) * (1.0 - 2.0 * scr)
                pos, good = 0, True
                bits_tb = []
                for r in range(Cn):
                    seq = o.rm_sequence(Kp, 0, rv)
                    soft_f[r] = fr.rate_dematch(llr[pos:pos + Es[r]], seq, Kp, 0, soft_f[r])
                    pos += Es[r]
                    b = fr.turbo_decode(soft_f[r], Kp, f1, f2, iters)
                    good &= o.crc_bits(b, o.CRC24B) == 0              # remainder over data + CRC is zero
                    bits_tb.append(b[:Kp - 24])
                a = np.concatenate(bits_tb)
                ok_f = bool(good) and o.crc_bits(a, o.CRC24A) == 0 and np.array_equal(np.packbits(a[:75376]), tb)
            err_i += int(not ok_i)
            err_f += int(not ok_f)
        rows.append(dict(snr_db=float(snr), blocks=n, bler_int16=err_i / n, bler_float=err_f / n))
        print("  chain%s %.2f dB: int16 %.4f float %.4f" % (" rv0+rv2" if combine_rv2 else "", snr, err_i / n, err_f / n), flush=True)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=400)
    ap.add_argument("--subframes", type=int, default=160)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "bler_r02.json"))
    a = ap.parse_args()
    if o.have_avx2():
        o.select("avx2")
    t0 = time.time()
    out = {"what": "BLER of the int16 windowed receiver (oracle/, SPEC 5-7) vs the independent float full-length receiver "
                   "(tests/float_ref) on identical inputs, 4 turbo iterations; snr_db is Es/N0 per coded BPSK bit for the "
                   "turbo scenarios and the channel SNR for the chain",
           "generated_by": "tests/bler_sweep.py", "scenarios": {}}
    sc = out["scenarios"]
    sc["turbo_K6144_rate_1_3"] = dict(K=6144, E=3 * 6144 + 12, llr_scale=64, rows=turbo_scenario(6144, 3 * 6144 + 12, np.arange(-4.4, -3.39, 0.15), a.blocks, 64.0, 11_000_000))
    sc["turbo_K5824_rate_0_84"] = dict(K=5824, E=6924, llr_scale=64, rows=turbo_scenario(5824, 6924, np.arange(2.4, 4.01, 0.2), a.blocks, 64.0, 12_000_000))
    sc["turbo_K40_rate_1_3"] = dict(K=40, E=132, llr_scale=64, rows=turbo_scenario(40, 132, np.arange(-6.0, -0.99, 0.5), 4 * a.blocks, 64.0, 13_000_000))
    sc["chain_20MHz_MCS28"] = dict(rows=chain_scenario(np.arange(20.75, 22.51, 0.25), a.subframes, 14_000_000))
    sc["chain_20MHz_MCS28_rv0_rv2"] = dict(rows=chain_scenario(np.arange(13.0, 14.51, 0.25), a.subframes, 15_000_000, combine_rv2=True))
    for name, s in sc.items():
        snrs = [r["snr_db"] for r in s["rows"]]
        s["snr_at_bler_0.1_int16"] = crossing(snrs, [r["bler_int16"] for r in s["rows"]])
        s["snr_at_bler_0.1_float"] = crossing(snrs, [r["bler_float"] for r in s["rows"]])
        if s["snr_at_bler_0.1_int16"] is not None and s["snr_at_bler_0.1_float"] is not None:
            s["gap_db"] = s["snr_at_bler_0.1_int16"] - s["snr_at_bler_0.1_float"]
        print(name, "gap", s.get("gap_db"))
    out["seconds"] = time.time() - t0
    json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
