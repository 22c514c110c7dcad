"""Generates tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference has no fixture for this path and srsLTE cannot be run here, so these vectors do not pin the
oracle to the reference; they freeze the oracle (= oracle/SPEC.md) so that neither it nor the CUDA kernels can
drift silently.  Small on purpose (a few hundred KB)."""
import os
import sys
import hashlib
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    # turbo decoder: a few code blocks per size, noisy, fixed iterations and CRC early stop
    tv = {}
    for K, ebn0 in ((40, 2.0), (512, 1.0), (1056, 1.0), (5824, 1.2)):
        llrs, bits4, its = [], [], []
        for i in range(3):
            rng = np.random.default_rng(1000 * K + i)
            c = rng.integers(0, 2, K, dtype=np.uint8)
            crc = o.crc_bits(c[:K - 24], o.CRC24B)
            c[K - 24:] = [(crc >> (23 - b)) & 1 for b in range(24)]
            d = o.turbo_encode(c).astype(np.float64) * 2 - 1
            sigma2 = 1.0 / (2.0 / 3.0 * 10 ** (ebn0 / 10))
            d += np.random.default_rng(1000 * K + i + 5_000_000).standard_normal(len(d)) * np.sqrt(sigma2)
            llr = np.clip(np.trunc(64 * d), -2048, 2047).astype(np.int16)
            b, it, ok, _ = o.tdec(llr, K, 4, 0)
            b2, it2, ok2, _ = o.tdec(llr, K, 6, 2)
            llrs.append(llr); bits4.append(np.packbits(b)); its.append((it2, ok2, int(np.packbits(b2).sum())))
        tv["llr_%d" % K] = np.stack(llrs)
        tv["bits4_%d" % K] = np.stack(bits4)
        tv["crcrun_%d" % K] = np.array(its, np.int32)
    np.savez_compressed(os.path.join(OUT, "turbo.npz"), **tv)
    # PDSCH chain: config 1 complete (IQ in, everything out), configs 2/3 as digests of the intermediates
    pv = {}
    cell = o.make_cell(6, 1, 1)
    cfg = o.make_cfg(cell, sf_idx=1, cfi=1, qm=2, tbs=152)
    tb, iq, _ = o.gen_subframe(cell, cfg, 1, 10.0)
    sf = o.ofdm_rx(6, iq); ce, meas = o.chest(cell, 1, sf)
    rc, pl, dbg = o.pdsch_decode(cell, cfg, sf, ce, 0.01, 4, want=True)
    pv.update(cfg1_iq=iq, cfg1_tb=tb, cfg1_sf=sf, cfg1_ce=ce, cfg1_meas=meas, cfg1_e=dbg["e"][:1656],
              cfg1_softbuf=dbg["softbuf"][0, :3 * 176 + 12], cfg1_payload=pl, cfg1_rc=np.array([rc]))
    digests = []
    rng = np.random.default_rng(77)
    taps = (rng.standard_normal((2, 6)) + 1j * rng.standard_normal((2, 6))) * np.array([1, .7, .5, .3, .2, .1])
    taps /= np.sqrt((abs(taps) ** 2).sum(1, keepdims=True))
    for name, prb, ports, qm, tbs, tm, snr, tp in (("cfg2", 100, 1, 6, 75376, 1, 30.0, None), ("cfg3", 100, 2, 4, 30576, 2, 15.0, taps)):
        cell = o.make_cell(prb, ports, 1)
        cfg = o.make_cfg(cell, sf_idx=1, cfi=1, qm=qm, tbs=tbs, tm=tm)
        tb, iq, _ = o.gen_subframe(cell, cfg, 20000, snr, tp)
        sf = o.ofdm_rx(prb, iq); ce, meas = o.chest(cell, 1, sf)
        rc, pl, dbg = o.pdsch_decode(cell, cfg, sf, ce, 0.01, 4, want=True)
        s = o.cbsegm(tbs)
        digests.append("%s iq=%s sf=%s ce=%s e=%s sb=%s payload=%s rc=%d iters=%s" % (
            name, sha(iq), sha(sf), sha(ce), sha(dbg["e"][:len(o.pdsch_re_list(cell, cfg)) * qm]),
            sha(dbg["softbuf"][:s.C, :3 * s.Kp + 12]), sha(pl), rc, ",".join(map(str, dbg["iters"]))))
    pv["digests"] = np.array(digests)
    np.savez_compressed(os.path.join(OUT, "pdsch.npz"), **pv)
    print("\n".join(digests))
    control()


def control():
    """control region: PCFICH (CFI + correlations) and PDCCH (LLRs of every CCE, the RNTI every search-space candidate
    decodes to, the blind-search result) of one 25-PRB, 2-port subframe carrying two DCIs at 6 dB"""
    prb, ports, cid, cfi, sf_idx, rnti, nb = 25, 2, 77, 2, 3, 0x4601, 25
    cell = o.make_cell(prb, ports, cid)
    cfg = o.make_cfg(cell, sf_idx=sf_idx, cfi=cfi, rnti=rnti, qm=4, tbs=4968, tm=2)
    rk, _ = o.pdcch_regs(cell, cfi, 6)
    ncce = len(rk) // 9
    ss = o.pdcch_search_space(ncce, sf_idx, rnti)
    bits = np.random.default_rng(25).integers(0, 2, nb, dtype=np.uint8)
    other = np.random.default_rng(26).integers(0, 2, nb, dtype=np.uint8)
    L0, n0 = ss[-1]
    dcis = [(bits, rnti, L0, n0), (other, 0x0999, 1, 0 if n0 >= 1 else ncce - 1)]
    tb, iq, _ = o.gen_subframe(cell, cfg, 31337, 6.0, None, pcfich=True, dcis=dcis)
    sf = o.ofdm_rx(prb, iq)
    ce, meas = o.chest(cell, sf_idx, sf)
    got_cfi, corr = o.pcfich_decode(cell, sf_idx, sf, ce, meas[0])
    llr, nc = o.pdcch_extract_llr(cell, sf_idx, cfi, sf, ce, meas[0])
    rem = np.array([o.pdcch_decode_candidate(llr[72 * n:], L, nb)[1] for L, n in ss], np.int32)
    f, out, L1, n1 = o.pdcch_find_dci(llr, nc, sf_idx, rnti, nb)
    np.savez_compressed(os.path.join(OUT, "control.npz"), iq=iq, cfi=np.array([got_cfi]), corr=corr, llr=llr[:8 * len(rk)],
                        cand=np.array(ss, np.int32), rem=rem, found=np.array([f, L1, n1]), bits=out, sent=bits)
    print("control: cfi", got_cfi, "corr", corr, "found", f, L1, n1, "sent at", L0, n0)


def cfo():
    """carrier-offset correction (SPEC.md 14): 1.4 MHz subframe of noise, +0.37 and -0.081 subcarrier spacings"""
    rng = np.random.default_rng(14)
    x = (rng.standard_normal(1920) + 1j * rng.standard_normal(1920)).astype(np.complex64)
    steps = np.array([o.cfo_step(0.37, 128), o.cfo_step(-0.081, 128)], np.int32)
    y = np.stack([o.cfo_correct(x, int(s)) for s in steps])
    sf = np.stack([o.ofdm_rx(6, v) for v in y])
    np.savez_compressed(os.path.join(OUT, "cfo.npz"), x=x, steps=steps, y=y, sf=sf, tab=o.cfo_table()[::64])
    print("cfo: steps", steps)


EXT = dict(prb=6, ports=2, cid=77, cfi=2, rnti=0x1234, qm=2, tbs=56, sfn=413, phich=((0, 1, 1), (1, 2, 0), (1, 3, 1)))


def extcp():
    """extended cyclic prefix (SPEC.md 15b): a subframe 0 of a 1.4 MHz two-port cell with PSS / SSS, the 216-element PBCH,
    PCFICH, three HARQ indicators in the two groups of one mapping unit and a PDSCH, behind 300 samples of lead-in"""
    c = EXT
    cell = o.make_cell(c["prb"], c["ports"], c["cid"], cp=1)
    cfg = o.make_cfg(cell, sf_idx=0, cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=c["tbs"], tm=2)
    mib = o.mib_pack(25, 0, 6, c["sfn"])
    tb, iq, _ = o.gen_subframe(cell, cfg, 1515, 9.0, None, pcfich=True, sync=True, mib=(mib, c["sfn"] % 4), phichs=list(c["phich"]))
    sf = o.ofdm_rx(c["prb"], iq, cp=1)
    ce, meas = o.chest(cell, 0, sf)
    cfi, corr = o.pcfich_decode(cell, 0, sf, ce, meas[0])
    ph = [o.phich_decode(cell, 0, sf, ce, g, q, float(meas[0])) for g, q, _ in c["phich"]]
    f, bits, ports, off = o.pbch_decode(cell, sf, ce, float(meas[0]))
    rc, pl, dbg = o.pdsch_decode(cell, cfg, sf, ce, float(meas[0]), 4, want=True)
    x = np.concatenate([iq[-300:], iq])
    pk = o.pss_search(x)
    n1, sf5, scorr, cp = o.sss_detect_cp(x, pk["pos"], pk["n_id_2"], 128, 2)
    nre = len(o.pdsch_re_list(cell, cfg))
    np.savez_compressed(os.path.join(OUT, "extcp.npz"), iq=iq, tb=tb, sf=sf[:12 * 72], ce=ce[:, :12 * 72], meas=meas, cfi=np.array([cfi]),
                        corr=corr, phich_ack=np.array([a for a, _ in ph], np.int32), phich_metric=np.array([m for _, m in ph], np.float32),
                        pbch=np.array([f, ports, off], np.int32), mib=bits, mib_sent=mib, e=dbg["e"][:nre * c["qm"]], payload=pl,
                        rc=np.array([rc]), sync=np.array([pk["pos"], pk["n_id_2"], n1, sf5, cp], np.int32),
                        sync_f=np.array([pk["peak"], scorr], np.float32))
    print("extcp: cfi", cfi, "phich", ph, "pbch", f, ports, off, "rc", rc, "sync", pk["pos"], pk["n_id_2"], n1, sf5, cp)


P4 = dict(prb=6, cid=101, cfi=2, rnti=0x2345, qm=4, tbs=208, sfn=822, phich=((0, 1, 1), (0, 5, 0)))


def _taps4():
    rng = np.random.default_rng(44)
    t = (rng.standard_normal((4, 5)) + 1j * rng.standard_normal((4, 5))) * np.array([1, .6, .4, .2, .1])
    return t / np.sqrt((abs(t) ** 2).sum(1, keepdims=True))


def fourports():
    """four-port cell (SPEC.md 15c): a subframe 0 of a 1.4 MHz cell over four independent channels with PBCH (CRC mask
    0x5555), PCFICH, a PDCCH in the common search space, two HARQ indicators and a 16QAM PDSCH"""
    c = P4
    cell = o.make_cell(c["prb"], 4, c["cid"])
    cfg = o.make_cfg(cell, sf_idx=0, cfi=c["cfi"], rnti=c["rnti"], qm=c["qm"], tbs=c["tbs"], tm=2)
    mib = o.mib_pack(6, 0, 6, c["sfn"])
    nb = 21
    bits = np.random.default_rng(4).integers(0, 2, nb, dtype=np.uint8)
    rk, _ = o.pdcch_regs(cell, c["cfi"], 6)
    ss = o.pdcch_search_space(len(rk) // 9, 0, c["rnti"])
    tb, iq, _ = o.gen_subframe(cell, cfg, 1616, 14.0, _taps4(), pcfich=True, mib=(mib, c["sfn"] % 4), phichs=list(c["phich"]),
                               dcis=[(bits, c["rnti"]) + ss[0]])
    sf = o.ofdm_rx(c["prb"], iq)
    ce, meas = o.chest(cell, 0, sf)
    cfi, corr = o.pcfich_decode(cell, 0, sf, ce, meas[0])
    ph = [o.phich_decode(cell, 0, sf, ce, g, q, float(meas[0])) for g, q, _ in c["phich"]]
    f, mbits, ports, off = o.pbch_decode(cell, sf, ce, float(meas[0]))
    llr, nc = o.pdcch_extract_llr(cell, 0, cfi, sf, ce, meas[0])
    fd, out, L1, n1 = o.pdcch_find_dci(llr, nc, 0, c["rnti"], nb)
    rc, pl, dbg = o.pdsch_decode(cell, cfg, sf, ce, float(meas[0]), 4, want=True)
    nre = len(o.pdsch_re_list(cell, cfg))
    np.savez_compressed(os.path.join(OUT, "fourports.npz"), iq=iq, tb=tb, sf=sf, ce=ce, meas=meas, cfi=np.array([cfi]), corr=corr,
                        phich_ack=np.array([a for a, _ in ph], np.int32), phich_metric=np.array([m for _, m in ph], np.float32),
                        pbch=np.array([f, ports, off], np.int32), mib=mbits, mib_sent=mib, llr=llr[:8 * len(rk)],
                        dci=np.array([fd, L1, n1], np.int32), dci_bits=out, dci_sent=bits, e=dbg["e"][:nre * c["qm"]], payload=pl,
                        rc=np.array([rc]))
    print("fourports: cfi", cfi, "phich", ph, "pbch", f, ports, off, "dci", fd, L1, n1, "rc", rc)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cfo":
        cfo()
    elif len(sys.argv) > 1 and sys.argv[1] == "extcp":
        extcp()
    elif len(sys.argv) > 1 and sys.argv[1] == "fourports":
        fourports()
    else:
        main()
        cfo()
        extcp()
        fourports()
