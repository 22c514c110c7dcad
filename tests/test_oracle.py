"""CPU-only checks that pin the oracle (oracle/SPEC.md): known-answer values, independent models, and
TX->RX round trips for the BASELINE.json configs.  The reference's own tests hold no vector for this path
(ue/test/phy/CMakeLists.txt:20-24), so these anchors are what "oracle checked" means here."""
import numpy as np
import pytest


def test_crc_check_values(oracle):
    # CRC catalogue check values over "123456789": CRC-24/LTE-A, CRC-24/LTE-B, CRC-16/XMODEM (36.212 5.1.1 polynomials)
    msg = np.unpackbits(np.frombuffer(b"123456789", np.uint8))
    assert oracle.crc_bits(msg, oracle.CRC24A) == 0xCDE703
    assert oracle.crc_bits(msg, oracle.CRC24B) == 0x23EF52
    assert oracle.crc_bits(msg, oracle.CRC16, 16) == 0x31C3
    # CRC of (message || crc) is zero
    rng = np.random.default_rng(1)
    for poly in (oracle.CRC24A, oracle.CRC24B):
        m = rng.integers(0, 2, 1000, dtype=np.uint8)
        c = oracle.crc_bits(m, poly)
        full = np.concatenate([m, [(c >> (23 - i)) & 1 for i in range(24)]]).astype(np.uint8)
        assert oracle.crc_bits(full, poly) == 0


def test_gold_sequence_against_numpy_model(oracle):
    def gold_np(c_init, n):
        x1 = np.zeros(n + 1600 + 31, np.uint8); x2 = np.zeros_like(x1)
        x1[0] = 1
        for i in range(31):
            x2[i] = (c_init >> i) & 1
        for i in range(n + 1600):
            x1[i + 31] = x1[i + 3] ^ x1[i]
            x2[i + 31] = x2[i + 3] ^ x2[i + 2] ^ x2[i + 1] ^ x2[i]
        return x1[1600:1600 + n] ^ x2[1600:1600 + n]
    for c_init in (0, 1, 0x1234 << 14 | 1 << 9 | 1, 0x7FFFFFFF):
        assert np.array_equal(oracle.gold(c_init, 500), gold_np(c_init, 500))


def test_qpp_all_188_are_permutations(oracle):
    Ks = oracle.qpp_Ks()
    assert len(Ks) == 188 and Ks[0] == 40 and Ks[-1] == 6144
    for K in Ks:
        assert np.array_equal(np.sort(oracle.qpp_perm(K)), np.arange(K)), K
    # spot values of 36.212 Table 5.1.3-3
    assert oracle.qpp_params(40) == (3, 10) and oracle.qpp_params(6144) == (263, 480) and oracle.qpp_params(5824) == (89, 182)


def test_window_rule(oracle):
    for K in oracle.qpp_Ks():
        W = oracle.window_len(K)
        assert K % W == 0 and W % 8 == 0
        assert W >= min(K, 64)
    assert oracle.window_len(5824) == 112 and oracle.window_len(6144) == 128 and oracle.window_len(40) == 40


def test_segmentation_of_baseline_configs(oracle):
    # C x K, filler
    s = oracle.cbsegm(152);   assert (s.C, s.Kp, s.F) == (1, 176, 0)
    s = oracle.cbsegm(75376); assert (s.C, s.Kp, s.Cm, s.F) == (13, 5824, 0, 0)
    s = oracle.cbsegm(30576); assert (s.C, s.Kp, s.Cm, s.F) == (5, 6144, 0, 0)
    s = oracle.cbsegm(6208);  assert (s.C, s.Kp, s.Km, s.Cp, s.Cm, s.F) == (2, 3200, 3136, 1, 1, 56)


def test_pdsch_re_counts(oracle):
    c6 = oracle.make_cell(6, 1, 1); c100 = oracle.make_cell(100, 1, 1); c100b = oracle.make_cell(100, 2, 1)
    assert len(oracle.pdsch_re_list(c6, oracle.make_cfg(c6, sf_idx=1, cfi=1))) == 828
    assert len(oracle.pdsch_re_list(c100, oracle.make_cfg(c100, sf_idx=1, cfi=1))) == 15000
    assert len(oracle.pdsch_re_list(c100b, oracle.make_cfg(c100b, sf_idx=1, cfi=1, tm=2))) == 14400
    # subframe 0 loses PSS/SSS/PBCH REs of the six central PRBs
    n0 = len(oracle.pdsch_re_list(c100, oracle.make_cfg(c100, sf_idx=0, cfi=1)))
    assert n0 == 15000 - 72 * 2 - (60 + 3 * 72)      # SSS+PSS symbols, PBCH symbols 7-10 (symbol 7 carries CRS: 10 data RE per PRB)


def test_ofdm_against_numpy_fft(oracle):
    rng = np.random.default_rng(5)
    for prb, n in ((6, 128), (15, 256), (25, 512), (50, 1024), (75, 1536), (100, 2048)):
        iq = (rng.standard_normal(15 * n) + 1j * rng.standard_normal(15 * n)).astype(np.complex64)
        sf = oracle.ofdm_rx(prb, iq).reshape(14, -1)
        pos, nsc = 0, 12 * prb
        for l in range(14):
            pos += (160 if l % 7 == 0 else 144) * n // 2048
            X = np.fft.fft(iq[pos:pos + n].astype(np.complex128)) / np.sqrt(n)
            pos += n
            ref = np.concatenate([X[n - nsc // 2:], X[1:nsc // 2 + 1]])
            assert np.max(np.abs(sf[l] - ref)) / np.sqrt(np.mean(abs(ref) ** 2)) < 1e-5


def test_turbo_encoder_and_noiseless_decode_all_K(oracle):
    rng = np.random.default_rng(3)
    for K in oracle.qpp_Ks():
        c = rng.integers(0, 2, K, dtype=np.uint8)
        d = oracle.turbo_encode(c)
        assert np.array_equal(d[0:3 * K:3], c)                    # systematic stream
        llr = ((d.astype(np.int16) * 2 - 1) * 64).astype(np.int16)
        bits, it, ok, _ = oracle.tdec(llr, K, 2, 0)
        assert np.array_equal(bits, c), K


def test_rate_matching_roundtrip_and_filler(oracle):
    for K, F, rv, E in ((176, 0, 0, 1656), (5824, 0, 0, 6924), (1056, 24, 2, 3000), (6144, 0, 3, 11520)):
        seq = oracle.rm_sequence(K, F, rv)
        assert len(seq) == 3 * (K + 4) - 2 * F and len(set(seq.tolist())) == len(seq)
        e = np.arange(E, dtype=np.int16) % 200 - 100
        w = oracle.rm_rx(e, K, F, rv)
        ref = np.zeros(3 * K + 12, np.int64)
        for i in range(E):
            ref[seq[i % len(seq)]] = np.clip(ref[seq[i % len(seq)]] + int(e[i]), -511, 511)
        for k in range(F):
            ref[3 * k] = ref[3 * k + 1] = -511
        assert np.array_equal(w, ref)


def test_no_wrap_in_decoder_even_for_extreme_inputs(oracle):
    """SPEC.md 7.6: with the input clamp the wrapping adds never wrap, so wrapping == saturating."""
    L = oracle.lib()
    L.lteo_wrap_events.restype = __import__("ctypes").c_long
    L.lteo_wrap_events(1)
    rng = np.random.default_rng(9)
    for K in (40, 512, 2048, 6144):
        for mode in range(4):
            if mode == 0: x = rng.integers(-32768, 32768, 3 * K + 12)
            elif mode == 1: x = np.full(3 * K + 12, 32767)
            elif mode == 2: x = np.full(3 * K + 12, -32768)
            else: x = rng.integers(0, 2, 3 * K + 12) * 65535 - 32768
            oracle.tdec(x.astype(np.int16), K, 6, 0)
    assert L.lteo_wrap_events(0) == 0


TAPS = None


def _taps():
    rng = np.random.default_rng(77)
    t = (rng.standard_normal((2, 6)) + 1j * rng.standard_normal((2, 6))) * np.array([1, .7, .5, .3, .2, .1])
    return t / np.sqrt((abs(t) ** 2).sum(1, keepdims=True))


@pytest.mark.parametrize("prb,ports,qm,tbs,tm,snr,sf", [
    (6, 1, 2, 152, 1, 10.0, 1),          # BASELINE config 1
    (100, 1, 6, 75376, 1, 30.0, 1),      # config 2
    (100, 2, 4, 30576, 2, 15.0, 1),      # config 3 (frequency-selective channel)
    (25, 1, 4, 4968, 1, 20.0, 5),
    (50, 1, 4, 6208, 1, 18.0, 0),        # two code-block sizes + filler
    (15, 1, 2, 1008, 1, 8.0, 3),
])
def test_tx_rx_round_trip(oracle, prb, ports, qm, tbs, tm, snr, sf):
    cell = oracle.make_cell(prb, ports, 1)
    cfg = oracle.make_cfg(cell, sf_idx=sf, cfi=1, qm=qm, tbs=tbs, tm=tm)
    tb, iq, _ = oracle.gen_subframe(cell, cfg, 7, snr, _taps() if ports == 2 else None)
    rc, pl, meas, it = oracle.ue_dl_decode(cell, cfg, iq, 0.01, 0, 4)
    assert rc == 0 and np.array_equal(pl, tb)
    assert 0.3 < meas[1] < 3.0 and meas[0] > 0       # rsrp around 1, positive noise estimate


def test_harq_combining_improves(oracle):
    """first transmission too noisy, soft combining with rv 2 decodes (dl_harq.cc:191-259 behaviour)"""
    cell = oracle.make_cell(25, 1, 1)
    tbs = 11448
    sb = oracle.new_softbuf(2)
    res = []
    for rv in (0, 2):
        cfg = oracle.make_cfg(cell, sf_idx=2, cfi=2, qm=6, tbs=tbs, rv=rv)
        tb, iq, _ = oracle.gen_subframe(cell, cfg, 42, 11.0)
        sf = oracle.ofdm_rx(25, iq)
        ce, _ = oracle.chest(cell, 2, sf)
        rc, pl = oracle.pdsch_decode(cell, cfg, sf, ce, 0.01, 4, softbuf=sb)
        res.append((rc, np.array_equal(pl, tb)))
    assert res[0][0] != 0 and res[1] == (0, True)


def test_avx2_build_is_bit_identical_to_portable():
    """bench.py times oracle/_build/liblteoracle_avx2.so (window-parallel AVX2 turbo decoder, -O3) as the CPU baseline;
    it must reproduce the portable checker bit for bit: turbo hard bits and iteration counts over code-block sizes
    with 1..96 windows, and the whole chain on the three PDSCH configs."""
    import numpy as np
    from oracle import oracle as o
    if not o.have_avx2():
        pytest.skip("host without AVX2")
    prev = o.select("portable")
    try:
        for K in (40, 104, 176, 512, 1056, 2048, 3136, 5824, 6144):
            for seed, eb in ((1, 0.5), (2, 1.5), (3, None)):
                llr = o.gen_turbo_llrs(K, seed, ebn0_db=eb)[1]
                for crc in (0, 1):
                    o.select("portable")
                    a = o.tdec_mt(llr[None], K, 1, 5, crc)
                    o.select("avx2")
                    assert o.lib().lteo_simd_build() == 1
                    b = o.tdec_mt(llr[None], K, 1, 5, crc)
                    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (K, seed, crc)
        for prb, ports, qm, tbs, tm, snr in ((6, 1, 2, 152, 1, 10.0), (100, 1, 6, 75376, 1, 21.0), (100, 2, 4, 30576, 2, 15.0)):
            o.select("portable")
            cell = o.make_cell(prb, ports, 1)
            cfg = o.make_cfg(cell, sf_idx=1, cfi=1, qm=qm, tbs=tbs, tm=tm)
            iq = o.gen_subframe(cell, cfg, 7, snr)[1]
            a = o.ue_dl_decode(cell, cfg, iq, 0.01, 1, 4)
            o.select("avx2")
            b = o.ue_dl_decode(cell, cfg, iq, 0.01, 1, 4)
            assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
    finally:
        o.select(prev)


def test_turbo_rate_matching_order_equals_matrix_description(oracle):
    """36.212 5.1.4.1 built literally: <NULL>-padded R x 32 matrices written row by row, columns permuted with the pattern
    of Table 5.1.4-1 and read column by column (streams 0 and 1), the shifted permutation for stream 2, the circular
    buffer v0 | v1/v2 interlaced, the start k0 of each redundancy version, NULLs (padding and filler bits of d0, d1)
    skipped -- an independent construction of the closed-form read order the oracle and the GPU tables use."""
    P = [0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30, 1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31]
    for K, F in ((40, 0), (176, 16), (1056, 24), (3136, 56), (6144, 0)):
        D = K + 4
        R = -(-D // 32)
        Kpi = 32 * R
        nd = Kpi - D
        streams = []
        for i in range(3):
            y = [None] * nd + [None if (i < 2 and k < F) else 3 * k + i for k in range(D)]
            if i < 2:
                mat = [y[r * 32:(r + 1) * 32] for r in range(R)]
                streams.append([mat[r][P[c]] for c in range(32) for r in range(R)])
            else:
                streams.append([y[(P[k // R] + 32 * (k % R) + 1) % Kpi] for k in range(Kpi)])
        w = list(streams[0])
        for k in range(Kpi):
            w += [streams[1][k], streams[2][k]]
        Kw = 3 * Kpi
        for rv in range(4):
            k0 = R * (2 * -(-Kw // (8 * R)) * rv + 2)
            want = [w[(k0 + j) % Kw] for j in range(Kw) if w[(k0 + j) % Kw] is not None]
            assert oracle.rm_sequence(K, F, rv).tolist() == want, (K, F, rv)


def test_turbo_encoder_equals_transfer_function(oracle):
    """36.212 5.1.3.2 from its definition: G(D) = [1, (1 + D + D^3) / (1 + D^2 + D^3)] for both constituent encoders, the
    second fed through the QPP interleaver, trellis termination by feeding back the register taps, and the tail multiplexing
    of 5.1.3.2.2 -- written here as shift-register recurrences, independently of the oracle's state-machine step."""
    o = oracle
    rng = np.random.default_rng(8)

    def rsc(bits):
        a = [0, 0, 0]                                    # a[k-1], a[k-2], a[k-3]
        z = []
        for x in bits:
            ak = x ^ a[1] ^ a[2]
            z.append(ak ^ a[0] ^ a[2])
            a = [ak, a[0], a[1]]
        xt, zt = [], []
        for _ in range(3):                               # termination: the input that makes the feedback sum zero
            x = a[1] ^ a[2]
            xt.append(x)
            zt.append(0 ^ a[0] ^ a[2])
            a = [0, a[0], a[1]]
        assert a == [0, 0, 0]
        return z, xt, zt

    for K in (40, 512, 6144):
        c = rng.integers(0, 2, K, dtype=np.uint8)
        pi = o.qpp_perm(K)
        z, x, zt = rsc(c.tolist())
        zp, xp, zpt = rsc(c[pi].tolist())
        d = np.zeros((K + 4, 3), np.uint8)
        d[:K, 0], d[:K, 1], d[:K, 2] = c, z, zp
        d[K] = [x[0], zt[0], x[1]]
        d[K + 1] = [zt[1], x[2], zt[2]]
        d[K + 2] = [xp[0], zpt[0], xp[1]]
        d[K + 3] = [zpt[1], xp[2], zpt[2]]
        assert np.array_equal(o.turbo_encode(c), d.reshape(-1)), K


def test_dci_encoding_equals_standard_description(oracle):
    """36.212 5.3.3.2-5.3.3.4 built literally: CRC16 with the RNTI XORed onto it (MSB first), the rate-1/3 tail-biting
    convolutional code with generators 133, 171, 165 (octal) written as circular convolutions, the sub-block interleaver of
    5.1.4.2.1 per stream (<NULL>s in front, column pattern of Table 5.1.4-2), the circular buffer v0 | v1 | v2 read from 0
    without the NULLs -- an independent construction of what the oracle's PDCCH transmitter produces."""
    o = oracle
    rng = np.random.default_rng(21)
    P = [1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31, 0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30]
    taps = ([0, 2, 3, 5, 6], [0, 1, 2, 3, 6], [0, 1, 2, 4, 6])          # 133, 171, 165 octal, MSB = current bit
    for nb, rnti, E in ((21, 0x4601, 72), (27, 0xFFFF, 144), (39, 0x0003, 288), (31, 0x1234, 576)):
        a = rng.integers(0, 2, nb, dtype=np.uint8)
        crc = o.crc_bits(a, o.CRC16, 16) ^ rnti
        c = a.tolist() + [(crc >> (15 - i)) & 1 for i in range(16)]
        K = len(c)
        streams = [[0] * K for _ in range(3)]
        for i, t in enumerate(taps):
            for k in range(K):
                streams[i][k] = sum(c[(k - j) % K] for j in t) & 1
        R = -(-K // 32)
        v = []
        for i in range(3):
            y = [None] * (32 * R - K) + streams[i]
            mat = [y[r * 32:(r + 1) * 32] for r in range(R)]
            v += [mat[r][P[col]] for col in range(32) for r in range(R)]
        w = [x for x in v if x is not None]                             # reading skips the NULLs, so drop them up front
        want = [w[k % len(w)] for k in range(E)]
        assert o.dci_encode(a, rnti, E).tolist() == want, (nb, rnti, E)


def test_sync_sequences_equal_standard_definition(oracle):
    """36.211 6.11 literally: PSS Zadoff-Chu roots 25/29/34; SSS from the three m-sequences, (m0, m1) taken from the block
    structure of Table 6.11.2.1-1 (differences 1..7 with 30, 29, ... entries; last row 167 -> (2, 9)) rather than from the
    closed form the oracle uses."""
    import ctypes as C
    lib = oracle.lib()

    def mseq(taps):
        x = [0, 0, 0, 0, 1]
        for i in range(26):
            x.append(sum(x[i + t] for t in taps) % 2)
        return [1 - 2 * v for v in x]

    s_t, c_t, z_t = mseq((2, 0)), mseq((3, 0)), mseq((4, 2, 1, 0))
    pairs = [(m0, m0 + d) for d in range(1, 8) for m0 in range(31 - d)][:168]
    assert pairs[0] == (0, 1) and pairs[29] == (29, 30) and pairs[30] == (0, 2) and pairs[167] == (2, 9)
    got = np.zeros(62, np.int8)
    for n1 in range(168):
        m0, m1 = pairs[n1]
        for n2 in range(3):
            for sf5 in (0, 1):
                a, b = (m0, m1) if sf5 == 0 else (m1, m0)
                want = []
                for n in range(31):
                    c0, c1 = c_t[(n + n2) % 31], c_t[(n + n2 + 3) % 31]
                    want.append(s_t[(n + a) % 31] * c0)
                    want.append(s_t[(n + b) % 31] * c1 * z_t[(n + (a % 8)) % 31])
                lib.lteo_sss_seq(n1, n2, sf5, got.ctypes.data_as(C.c_void_p))
                assert got.tolist() == want, (n1, n2, sf5)
    t = np.zeros(128, np.complex64)
    for n2, u in enumerate((25, 29, 34)):
        lib.lteo_pss_time(n2, t.ctypes.data_as(C.c_void_p))
        n = np.arange(62)
        d = np.where(n < 31, np.exp(-1j * np.pi * u * n * (n + 1) / 63), np.exp(-1j * np.pi * u * (n + 1) * (n + 2) / 63))
        X = np.zeros(128, np.complex128)
        X[128 - 31:] = d[:31]                                  # subcarriers -31 .. -1
        X[1:32] = d[31:]                                       # subcarriers +1 .. +31, DC empty
        ref = np.fft.ifft(X) * 128 / np.sqrt(128)
        assert np.max(np.abs(t - ref)) < 1e-6


def test_cell_reference_signal_equals_standard_definition(oracle):
    """36.211 6.10.1 literally: r(m) from the Gold sequence with c_init = 2^10 (7 (n_s + 1) + l + 1)(2 N_ID + 1) + 2 N_ID + 1
    (normal CP), the centred window m' = m + 110 - N_RB, and k = 6 m + (v + N_ID mod 6) mod 6 with v = 0 / 3 for port 0 at
    l = 0 / 4 and the opposite for port 1."""
    import ctypes as C
    lib = oracle.lib()
    for prb, cid in ((6, 0), (25, 77), (100, 503)):
        for ports in (1, 2):
            cell = oracle.make_cell(prb, ports, cid)
            for sf in (0, 3, 9):
                for l in (0, 4, 7, 11):
                    ns, ls = 2 * sf + l // 7, l % 7
                    c_init = (1 << 10) * (7 * (ns + 1) + ls + 1) * (2 * cid + 1) + 2 * cid + 1
                    c = oracle.gold(c_init, 440)
                    re, im = np.zeros(2 * prb, np.int8), np.zeros(2 * prb, np.int8)
                    lib.lteo_crs_values(C.byref(cell), sf, l, re.ctypes.data_as(C.c_void_p), im.ctypes.data_as(C.c_void_p))
                    mp = np.arange(2 * prb) + 110 - prb
                    assert np.array_equal(re, 1 - 2 * c[2 * mp].astype(np.int8)) and np.array_equal(im, 1 - 2 * c[2 * mp + 1].astype(np.int8))
                    for port in range(ports):
                        k = np.zeros(2 * prb, np.int32)
                        n = lib.lteo_crs_positions(C.byref(cell), port, l, k.ctypes.data_as(C.c_void_p))
                        v = (0 if ls == 0 else 3) if port == 0 else (3 if ls == 0 else 0)
                        assert n == 2 * prb and k.tolist() == [6 * m + (v + cid % 6) % 6 for m in range(2 * prb)]


def test_pdsch_bit_chain_equals_standard_description(oracle):
    """36.212 5.1.1-5.1.5 + 36.211 6.3.1 for a whole transport block, written out in numpy: CRC24A, segmentation (B', K+, K-,
    C+, C-, F filler bits in front of the first block), CRC24B per block when C > 1, turbo coding, rate matching with the
    E_r split of 5.1.4.1.2 (G' = G / (N_L Q_m), gamma = G' mod C), concatenation, scrambling with
    c_init = n_RNTI 2^14 + floor(n_s / 2) 2^9 + N_ID.  The turbo encoder and the rate-matching order have their own
    independent constructions above; here they are used as given."""
    import ctypes as C
    o = oracle
    lib = o.lib()
    Ks = o.qpp_Ks()
    rng = np.random.default_rng(33)
    for prb, ports, qm, tbs, tm, rv, sf in ((6, 1, 2, 152, 1, 0, 1), (50, 2, 4, 6208, 2, 0, 4), (100, 1, 6, 75376, 1, 0, 1),
                                             (25, 1, 6, 11448, 1, 2, 0), (75, 1, 6, 55056, 1, 3, 7)):
        cell = o.make_cell(prb, ports, 21)
        cfg = o.make_cfg(cell, sf_idx=sf, cfi=1, rnti=0x3A7B, qm=qm, tbs=tbs, rv=rv, tm=tm)
        tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        G = len(o.pdsch_re_list(cell, cfg)) * qm
        e = np.zeros(G, np.uint8)
        g_out = C.c_int()
        assert lib.lteo_pdsch_encode_bits(C.byref(cell), C.byref(cfg), tb.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p),
                                          C.byref(g_out)) == 0 and g_out.value == G
        # --- the standard, step by step
        a = np.unpackbits(tb)
        crc = o.crc_bits(a, o.CRC24A)
        b = np.concatenate([a, [(crc >> (23 - i)) & 1 for i in range(24)]]).astype(np.uint8)
        B, Z = len(b), 6144
        if B <= Z:
            L, Cn, Bp = 0, 1, B
        else:
            L, Cn = 24, -(-B // (Z - 24))
            Bp = B + Cn * 24
        Kp = min(k for k in Ks if Cn * k >= Bp)
        if Cn == 1:
            Cp, Km, Cm = 1, 0, 0
        else:
            Km = max(k for k in Ks if k < Kp)
            Cm = (Cn * Kp - Bp) // (Kp - Km)
            Cp = Cn - Cm
        F = Cp * Kp + Cm * Km - Bp
        nl = 2 if tm == 2 else 1
        Gp = G // (nl * qm)
        gamma = Gp % Cn
        out, rp = [], 0
        for r in range(Cn):
            K = Km if r < Cm else Kp
            cb = np.zeros(K, np.uint8)
            nfill = F if r == 0 else 0
            take = K - L - nfill
            cb[nfill:nfill + take] = b[rp:rp + take]
            rp += take
            if L:
                c24 = o.crc_bits(cb[:K - L], o.CRC24B)
                cb[K - L:] = [(c24 >> (23 - i)) & 1 for i in range(24)]
            d = o.turbo_encode(cb)
            seq = o.rm_sequence(K, nfill, rv)
            E = nl * qm * (Gp // Cn) if r <= Cn - gamma - 1 else nl * qm * -(-Gp // Cn)
            out.append(d[seq[np.arange(E) % len(seq)]])
        assert rp == B
        bits = np.concatenate(out)
        c_init = (0x3A7B << 14) + (sf << 9) + 21
        want = bits ^ o.gold(c_init, len(bits))
        assert len(want) == G and np.array_equal(e, want), (prb, tbs)


def test_modulation_and_transmit_diversity_equal_standard_description(oracle):
    """36.211 7.1 (QPSK / 16QAM / 64QAM tables as their closed forms), 6.3.3.3 + 6.3.4.3 (two-port transmit diversity:
    y0(2i) = x0(i), y0(2i+1) = x1(i), y1(2i) = -conj(x1(i)), y1(2i+1) = conj(x0(i)), all over sqrt 2) and 6.3.5 (mapping in
    the RE order) against the grid the oracle's transmitter builds, and the CRS it puts around the data."""
    import ctypes as C
    o = oracle
    lib = o.lib()
    rng = np.random.default_rng(35)
    for prb, ports, qm, tbs, tm, sf in ((6, 1, 2, 152, 1, 1), (25, 1, 4, 4968, 1, 3), (50, 1, 6, 30576, 1, 5), (25, 2, 4, 4968, 2, 2),
                                        (100, 2, 6, 46888, 2, 0)):
        cell = o.make_cell(prb, ports, 31)
        cfg = o.make_cfg(cell, sf_idx=sf, cfi=2, rnti=0x0ABC, qm=qm, tbs=tbs, tm=tm)
        tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        re = o.pdsch_re_list(cell, cfg)
        G = len(re) * qm
        e = np.zeros(G, np.uint8)
        g_out = C.c_int()
        assert lib.lteo_pdsch_encode_bits(C.byref(cell), C.byref(cfg), tb.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p),
                                          C.byref(g_out)) == 0
        b = (1 - 2 * e.astype(np.float64)).reshape(-1, qm)           # 1 - 2 b_i
        if qm == 2:
            d = (b[:, 0] + 1j * b[:, 1]) / np.sqrt(2)
        elif qm == 4:
            d = (b[:, 0] * (2 - b[:, 2]) + 1j * b[:, 1] * (2 - b[:, 3])) / np.sqrt(10)
        else:
            d = (b[:, 0] * (4 - b[:, 2] * (2 - b[:, 4])) + 1j * b[:, 1] * (4 - b[:, 3] * (2 - b[:, 5]))) / np.sqrt(42)
        grid = o.pdsch_tx_grid(cell, cfg, tb).reshape(ports, -1)
        if tm == 1:
            assert np.allclose(grid[0][re], d, atol=1e-12)
        else:
            x0, x1 = d[0::2], d[1::2]
            y0 = np.empty(len(d), np.complex128)
            y1 = np.empty(len(d), np.complex128)
            y0[0::2], y0[1::2] = x0, x1
            y1[0::2], y1[1::2] = -np.conj(x1), np.conj(x0)
            assert np.allclose(grid[0][re], y0 / np.sqrt(2), atol=1e-12) and np.allclose(grid[1][re], y1 / np.sqrt(2), atol=1e-12)
        # reference signals: unit-power QPSK on the CRS positions of each port, nothing on the other port's positions
        nsc = 12 * prb
        for port in range(ports):
            for l in (0, 4, 7, 11):
                k = np.zeros(2 * prb, np.int32)
                lib.lteo_crs_positions(C.byref(cell), port, l, k.ctypes.data_as(C.c_void_p))
                sr, si = np.zeros(2 * prb, np.int8), np.zeros(2 * prb, np.int8)
                lib.lteo_crs_values(C.byref(cell), sf, l, sr.ctypes.data_as(C.c_void_p), si.ctypes.data_as(C.c_void_p))
                assert np.allclose(grid[port][l * nsc + k], (sr + 1j * si) / np.sqrt(2), atol=1e-12)
                if ports == 2:
                    assert np.all(grid[1 - port][l * nsc + k] == 0)


def test_pcfich_transmission_equals_standard_description(oracle):
    """36.212 5.3.4 (CFI code words '011', '101', '110' repeated to 32 bits), 36.211 6.7.1 scrambling with
    c_init = (floor(n_s/2) + 1)(2 N_ID + 1) 2^9 + N_ID, QPSK, and the 16 resource elements of 6.7.4 (single port)."""
    import ctypes as C
    o = oracle
    lib = o.lib()
    words = {1: [0, 1, 1], 2: [1, 0, 1], 3: [1, 1, 0]}
    for prb, cid, sf in ((6, 0, 0), (25, 77, 3), (100, 503, 9)):
        cell = o.make_cell(prb, 1, cid)
        k16 = o.pcfich_re(cell)
        for cfi in (1, 2, 3):
            grid = np.zeros((1, 14, 12 * prb), np.complex128)
            lib.lteo_pcfich_tx(C.byref(cell), sf, cfi, grid.ctypes.data_as(C.c_void_p))
            b = np.array([words[cfi][i % 3] for i in range(32)], np.uint8) ^ o.gold((sf + 1) * (2 * cid + 1) * 512 + cid, 32)
            d = ((1 - 2.0 * b[0::2]) + 1j * (1 - 2.0 * b[1::2])) / np.sqrt(2)
            assert np.allclose(grid[0, 0, k16], d, atol=1e-12)
            assert np.count_nonzero(grid) == 16


def test_phich_and_pbch_transmission_equal_standard_description(oracle):
    """PHICH (36.212 5.3.5, 36.211 6.9.1-6.9.3, single port): HI repeated three times, BPSK, spread with the orthogonal
    sequence of Table 6.9.1-2, cell-specific scrambling, three quadruplets on the group's REGs.  PBCH (36.212 5.3.1,
    36.211 6.6): 24 MIB bits + CRC16 XOR the antenna mask, tail-biting code, rate matching to 1920 bits, scrambling with
    c_init = N_ID, QPSK, the quarter of the frame on the 240 resource elements of slot 1."""
    import ctypes as C
    o = oracle
    lib = o.lib()
    W = [[1, 1, 1, 1], [1, -1, 1, -1], [1, 1, -1, -1], [1, -1, -1, 1], [1j, 1j, 1j, 1j], [1j, -1j, 1j, -1j], [1j, 1j, -1j, -1j],
         [1j, -1j, -1j, 1j]]
    for prb, cid, sf in ((6, 5, 0), (25, 77, 3), (100, 503, 8)):
        cell = o.make_cell(prb, 1, cid)
        for ng_x6 in (1, 6, 12):
            g = lib.lteo_phich_groups(prb, ng_x6) - 1
            k12 = o.phich_res(cell, g, ng_x6)
            c = o.gold((sf + 1) * (2 * cid + 1) * 512 + cid, 12)
            for seq in (0, 3, 5, 7):
                for ack in (0, 1):
                    grid = np.zeros((1, 14, 12 * prb), np.complex128)
                    lib.lteo_phich_tx(C.byref(cell), sf, ng_x6, g, seq, ack, grid.ctypes.data_as(C.c_void_p))
                    z = (1 - 2 * ack) * (1 + 1j) / np.sqrt(2)                # BPSK of the repeated indicator
                    d = np.array([W[seq][i % 4] * (1 - 2 * int(c[i])) * z for i in range(12)])
                    assert np.allclose(grid[0, 0, k12], d, atol=1e-12) and np.count_nonzero(grid) == 12
    P = [1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31, 0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30]
    taps = ([0, 2, 3, 5, 6], [0, 1, 2, 3, 6], [0, 1, 2, 4, 6])
    for prb, cid in ((6, 9), (50, 301)):
        cell = o.make_cell(prb, 1, cid)
        g240 = np.zeros(240, np.int32)
        lib.lteo_pbch_res(C.byref(cell), g240.ctypes.data_as(C.c_void_p))
        mib = o.mib_pack(prb, 0, 6, 516)
        crc = o.crc_bits(mib, o.CRC16, 16) ^ 0x0000                          # one antenna port: mask 0
        cbits = mib.tolist() + [(crc >> (15 - i)) & 1 for i in range(16)]
        K = 40
        v = []
        for t in taps:
            y = [None] * (64 - K) + [sum(cbits[(k - j) % K] for j in t) & 1 for k in range(K)]
            mat = [y[r * 32:(r + 1) * 32] for r in range(2)]
            v += [mat[r][P[col]] for col in range(32) for r in range(2)]
        w = [x for x in v if x is not None]
        e = np.array([w[k % 120] for k in range(1920)], np.uint8) ^ o.gold(cid, 1920)
        for q in range(4):
            grid = np.zeros((1, 14, 12 * prb), np.complex128)
            lib.lteo_pbch_tx(C.byref(cell), mib.ctypes.data_as(C.c_void_p), q, grid.ctypes.data_as(C.c_void_p))
            b = e[480 * q:480 * (q + 1)]
            d = ((1 - 2.0 * b[0::2]) + 1j * (1 - 2.0 * b[1::2])) / np.sqrt(2)
            assert np.allclose(grid.reshape(-1)[g240], d, atol=1e-12) and np.count_nonzero(grid) == 240


def test_pdcch_reg_order_equals_standard_description(oracle):
    """36.211 6.8.5 step 10 literally: walk k' upwards and, for each k', l' = 0 .. L-1; wherever (k', l') starts a
    resource-element group (6 subcarriers in symbols with cell-specific reference signals, 4 elsewhere) that is not taken
    by the PCFICH or a PHICH group, it receives the next quadruplet."""
    import ctypes as C
    o = oracle
    lib = o.lib()
    for prb, ports, cid in ((6, 1, 3), (25, 2, 77), (100, 1, 301)):
        cell = o.make_cell(prb, ports, cid)
        pc = {int(k) // 6 * 6 for k in o.pcfich_re(cell)[::4]}
        for cfi in (1, 2, 3):
            nsym = cfi + (1 if prb <= 10 else 0)
            for ng_x6 in (1, 6, 12):
                taken = {(k, 0) for k in pc}
                for g in range(lib.lteo_phich_groups(prb, ng_x6)):
                    taken |= {(int(k) // 6 * 6, 0) for k in o.phich_res(cell, g, ng_x6)[::4]}
                want = []
                for k in range(12 * prb):
                    for l in range(nsym):
                        size = 6 if l == 0 else 4                      # up to two antenna ports: only symbol 0 carries CRS
                        if k % size == 0 and (k, l) not in taken:
                            want.append((k, l))
                rk, rl = o.pdcch_regs(cell, cfi, ng_x6)
                assert list(zip(rk.tolist(), rl.tolist())) == want, (prb, cfi, ng_x6)


def test_pdcch_transmission_equals_standard_description(oracle):
    """36.211 6.8.2-6.8.5 for a whole control region (one port): the coded DCIs at bit 72 n_CCE of the multiplexed block,
    <NIL> elsewhere, scrambling with c_init = floor(n_s/2) 2^9 + N_ID over the whole block, QPSK, quadruplets permuted
    (interleaver + cyclic shift by N_ID) and written to the REGs in mapping order, four data REs per REG."""
    o = oracle
    rng = np.random.default_rng(41)
    for prb, cid, sf, cfi in ((6, 3, 2, 3), (25, 77, 0, 2), (50, 301, 9, 1)):
        cell = o.make_cell(prb, 1, cid)
        rk, rl = o.pdcch_regs(cell, cfi, 6)
        n_reg = len(rk)
        n_cce = n_reg // 9
        dcis = []
        for L, ncce, nb, rnti in ((1, 0, 25, 0x0101), (2, 2, 27, 0x4601), (4, 4, 21, 0xFFFF)):
            if ncce + L <= n_cce:
                dcis.append((rng.integers(0, 2, nb, dtype=np.uint8), rnti, L, ncce))
        grid = np.zeros((1, 14, 12 * prb), np.complex128)
        assert o.pdcch_tx(cell, sf, cfi, dcis, grid, 6) == n_cce
        c = o.gold(sf * 512 + cid, 8 * n_reg)
        sym = np.zeros(4 * n_reg, np.complex128)                           # <NIL> elements transmit nothing
        for bits, rnti, L, ncce in dcis:
            e = o.dci_encode(bits, rnti, 72 * L) ^ c[72 * ncce:72 * (ncce + L)]
            sym[36 * ncce:36 * (ncce + L)] = ((1 - 2.0 * e[0::2]) + 1j * (1 - 2.0 * e[1::2])) / np.sqrt(2)
        src = o.pdcch_quad_perm(n_reg, cid)
        want = np.zeros_like(grid)
        for j in range(n_reg):
            size = 6 if rl[j] == 0 else 4
            ks = [k for k in range(rk[j], rk[j] + size) if not (rl[j] == 0 and k % 3 == cid % 3)]
            assert len(ks) == 4
            want[0, rl[j], ks] = sym[4 * src[j]:4 * src[j] + 4]
        assert np.allclose(grid, want, atol=1e-12)


# ---- uplink shared-channel encoder (lteo_ulsch_encode): pinned by the receive-side functions above, which are themselves
# pinned by the 36.212 constructions of this file: undo scrambling and the channel interleaver, de-match every code block,
# decode it, and the transport block with a good CRC24A must come back, for every redundancy version
@pytest.mark.parametrize("tbs,qm,nof_prb,rv,n_symb", [(152, 2, 6, 0, 12), (2216, 4, 15, 1, 12), (11448, 6, 25, 2, 11), (30576, 4, 100, 3, 12)])
def test_ulsch_encoder_round_trip_through_the_receive_side(oracle, tbs, qm, nof_prb, rv, n_symb):
    o = oracle
    rng = np.random.default_rng(tbs)
    tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    rnti, sf_idx, cell_id = 0x77, 4, 300
    h = o.ulsch_encode(tbs, qm, nof_prb, tb, rv=rv, rnti=rnti, sf_idx=sf_idx, cell_id=cell_id, n_symb=n_symb)
    rows, G = 12 * nof_prb, 12 * nof_prb * n_symb * qm
    assert len(h) == G
    # 36.211 5.3.1 scrambling, then the inverse of the 36.212 5.2.2.8 interleaver written as a matrix transpose
    b = h ^ o.gold((rnti << 14) | (sf_idx << 9) | cell_id, G)
    g = b.reshape(n_symb, rows, qm).transpose(1, 0, 2).reshape(-1)
    llr = (1 - 2 * g.astype(np.int16)) * -40            # bit 1 -> positive LLR, the receive chain's convention
    s = o.cbsegm(tbs)
    stream, pos = [], 0
    for r in range(s.C):
        K, F, E = o.cb_len(s, r), (s.F if r == 0 else 0), o.cb_E(s, G, qm, 1, r)
        w = o.rm_rx(llr[pos:pos + E], K, F, rv)
        pos += E
        bits, it, ok = o.tdec(w, K, 4, 2 if s.C > 1 else 1)[:3]
        assert ok, "code block %d: CRC failed" % r
        stream.append(bits[F:K - (24 if s.C > 1 else 0)])
    assert pos == G
    stream = np.concatenate(stream)
    assert len(stream) == tbs + 24
    assert np.array_equal(np.packbits(stream[:tbs]), tb)
    assert o.crc_bits(stream, o.CRC24A) == 0
