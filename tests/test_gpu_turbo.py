"""K5/K6a parity: the CUDA turbo decoder (through the C ABI) against the CPU oracle, bit for bit.

Mirrors BASELINE config 4 (turbo sweep): int16 LLRs in srsLTE decoder-input order, K over the table,
fixed iteration counts and CRC early stop."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_gpu(sg, ctx, llrs, K, max_iter, crc_type):
    import torch
    n = llrs.shape[0]
    d_in = torch.from_numpy(llrs).cuda()
    d_bits = torch.zeros((n, K // 8), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx.tdec_run_all(d_in, n, K, max_iter, crc_type, d_bits, d_st)
    torch.cuda.synchronize()
    bits = np.unpackbits(d_bits.cpu().numpy(), axis=1)
    st = d_st.cpu().numpy()
    return bits, st & 0xFF, (st >> 8) & 1


def _batch(oracle, K, n, ebn0, seed0, with_crc=False, scale=64.0):
    llrs = np.zeros((n, 3 * K + 12), np.int16)
    info = np.zeros((n, K), np.uint8)
    for i in range(n):
        rng = np.random.default_rng(seed0 + i)
        c = rng.integers(0, 2, K, dtype=np.uint8)
        if with_crc:
            crc = oracle.crc_bits(c[:K - 24], oracle.CRC24B)
            c[K - 24:] = [(crc >> (23 - b)) & 1 for b in range(24)]
        d = oracle.turbo_encode(c).astype(np.float64) * 2 - 1
        if ebn0 is not None:
            sigma2 = 1.0 / (2.0 * (1.0 / 3.0) * 10.0 ** (ebn0 / 10.0))
            d = d + np.random.default_rng(seed0 + i + 5_000_000).standard_normal(len(d)) * np.sqrt(sigma2)
        llrs[i] = np.clip(np.trunc(scale * d), -2048, 2047).astype(np.int16)
        info[i] = c
    return info, llrs


@pytest.mark.parametrize("K", [40, 48, 104, 128, 176, 264, 512, 528, 1008, 1056, 2048, 2112, 3136, 4992, 5056, 5824, 6144])
def test_turbo_fixed_iterations_bit_exact(gpu, oracle, K):
    sg, ctx = gpu
    info, llrs = _batch(oracle, K, 6, 1.0, 4_0000 + K)
    for iters in (1, 4):
        bits, it, ok = _run_gpu(sg, ctx, llrs, K, iters, 0)
        for i in range(llrs.shape[0]):
            ob, oit, ook, _ = oracle.tdec(llrs[i], K, iters, 0)
            assert np.array_equal(bits[i], ob), "K=%d cb=%d iters=%d: hard bits differ from oracle" % (K, i, iters)
            assert it[i] == oit


def test_turbo_all_K_noiseless_and_noisy(gpu, oracle):
    sg, ctx = gpu
    for K in oracle.qpp_Ks():
        info, llrs = _batch(oracle, K, 2, None, 7_0000 + K)
        _, l2 = _batch(oracle, K, 2, 1.5, 8_0000 + K)
        llrs = np.concatenate([llrs, l2])
        bits, it, ok = _run_gpu(sg, ctx, llrs, K, 3, 0)
        for i in range(4):
            ob, _, _, _ = oracle.tdec(llrs[i], K, 3, 0)
            assert np.array_equal(bits[i], ob), "K=%d cb=%d" % (K, i)
        assert np.array_equal(bits[:2], info)      # noiseless blocks decode to the sent bits


@pytest.mark.parametrize("K,ebn0", [(5824, 0.8), (5824, 2.0), (6144, 1.0), (1024, 1.0), (40, 2.0)])
def test_turbo_crc_early_stop_matches_oracle(gpu, oracle, K, ebn0):
    sg, ctx = gpu
    info, llrs = _batch(oracle, K, 24, ebn0, 9_0000 + K, with_crc=True)
    bits, it, ok = _run_gpu(sg, ctx, llrs, K, 6, 2)
    seen = set()
    for i in range(llrs.shape[0]):
        ob, oit, ook, _ = oracle.tdec(llrs[i], K, 6, 2)
        assert np.array_equal(bits[i], ob), "K=%d cb=%d" % (K, i)
        assert it[i] == oit and ok[i] == ook
        seen.add(int(oit))
    if K >= 1024:
        assert len(seen) > 1, "test should exercise different stopping iterations"


def test_turbo_extreme_inputs_wrap_free_and_exact(gpu, oracle):
    """Inputs far outside the nominal range (clamped to +-511 on import) and adversarial signs."""
    sg, ctx = gpu
    K = 2048
    rng = np.random.default_rng(123)
    llrs = np.stack([
        rng.integers(-32768, 32768, 3 * K + 12).astype(np.int16),
        np.full(3 * K + 12, 32767, np.int16),
        np.full(3 * K + 12, -32768, np.int16),
        (rng.integers(0, 2, 3 * K + 12) * 65535 - 32768).astype(np.int16),
    ])
    bits, it, ok = _run_gpu(sg, ctx, llrs, K, 5, 0)
    for i in range(llrs.shape[0]):
        ob, _, _, _ = oracle.tdec(llrs[i], K, 5, 0)
        assert np.array_equal(bits[i], ob)


def test_turbo_large_batch_property(gpu, oracle):
    """BASELINE-size batch (10^4 code blocks of K=5824): noiseless blocks must decode to what was sent;
    spot-check a sample against the oracle."""
    import torch
    sg, ctx = gpu
    K, pool, n = 5824, 16, 10_000
    info, llrs = _batch(oracle, K, pool, 2.0, 11_0000)
    idx = np.arange(n) % pool
    big = llrs[idx]
    bits, it, ok = _run_gpu(sg, ctx, big, K, 4, 0)
    ref = [oracle.tdec(llrs[i], K, 4, 0)[0] for i in range(pool)]
    for i in range(n):
        assert np.array_equal(bits[i], ref[idx[i]])
    assert np.array_equal(bits[:pool], info)


def test_tdec_layout_roundtrip(gpu):
    import torch
    sg, ctx = gpu
    for K in (40, 5824, 5056):
        W, P, n = ctx.tdec_geometry(K)
        x = torch.randint(-511, 512, (3, 3 * K + 12), dtype=torch.int16, device="cuda")
        t = torch.zeros((3, n), dtype=torch.int16, device="cuda")
        y = torch.zeros_like(x)
        ctx.tdec_import(x, 3, K, t)
        ctx.tdec_export(t, 3, K, y)
        torch.cuda.synchronize()
        assert torch.equal(x, y)
