"""Uplink shared-channel encoder (srsue_gpu_ulsch_*) against the oracle, bit for bit.

Replaces the bit chain of srslte_ue_ul_pusch_encode_rnti_softbuffer (ue/src/phy/phch_worker.cc:545-590): CRC24A,
segmentation + CRC24B, turbo encoder, rate matching, channel interleaver, scrambling."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (tbs, qm, nof_prb, rv, n_symb): one code block, filler bits, two sizes of code blocks, many code blocks, repetition
# (E > circular buffer), all redundancy versions, SRS-shortened subframe
CASES = [
    (152, 2, 1, 0, 12),
    (152, 2, 6, 1, 12),         # E above the circular buffer: repetition
    (2216, 4, 15, 0, 12),
    (2216, 4, 15, 2, 12),
    (6200, 2, 50, 1, 12),
    (11448, 6, 25, 3, 12),
    (30576, 4, 100, 0, 11),
    (75376, 6, 100, 0, 12),
    (75376, 6, 100, 2, 12),
    (61664, 6, 96, 1, 12),
]


@pytest.mark.parametrize("tbs,qm,nof_prb,rv,n_symb", CASES)
def test_ulsch_encode_matches_oracle(gpu, oracle, tbs, qm, nof_prb, rv, n_symb):
    sg, ctx = gpu
    n = 5
    rng = np.random.default_rng(tbs + rv)
    tb = rng.integers(0, 256, (n, tbs // 8), dtype=np.uint8)
    plan = sg.UlschPlan(ctx, tbs, qm, nof_prb, rv=rv, rnti=0x4321, sf_idx=(3 + rv) % 10, cell_id=77, n_symb=n_symb, max_batch=8)
    out = np.zeros((n, plan.G // 8), np.uint8)
    plan.encode_host(n, tb, out)
    assert plan.G == 12 * nof_prb * n_symb * qm
    for i in range(n):
        ref = oracle.ulsch_encode(tbs, qm, nof_prb, tb[i], rv=rv, rnti=0x4321, sf_idx=(3 + rv) % 10, cell_id=77, n_symb=n_symb)
        assert np.array_equal(np.unpackbits(out[i]), ref), "transport block %d differs" % i
    plan.close()


def test_ulsch_device_pointers_and_batch_limit(gpu, oracle):
    import torch
    sg, ctx = gpu
    tbs, qm, nof_prb = 11448, 4, 50
    plan = sg.UlschPlan(ctx, tbs, qm, nof_prb, max_batch=16)
    tb = np.random.default_rng(9).integers(0, 256, (16, tbs // 8), dtype=np.uint8)
    d_tb = torch.from_numpy(tb).cuda()
    d_out = torch.zeros((16, plan.G // 8), dtype=torch.uint8, device="cuda")
    plan.encode(16, d_tb, d_out)
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    for i in (0, 7, 15):
        assert np.array_equal(np.unpackbits(out[i]), oracle.ulsch_encode(tbs, qm, nof_prb, tb[i]))
    with pytest.raises(sg.GpuError):
        plan.encode(17, d_tb, d_out)
    plan.close()
    with pytest.raises(sg.GpuError):
        sg.UlschPlan(ctx, 1001, 2, 6)          # not a multiple of 8
