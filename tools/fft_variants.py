"""Times the OFDM-demodulation kernel variants at the headline shape (4096 subframes, N = 2048, 100 PRB):
plain cf32 (ofdm_rx_inplace_kernel), cf32 with the carrier-offset rotation on the loads (ofdm_rx_cfo_kernel) and int16
input (ofdm_rx_inplace_iq16_kernel).  Prints one JSON line with ms per launch and algorithmic GB/s."""
import json
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srsue_b200 as sg  # noqa: E402


def main():
    B, prb, n = 4096, 100, 2048
    ctx = sg.Context(0)
    cell = sg.make_cell(prb, 1, 1)
    plan = sg.PdschPlan(ctx, cell, sg.make_cfg(cell, sf_idx=1, cfi=1, qm=6, tbs=75376), B)
    g = torch.Generator(device="cuda").manual_seed(1)
    d_iq = torch.randn((B, 15 * n * 2), dtype=torch.float32, device="cuda", generator=g)
    d_q = torch.randint(-30000, 30000, (B, 15 * n * 2), dtype=torch.int16, device="cuda", generator=g)
    d_sf = torch.empty((B, 14 * 12 * prb * 2), dtype=torch.float32, device="cuda")
    steps = torch.from_numpy(np.array([sg.host_cfo_step(0.01 * (i % 50) - 0.2, n) or 1 for i in range(B)], np.int32)).cuda()
    out_bytes = B * 14 * 12 * prb * 8
    variants = {
        "cf32": (lambda: plan.ofdm_rx(B, d_iq, d_sf), B * 14 * n * 8 + out_bytes),
        "cf32_cfo": (lambda: plan.ofdm_rx(B, d_iq, d_sf, d_cfo_steps=steps), B * 14 * n * 8 + out_bytes),
        "sc16": (lambda: plan.ofdm_rx_sc16(B, d_q, 1.0 / 32768.0, d_sf), B * 14 * n * 4 + out_bytes),
    }
    res = {}
    for name, (fn, nbytes) in variants.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 10
        a.record()
        for _ in range(K):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / K
        res[name] = {"ms": round(ms, 4), "algorithmic_GBps": round(nbytes / ms / 1e6, 1)}
    print(json.dumps({"shape": "4096 subframes x 14 symbols, N=2048, 1200 bins kept", "variants": res}))


if __name__ == "__main__":
    main()
