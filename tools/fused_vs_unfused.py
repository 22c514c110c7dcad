import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import srsue_b200 as sg
from oracle import oracle as o
ctx = sg.Context(0)
B = 4096
ocell = o.make_cell(100, 1, 1); ocfg = o.make_cfg(ocell, sf_idx=1, cfi=1, qm=6, tbs=75376)
pool = np.stack([o.gen_subframe(ocell, ocfg, 20000 + i, 30.0)[1] for i in range(8)])
cell = sg.make_cell(100, 1, 1); cfg = sg.make_cfg(cell, sf_idx=1, cfi=1, qm=6, tbs=75376)
plan = sg.PdschPlan(ctx, cell, cfg, B); I = plan.info
d_iq = torch.from_numpy(pool.view(np.float32).reshape(8, -1)).cuda()[torch.arange(B, device="cuda") % 8].contiguous()
d_sf = torch.empty((B, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
d_ce = torch.empty((B, 14 * I.nsc * 2), dtype=torch.float32, device="cuda")
d_pil = torch.empty((B, 4 * 200 * 2), dtype=torch.float32, device="cuda")
d_meas = torch.empty((B, 5), dtype=torch.float32, device="cuda")
d_sb = torch.empty((B, I.sb_sf_stride), dtype=torch.int16, device="cuda")
d_sb2 = torch.empty((B, I.sb_sf_stride), dtype=torch.int16, device="cuda")
plan.ofdm_rx(B, d_iq, d_sf)
def t(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
a = t(lambda: (plan.chest(B, d_sf, d_ce, d_meas), plan.pdsch_llr(B, d_sf, d_ce, d_meas, 0.01, 0, 0, d_sb)))
b = t(lambda: (plan.chest_pilots(B, d_sf, d_pil, d_meas), plan.pdsch_llr_fused(B, d_sf, d_pil, d_meas, 0.01, 0, 0, d_sb2)))
c1 = t(lambda: plan.chest_pilots(B, d_sf, d_pil, d_meas))
print("unfused chest+llr %.3f ms, fused pilots+llr %.3f ms (pilots alone %.3f)" % (a, b, c1), "identical", bool(torch.equal(d_sb, d_sb2)))
