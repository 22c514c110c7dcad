"""ad-hoc timing of the turbo kernel (device-resident tcb input), used while optimising"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import srsue_b200 as sg
from oracle import oracle as o

K = int(sys.argv[1]) if len(sys.argv) > 1 else 5824
n_cb = int(sys.argv[2]) if len(sys.argv) > 2 else 13 * 4096
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
crc = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ebn0 = float(sys.argv[5]) if len(sys.argv) > 5 else 2.0
ctx = sg.Context(0)
pool = 64
llrs = []
for i in range(pool):
    rng = np.random.default_rng(i)
    c = rng.integers(0, 2, K, dtype=np.uint8)
    crcv = o.crc_bits(c[:K - 24], o.CRC24B)
    c[K - 24:] = [(crcv >> (23 - b)) & 1 for b in range(24)]
    d = o.turbo_encode(c).astype(np.float64) * 2 - 1
    sigma2 = 1.0 / (2.0 * (1.0 / 3.0) * 10.0 ** (ebn0 / 10.0))
    d = d + np.random.default_rng(i + 5_000_000).standard_normal(len(d)) * np.sqrt(sigma2)
    llrs.append(np.clip(np.trunc(64 * d), -2048, 2047).astype(np.int16))
llrs = np.stack(llrs)
W, P, elems = ctx.tdec_geometry(K)
d_tri = torch.from_numpy(llrs[np.arange(n_cb) % pool]).cuda()
d_tcb = torch.zeros((n_cb, elems), dtype=torch.int16, device="cuda")
ctx.tdec_import(d_tri, n_cb, K, d_tcb)
d_bits = torch.zeros((n_cb, K // 8), dtype=torch.uint8, device="cuda")
d_st = torch.zeros(n_cb, dtype=torch.int32, device="cuda")
for _ in range(3):
    ctx.tdec_decode(d_tcb, n_cb, K, iters, crc, d_bits, d_st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    ctx.tdec_decode(d_tcb, n_cb, K, iters, crc, d_bits, d_st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
st = d_st.cpu().numpy()
avg_it = float((st & 0xFF).mean())
ops = 168.0 * K * avg_it * n_cb
print(json.dumps(dict(K=K, W=W, P=P, n_cb=n_cb, max_iter=iters, crc=crc, ms=ms, cb_per_s=n_cb / ms * 1e3,
                      mbit_per_s=n_cb * K / ms / 1e3, avg_iters=avg_it, int16_tops=ops / ms / 1e9,
                      launch=ctx.tdec_last_launch())))
