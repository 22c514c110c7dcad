"""per-subframe latency of the phch_worker call sequence through the srsLTE-shaped symbols (driver/pdsch_offline worker)
usage: python tools/worker_latency.py [prb qm tbs n_sf]"""
import os
import struct
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as o

prb, qm, tbs, n = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (100, 6, 75376, 40)))
ocell = o.make_cell(prb, 1, 1)
ocfg = o.make_cfg(ocell, sf_idx=1, cfi=1, qm=qm, tbs=tbs)
pool = [o.gen_subframe(ocell, ocfg, 100 + i, 30.0)[1] for i in range(4)]
path_in, path_out = "/tmp/worker_in.bin", "/tmp/worker_out.bin"
with open(path_in, "wb") as f:
    f.write(struct.pack("12i", 0x53525355, prb, 1, 1, 1, 1, 0x1234, qm, tbs, 0, n, 4))
    for i in range(n):
        f.write(pool[i % 4].tobytes())
for mode in ("worker", "batch"):
    r = subprocess.run([os.path.join(ROOT, "build", "pdsch_offline"), mode, path_in, path_out], capture_output=True, text=True)
    print(mode, "rc", r.returncode, r.stderr.strip())
out = np.fromfile(path_out, np.uint8)
rec = 12 + tbs // 8
acks = [struct.unpack("i", out[i * rec:i * rec + 4].tobytes())[0] for i in range(n)]
print("acks", sum(acks), "of", n)
