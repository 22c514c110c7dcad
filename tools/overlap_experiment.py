"""Does running the HBM-bound front end of one half-batch next to the ALU-bound turbo decoder of the other pay?
Two plans / two streams, half batches pipelined: front(k+1) || turbo(k).  Prints ms per 4096 subframes.
  python tools/overlap_experiment.py            (serial reference and the two-stream schedule)
Env: SRSUE_TURBO_CTAS_PER_SM=1|2 selects the decoder's footprint."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srsue_b200 as sg  # noqa: E402
from oracle import oracle as o  # noqa: E402


def main():
    B, H, pool = 4096, 2048, 16
    ocell = o.make_cell(100, 1, 1)
    ocfg = o.make_cfg(ocell, sf_idx=1, cfi=1, qm=6, tbs=75376)
    iqs = np.stack([o.gen_subframe(ocell, ocfg, 100 + i, 30.0)[1] for i in range(pool)])
    ctx = sg.Context(0)
    cell = sg.make_cell(100, 1, 1)
    cfg = sg.make_cfg(cell, sf_idx=1, cfi=1, qm=6, tbs=75376)
    d_pool = torch.from_numpy(iqs.view(np.float32).reshape(pool, -1)).cuda()
    halves = []
    for h in range(2):
        plan = sg.PdschPlan(ctx, cell, cfg, H)
        I = plan.info
        t = dict(plan=plan,
                 iq=d_pool[torch.arange(H, device="cuda") % pool].contiguous(),
                 sf=torch.empty((H, 14 * I.nsc * 2), dtype=torch.float32, device="cuda"),
                 ce=torch.empty((H, 14 * I.nsc * 2), dtype=torch.float32, device="cuda"),
                 meas=torch.empty((H, 5), dtype=torch.float32, device="cuda"),
                 sb=torch.empty((H, I.sb_sf_stride), dtype=torch.int16, device="cuda"),
                 pl=torch.zeros((H, I.payload_stride), dtype=torch.uint8, device="cuda"),
                 st=torch.zeros((H, 4), dtype=torch.int32, device="cuda"))
        halves.append(t)

    def front(t):
        t["plan"].ofdm_rx(H, t["iq"], t["sf"])
        t["plan"].chest(H, t["sf"], t["ce"], t["meas"])
        t["plan"].pdsch_llr(H, t["sf"], t["ce"], t["meas"], 0.01, 0, 0, t["sb"])

    def turbo(t):
        t["plan"].pdsch_turbo(H, t["sb"], 4, t["pl"], t["st"])

    def serial(steps):
        for _ in range(steps):
            for t in halves:
                front(t)
                turbo(t)

    s_front, s_turbo = torch.cuda.Stream(), torch.cuda.Stream()

    def overlapped(steps):
        # front(k) on s_front; turbo(k) on s_turbo after front(k); front(k+2) must wait for turbo(k) (same soft buffer)
        done_front = [None, None]
        done_turbo = [None, None]
        for k in range(2 * steps):
            t = halves[k % 2]
            with torch.cuda.stream(s_front):
                if done_turbo[k % 2] is not None:
                    s_front.wait_event(done_turbo[k % 2])
                front(t)
                e = torch.cuda.Event()
                e.record(s_front)
                done_front[k % 2] = e
            with torch.cuda.stream(s_turbo):
                s_turbo.wait_event(done_front[k % 2])
                turbo(t)
                e = torch.cuda.Event()
                e.record(s_turbo)
                done_turbo[k % 2] = e

    for name, fn in (("serial", serial), ("two_streams", overlapped)):
        fn(2)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 8
        a.record()
        fn(steps)
        s_front.synchronize(); s_turbo.synchronize()
        b.record()
        torch.cuda.synchronize()
        ok = all(bool((t["st"][:, 0] == 1).all().item()) for t in halves)
        print(name, "ms per 4096 subframes: %.3f" % (a.elapsed_time(b) / steps), "all CRC ok:", ok,
              "turbo CTAs/SM:", os.environ.get("SRSUE_TURBO_CTAS_PER_SM", "default"), flush=True)


if __name__ == "__main__":
    main()
