"""Host-side sharding of independent subframe batches across the GPUs of one box: every rank
decodes a contiguous share, there is no data-path collective; only the timing (max) and the unit counts
(sum) are reduced.  Replaces, for offline use, the hand-off of one subframe per worker thread that
phch_recv -> thread_pool::start_worker performs (/root/reference/ue/src/phy/phch_recv.cc:309-369,
/root/reference/ue/src/common/thread_pool.cc:206-254)."""


def shard_range(n_units, rank, world):
    """contiguous [lo, hi) share of n_units for `rank`; shares differ by at most one unit"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balance_by_work(work, world):
    """greedy longest-first assignment of heterogeneous batches (mixed bandwidths) to ranks by their
    estimated turbo work sum(C*K*iterations); returns a list of index lists, one per rank"""
    order = sorted(range(len(work)), key=lambda i: -work[i])
    loads, out = [0.0] * world, [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: loads[k])
        out[r].append(i)
        loads[r] += work[i]
    return out


def reduce_metrics(elapsed_s, units, dist=None):
    """whole-job throughput inputs: max over ranks of the elapsed time, sum over ranks of the units"""
    if dist is None or not dist.is_initialized():
        return elapsed_s, units
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([elapsed_s], dtype=torch.float64, device=dev)
    u = torch.tensor([float(units)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def bind_rank_to_gpu_numa(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs local_cpulist of the GPU's PCI function),
    before any pinned host buffer is allocated: pages are then first-touched on that node and the H2D / D2H copies of
    the rank do not cross the socket interconnect.  Host-side dispatch detail of SURVEY 8e (one process per GPU with its
    own pinned staging); returns the CPU set or None when the topology cannot be read."""
    import os
    try:
        import pynvml as nv
        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # nvml prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bus) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001
        return None
