"""Plain-copy control for the end-to-end leg: what host <-> device rate does the
platform give N ranks that do nothing but copy?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/pcie_ctl.py [--up-mb 304] [--down-mb 257]

Every rank (one per GPU) moves --up-mb host->device and --down-mb device->host concurrently on two
streams, timed with CUDA events between barriers, in three host-memory variants:
  hostalloc      cudaHostAlloc (torch pin_memory), no CPU affinity           -- what round 1 did
  numa_register  rank bound to the cores of its GPU's NUMA node (sysfs), pages first-touched
                 there, then cudaHostRegister                               -- NUMA-local pinning
  spread_cores   hostalloc + each rank bound to its own slice of the visible cores
Rank 0 prints one JSON line per variant: per-rank ms, aggregate GB/s (sum of bytes / max time)."""
import argparse
import ctypes
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def gpu_numa_node(index: int):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # 00000000:1b:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            return int(f.read().strip()), bus
    except Exception as e:          # noqa: BLE001
        return -1, repr(e)


def node_cpus(node: int):
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            txt = f.read().strip()
        cpus = []
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        return cpus
    except Exception:               # noqa: BLE001
        return []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--up-mb", type=float, default=304.0)
    ap.add_argument("--down-mb", type=float, default=257.0)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nu, nd = int(args.up_mb * 1e6), int(args.down_mb * 1e6)
    dev = torch.device("cuda", local)
    d_up = torch.empty(nu, dtype=torch.uint8, device=dev)
    d_dn = torch.zeros(nd, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    all_cpus = sorted(os.sched_getaffinity(0))
    node, bus = gpu_numa_node(local)
    nodes_present = sorted(int(x[4:]) for x in os.listdir("/sys/devices/system/node") if x.startswith("node") and x[4:].isdigit()) \
        if os.path.isdir("/sys/devices/system/node") else []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(h_up, h_dn, mode):
        def once():
            if mode in ("both", "up"):
                with torch.cuda.stream(s1):
                    d_up.copy_(h_up, non_blocking=True)
            if mode in ("both", "down"):
                with torch.cuda.stream(s2):
                    h_dn.copy_(d_dn, non_blocking=True)
        once()
        barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(s1); s2.wait_event(e0)
        for _ in range(args.reps):
            once()
        e1.record(s1); e2.record(s2)
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), e0.elapsed_time(e2)) / args.reps
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            lst = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(lst, t)
            per = [float(x.item()) for x in lst]
        else:
            per = [ms]
        barrier()
        return per

    def report(name, h_up, h_dn, extra):
        out = {"variant": name, "n_gpus": world, "up_mb": args.up_mb, "down_mb": args.down_mb}
        for mode in ("both", "up", "down"):
            per = measure(h_up, h_dn, mode)
            nbytes = (nu if mode != "down" else 0) + (nd if mode != "up" else 0)
            out[mode] = {"ms_per_rank": [round(x, 3) for x in per], "ms_max": round(max(per), 3),
                         "aggregate_gbs": round(world * nbytes / (max(per) * 1e-3) / 1e9, 1),
                         "per_gpu_gbs": round(nbytes / (max(per) * 1e-3) / 1e9, 1)}
        out.update(extra)
        if rank == 0:
            print(json.dumps(out), flush=True)

    # ---- variant 1: cudaHostAlloc, no affinity
    h_up = torch.empty(nu, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(nd, dtype=torch.uint8).pin_memory()
    h_up.fill_(65)
    report("hostalloc", h_up, h_dn, {"cpus_visible": len(all_cpus), "numa_nodes": nodes_present})
    # ---- variant 3: hostalloc + per-rank core slice
    per = max(1, len(all_cpus) // world)
    mine = all_cpus[local * per:(local + 1) * per] or all_cpus
    os.sched_setaffinity(0, mine)
    report("spread_cores", h_up, h_dn, {"cores_per_rank": len(mine)})
    del h_up, h_dn
    # ---- variant 2: NUMA-local first touch + cudaHostRegister
    cpus = [c for c in node_cpus(node) if c in all_cpus] if node >= 0 else []
    os.sched_setaffinity(0, cpus or all_cpus)
    a_up = np.empty(nu, np.uint8); a_up[:] = 65          # first touch on the bound cores
    a_dn = np.empty(nd, np.uint8); a_dn[:] = 0
    rt = torch.cuda.cudart()
    r1 = rt.cudaHostRegister(a_up.ctypes.data, nu, 0)
    r2 = rt.cudaHostRegister(a_dn.ctypes.data, nd, 0)
    t_up, t_dn = torch.from_numpy(a_up), torch.from_numpy(a_dn)
    report("numa_register", t_up, t_dn, {"gpu_numa_node_rank0": node, "gpu_bus_rank0": bus, "bound_cpus_rank0": len(cpus),
                                         "register_rc": [int(r1), int(r2)], "is_pinned": bool(t_up.is_pinned())})
    rt.cudaHostUnregister(a_up.ctypes.data); rt.cudaHostUnregister(a_dn.ctypes.data)
    os.sched_setaffinity(0, all_cpus)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
