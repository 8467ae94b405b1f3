#!/usr/bin/env python
"""Benchmark of the DBGA pairwise hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the whole hot path (stage 1 k-mer hashing, stage 2 anchors,
stage 3 DP, assembly) over one batch of synthetic pairs.  Workload (N=1): BASELINE config
[2], "100k synthetic 1.5 kb pairs at 2 % divergence, k=9" (config [1], a single 10 kb pair at
k=7, is parity-correctly a "Cycles detected" error for every seed and is covered in tests).
N>1: every rank gets its own 100k-pair batch (weak scaling), no collective on the data
path; time = max over ranks.

JSON keys beyond the base contract: `roofline` (dominant kernel vs measured HBM peak),
`cpu_baseline` (oracle C port on the host cores), `dp` (stage-3 GCUPS vs the measured
integer-pipe rate on the workload, and `dp.microbench`: the same on 1024 x 1024 segments),
`stages_ms`, `cycle_frac`.
"""
from __future__ import annotations

import argparse
import os as _os
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before CUDA initialises: one hardware queue per stream
# stdout carries exactly ONE JSON line: everything else that libraries print there (NCCL's version
# banner at the first communicator, for one) is sent to stderr; emit() writes to the real stdout
_REAL_STDOUT = _os.dup(1)
_os.dup2(2, 1)


def emit(line: dict) -> None:
    _os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned pairs/s + DP GCUPS at 1/2/4/8 B200 vs ref CPU on host cores"
WORKLOADS = {
    # name: (pairs, n, p_sub, p_indel, k)
    "cfg3": (100_000, 1500, 0.018, 0.002, 9),
    "cfg4": (10_000, 30_000, 0.004, 0.001, 11),
}


def workload_desc(name, P):
    _, n, ps, pi, k = WORKLOADS[name]
    return (f"{name}: {P} synthetic {n} nt pairs, p_sub={ps}, p_indel={pi} (indel len ~Geom(mean 3)), k={k}, "
            "default scores (match 10, transition -1, transversion -8, d 10, e 2)")


# ------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return {}
    return {}


# ------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU implementation of the path.  The reference is pure Python +
    cogent3 (absent here and on the GPU box) and cannot travel, so this arm times its C
    restatement (oracle/dbga_oracle.c, kind "port") with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dbga_b200.synth import synth_batch
    from oracle import c_oracle as CO
    name = args.workload
    P_full, n, ps, pi, k = WORKLOADS[name]
    P = min(args.pairs or P_full, args.ref_sample)
    buf, offs = synth_batch(P, n, ps, pi, seed=args.seed)
    cores = os.cpu_count() or 1
    for _ in range(max(1, args.warmup)):
        CO.align_batch(buf, offs, k, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = CO.align_batch(buf, offs, k, threads=cores)
    dt = time.perf_counter() - t0
    v = P * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": {"workload": workload_desc(name, P_full), "sample": f"{P} pairs per step"},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"{P} pairs of the workload per step, {args.steps} steps; C restatement of the "
                                   "reference path (reference = Python + cogent3, not runnable here)"},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cycle_frac": float((out["status"] == 1).mean()),
    }
    emit(line)


# ------------------------------------------------------------------------ our arm
def run_ours(args):
    import ctypes as C
    import torch
    from dbga_b200 import Aligner, _lib
    from dbga_b200.synth import synth_batch, pairs_from_buffers

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    name = args.workload
    P_full, n, ps, pi, k = WORKLOADS[name]
    P = args.pairs or P_full
    al = Aligner(local)
    lib = al.lib
    # pinned host input (the e2e leg copies from here every step)
    cap = P * (2 * n + 64)
    hptr = lib.dbga_host_alloc(cap)
    hbuf = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_uint8)), shape=(cap,))
    buf, offs = synth_batch(P, n, ps, pi, seed=args.seed + 7919 * rank, out=hbuf)
    total_bytes = int(offs[-1])
    prm = al.params(k)

    stream = torch.cuda.Stream(device=dev)      # the library launches on this stream; events too
    torch.cuda.set_stream(stream)
    al.set_stream(stream.cuda_stream)
    d_seqs = torch.empty(total_bytes + 16, dtype=torch.uint8, device=dev)
    d_seqs[:total_bytes].copy_(torch.from_numpy(buf[:total_bytes]), non_blocking=False)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg (`value`): plan once, K x kernels only
    al.plan(offs, prm)
    for _ in range(max(3, args.warmup)):
        al.run(d_seqs.data_ptr())
        res = al.fetch(unpack=False)      # also sizes the pools (re-runs stage 2.. on overflow)
    launches_per_step = int(res.kernel_launches)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        al.run(d_seqs.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    barrier()
    value = world * P * args.steps / (ms * 1e-3)

    # ---------------- per-stage device times (CUDA events inside the library, same stream)
    st_acc = np.zeros(4)
    for _ in range(args.steps):
        al.run(d_seqs.data_ptr())
        r = al.fetch(unpack=False)
        st_acc += [r.ms_stage1, r.ms_stage2, r.ms_stage3, r.ms_assemble]
    st_ms = st_acc / args.steps
    full = al.fetch()
    status = full.status
    ok = status == 0
    cells = int(full.timing["dp_cells"])
    M_total = int(full.n_anchors[ok | (status == 1)].sum())
    # sanity on the timed workload (size-independent properties; the oracle is not used here)
    rs = np.random.default_rng(1)
    for i in rs.choice(P, size=min(P, 64), replace=False):
        if status[i] != 0:
            continue
        a0, a1 = full.alignment(int(i))
        s0, s1 = pairs_from_buffers(buf, offs, [int(i)])[0]
        assert len(a0) == len(a1) and a0.replace("-", "") == s0 and a1.replace("-", "") == s1, "degap check failed"

    # ---------------- end-to-end leg: pinned host in -> C ABI -> pinned host out
    al2 = Aligner(local)          # own streams: dbga_align_batch pipelines chunks over three lanes
    for _ in range(2):
        raw = al2.align_buffers(hptr, offs, prm, unpack=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        raw = al2.align_buffers(hptr, offs, prm, unpack=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_value = world * P * args.steps / dt
    n_aln = int(np.ctypeslib.as_array(raw.aln_off, shape=(P + 1,))[-1])
    n_runs = int(np.ctypeslib.as_array(raw.run_off, shape=(P + 1,))[-1])
    h2d = total_bytes + 40 * P + 4 * P
    d2h = 2 * n_aln + 12 * n_runs + P * (4 + 8 + 8 + 4 + 8) + 16 * (P + 1)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (stage 1) and DP GCUPS
    peaks, peak_src = measured_peaks()
    stage_names = ["stage12_kmer_anchor", "stage2_large_pairs", "stage3_dp", "assemble"]
    dom = int(np.argmax(st_ms))
    n_runs_total = int(full.run_off[-1])
    alg_bytes_12 = total_bytes + 12 * n_runs_total                 # (n0+n1) read + anchors written as (a, b, len) runs
    alg_bytes = {0: alg_bytes_12, 1: 4 * int((offs[2::2] - offs[1::2]).sum()) + 12 * n_runs_total,
                 2: 2 * n_aln, 3: total_bytes + 2 * n_aln}[dom]
    achieved = alg_bytes / (st_ms[dom] * 1e-3) / 1e9
    traffic = ncu_traffic() if name == "cfg3" else {}        # the committed ncu capture is of the cfg3 workload
    roofline = {"bound": "hbm", "kernel": stage_names[dom], "achieved": achieved, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic.get(stage_names[dom]),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "launch_ms": float(st_ms[dom])}
    intpk = al.int_peak()
    gcups = cells / (st_ms[2] * 1e-3) / 1e9 if st_ms[2] > 0 else 0.0
    roof_gcups = intpk["mix_ops_per_s"] / 9 / 1e9
    dp = {"gcups": gcups, "cells_per_step": cells, "roofline_gcups": roof_gcups, "frac": gcups / roof_gcups,
          "int_ops_per_s": intpk, "note": "cells = sum |x||y| over segments handed to the DP (SURVEY 8d); "
          "roofline = measured int32 5add:4max mix rate / 9 ops per cell; cfg3 segments are ~10x10 so this "
          "stage is latency/launch bound here (see DESIGN.md DP microbenchmark)"}

    # ---------------- DP microbenchmark (the integer roofline is only testable on long
    # segments): 3814 uniform 1024 x 1024 global alignments, device-resident, stage-3 time from the
    # library's CUDA events on the launching stream (forward kernels + traceback kernel)
    if not args.no_dp_micro:
        L, Pm = 1024, 3814
        mb, mo = synth_batch(Pm, L, 0.05, 0.01, seed=L)
        d_m = torch.from_numpy(np.concatenate([mb, np.zeros(64, np.uint8)])).to(dev)
        torch.cuda.synchronize()
        alm = Aligner(local)
        alm.plan(mo, alm.params(1), global_only=True)
        t3 = []
        for it in range(6):
            alm.run(d_m.data_ptr())
            rm = alm.fetch(unpack=False)
            if it >= 3:
                t3.append(rm.ms_stage3)
        mcells = int(rm.dp_cells_total)
        mg = mcells / (min(t3) * 1e-3) / 1e9
        dp["microbench"] = {"workload": f"{Pm} global alignments of {L} x {L} (5 % substitutions, 1 % indels)",
                            "cells": mcells, "stage3_ms": min(t3), "gcups": mg, "frac": mg / roof_gcups,
                            "kernel": "dp_wf16_kernel<16> (packed int16x2 wavefront) + dp_tb16_kernel"}
        alm.close()

    # ---------------- CPU baseline on this box's host cores (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import c_oracle as CO
        cores = os.cpu_count() or 1
        sample = min(P, args.ref_sample)
        sb, so = buf[:int(offs[2 * sample])], offs[:2 * sample + 1]
        CO.align_batch(sb[:int(so[min(2 * 512, 2 * sample)])], so[:min(2 * 512, 2 * sample) + 1], k, threads=cores)
        passes, cdt = 0, 0.0
        while cdt < args.cpu_seconds:                 # whole passes over the sample until the budget is spent
            t0 = time.perf_counter()
            CO.align_batch(sb, so, k, threads=cores)
            cdt += time.perf_counter() - t0
            passes += 1
        cpu = {"value": sample * passes / cdt, "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": f"first {sample} pairs of the same batch, {passes} passes, {cdt:.1f} s of CPU time; C restatement "
                         "(oracle/dbga_oracle.c) of the reference path, pthreads over all host cores"}

    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_desc(name, P), "pairs_per_gpu": P,
                   "l2": f"inputs {total_bytes / 1e6:.0f} MB + outputs {2 * n_aln / 1e6:.0f} MB per step exceed the 126 MB L2",
                   "parallelism": f"pairs sharded by index over {world} GPU(s), no collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": dt / args.steps * 1e3},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "dp": dp,
        "stages_ms": dict(zip(stage_names, [float(x) for x in st_ms])),
        "cycle_frac": float((status == 1).mean()),
        "ok_frac": float(ok.mean()),
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (default: the workload's)")
    ap.add_argument("--ref-sample", type=int, default=50_000, help="pairs per CPU-baseline pass")
    ap.add_argument("--seed", type=int, default=20261019)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU time spent on the cpu_baseline leg")
    ap.add_argument("--no-dp-micro", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
