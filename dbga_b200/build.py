"""Builds dbga_b200/libdbga_b200.so (hand-written CUDA for sm_100a) with nvcc, in-tree."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdbga_b200.so")
SOURCES = ["hostpack.cpp", "stage12_small.cu", "stage1_kmer.cu", "stage2_anchor.cu", "stage3_dp.cu", "stage3_dp16.cu", "assemble.cu", "pack.cu", "fasta_io.cu", "kmer_stats.cu", "intpeak.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "dbga_b200.h"))
    nvcc = _nvcc()
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(6, max(1, len(jobs)))) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    if jobs or force or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lpthread"])
    # issue-rate microbenchmark of the DP cell's instructions (tools/pipe_rates.cu), a standalone binary
    tool_src = os.path.join(os.path.dirname(HERE), "tools", "pipe_rates.cu")
    tool_bin = os.path.join(os.path.dirname(HERE), "tools", "bin", "pipe_rates")
    if os.path.exists(tool_src) and (force or _stale(tool_bin, [tool_src])):
        os.makedirs(os.path.dirname(tool_bin), exist_ok=True)
        run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", tool_bin, tool_src])
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
