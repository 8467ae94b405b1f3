"""Summarise an ncu --set full report of the DP microbenchmark kernels into a small CSV under profiles/.
usage: python tools/ncu_dp_summary.py <tag> <report.ncu-rep> "<command profiled>" """
import csv, subprocess, sys
tag, rep, cmd = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
want = [w for w in want if w in idx]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
with open(f"profiles/{tag}_ncu_full_summary.csv", "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on  {cmd}\n")
    f.write("kernel," + ",".join(want) + ",top_stalls(per issue)\n")
    f.write("unit," + ",".join(units[idx[w]] for w in want) + ",\n")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(const")[0].replace("void ", "").replace("(int)", "").replace("(bool)", "").strip()
        st = sorted(((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stalls), reverse=True)[:4]
        f.write('"' + name + '",' + ",".join(r[idx[w]] for w in want) + "," + " ".join(f"{n}={v:.2f}" for v, n in st) + "\n")
print(open(f"profiles/{tag}_ncu_full_summary.csv").read())
