"""Turn ncu outputs under gpurun_out/ into the small CSV summaries committed under profiles/.
usage: python tools/ncu_summaries.py <tag> <launches.csv> <report.ncu-rep> "<command profiled>" """
import collections, csv, json, os, re, subprocess, sys

tag, launches, rep, cmd = sys.argv[1:5]
os.makedirs("profiles", exist_ok=True)
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"]
    m = re.match(r"(?:void )?(?:dbga::)?([A-Za-z0-9_]+)", name)
    short = m.group(1) if m else name
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    agg.setdefault(short, []).append(v)
tot = sum(sum(v) for k, v in agg.items() if k != "int_peak_kernel")
with open(f"profiles/{tag}_launches_summary.csv", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 400  {cmd}\n")
    f.write("# shares exclude int_peak_kernel (integer-pipe microbenchmark run once for the DP roofline denominator)\n")
    f.write("kernel,launches,total_us,avg_us,share_of_hot_path\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        share = "" if k == "int_peak_kernel" else f"{sum(v) / tot:.3f}"
        f.write(f"{k},{len(v)},{sum(v):.1f},{sum(v) / len(v):.1f},{share}\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
want = [w for w in want if w in idx]
traffic = {}
with open(f"profiles/{tag}_ncu_full_summary.csv", "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on  {cmd}\n")
    f.write("kernel," + ",".join(want) + "\n")
    f.write("unit," + ",".join(units[idx[w]] for w in want) + "\n")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        f.write(name + "," + ",".join(r[idx[w]] for w in want) + "\n")
        def to_bytes(col):
            v = float(r[idx[col]].replace(",", "")); u = units[idx[col]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        key = re.match(r"([A-Za-z0-9_]+)", name).group(1)
        traffic.setdefault(key, []).append(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
print(open(f"profiles/{tag}_launches_summary.csv").read())
print(json.dumps({k: max(v) for k, v in traffic.items()}))
