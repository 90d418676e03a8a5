"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/r01_launches.txt
  python tools/summarize_ncu.py full gpurun_out/prof_score.ncu-rep > profiles/r01_k_score_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        k = row["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache, serialised)")
    print("%-40s %6s %12s %12s %7s" % ("kernel", "n", "total_ms", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-40s %6d %12.3f %12.3f %6.1f%%" % (k[:40], v[0], v[1] / 1e6, v[1] / v[0] / 1e3, 100 * v[1] / tot))


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("# ncu --set full --clock-control none --import-source on : %s" % path)
    kn = hdr.index("Kernel Name")
    for d in data:
        print("kernel:", d[kn].split("(")[0])
    for i, h in enumerate(hdr):
        if h in KEYS:
            print("%-70s %-14s %s" % (h, units[i], "  ".join(d[i] for d in data)))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, smp, tot = collections.Counter(), collections.Counter(), 0
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] == "Kernel Name":
            break
        try:
            ex = int(r[iex])
        except ValueError:
            continue
        parts = r[isrc].split()
        op = (parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")).split(".")[0]
        ops[op] += ex
        smp[op] += int(r[ismp] or 0)
        tot += ex
    print("\n# executed SASS opcode mix, first captured launch (warp-level instructions)")
    print("total %d" % tot)
    for op, c in ops.most_common(16):
        print("%-8s %14d %5.1f%%  stall-samples %d" % (op, c, 100.0 * c / tot, smp[op]))
    fp64 = sum(c for op, c in ops.items() if op in ("DFMA", "DADD", "DMUL", "DSETP"))
    print("FP64-pipe instructions: %d (%.1f%% of issued)" % (fp64, 100.0 * fp64 / tot))
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    st = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] == "Kernel Name":
            break
        for i in stall_cols:
            try:
                st[hdr[i]] += int(r[i] or 0)
            except ValueError:
                pass
    ts = sum(st.values()) or 1
    print("\n# warp stall samples (all)")
    for k, v in st.most_common(8):
        print("%-28s %9d %5.1f%%" % (k, v, 100.0 * v / ts))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
