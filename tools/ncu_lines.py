"""Per-source-line view of one ncu capture (--import-source on, -lineinfo build): warp-level instructions
executed, FP64-pipe instructions and stall samples aggregated by CUDA source line.
  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import collections, csv, io, os, subprocess, sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
fname, hdr = None, None


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


lines = {}  # (file, line) -> [text, samples, inst, fp64, lds_excess]
cur = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if r and r[0] == "Line No":
        hdr = r
        iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
        iX = hdr.index("L1 Wavefronts Shared Excessive")
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] != "":
        cur = (fname, num(r[0]))
        lines[cur] = [r[1].strip(), num(r[iS]), num(r[iI]), 0, num(r[iX])]
    elif cur is not None:
        parts = r[3].split()
        op = (parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")).split(".")[0]
        if op in ("DFMA", "DADD", "DMUL", "DSETP"):
            lines[cur][3] += num(r[iI])
tot_s = sum(v[1] for v in lines.values()); tot_i = sum(v[2] for v in lines.values()); tot_f = sum(v[3] for v in lines.values())
print("total samples %d, warp instructions %d, FP64-pipe %d (%.1f%%)" % (tot_s, tot_i, tot_f, 100.0 * tot_f / max(tot_i, 1)))
print("%-22s %7s %6s %12s %12s %9s  %s" % ("file:line", "samples", "%", "inst", "fp64", "lds_exc", "source"))
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-22s %7d %5.1f%% %12d %12d %9d  %s" % ("%s:%d" % k, v[1], 100.0 * v[1] / max(tot_s, 1), v[2], v[3], v[4], v[0][:90]))
if len(sys.argv) > 3:  # bucket score.cu lines into regions: "name:lo-hi,name:lo-hi,..."
    regs = [(a.split(":")[0], int(a.split(":")[1].split("-")[0]), int(a.split(":")[1].split("-")[1])) for a in sys.argv[3].split(",")]
    agg = collections.OrderedDict((r[0], [0, 0, 0]) for r in regs)
    other = collections.Counter()
    for (f, ln), v in lines.items():
        hit = False
        if f == "score.cu":
            for name, lo, hi in regs:
                if lo <= ln <= hi:
                    agg[name][0] += v[1]; agg[name][1] += v[2]; agg[name][2] += v[3]; hit = True
                    break
        if not hit:
            other[f] += v[1]
    for name, v in agg.items():
        print("%-14s samples %6d %5.1f%%  inst %11d  fp64 %11d" % (name, v[0], 100.0 * v[0] / tot_s, v[1], v[2]))
    for f, c in other.most_common():
        print("other %-28s samples %6d %5.1f%%" % (f, c, 100.0 * c / tot_s))
if len(sys.argv) > 4 and sys.argv[4] == "stalls":  # per-region warp-stall reasons (SASS-level samples)
    fname, hdr, cur = None, None, None
    reg_of = {}
    tot = collections.Counter()
    per = collections.defaultdict(collections.Counter)
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = os.path.basename(r[1]); continue
        if r and r[0] == "Line No":
            hdr = r
            cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        if r[0] != "":
            ln = num(r[0])
            name = fname
            if fname == "score.cu":
                name = "score.cu:other"
                for nm, lo, hi in regs:
                    if lo <= ln <= hi:
                        name = nm
                        break
            cur = name
        elif cur is not None:
            for i, h in cols:
                v = num(r[i])
                if v:
                    per[cur][h] += v
                    tot[h] += v
    keys = [k for k, _ in tot.most_common(8)]
    print("%-22s %8s " % ("region", "samples") + " ".join("%14s" % k[6:20] for k in keys))
    for name, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values())):
        sm = sum(c.values())
        print("%-22s %8d " % (name[:22], sm) + " ".join("%13.1f%%" % (100.0 * c[k] / max(sm, 1)) for k in keys))
