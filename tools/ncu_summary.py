"""Summarises an .ncu-rep (read here, no GPU): one line per launch + stall ratios.
    python tools/ncu_summary.py file.ncu-rep [kernel-substring-for-source-page]"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, rows = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
def col(prefix):
    for h in hdr:
        if h.startswith(prefix):
            return h
keys = [("time_us", "gpu__time_duration.sum"), ("fma%", "sm__pipe_fma_cycles_active.avg.pct"), ("issue%", "sm__inst_issued.avg.pct"),
        ("inst_M", "smsp__inst_executed.sum"), ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"),
        ("dram_rd_MB", "dram__bytes_read.sum "), ("dram_wr_MB", "dram__bytes_write.sum "), ("l2hit%", "lts__t_sector_hit_rate"),
        ("smem_conf", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")]
for r in rows:
    out = [r[idx["Kernel Name"]][:44]]
    for name, k in keys:
        h = col(k)
        v = r[idx[h]] if h else "?"
        try:
            v = float(v)
            v = v / 1e6 if name == "inst_M" else v
            out.append(f"{name}={v:.4g}")
        except Exception:
            out.append(f"{name}={v}")
    print("  ".join(out))
print("stall cycles per issued instruction:")
for h in hdr:
    if "issue_stalled" in h and "ratio" in h:
        vals = [float(r[idx[h]] or 0) for r in rows]
        if max(vals) >= 0.03:
            print("  %-22s" % h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""),
                  " ".join("%.2f" % v for v in vals))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    for part in src.split('"Kernel Name",')[1:]:
        lines = part.split("\n")
        if sys.argv[2] not in lines[0]:
            continue
        rd = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        h2 = rd[0]; ix = {h: i for i, h in enumerate(h2)}
        rr = [r for r in rd[1:] if len(r) == len(h2)]
        print(lines[0][:80], "instructions", len(rr), "samples", sum(int(r[ix["# Samples"]]) for r in rr))
        c, cs = Counter(), {}
        for r in rr:
            t = r[ix["Source"]].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            c[op] += int(r[ix["# Samples"]])
            d = cs.setdefault(op, Counter())
            for k in h2:
                if k.startswith("stall_") and "(" not in k:
                    d[k[6:]] += int(r[ix[k]])
        for op, n in c.most_common(12):
            print("  %-8s %5d  %s" % (op, n, dict(cs[op].most_common(4))))
        top = sorted(rr, key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 12]
        for r in top:
            st = {k[6:]: int(r[ix[k]]) for k in h2 if k.startswith("stall_") and "(" not in k and int(r[ix[k]]) > 0}
            print("  ", r[ix["Address"]][-5:], r[ix["Source"]][:64].ljust(64), r[ix["# Samples"]], st)
        break
