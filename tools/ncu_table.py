"""Markdown table of the kernels in an ncu report: python tools/ncu_table.py rep.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time us", 1.0), ("launch__grid_size", "grid", 1.0), ("launch__registers_per_thread", "regs", 1.0),
        ("dram__bytes_read.sum", "DRAM rd MB", 1.0), ("dram__bytes_write.sum", "DRAM wr MB", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %", 1.0),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0)]
units = rows[1]
print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    name = name.replace("void ", "").split("(")[0][:44]
    vals = []
    for key, _, _ in cols:
        v = r[ix[key]].replace(",", "") if key in ix else ""
        u = units[ix[key]] if key in ix else ""
        try:
            f = float(v)
            if u == "Gbyte": f *= 1000
            if u == "Kbyte": f /= 1000
            if u == "byte": f /= 1e6
            if u == "ms": f *= 1000
            if u == "ns": f /= 1000
            vals.append(f"{f:.1f}" if f < 1000 else f"{f:.0f}")
        except ValueError:
            vals.append(v)
    print(f"| {name} | " + " | ".join(vals) + " |")
