"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: the last frame of
tools/one_frame.py (or any run whose last 1/n of the launches is one frame).
    python tools/launch_summary.py launches.csv [n_frames]"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hdr, out = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        x = dict(zip(hdr, r))
        out.append((x["Kernel Name"], x.get("Grid Size"), float(x["Metric Value"]) / 1000))
n = len(out) // n_frames
fr = out[-n:]
agg = OrderedDict()
for k, g, t in fr:
    key = k.split("(")[0].replace("void ", "").replace("sb::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(t for _, _, t in fr)
print("| kernel | launches | time us | share |\n|---|---|---|---|")
for k, (c, t) in agg.items():
    print(f"| `{k}` | {c} | {t:.1f} | {100 * t / tot:.1f} % |")
print(f"| total | {n} | {tot:.1f} | |")
if "-v" in sys.argv:
    for k, g, t in fr:
        print(k[:60], g, round(t, 1))
