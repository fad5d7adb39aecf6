"""Summarise the source page of an ncu report: top stalled SASS instructions.
    ncu -i rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
opc = {}
for r in body:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    e = opc.setdefault(op, [0, 0])
    e[0] += int(r[ix["# Samples"]] or 0)
    e[1] += int(r[ix["Instructions Executed"]] or 0)
print("by opcode (samples, warp-instructions):")
for op, (s, c) in sorted(opc.items(), key=lambda kv: -kv[1][0])[:16]:
    print(f"  {op:10s} {s:7d} {100*s/tot:5.1f}%  {c}")
print("hottest instructions:")
for idx, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][ix["# Samples"]] or 0))[:n]:
    top = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"  #{idx:5d} {int(r[ix['# Samples']]):6d} {r[ix['Source']].strip()[:70]:70s} {top}")
