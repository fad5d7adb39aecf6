"""Per-kernel times (us) of the last frame in ncu launch lists: python tools/launch_cmp.py a.csv [b.csv ...]"""
import collections, csv, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault(r[ki].split('(')[0][-48:], []).append(float(r[vi].replace(',', '')) / 1000)
    print(f)
    tot = 0
    for k, v in d.items():
        n = len(v) // 3
        last = v[-n:]
        tot += sum(last)
        print(f"  {k:50s} {n:2d} {sum(last):7.1f}  {[round(x, 1) for x in last]}")
    print("  total", round(tot, 1))
