"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py launches.csv "title" """
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
print("(cold-cache, serialised launches under ncu: compare SHARES, not absolute times)\n")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} n={n:4d} total={us:10.1f} us avg={us/n:8.1f} us share={us/tot*100:5.1f}%")
