"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name.  usage: python tools/launch_summary.py file.csv"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
agg = collections.OrderedDict(); seq = []
for r in rows[1:]:
    k = r[ik].split("(")[0][:90]; v = float(r[iv].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; seq.append((k, v))
tot = sum(a[1] for a in agg.values())
unit = 1e6 if rows[1][hdr.index("Metric Unit")] in ("ns", "nsecond") else 1e3
print(f"{len(seq)} launches, {tot / unit:.3f} ms in kernels")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1] / unit:9.3f} ms {100 * a[1] / tot:5.1f}% x{a[0]:4d}  {k}")
