"""Summarise an ncu --csv launch list (gpu__time_duration.sum [+ dram bytes]) per kernel."""
import collections, csv, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = collections.OrderedDict()
for row in csv.DictReader(lines):
    d = rows.setdefault(int(row["ID"]), {"name": row["Kernel Name"], "grid": row["Grid Size"]})
    d[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
def short(n):
    return re.sub(r"\(.*", "", n).replace("pg::", "").replace("void ", "")[:44]
agg = collections.OrderedDict()
for d in rows.values():
    a = agg.setdefault(short(d["name"]) + " " + d["grid"], [0, 0.0, 0.0])
    a[0] += 1; a[1] += d["gpu__time_duration.sum"]; a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {len(rows)} launches, {tot/1e3:.1f} us total (ncu: cold-cache, serialised - compare shares)")
print(f"{'kernel grid':66s} {'n':>4s} {'us':>9s} {'share':>6s} {'us/launch':>9s} {'DRAM MB/launch':>14s} {'GB/s':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (a[2] / a[1]) if a[1] else 0.0
    print(f"{k:66s} {a[0]:4d} {a[1]/1e3:9.1f} {100*a[1]/tot:5.1f}% {a[1]/a[0]/1e3:9.2f} {a[2]/a[0]/1e6:14.2f} {gbs:7.0f}")
