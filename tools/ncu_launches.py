"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_*]` launch list: per-kernel count / time / share.
usage: python tools/ncu_launches.py launches.csv [first_n]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
byid = collections.OrderedDict()
for r in rows:
    d = byid.setdefault(r[0], {"name": r[4], "grid": r[8], "block": r[7]})
    d[r[12]] = float(r[14])
ids = list(byid.values())
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(ids)
ids = ids[:n]
agg = collections.OrderedDict()
for d in ids:
    nm = re.sub(r"\(.*", "", d["name"]).replace("<unnamed>::", "").replace("void ", "")
    a = agg.setdefault(nm, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d["gpu__time_duration.sum"]; a[2] += d.get("dram__bytes_read.sum", 0); a[3] += d.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
print(f"{len(ids)} launches, total {tot/1000:.1f} us")
print("| kernel | launches | total us | share | avg us | dram rd MB | dram wr MB |\n|---|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]/1000:.1f} | {100*a[1]/tot:.1f}% | {a[1]/a[0]/1000:.2f} | {a[2]/1e6:.1f} | {a[3]/1e6:.1f} |")
if "--detail" in sys.argv:
    for d in ids:
        print(re.sub(r"\(.*", "", d["name"])[-40:], d["grid"], d["block"], d["gpu__time_duration.sum"])
