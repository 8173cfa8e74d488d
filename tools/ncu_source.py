"""Dev: summarise the SASS source page of one launch in an ncu report (stall reasons, hottest instructions).
    ncu -i rep.ncu-rep --page source --csv --launch-skip K --launch-count 1 > src.csv ; python tools/ncu_source.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; ix = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_")]
data = [r for r in rows[hi + 1:] if len(r) == len(h) and r[0] != "Address"]
num = lambda r, n: int(float(r[ix[n]] or 0))
tot = sum(num(r, "# Samples") for r in data)
print("kernel:", rows[0][1][:80] if len(rows[0]) > 1 else "", "| instructions", len(data), "| samples", tot)
agg = {n: sum(num(r, n) for r in data) for n in stall_cols}
print("stall reasons:", ", ".join(f"{n[6:]} {100 * v / max(1, sum(agg.values())):.1f}%" for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:N]:
    st = {n[6:]: num(r, n) for n in stall_cols if num(r, n) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{r[ix['Address']][-5:]} {num(r, '# Samples'):6d} {100 * num(r, '# Samples') / tot:5.1f}%  {r[ix['Source']][:64]:64s} {st}")
