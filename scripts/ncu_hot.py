"""Hot SASS lines (by warp-stall samples) of one kernel from `ncu -i X.ncu-rep --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) == len(rows[h])]
ia, iss, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
num = lambda s: int(s) if s.strip().isdigit() else 0
tot = sum(num(r[iss]) for r in data)
print("total samples", tot, "instructions", len(data))
top = sorted(((num(r[iss]), i) for i, r in enumerate(data)), reverse=True)[:topn]
for s, i in sorted(top, key=lambda x: x[1]):
    print(f"{i:5d} {s:6d} {100 * s / max(tot, 1):5.1f}% exec={data[i][ie]:>8} {data[i][ia][:120]}")
