"""Top stalled SASS instructions of an `ncu --page source --csv` dump:  python scripts/ncu_top.py file.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = rows[hi + 1:]
ci, cs = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.lower().startswith("stall")]
tot = sum(float(r[cs] or 0) for r in body if len(r) > cs)
print("kernel:", rows[0][1][:120], "total samples", tot)
order = sorted(range(len(body)), key=lambda i: -float(body[i][cs] or 0))
for i in order[:n]:
    r = body[i]
    v = float(r[cs] or 0)
    top = sorted(((float(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{v / tot * 100:5.1f}% #{i:5d} {r[ci].strip()[:70]:70s} {[(t[1], int(t[0])) for t in top if t[0] > 0]}")
