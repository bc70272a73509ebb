"""Print the hottest SASS lines (by warp-stall samples) of one kernel from an ncu source-page CSV:
    ncu -i X.ncu-rep --page source --csv --kernel-id ::name:N > src.csv ; python profiles/top_stalls.py src.csv [n]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
h = rows[1]
sc = h.index("# Samples")
src = h.index("Source")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
body = [r for r in rows[2:] if len(r) > sc]
tot = sum(float(r[sc]) for r in body)
agg = {h[i]: sum(float(r[i]) for r in body) for i in stall_cols}
print("total samples", tot)
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for idx, r in sorted(enumerate(body), key=lambda x: -float(x[1][sc]))[:n]:
    why = max(stall_cols, key=lambda i: float(r[i]))
    print("%6.2f%%  line %5d  %-12s %s" % (100 * float(r[sc]) / tot, idx, h[why][6:], r[src].strip()[:100]))
