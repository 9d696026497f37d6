"""Top stall-sample SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv ...` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia = hdr.index("Source"); iss = hdr.index("Warp Stall Sampling (All Samples)"); ie = hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[hi + 1:]):
    if len(r) <= ie or r[0] == "Address" or not r[0].startswith("0x"):
        continue
    data.append((int(r[iss] or 0), int(r[ie] or 0), r[ia].strip(), k))
tot = sum(d[0] for d in data)
print("total samples", tot, "warp-instr executed", sum(d[1] for d in data), "sass lines", len(data))
for s, e, src, k in sorted(data, reverse=True)[:top]:
    print("%6d %5.1f%% exec=%9d  #%4d %s" % (s, 100 * s / max(tot, 1), e, k, src[:100]))
