"""Top stall sites from an `ncu --page source --csv` export.  usage: python tools/ncu_top.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
iS_tmp = hdr.index("# Samples")
body = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    if r[iS_tmp] == "# Samples":   # header of the next launch in the same file: keep the first only
        break
    body.append(r)
iS = hdr.index("# Samples")
iSrc = hdr.index("Source")
iEx = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS] or 0) for r in body)
print("total samples", tot)
agg = {}
for h in stall_cols:
    agg[hdr[h]] = sum(int(r[h] or 0) for r in body)
print("by reason:", ", ".join("%s=%.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda t: -t[1]) if v))
order = sorted(range(len(body)), key=lambda i: -int(body[i][iS] or 0))[:n]
for i in sorted(order):
    r = body[i]
    reasons = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print("%5d  %5.1f%%  ex=%-9s %-70s %s" % (i, 100.0 * int(r[iS] or 0) / max(tot, 1), r[iEx], r[iSrc].strip()[:70], reasons))
