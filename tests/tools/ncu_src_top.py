"""Summarise an `ncu --page source --csv` export: top SASS lines by stall samples.
usage: ncu_src_top.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))[2:]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot = sum(int(r[2]) for r in rows); toti = sum(int(r[5]) for r in rows)
print("samples", tot, "warp instructions", toti, "sass lines", len(rows))
top = sorted(range(len(rows)), key=lambda i: -int(rows[i][2]))[:n]
for i in sorted(top):
    r = rows[i]
    print(i, r[1].strip()[:64].ljust(64), r[2], "%.1fM" % (int(r[5]) / 1e6), r[8])
