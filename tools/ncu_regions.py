"""Groups the SASS instructions of an `ncu --page source --csv --print-source sass` export into regions of equal
execution count and prints each region's share of executed warp instructions and of stall samples.
usage: python tools/ncu_regions.py export.csv [min_share]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows[hi + 1:] if len(r) > ie and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in data)
ts = sum(int(r[isamp]) for r in data)
print("total warp instructions", tot, "samples", ts, "static instructions", len(data))
reg, cur = [], None
for i, r in enumerate(data):
    e, s = int(r[ie]), int(r[isamp])
    t = r[ia].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    if cur is None or abs(e - cur["e"]) > 0.15 * max(e, cur["e"], 1):
        cur = {"start": i, "e": e, "n": 0, "inst": 0, "samp": 0, "ops": {}}
        reg.append(cur)
    cur["n"] += 1
    cur["inst"] += e
    cur["samp"] += s
    cur["ops"][op] = cur["ops"].get(op, 0) + 1
for c in reg:
    if c["inst"] > min_share * tot or c["samp"] > min_share * ts:
        top = sorted(c["ops"].items(), key=lambda kv: -kv[1])[:7]
        print(f"idx {c['start']:5d} n={c['n']:4d} exec={c['e']:9d} inst%={100 * c['inst'] / tot:5.1f} "
              f"samp%={100 * c['samp'] / ts:5.1f}", top)
