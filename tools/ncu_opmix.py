"""Per-opcode executed-instruction and stall-sample mix of one kernel from an ncu report's source page.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv ; python tools/ncu_opmix.py src.csv WARP_STEPS"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
norm = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows if len(r) == len(hdr) and r is not hdr and r[iE].isdigit()]
tot = sum(int(r[iE]) for r in data)
tots = sum(int(r[iSm]) for r in data)
print(f"total warp-instructions {tot}  ({tot / norm:.1f} per unit), stall samples {tots}")
byop, bys = collections.Counter(), collections.Counter()
for r in data:
    toks = r[iS].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    byop[op] += int(r[iE])
    bys[op] += int(r[iSm])
for op, c in byop.most_common(30):
    print(f"{op:10s} {c / norm:8.1f} per unit   stall-samples {100 * bys[op] / max(tots, 1):5.1f}%")
