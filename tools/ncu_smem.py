"""Shared-memory wavefronts per opcode, split by execution frequency, from the SASS source page of an ncu report.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv; python tools/ncu_smem.py src.csv UNITS [split_freq]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
U = float(sys.argv[2]); split = float(sys.argv[3]) if len(sys.argv) > 3 else 0.9
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
iW, iI = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
data = [r for r in rows if len(r) == len(hdr) and r is not hdr and r[iE].isdigit()]
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = 0
for r in data:
    w = float(r[iW] or 0)
    if w == 0: continue
    f = int(r[iE]) / U
    op = (r[iS].split()[1] if r[iS].startswith("@") else r[iS].split()[0])
    a = agg[("hot" if f >= split else "chunk", op)]
    a[0] += f; a[1] += w / U; a[2] += float(r[iI] or 0) / U; tot += w / U
print("total shared wavefronts per unit", round(tot, 1))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:6s}{k[1]:24s} instr {v[0]:6.1f}  wavefronts {v[1]:6.1f}  ideal {v[2]:6.1f}")
