"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list: one line per launch + per-kernel totals/shares.
usage: python tools/launches_summary.py launches.csv [skip_first_n]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = collections.OrderedDict()
def short(n):
    n = re.sub(r"\(.*", "", n.replace("void ", ""))
    return re.sub(r"<.*", "<...>", n) if n.startswith("at::") or "unnamed" in n else n
print("id  kernel  grid  block  us")
for r in rows:
    us = float(r[14]) / 1e3 if r[13] in ("ns", "nsecond") else float(r[14]) * (1e3 if r[13].startswith("ms") else 1.0)
    name = short(r[4])
    print(f"{r[0]:>4s}  {name[:90]:90s} {r[8]:>16s} {r[7]:>12s} {us:10.1f}")
    if int(r[0]) >= skip:
        t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += us
all_us = sum(v[1] for v in tot.values())
print(f"\nper-kernel totals (launch id >= {skip}); cold-cache serialised times: compare SHARES only")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / all_us * 100:6.2f}%  {n:4d} launches  {us:10.1f} us  {k[:100]}")
