"""Per-SASS-instruction view of one kernel from an ncu report: executed count per unit, stall samples by reason, and a
per-opcode / per-region roll-up.  usage: python tools/ncu_sass_stalls.py report.ncu-rep KERNEL_REGEX [units] [top]"""
import collections, csv, io, subprocess, sys
rep, rx = sys.argv[1:3]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r is not hdr and r[ix["Instructions Executed"]].isdigit()]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
print(f"{len(data)} SASS instructions, {tot_inst / units:.1f} executed per unit, {tot_s} samples")
agg = collections.Counter()
for r in data:
    for h in reasons:
        agg[h] += int(r[ix[h]])
print("stall samples by reason:", ", ".join(f"{h[6:]}={100 * v / tot_s:.1f}%" for h, v in agg.most_common(10)))
byop = collections.defaultdict(lambda: [0, 0])
for r in data:
    op = r[ix["Source"]].split()[0] if not r[ix["Source"]].strip().startswith("@") else r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    byop[op][0] += int(r[ix["Instructions Executed"]]); byop[op][1] += int(r[ix["# Samples"]])
print("opcode: executed/unit, share of samples")
for op, (e, s) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"  {op:10s} {e / units:8.1f}  {100 * s / tot_s:5.1f}%")
print(f"top {top} instructions by samples: idx, samples%, exec/unit, main reasons, SASS")
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    rs = sorted(((int(r[ix[h]]), h[6:]) for h in reasons), reverse=True)[:3]
    print(f"  {i:5d} {100 * int(r[ix['# Samples']]) / tot_s:5.2f}% {int(r[ix['Instructions Executed']]) / units:7.2f}  "
          f"{' '.join(f'{n}:{c}' for c, n in rs if c)}  | {r[ix['Source']].strip()[:100]}")
