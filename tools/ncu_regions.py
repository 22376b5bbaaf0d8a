"""Executed warp-instructions and stall samples of one kernel, grouped into regions of equal execution count
(loop bodies show up as plateaus), with the opcode mix of every region that matters.
usage: python tools/ncu_regions.py report.ncu-rep KERNEL_REGEX WARP_CHUNKS [min_share]
WARP_CHUNKS = (warp, 16-step chunk) pairs of the launch, e.g. cfg3: 16 * 1536 / 16 * 4096 / 16 = 393216."""
import collections, csv, io, subprocess, sys
rep, rx, wc = sys.argv[1], sys.argv[2], float(sys.argv[3])
min_share = float(sys.argv[4]) if len(sys.argv) > 4 else 0.03
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows if len(r) == len(hdr) and r is not hdr and r[iE].isdigit()]
ex = [int(r[iE]) for r in data]; sm = [int(r[iSm]) for r in data]; src = [r[iS].strip() for r in data]
tot, tots = sum(ex), sum(sm)
print(f"{len(data)} SASS instructions, {tot / wc:.0f} executed per warp-chunk, {tots} samples")
lvl = lambda x: round(x / wc, 2)
regs, start, cur = [], 0, lvl(ex[0])
for k in range(1, len(ex)):
    l = lvl(ex[k])
    if abs(l - cur) > 0.26 * max(cur, l, 0.05):
        regs.append((start, k, cur)); start, cur = k, l
regs.append((start, len(ex), cur))
acc = []
for a, b, l in regs:
    e, s = sum(ex[a:b]), sum(sm[a:b])
    if b - a < 6 and acc:
        pa, pb, pl, pe, ps = acc[-1]; acc[-1] = (pa, b, pl, pe + e, ps + s)
    else:
        acc.append((a, b, l, e, s))
for a, b, l, e, s in acc:
    print(f"[{a:4d},{b:4d}) n={b - a:4d} runs/wc={l:5.2f} instr/wc={e / wc:7.1f} ({100 * e / tot:4.1f}%)  samples {100 * s / tots:4.1f}%   {src[a][:48]}")
for a, b, l, e, s in acc:
    if e / tot < min_share: continue
    h, hs = collections.Counter(), collections.Counter()
    for r in data[a:b]:
        t = r[iS].split()
        if t[0].startswith("@"): t = t[1:]
        op = t[0].split(".")[0]
        h[op] += int(r[iE]) / wc; hs[op] += int(r[iSm])
    print(f"[{a},{b}) {e / wc:.0f} instr/wc:  " + "  ".join(f"{k}={v:.0f}({100 * hs[k] / tots:.1f}%)" for k, v in h.most_common(30)))
