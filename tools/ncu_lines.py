"""Executed warp-instructions and stall samples per CUDA source line for one kernel.
Joins the SASS-level source page of an ncu report with nvdisasm's line info of the same cubin (by instruction order).
usage: python tools/ncu_lines.py report.ncu-rep KERNEL_REGEX cubin MANGLED_SUBSTR [units]"""
import collections, csv, io, re, subprocess, sys
rep, rx, cubin, mangled = sys.argv[1:5]
units = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows if len(r) == len(hdr) and r is not hdr and r[iE].isdigit()]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
# find function section
lines, cur, infn = [], None, False
for l in dis:
    if l.startswith(".text."):
        infn = mangled in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        lines.append(cur)
print(f"{len(data)} profiled SASS instructions, {len(lines)} disassembled", file=sys.stderr)
n = min(len(data), len(lines))
by = collections.Counter(); st = collections.Counter()
for r, ln in zip(data[:n], lines[:n]):
    by[ln] += int(r[iE]); st[ln] += int(r[iSm])
tot = sum(by.values()); tots = sum(st.values())
src = {}
for (f, ln), c in sorted(by.items(), key=lambda kv: (kv[0] is None, kv[0])):
    if c == 0: continue
    if f not in src:
        try: src[f] = open(f"gfe_mamba_b200/csrc/{f}").read().split("\n")
        except Exception: src[f] = []
    text = src[f][ln - 1].strip()[:90] if ln - 1 < len(src[f]) else ""
    print(f"{f}:{ln:4d} {c / units:8.1f} {100 * c / tot:5.1f}%  stall {100 * st[(f, ln)] / max(tots, 1):5.1f}%  | {text}")
