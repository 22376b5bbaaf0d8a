"""Decode the per-instruction control fields (stall count, yield, barriers) of a kernel's SASS and summarise the
static issue cost of its largest straight-line blocks.  usage: python tools/sass_stalls.py obj kernel_regex [minblock] [dump]"""
import re, subprocess, sys, collections
obj, rx = sys.argv[1], re.compile(sys.argv[2])
minb = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dump = len(sys.argv) > 4
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
cur, ins = None, {}
i = 0
while i < len(out):
    l = out[i]
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); ins[cur] = []
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", l)
    if m and cur:
        hi = re.search(r"/\* (0x[0-9a-f]+) \*/", out[i + 1])
        w = int(hi.group(1), 16)
        ins[cur].append((m.group(2).strip(), (w >> 41) & 0xf, (w >> 45) & 1, (w >> 46) & 7, (w >> 49) & 7, (w >> 52) & 0x3f))
        i += 1
    i += 1
for k, v in ins.items():
    if not rx.search(k): continue
    blk = []
    def flush():
        if len(blk) >= minb:
            n = len(blk); st = sum(max(b[1], 1) for b in blk)
            byop = collections.Counter()
            for b in blk:
                o = b[0].split()[0] if not b[0].startswith("@") else b[0].split()[1]
                byop[o.split(".")[0]] += max(b[1], 1) - 1
            print(f"block n={n} static issue cycles={st} ({st / n:.2f}/instr); extra stall cycles by opcode:", dict(byop.most_common(8)))
            if dump:
                for b in blk: print(f"   s{b[1]:2d} y{b[2]} w{b[3]} r{b[4]} m{b[5]:02x}  {b[0]}")
    for t in v:
        o = t[0].split()[0] if not t[0].startswith("@") else t[0].split()[1]
        if o.split(".")[0] in ("BRA", "EXIT", "BSYNC", "BSSY", "WARPSYNC", "CALL", "RET", "NANOSLEEP"):
            flush(); blk = []
        else:
            blk.append(t)
    flush()
