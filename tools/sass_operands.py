"""Register-file read traffic of a kernel's largest straight-line blocks: for every SASS instruction count the 32-bit
source words it reads from the vector register file (a .F32x2 / .64 operand counts 2, an R.F32 scalar-broadcast operand 1,
uniform registers / immediates / constants 0).  With ~2 words per lane per clock per scheduler (tools/microbench_forms.cu)
words / 2 is the operand-delivery time of the block.   usage: python tools/sass_operands.py obj kernel_regex [minblock]"""
import collections, re, subprocess, sys
obj, rx = sys.argv[1], re.compile(sys.argv[2])
minb = int(sys.argv[3]) if len(sys.argv) > 3 else 300
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
cur, ins = None, {}
for l in out:
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); ins[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur: ins[cur].append(m.group(2).strip())
STORE = ("STS", "STG", "ST", "STL", "RED", "ATOMG", "ATOMS")
def words(i):
    t = i.split()
    if t[0].startswith("@"): t = t[1:]
    op = t[0]; base = op.split(".")[0]
    args = " ".join(t[1:])
    parts = [a.strip() for a in re.split(r",(?![^\[]*\])", args)] if args else []
    srcs = parts if base in STORE or base in ("BRA", "BAR", "EXIT", "NOP", "ISETP", "FSETP") else parts[1:]
    if base in ("ISETP", "FSETP", "PLOP3"): srcs = parts[2:] if len(parts) > 2 else []
    w = 0
    for a in srcs:
        for r in re.findall(r"(?<![U\w])R(\d+|Z)((?:\.\w+)*)", a):
            if r[0] == "Z": continue
            mod = r[1]
            wide = (".F32x2" in mod or ".64" in mod or (base in ("LDS", "LDG", "STS", "STG", "LDGSTS") and ".64" in a))
            w += 2 if wide else 1
        if base in STORE and a is parts[-1]:   # store data width
            m = re.search(r"\.(64|128)", op)
            if m: w += {"64": 1, "128": 3}[m.group(1)]
    if base in ("FFMA2", "FMUL2", "FADD2"):   # operands without a .F32 marker are register pairs
        w = 0
        for a in srcs:
            if re.search(r"(?<![U\w])R\d+", a): w += 1 if ".F32" in a and "x2" not in a else 2
    return base, w
for k, v in ins.items():
    if not rx.search(k): continue
    blk = []
    def flush():
        if len(blk) >= minb:
            tot = collections.Counter(); cnt = collections.Counter()
            for b in blk:
                base, w = words(b); tot[base] += w; cnt[base] += 1
            W = sum(tot.values())
            print(f"block n={len(blk)}: {W} source words -> {W / 2:.0f} clk of operand delivery; MUFU {cnt['MUFU']} x 8 = {cnt['MUFU'] * 8} clk")
            print("   words by opcode:", ", ".join(f"{o}={tot[o]}({cnt[o]})" for o, _ in tot.most_common(14)))
    for t in v:
        o = t.split()[1] if t.startswith("@") else t.split()[0]
        if o.split(".")[0] in ("BRA", "EXIT", "BSYNC", "BSSY", "WARPSYNC", "CALL", "RET", "NANOSLEEP"):
            flush(); blk = []
        else:
            blk.append(t)
    flush()
