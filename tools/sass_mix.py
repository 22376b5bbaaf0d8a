"""Static SASS opcode mix of one kernel.  usage: python tools/sass_mix.py obj_or_so kernel_regex [min_block]
Prints the opcode histogram of the whole kernel and of every straight-line region (no labels / branches inside)
with at least min_block instructions -- the unrolled chunk bodies of the scan kernels."""
import collections, re, subprocess, sys
obj, rx = sys.argv[1], re.compile(sys.argv[2])
minb = int(sys.argv[3]) if len(sys.argv) > 3 else 300
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
kern, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kern[cur] = []; continue
    if cur is None: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        kern[cur].append(m.group(2).strip())
def op(ins):
    t = ins.split()
    if t[0].startswith("@"): t = t[1:]
    return t[0]
def klass(o):
    b = o.split(".")[0]
    if b in ("FFMA2", "FMUL2", "FADD2"): return "fma2"
    if b in ("FFMA", "FMUL", "FADD"): return "fma1"
    if b == "MUFU": return "mufu"
    if b in ("LDS", "STS", "LDG", "STG", "LDGSTS", "LDSM", "ATOMG", "RED", "LD", "ST"): return b.lower()
    if b in ("SHFL",): return "shfl"
    if b in ("MOV", "IMAD", "IADD3", "IADD", "LEA", "SHF", "LOP3", "PRMT", "SEL", "FSEL", "FMNMX", "ISETP", "FSETP", "I2F", "F2F", "F2FP", "VOTE", "IMNMX", "UMOV", "UIADD3", "ULEA", "USHF", "UIMAD", "ULOP3", "S2R", "CS2R", "R2UR", "PLOP3", "FSET", "F2I", "I2FP", "VIADD", "VIMNMX", "UISETP", "USEL", "UPRMT", "FMNMX3", "FCHK"): return "alu:" + b
    return "other:" + b
for name, ins in kern.items():
    if not rx.search(name): continue
    print("==", name, len(ins), "instructions")
    def hist(lst, title):
        h = collections.Counter(klass(op(i)) for i in lst)
        tot = len(lst)
        grp = collections.Counter()
        for k, v in h.items(): grp[k.split(":")[0]] += v
        print(f"  [{title}] n={tot}  " + "  ".join(f"{k}={v}" for k, v in sorted(grp.items(), key=lambda kv: -kv[1])))
        det = {k: v for k, v in h.items() if ":" in k}
        print("      " + "  ".join(f"{k.split(':')[1]}={v}" for k, v in sorted(det.items(), key=lambda kv: -kv[1])))
    hist(ins, "kernel")
    blk = []
    for i in ins + ["BRA end"]:
        o = op(i).split(".")[0]
        if o in ("BRA", "EXIT", "RET", "BSYNC", "BSSY", "CALL", "WARPSYNC", "BRX", "JMP", "NANOSLEEP", "YIELD") :
            if len(blk) >= minb: hist(blk, "block")
            blk = []
        else:
            blk.append(i)
