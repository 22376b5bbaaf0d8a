"""Summarise registers / spills per kernel from the nvcc -Xptxas -v logs under gfe_mamba_b200/build/."""
import glob, re, subprocess, sys
for log in sorted(glob.glob("gfe_mamba_b200/build/*.o.log")):
    txt = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(gfe::\w+\)|void |gfe::", "", name)
        if len(sys.argv) > 1 and sys.argv[1] not in name:
            continue
        print(f"{int(m.group(5)):4d} regs  spill st/ld {int(m.group(3)):4d}/{int(m.group(4)):4d}  {name}")
