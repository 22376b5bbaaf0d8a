"""Per-launch DRAM traffic of the scan kernels from an `ncu --set full` report -> profiles/traffic.json (read by bench.py).
usage: python tools/ncu_traffic.py report.ncu-rep workload batch_per_gpu [source-note]"""
import csv, io, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_source_hash
rep, wl, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
note = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(rep)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
iK, iR, iW = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
units = rows[1]
def to_bytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
out = {}
for r in rows[2:]:
    name = "selscan_fwd" if "selscan_fwd" in r[iK] else "selscan_bwd" if ("selscan_bwd_chain" in r[iK] or "selscan_bwd_fast" in r[iK]) else None
    if name is None or name in out:
        continue
    out[name] = int(to_bytes(r[iR], units[iR]) + to_bytes(r[iW], units[iW]))
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
try:
    allrec = json.load(open(path))
except Exception:
    allrec = {}
allrec[wl] = {"batch_per_gpu": B, "kernels": out, "source": note, "source_hash": kernel_source_hash(),
              "metric": "dram__bytes_read.sum + dram__bytes_write.sum per launch"}
json.dump(allrec, open(path, "w"), indent=1)
print(allrec[wl])
