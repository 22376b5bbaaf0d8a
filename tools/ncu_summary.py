"""Key metrics per kernel from an ncu report.  usage: python tools/ncu_summary.py X.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("----")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:72s} {r[i]} {units[i]}")
    st = sorted(((float(r[hdr.index(h)]), h) for h in stalls), reverse=True)[:7]
    print("  top stalls (warps per issue):", ", ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}={v:.2f}" for v, h in st))
