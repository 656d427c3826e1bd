"""Condense an .ncu-rep (read here, no GPU needed) into the few numbers DESIGN.md / profiles/ quote:
    python tools/ncu_summary.py gpurun_out/secondary.ncu-rep > profiles/r01_secondary_ncu.txt"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "local_load_requests", "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum",
]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
print(f"# {rep}: ncu --set full --clock-control none (per-launch, cold cache, serialised)")
for r in body:
    print(f"kernel: {r[col['Kernel Name']]}")
    for kname in KEEP:
        if kname in col and r[col[kname]] != "":
            print(f"  {kname:<72} {r[col[kname]]} {units[col[kname]]}")
    stalls = []
    for h in hdr:
        if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio") is False:
            continue
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[col[h]])
            except ValueError:
                continue
            if v >= 0.3:
                stalls.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    for v, name in sorted(stalls, reverse=True):
        print(f"  STALL {name:<40} {v:.2f} warps per issue-active cycle")
    print()
