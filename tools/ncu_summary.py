"""Summarise an .ncu-rep (one or more kernels) into the handful of counters DESIGN.md / bench.py argue from.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.max"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"== {name[:110]}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:82s} {vals[i]:>18s} {units[i]}")
    stalls = [(h, float(vals[i].replace(',', ''))) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_warp_active.pct")
              and vals[i].replace('.', '').replace(',', '').isdigit()]
    for h, v in sorted(stalls, key=lambda t: -t[1])[:5]:
        print(f"   stall {h.split('issue_stalled_')[1].split('_per_warp')[0]:74s} {v:18.2f} %")
