"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py file.ncu-rep [metric-prefix ...]"""
import csv
import subprocess
import sys

DEFAULT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active",
           "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__inst_executed_pipe_xu", "sm__cycles_active.avg", "launch__grid_size", "launch__block_size",
           "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum", "smsp__inst_executed.sum",
           "launch__occupancy_limit", "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active",
           "sm__inst_executed_pipe_xu", "smsp__issue_active.avg.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
           "sm__pipe_shared_cycles_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
rep = sys.argv[1]
want = sys.argv[2:] or DEFAULT
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = [i for i, h in enumerate(hdr) if any(h.startswith(w) for w in want)]
for row in rows[2:]:
    print("---")
    for i in idx:
        print(f"  {hdr[i]} = {row[i]} {units[i]}")
