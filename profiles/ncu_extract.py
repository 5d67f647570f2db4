#!/usr/bin/env python
"""Extract the per-launch summary the roofline discussion uses from `ncu -i X.ncu-rep --page raw --csv`.

    ncu -i gpurun_out/r2_full_c5tile.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/ncu_extract.py /tmp/raw.csv profiles/r2/r2_ncu_full_c5tile_summary.csv
"""
import csv
import sys

WANT = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
cols = [w for w in WANT if w in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[hdr.index(c)] for c in cols])
    for r in rows[2:]:
        w.writerow([r[hdr.index(c)][:60] if c == "Kernel Name" else r[hdr.index(c)] for c in cols])
print("wrote", sys.argv[2], len(rows) - 2, "launches")
