"""Key metrics of every launch in an ncu report dump: ncu -i rep.ncu-rep --page raw --csv > raw.csv;
python tools/ncu_metrics_summary.py raw.csv"""
import csv
import sys

WANT = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block", "smem/block"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM by smem"), ("launch__occupancy_limit_registers", "CTAs/SM by regs"),
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 (non-tensor) pipe active %"),
    ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "shared FP64 pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct", "issue active %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for n, r in enumerate(rows[2:]):
    print("launch %d" % n)
    for key, label in WANT:
        if key in col:
            print("  %-36s %s %s" % (label, r[col[key]], units[col[key]]))
