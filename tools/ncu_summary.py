"""Reduce `ncu -i X.ncu-rep --page raw --csv` of one kernel launch to the counters DESIGN.md / profiles/README.md quote:
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/NAME_ncu_full_summary.csv"""
import csv
import sys

KEEP = """ID,Kernel Name,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,launch__block_size,launch__grid_size,
launch__cluster_dim_x,launch__registers_per_thread,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,
sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active""".replace("\n", "").split(",")
rows = [r for r in csv.reader(sys.stdin) if r]
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units, vals = rows[h], rows[h + 1], rows[h + 2]
out = csv.writer(sys.stdout)
out.writerow(["metric", "unit", "value"])
for n, u, v in zip(names, units, vals):
    if n in KEEP or n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
        out.writerow([n, u, v])
