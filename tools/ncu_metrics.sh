#!/bin/bash
# targeted counters of the scan kernel launches of one probe run on config 2 (pre-seed, seed, full pass).  usage: bash tools/ncu_metrics.sh TAG
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:rr_k_scan_umma -s 3 -c 3 --csv --log-file gpurun_out/${TAG}_scan_metrics.csv \
    python tools/probe_umma.py Tree_1perc_30000 0 umma_mxf4 2 > gpurun_out/${TAG}_ncu_metrics.log 2>&1
echo "ncu metrics rc=$?"
