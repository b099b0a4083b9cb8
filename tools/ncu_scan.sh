#!/bin/bash
# one `ncu --set full` capture of the full pass of the fused scan kernel on config 2 (after the same command has run without
# ncu), plus the launch list of a bench step.  usage (under gpurun): bash tools/ncu_scan.sh TAG
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
python tools/probe_umma.py Tree_1perc_30000 0 umma_mxf4 1 > gpurun_out/${TAG}_probe_plain.log 2>&1 || exit 1
tail -1 gpurun_out/${TAG}_probe_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rr_k_scan_umma -s 2 -c 1 -f -o gpurun_out/${TAG}_scan_full \
    python tools/probe_umma.py Tree_1perc_30000 0 umma_mxf4 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu rc=$?"
