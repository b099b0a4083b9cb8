#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's last GPU call and has never run on a GPU.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/validate_experimental.sh'
# Writes gpurun_out/exp_*.log|json.  Nothing here is a bench value.
#   1. the ungated Relative_Vars test (composition of validated device steps) and the real-pipeline golden case
#   2. the opt-in tests: Cliquer count kernel 3, the tiled Relative_Vars kernel, the Kmeans sweeps
#   3. wall times of Relative_Vars (both paths, which must agree) and Kmeans on a generated MSA
#   4. timing of the three Cliquer count kernels side by side, and one ncu --set full launch of each
set -u
mkdir -p gpurun_out
python -m pytest tests/test_zz_gpu_relvars.py -m gpu -q -p no:cacheprovider > gpurun_out/exp_relvars_default.log 2>&1
echo "relvars default rc=$?" | tee -a gpurun_out/exp_relvars_default.log
python -m pytest tests/test_zz_gpu_real_pipeline.py -m gpu -q -p no:cacheprovider > gpurun_out/exp_real_pipeline.log 2>&1
echo "real pipeline golden rc=$?" | tee -a gpurun_out/exp_real_pipeline.log
RR_TEST_UNVALIDATED=1 python -m pytest tests/test_zz_gpu_cliquer.py tests/test_zz_gpu_relvars.py tests/test_zz_gpu_kmeans.py -m gpu -q \
    -p no:cacheprovider > gpurun_out/exp_unvalidated.log 2>&1
echo "unvalidated kernels rc=$?" | tee -a gpurun_out/exp_unvalidated.log
tail -5 gpurun_out/exp_unvalidated.log
timeout 300 python tools/probe_rows.py 40 4000 gpurun_out/exp_rows_probe.json | cut -c1-900
RR_TEST_UNVALIDATED=1 python tools/probe_cliquer.py 100 6000 1024 gpurun_out/exp_cliquer_probe.json 3 | cut -c1-600
RR_TEST_UNVALIDATED=1 timeout 120 ncu --set full --clock-control none --import-source on -k regex:rr_k_cliquer_counts -c 3 -f \
    -o gpurun_out/exp_prof_cliquer python tools/probe_cliquer.py 100 6000 1024 /dev/null 1 > gpurun_out/exp_ncu_cliquer.log 2>&1
echo "ncu rc=$?"
