#!/bin/bash
# the launch list of one default bench run (every kernel with its device time), after the same command has run without ncu.
# usage (under gpurun): bash tools/ncu_launches.sh TAG
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-file --e2e-steps 0 > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-file --e2e-steps 0 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
echo "ncu launches rc=$?"
