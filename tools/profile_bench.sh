#!/bin/bash
# ncu evidence for the headline bench (run on the GPU box through gpurun, one call):
#   gpurun --timeout 1500 -- 'bash tools/profile_bench.sh r02'
# 1. plain run (must exit 0), 2. launch list of the same command (cold-cache, serialised: compare SHARES),
# 3. `--set full` of one forward + one backtrace launch of a timed step.  Post-process here with
#    python tools/ncu_launch_summary.py / tools/measure_traffic.py (no GPU needed).
set -u
TAG=${1:-r02}
CMD="python bench.py --no-cpu --no-other --steps 2 --warmup 3"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench_pos.csv \
    $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'decode_small_fwd|backtrace_small' -s 8 -c 2 \
    -o gpurun_out/${TAG}_fwd_bt $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -8
