#!/bin/bash
# Profile capture for the default bench workload (256^3 ioniz_sphere, one GPU).  Run on the GPU box:
#   gpurun --timeout 1200 -- 'bash profiles/capture.sh r1c'
# then here:  python profiles/summarize.py r1c
# Order follows /opt/skills/guides/B200_PROFILING.md: the plain command first (must exit 0), then the launch list,
# then ONE --set full capture of one main-loop step's kernels.  Reports stay in /tmp on the box (they exceed the
# 64 MiB that comes back); only the csv exports land in gpurun_out/.
set -e
TAG=${1:-cap}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/${TAG}_prof_bench.json 2> gpurun_out/${TAG}_prof_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > /tmp/ncu_launch.log 2>&1
# skip set-up + 3 warm-up steps' worth of launches, then take ~1.3 steps so every kernel of the step is in
ncu --set full --clock-control none --import-source on -k regex:'^k_' -s 190 -c 60 -f -o /tmp/${TAG}_full $CMD > /tmp/ncu_full.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_raw.csv
ls -la gpurun_out/${TAG}_*
