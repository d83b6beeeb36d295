#!/bin/bash
# Round-2h profile capture of the sparse ion sub-cycle (run on the GPU box:  gpurun --timeout 900 -- 'bash profiles/capture4.sh r2h'
# then here:  NCU_TRAFFIC_MERGE=override python profiles/summarize.py r2h multi ion).  Plain run first (must exit 0), then the
# launch list of the same command, then ONE --set full capture restricted to the ion kernels of one timed step.
TAG=${1:-r2h}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --e2e-members 1 --no-parity-gate"
$B > gpurun_out/${TAG}_ion_bench.json 2> gpurun_out/${TAG}_ion_bench.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_ion_launches.csv $B > /tmp/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^k_ion' -s 45 -c 45 -f -o /tmp/${TAG}_ion_full $B > /tmp/ncu_full.log 2>&1
ncu -i /tmp/${TAG}_ion_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ion_ncu_full_raw.csv
ls -la gpurun_out/${TAG}_*
