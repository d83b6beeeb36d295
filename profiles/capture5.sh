#!/bin/bash
# Round-2m supplementary capture: k_ion_rates[first] (the first rates pass with the call's prologue folded in) and the full
# k_ion_update behind it.   gpurun --timeout 600 -- 'bash profiles/capture5.sh r2m'
#   then:  NCU_TRAFFIC_MERGE=override python profiles/summarize.py r2m multi first
TAG=${1:-r2m}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --e2e-members 1 --no-parity-gate"
$B > gpurun_out/${TAG}_first_bench.json 2> gpurun_out/${TAG}_first_bench.err || { echo "plain run failed"; exit 1; }
# launches of a step: [first], update, then walk / rates[memo] / update[memo] per further iteration; steps 2.. have ~5 iterations
ncu --set full --clock-control none --import-source on -k regex:'^k_ion_(rates|update)' -s 40 -c 30 -f -o /tmp/${TAG}_first_full $B > /tmp/ncu_full.log 2>&1
ncu -i /tmp/${TAG}_first_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_first_ncu_full_raw.csv
ls -la gpurun_out/${TAG}_*
