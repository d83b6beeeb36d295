#!/bin/bash
# Round-2 supplementary capture (run on the GPU box: gpurun --timeout 900 -- 'bash profiles/capture3.sh r2b'):
# the kernels capture2.sh did not reach -- the four SMR operators, k_correct3 (the default pass C since r2) and k_update with
# the folded d^{n+1/2} -- from the shipped 5-level deck, plus k_bvals_table from the jet drop-in.
TAG=${1:-cap}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-parity-gate"
$B --workload ioniz_smr > gpurun_out/${TAG}_smr_bench.json 2> gpurun_out/${TAG}_smr_bench.err || echo "smr plain run failed"
ncu --set full --clock-control none --import-source on -k regex:'k_smr|k_correct3|k_update' -s 40 -c 48 -f -o /tmp/${TAG}_smr_full $B --workload ioniz_smr > /tmp/ncu_smr.log 2>&1
ncu -i /tmp/${TAG}_smr_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_smr_ncu_full_raw.csv
$B > gpurun_out/${TAG}_default_bench.json 2> gpurun_out/${TAG}_default_bench.err
ncu --set full --clock-control none --import-source on -k regex:'k_correct3|k_update|k_final' -s 20 -c 6 -f -o /tmp/${TAG}_c3_full $B > /tmp/ncu_c3.log 2>&1
ncu -i /tmp/${TAG}_c3_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_default_ncu_full_raw.csv
ls -la gpurun_out/${TAG}_*
