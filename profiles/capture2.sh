#!/bin/bash
# Round-2 profile capture, run on the GPU box:   gpurun --timeout 1500 -- 'bash profiles/capture2.sh r2a'
# then here:  python profiles/summarize.py r2a ...
# Per /opt/skills/guides/B200_PROFILING.md: every command first runs plain (must exit 0), then its launch list, then ONE
# --set full capture restricted to our kernels.  Reports stay in /tmp on the box; only csv exports come back.
#   default : bench.py default workload (256^3 ioniz_sphere, unfused correct/final pair + the separate ion kernels)
#   fused   : the same with AB2_HYDRO_FUSE=1 (k_cfinal) and AB2_ION_FUSE=1 (k_ion_rates<fused update>)
#   smr     : --workload ioniz_smr (the four SMR kernels)
#   ppm     : --workload blast --order 3 (the three PPM kernels)
TAG=${1:-cap}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # name, skip, count, env..., -- command
  local name=$1 skip=$2 count=$3; shift 3
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" "$@" > gpurun_out/${TAG}_${name}_bench.json 2> gpurun_out/${TAG}_${name}_bench.err || { echo "$name: plain run failed"; return; }
  env "${envs[@]}" ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_${name}_launches.csv "$@" > /tmp/ncu_launch_$name.log 2>&1
  env "${envs[@]}" ncu --set full --clock-control none --import-source on -k regex:'^k_' -s $skip -c $count -f -o /tmp/${TAG}_${name}_full "$@" > /tmp/ncu_full_$name.log 2>&1
  ncu -i /tmp/${TAG}_${name}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_ncu_full_raw.csv
}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-parity-gate"
cap default 200 60 AB2_HYDRO_FUSE=0 -- $B
cap fused 120 40 AB2_HYDRO_FUSE=1 AB2_ION_FUSE=1 -- $B
cap ppm 60 40 AB2_HYDRO_FUSE=0 -- $B --workload blast --order 3
cap smr 700 260 AB2_HYDRO_FUSE=0 -- $B --workload ioniz_smr
ls -la gpurun_out/${TAG}_*
