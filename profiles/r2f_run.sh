#!/bin/bash
# Round-2f: the whole GPU suite on the sparse sub-cycle, then where the fused kernel / the pair with the memo pays
# (shipped 5-level SMR deck, 128^3 and 64^3 unigrid).   gpurun --timeout 1500 -- 'bash profiles/r2f_run.sh r2f'
TAG=${1:-r2f}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
B="python bench.py --no-cpu-baseline --e2e-steps 1 --e2e-members 1 --no-parity-gate"
for f in default 0 1; do
  if [ $f = default ]; then unset AB2_ION_FUSE; else export AB2_ION_FUSE=$f; fi
  timeout 300 $B --workload ioniz_smr --steps 10 --warmup 3 > gpurun_out/${TAG}_smr_fuse_$f.json 2> gpurun_out/${TAG}_smr_fuse_$f.err; echo "smr fuse=$f rc=$?"
  for n in 128 64; do
    timeout 200 $B --size $n --steps 20 --warmup 5 > gpurun_out/${TAG}_n${n}_fuse_$f.json 2> gpurun_out/${TAG}_n${n}_fuse_$f.err; echo "n=$n fuse=$f rc=$?"
  done
done
unset AB2_ION_FUSE
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*fuse*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], round(d["ms_per_step"], 3), "%.4g" % d["value"], d["config"].get("ion_iterations_per_step", d["config"].get("ion_iterations_per_level_step")))
PY
