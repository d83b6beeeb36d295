#!/bin/bash
# Round-2c (sparse ion sub-cycle): parity of the memo against the reference's fingerprints, then the 256^3 bench with the
# memo on / off and register-cap variants of the two ion kernels, back to back on one box.
#   gpurun --timeout 900 -- 'bash profiles/r2c_memo_run.sh r2c'
TAG=${1:-r2c}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_longrun.py -x -q -k "sparse" > gpurun_out/${TAG}_pytest_sparse.log 2>&1; echo "pytest sparse rc=$?" | tee -a gpurun_out/${TAG}_pytest_sparse.log
tail -3 gpurun_out/${TAG}_pytest_sparse.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "ion" > gpurun_out/${TAG}_pytest_ion.log 2>&1; echo "pytest ion rc=$?" | tee -a gpurun_out/${TAG}_pytest_ion.log
tail -3 gpurun_out/${TAG}_pytest_ion.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2"
timeout 300 $B > gpurun_out/${TAG}_bench_memo.json 2> gpurun_out/${TAG}_bench_memo.err; echo "bench memo rc=$?"
AB2_ION_SKIP=0 timeout 300 $B --no-parity-gate > gpurun_out/${TAG}_bench_nomemo.json 2> gpurun_out/${TAG}_bench_nomemo.err; echo "bench nomemo rc=$?"
for v in u10 u8 r6u10; do
  L=$PWD/atmospheric-athena_b200/csrc/libv_$v.so
  [ -f $L ] || continue
  AB2_LIB=$L timeout 300 $B --no-parity-gate > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err; echo "bench $v rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    k = {n: round(v["ms_total"] / d["steps"], 3) for n, v in d["kernels"].items() if "ion" in n}
    print(f.split("/")[-1], round(d["ms_per_step"], 3), d["value"], k, (d.get("roofline") or {}).get("ion_memo"))
PY
