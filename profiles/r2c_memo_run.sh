#!/bin/bash
# Round-2c (sparse ion sub-cycle): parity of the memo against the reference's fingerprints, then the 256^3 bench with the
# memo on / off and register-cap variants of k_ion_update, back to back on one box.
#   gpurun --timeout 900 -- 'bash profiles/r2c_memo_run.sh r2c'
TAG=${1:-r2c}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_longrun.py -x -q -k "sparse" > gpurun_out/${TAG}_pytest_sparse.log 2>&1; echo "pytest sparse rc=$?" | tee -a gpurun_out/${TAG}_pytest_sparse.log
tail -3 gpurun_out/${TAG}_pytest_sparse.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "ion" > gpurun_out/${TAG}_pytest_ion.log 2>&1; echo "pytest ion rc=$?" | tee -a gpurun_out/${TAG}_pytest_ion.log
tail -3 gpurun_out/${TAG}_pytest_ion.log
B="python bench.py --steps 20 --warmup 5"
timeout 400 $B > gpurun_out/${TAG}_bench_memo.json 2> gpurun_out/${TAG}_bench_memo.err; echo "bench memo rc=$?"
B="$B --no-cpu-baseline --e2e-steps 2 --e2e-members 1 --no-parity-gate"
AB2_ION_SKIP=0 timeout 300 $B > gpurun_out/${TAG}_bench_nomemo.json 2> gpurun_out/${TAG}_bench_nomemo.err; echo "bench nomemo rc=$?"
for v in u10; do
  L=$PWD/atmospheric-athena_b200/csrc/libv_$v.so
  [ -f $L ] || continue
  AB2_LIB=$L timeout 300 $B > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err; echo "bench $v rc=$?"
done
AB2_X=1 timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-parity-gate --e2e-steps 4 --e2e-members 4 > gpurun_out/${TAG}_bench_e2e4.json 2> gpurun_out/${TAG}_bench_e2e4.err; echo "bench e2e4 rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    k = {n: (round(v["ms_total"] / v["launches"], 4), v["launches"]) for n, v in d["kernels"].items() if "ion" in n}
    print(f.split("/")[-1], round(d["ms_per_step"], 3), d["value"], k, (d.get("roofline") or {}).get("ion_memo"), d.get("e2e"))
PY
