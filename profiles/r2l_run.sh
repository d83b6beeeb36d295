#!/bin/bash
# Round-2l: the call's prologue folded into the first rates pass (k_ion_rates[first]); whole GPU suite, smoke, then the
# bench with the fold on / off (AB2_ION_PREP=1: separate k_ion_prep pass).   gpurun --timeout 1500 -- 'bash profiles/r2l_run.sh r2l'
TAG=${1:-r2l}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
B="python bench.py --no-cpu-baseline --e2e-steps 1 --e2e-members 1 --steps 20 --warmup 5"
timeout 300 $B > gpurun_out/${TAG}_bench_fold.json 2> gpurun_out/${TAG}_bench_fold.err; echo "bench fold rc=$?"
AB2_ION_PREP=1 timeout 300 $B --no-parity-gate > gpurun_out/${TAG}_bench_prep.json 2> gpurun_out/${TAG}_bench_prep.err; echo "bench prep rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    k = {n: (round(v["ms_total"] / v["launches"], 4), v["launches"]) for n, v in d.get("kernels", {}).items() if "ion" in n}
    print(f.split("/")[-1], round(d["ms_per_step"], 3), "%.4g" % d["value"], (d.get("parity") or {}).get("bit_equal"), k)
PY
