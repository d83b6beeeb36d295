#!/bin/bash
# Round-2i: the sparse sub-cycle with re-use inside listed rays (chunks behind the ionisation front) and the dt-bounded inert test:
# parity (fingerprints, small goldens through the pair, SMR goldens, five-level deck), then the unigrid and SMR benches.
#   gpurun --timeout 1200 -- 'bash profiles/r2i_run.sh r2i'
TAG=${1:-r2i}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_longrun.py tests/test_gpu_parity.py tests/test_gpu_smr.py tests/test_gpu_dropin.py -q -k "sparse or ion or five_level or fingerprint" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
B="python bench.py --no-cpu-baseline --e2e-steps 1 --e2e-members 1"
timeout 300 $B --steps 20 --warmup 5 > gpurun_out/${TAG}_bench256.json 2> gpurun_out/${TAG}_bench256.err; echo "bench256 rc=$?"
timeout 300 $B --workload ioniz_smr --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_smr.json 2> gpurun_out/${TAG}_bench_smr.err; echo "bench smr rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    k = {n: (round(v["ms_total"] / v["launches"], 4), v["launches"]) for n, v in d.get("kernels", {}).items() if "ion" in n}
    print(f.split("/")[-1], round(d["ms_per_step"], 3), "%.4g" % d["value"], d["parity"]["bit_equal"], k, (d.get("roofline") or {}).get("ion_memo"))
PY
