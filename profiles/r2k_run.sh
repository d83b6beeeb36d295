#!/bin/bash
# Round-2k: 4 GPUs -- weak scaling (one 256^3 slab per GPU) and 768^3 strong, both behind the parity gate.
#   gpurun --gpus 4 --timeout 900 -- 'bash profiles/r2k_run.sh r2k'
TAG=${1:-r2k}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $T --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/${TAG}_weak_n4.json 2> gpurun_out/${TAG}_weak_n4.err; echo "weak n4 rc=$?"
timeout 500 $T --master-port 29522 bench.py --gpus 4 --strong --size 768 --steps 10 --warmup 5 > gpurun_out/${TAG}_strong768_n4.json 2> gpurun_out/${TAG}_strong768_n4.err; echo "strong768 n4 rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], round(d["ms_per_step"], 3), "%.4g" % d["value"], d["parity"]["bit_equal"], d["config"].get("ion_iterations_per_step"), d["e2e"]["value"], (d.get("roofline") or {}).get("ion_memo"))
PY
