#!/bin/bash
# Round-2g: two GPUs -- the multi-GPU tests (x3 slabs through NCCL inside the library, C driver on 2 ranks, SMR slabs) and the
# weak-scaling bench with its parity gate, on the sparse sub-cycle.   gpurun --gpus 2 --timeout 1500 -- 'bash profiles/r2g_run.sh r2g'
TAG=${1:-r2g}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/${TAG}_pytest_multi.log
tail -5 gpurun_out/${TAG}_pytest_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench_n2.json | head -c 600; echo
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${TAG}_bench_n2.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["parity"], d["e2e"], (d.get("roofline") or {}).get("ion_memo"))
PY
