#!/bin/bash
# round 2, GPU run AJ (2 GPUs): the driver's torchrun launch on the final tree, both arms
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2aj_bench_2gpu.json 2> gpurun_out/r2aj_bench_2gpu.err; echo "2gpu rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2aj_bench_2gpu_reference.json 2> gpurun_out/r2aj_bench_2gpu_reference.err; echo "2gpu reference rc=$?"
python - <<'PY'
import json
for n in ("2gpu","2gpu_reference"):
    try:
        t=open(f"gpurun_out/r2aj_bench_{n}.json").read()
        d=json.loads(t[t.index('{"'):].splitlines()[0])
        print(n, d.get("n_gpus"), round(d["value"],2), d.get("e2e",{}).get("value"), d.get("ms_per_step"), d.get("parity"), (d.get("roofline") or {}).get("kernel"), d.get("impl"))
    except Exception as e: print(n, "failed", e)
PY
