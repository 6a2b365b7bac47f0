#!/bin/bash
# round 2, GPU run AF (4 GPUs): the torchrun launch the driver uses, N = 4
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2af_bench_4gpu.json 2> gpurun_out/r2af_bench_4gpu.err; echo "4gpu rc=$?"
python - <<'PY'
import json
t=open("gpurun_out/r2af_bench_4gpu.json").read()
d=json.loads(t[t.index('{"'):].splitlines()[0])
print(d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d.get("parity"), d.get("clocks"), d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["roofline"].get("latency_bound_peer",{}).get("kernel"))
PY
nproc; lscpu | grep -E "NUMA node|Socket" | head
