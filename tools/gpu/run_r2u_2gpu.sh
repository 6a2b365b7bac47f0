#!/bin/bash
# round 2, GPU run U (2 GPUs): the torchrun launch the driver uses, both arms
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2u_bench_2gpu.json 2> gpurun_out/r2u_bench_2gpu.err; echo "2gpu rc=$?"
tail -c 600 gpurun_out/r2u_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2u_bench_2gpu.json"))
print(d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d.get("parity"), d.get("clocks"))
PY
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2u_bench_1gpu.json 2> gpurun_out/r2u_bench_1gpu.err; echo "1gpu rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2u_bench_1gpu.json"))
print(d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d.get("parity"), d.get("cpu_baseline"), d.get("roofline"))
PY
