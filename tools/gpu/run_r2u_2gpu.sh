#!/bin/bash
# round 2, GPU run U (2 GPUs): the torchrun launch the driver uses; blocking vs spinning worker waits; 1-GPU reference
mkdir -p gpurun_out
run2() {  # $1 = tag, env in front
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_2gpu_$1.json 2> gpurun_out/r2u_bench_2gpu_$1.err; echo "2gpu $1 rc=$?"
}
run2 block
OCRS_B200_SPIN_SYNC=1 run2 spin
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_1gpu_block.json 2> gpurun_out/r2u_bench_1gpu_block.err; echo "1gpu block rc=$?"
OCRS_B200_SPIN_SYNC=1 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_1gpu_spin.json 2> gpurun_out/r2u_bench_1gpu_spin.err; echo "1gpu spin rc=$?"
CUDA_VISIBLE_DEVICES=1 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_1gpu_dev1.json 2> gpurun_out/r2u_bench_1gpu_dev1.err; echo "1gpu dev1 rc=$?"
python - <<'PY'
import json
for n in ("2gpu_block","2gpu_spin","1gpu_block","1gpu_spin","1gpu_dev1"):
    try:
        t=open(f"gpurun_out/r2u_bench_{n}.json").read()
        d=json.loads(t[t.index('{"'):].splitlines()[0])
        print(n, d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d.get("parity"), d.get("clocks"))
    except Exception as e: print(n, "failed", e)
PY
