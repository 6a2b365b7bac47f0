#!/bin/bash
# round 2, GPU run Z: batches in flight per GPU
mkdir -p gpurun_out
for IF in 6 8 10 12; do
timeout -s ABRT 240 python -X faulthandler bench.py --steps 32 --warmup 3 --in-flight $IF --no-cpu-baseline > gpurun_out/r2z2_bench_if$IF.json 2> gpurun_out/r2z2_bench_if$IF.err; echo "bench if$IF rc=$?"
done
python - <<'PY'
import json
for i in (6,8,10,12):
    try:
        d=json.load(open(f"gpurun_out/r2z2_bench_if{i}.json"))
        print(i, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["e2e"].get("ms_per_step"), d.get("parity"), d.get("clocks"))
    except Exception as e: print(i, "failed", e)
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
