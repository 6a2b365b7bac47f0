#!/bin/bash
# round 2, GPU run AL: batches in flight on the final tree
mkdir -p gpurun_out
for IF in 10 12 14; do
timeout -s ABRT 120 python -X faulthandler bench.py --steps 24 --warmup 3 --in-flight $IF --no-cpu-baseline --min-seconds 1.0 > gpurun_out/r2al_bench_if$IF.json 2> gpurun_out/r2al_bench_if$IF.err; echo "bench if$IF rc=$?"
done
python - <<'PY'
import json
for i in (10,12,14):
    try:
        d=json.load(open(f"gpurun_out/r2al_bench_if{i}.json"))
        print(i, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d.get("parity"), d["clocks"]["sm_mhz"])
    except Exception as e: print(i, "failed", e)
PY
