#!/bin/bash
# round 2, GPU run AG: default bench after the roofline selection change
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2ag_bench_default.json 2> gpurun_out/r2ag_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ag_bench_default.json"))
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d.get("parity"))
print(json.dumps(d["roofline"])[:1500])
PY
