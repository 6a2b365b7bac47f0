#!/bin/bash
# round 2, GPU run L: full GPU suite + bench on the make-built library
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2l_pytest.log
timeout 400 python bench.py --steps 20 --warmup 3 --in-flight 3 --no-cpu-baseline > gpurun_out/r2l_bench_if3.json 2> gpurun_out/r2l_bench_if3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2l_bench_if3.json"))
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["op_ms_per_step"], d["stage_ms_per_step"], d.get("parity"))
PY
