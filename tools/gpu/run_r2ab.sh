#!/bin/bash
# round 2, GPU run AB: GPU suite + default bench on the final library (blocking worker waits, ten batches in flight)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ab_pytest.log
timeout 900 python bench.py > gpurun_out/r2ab_bench_default.json 2> gpurun_out/r2ab_bench_default.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ab_bench_default.json"))
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["config"]["batches_in_flight"], d.get("parity"), d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["cpu_baseline"]["value"])
PY
