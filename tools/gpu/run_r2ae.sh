#!/bin/bash
# round 2, GPU run AE: final validation of the committed tree: GPU suite, smoke, default bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ae_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2ae_bench_default.json 2> gpurun_out/r2ae_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ae_bench_reference.json 2> gpurun_out/r2ae_bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ae_bench_default.json"))
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["config"]["batches_in_flight"], d.get("parity"), d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["roofline_by_op"], d["cpu_baseline"]["value"], d["clocks"])
r=json.load(open("gpurun_out/r2ae_bench_reference.json"))
print("reference", r["value"], r["unit"])
PY
