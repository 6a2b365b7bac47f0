#!/bin/bash
# round 2, GPU run R: two issuing threads in the halo kernel, three in the resident-weights kernel: suite, timers, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2r_pytest.log
OCRS_B200_CONV_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01 > gpurun_out/r2r_dbg.json 2> gpurun_out/r2r_dbg.err; grep -m5 "conv dbg" gpurun_out/r2r_dbg.err
timeout -s ABRT 200 python -X faulthandler bench.py --steps 20 --warmup 3 --in-flight 3 --no-cpu-baseline > gpurun_out/r2r_bench_if3.json 2> gpurun_out/r2r_bench_if3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2r_bench_if3.json"))
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["op_ms_per_step"], d["stage_ms_per_step"], d.get("parity"))
PY
