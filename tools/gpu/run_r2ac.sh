#!/bin/bash
# round 2, GPU run AC: two-pixel-per-thread dwpw kernel: GPU suite, bench against the one-pixel form
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ac_pytest.log
timeout -s ABRT 240 python -X faulthandler bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/r2ac_bench_two.json 2> gpurun_out/r2ac_bench_two.err; echo "bench two rc=$?"
OCRS_B200_DWPW_ONE_PIXEL=1 timeout -s ABRT 240 python -X faulthandler bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/r2ac_bench_one.json 2> gpurun_out/r2ac_bench_one.err; echo "bench one rc=$?"
python - <<'PY'
import json
for n in ("two","one"):
    try:
        d=json.load(open(f"gpurun_out/r2ac_bench_{n}.json"))
        print(n, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["op_ms_per_step"].get("det/Conv"), d["stage_ms_per_step"].get("stage/det_net"), d.get("parity"))
    except Exception as e: print(n, "failed", e)
PY
