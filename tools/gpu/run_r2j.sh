#!/bin/bash
# round 2, GPU run J: the CTA-pair + halo-reuse conv kernel (OCRS_B200_CONV_MODE=2): parity tests, phase timers, bench
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01"
OCRS_B200_CONV_MODE=2 timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_configs.py tests/test_gpu_e2e.py tests/test_gpu_recognition.py tests/test_gpu_goldens.py -m gpu -x -q > gpurun_out/r2j_pytest_halo.log 2>&1; echo "pytest halo rc=$?"; tail -3 gpurun_out/r2j_pytest_halo.log
OCRS_B200_CONV_MODE=2 OCRS_B200_CONV_DEBUG=1 timeout 300 $B > gpurun_out/r2j_dbg_halo.json 2> gpurun_out/r2j_dbg_halo.err; echo "dbg halo rc=$?"; grep -m5 "conv dbg" gpurun_out/r2j_dbg_halo.err
OCRS_B200_CONV_MODE=2 timeout 400 python bench.py --steps 20 --warmup 3 --in-flight 3 --no-cpu-baseline > gpurun_out/r2j_bench_halo.json 2> gpurun_out/r2j_bench_halo.err; echo "bench halo rc=$?"
timeout 400 python bench.py --steps 20 --warmup 3 --in-flight 3 --no-cpu-baseline > gpurun_out/r2j_bench_if3.json 2> gpurun_out/r2j_bench_if3.err; echo "bench rc=$?"
python - <<'PY'
import json
for n in ("halo","if3"):
    try:
        d=json.load(open(f"gpurun_out/r2j_bench_{n}.json"))
        print(n, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["op_ms_per_step"], d.get("parity"))
    except Exception as e: print(n, "failed", e)
PY
