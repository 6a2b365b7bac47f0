#!/bin/bash
# round 2, GPU run AH: weights-stationary transposed kernel, N = 128, no promotion: parity tests, bench against the resident-weights kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_configs.py tests/test_gpu_recognition.py tests/test_gpu_goldens.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2ah_pytest.log | cut -c1-300
timeout -s ABRT 240 python -X faulthandler bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/r2ah_bench_ws.json 2> gpurun_out/r2ah_bench_ws.err; echo "bench ws rc=$?"
OCRS_B200_CONV_WS=0 timeout -s ABRT 240 python -X faulthandler bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/r2ah_bench_res.json 2> gpurun_out/r2ah_bench_res.err; echo "bench res rc=$?"
tail -3 gpurun_out/r2ah_bench_ws.err | cut -c1-300
python - <<'PY'
import json
for n in ("ws","res"):
    try:
        d=json.load(open(f"gpurun_out/r2ah_bench_{n}.json"))
        print(n, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["op_ms_per_step"].get("rec/conv3x3_tc_kernel"), d.get("parity"))
    except Exception as e: print(n, "failed", e)
PY
