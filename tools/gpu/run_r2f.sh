#!/bin/bash
# round 2, GPU run F: CTA-pair conv kernel parity + timing, promotion-side timers, full test suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2f_pytest.log
OCRS_B200_CONV_PAIR=1 timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_configs.py tests/test_gpu_e2e.py tests/test_gpu_recognition.py tests/test_gpu_goldens.py -m gpu -x -q > gpurun_out/r2f_pytest_pair.log 2>&1; echo "pytest pair rc=$?"
tail -3 gpurun_out/r2f_pytest_pair.log
OCRS_B200_CONV_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01 > gpurun_out/r2f_dbg.json 2> gpurun_out/r2f_dbg.err; echo "dbg rc=$?"
grep -m5 "conv dbg" gpurun_out/r2f_dbg.err
OCRS_B200_CONV_PAIR=1 OCRS_B200_CONV_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01 > gpurun_out/r2f_dbg_pair.json 2> gpurun_out/r2f_dbg_pair.err; echo "dbg pair rc=$?"
grep -m5 "conv dbg" gpurun_out/r2f_dbg_pair.err
timeout 400 python bench.py --steps 20 --warmup 3 --in-flight 3 --no-cpu-baseline > gpurun_out/r2f_bench_if3.json 2> gpurun_out/r2f_bench_if3.err; echo "bench if3 rc=$?"
OCRS_B200_CONV_PAIR=1 timeout 400 python bench.py --steps 20 --warmup 3 --in-flight 3 --no-cpu-baseline > gpurun_out/r2f_bench_if3_pair.json 2> gpurun_out/r2f_bench_if3_pair.err; echo "bench pair rc=$?"
python - <<'PY'
import json
for n in ["if3","if3_pair"]:
    try:
        d=json.load(open(f"gpurun_out/r2f_bench_{n}.json"))
        print(n, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["op_ms_per_step"], d["stage_ms_per_step"], d["parity"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2f_bench_{n}.err").read()[-1500:])
PY
