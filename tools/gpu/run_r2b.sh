#!/bin/bash
# round 2, GPU run B: GRU v3 + pool parity, phase timers (GRU v3, conv both issue orders), bench through the pool
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2b_pytest.log
OCRS_B200_CONV_DEBUG=1 OCRS_B200_GRU_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01 > gpurun_out/r2b_dbg.json 2> gpurun_out/r2b_dbg.err; echo "dbg rc=$?"
OCRS_B200_CONV_HH_FIRST=1 OCRS_B200_CONV_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01 > gpurun_out/r2b_dbg_hhfirst.json 2> gpurun_out/r2b_dbg_hhfirst.err; echo "dbg2 rc=$?"
grep -m5 "conv dbg" gpurun_out/r2b_dbg.err; grep -m2 "gru dbg" gpurun_out/r2b_dbg.err; grep -m5 "conv dbg" gpurun_out/r2b_dbg_hhfirst.err
for f in 2 3; do
  timeout 400 python bench.py --steps 20 --warmup 3 --in-flight $f --no-cpu-baseline > gpurun_out/r2b_bench_if$f.json 2> gpurun_out/r2b_bench_if$f.err; echo "bench if$f rc=$?"
done
OCRS_B200_GRU_V2=1 timeout 400 python bench.py --steps 20 --warmup 3 --in-flight 2 --no-cpu-baseline > gpurun_out/r2b_bench_if2_gruv2.json 2> gpurun_out/r2b_bench_if2_gruv2.err; echo "bench v2 rc=$?"
python - <<'PY'
import json
for n in ["if2","if3","if2_gruv2"]:
    try:
        d=json.load(open(f"gpurun_out/r2b_bench_{n}.json"))
        print(n, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["region_ms"], d["op_ms_per_step"], d["parity"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2b_bench_{n}.err").read()[-1500:])
PY
