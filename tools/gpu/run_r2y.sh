#!/bin/bash
# round 2, GPU run Y: GEMM with eight epilogue warps; register cap of the recurrent kernel x batches in flight
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2y_pytest.log
for R in 96 80 72 64; do for IF in 3 4; do
OCRS_B200_GRU_REGS=$R timeout -s ABRT 200 python -X faulthandler bench.py --steps 20 --warmup 3 --in-flight $IF --no-cpu-baseline > gpurun_out/r2y_bench_r${R}_if$IF.json 2> gpurun_out/r2y_bench_r${R}_if$IF.err; echo "bench regs $R if$IF rc=$?"
done; done
python - <<'PY'
import json
for r in (96,80,72,64):
  for i in (3,4):
    try:
        d=json.load(open(f"gpurun_out/r2y_bench_r{r}_if{i}.json"))
        o=d["op_ms_per_step"]
        print(r, i, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), "rec", o.get("rec/GRU recurrence(packed)"), "proj", o.get("rec/GRU input projection(packed)"), "conv", o.get("rec/conv3x3_tc_kernel"), d.get("parity"))
    except Exception as e: print(r, i, "failed", e)
PY
