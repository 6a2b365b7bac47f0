#!/bin/bash
# round 2, GPU run AA: register cap of the recurrent kernel with ten batches in flight (GPU saturated)
mkdir -p gpurun_out
for R in 96 80 72 64; do
OCRS_B200_GRU_REGS=$R timeout -s ABRT 240 python -X faulthandler bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/r2aa_bench_r$R.json 2> gpurun_out/r2aa_bench_r$R.err; echo "bench regs $R rc=$?"
done
python - <<'PY'
import json
for r in (96,80,72,64):
    try:
        d=json.load(open(f"gpurun_out/r2aa_bench_r{r}.json"))
        print(r, d["config"]["batches_in_flight"], round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d.get("parity"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e: print(r, "failed", e)
PY
