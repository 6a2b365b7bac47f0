#!/bin/bash
# round 2, GPU run K: halo conv kernel (3 hi*hi buffers, shift-only epilogue) as default, warp-per-component rects,
# N-fastest GEMM raster: full GPU suite, phase timers, bench (default / single-CTA conv)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
OCRS_B200_CONV_DEBUG=1 timeout 300 $B > gpurun_out/r2k_dbg.json 2> gpurun_out/r2k_dbg.err; echo "dbg rc=$?"; grep -m5 "conv dbg" gpurun_out/r2k_dbg.err
timeout 400 python bench.py --steps 20 --warmup 3 --in-flight 3 --no-cpu-baseline > gpurun_out/r2k_bench_if3.json 2> gpurun_out/r2k_bench_if3.err; echo "bench rc=$?"
OCRS_B200_CONV_MODE=0 timeout 400 python bench.py --steps 20 --warmup 3 --in-flight 3 --no-cpu-baseline > gpurun_out/r2k_bench_mode0.json 2> gpurun_out/r2k_bench_mode0.err; echo "bench mode0 rc=$?"
python - <<'PY'
import json
for n in ("if3","mode0"):
    try:
        d=json.load(open(f"gpurun_out/r2k_bench_{n}.json"))
        print(n, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["op_ms_per_step"], d["stage_ms_per_step"], d.get("parity"))
    except Exception as e: print(n, "failed", e)
PY
