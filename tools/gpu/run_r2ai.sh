#!/bin/bash
# round 2, GPU run AI: final validation (weights-stationary kernel in): GPU suite, smoke, default bench, reference arm,
# ncu --set full of the conv kernels of one batch
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ai_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ai_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2ai_bench_default.json 2> gpurun_out/r2ai_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ai_bench_reference.json 2> gpurun_out/r2ai_bench_reference.err; echo "reference rc=$?"
B="python bench.py --steps 1 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01"
timeout 600 ncu --set full --clock-control none -k regex:conv3x3_ -s 10 -c 5 -f -o /tmp/r2ai_conv $B > gpurun_out/r2ai_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
ncu -i /tmp/r2ai_conv.ncu-rep --page raw --csv > gpurun_out/r2ai_conv_raw.csv 2>/dev/null; ls -la /tmp/r2ai_conv.ncu-rep
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ai_bench_default.json"))
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["config"]["batches_in_flight"], d.get("parity"), d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["cpu_baseline"]["value"], d["clocks"])
print(d["op_ms_per_step"]); print(d["stage_ms_per_step"])
r=json.load(open("gpurun_out/r2ai_bench_reference.json"))
print("reference", r["value"], r["unit"])
PY
