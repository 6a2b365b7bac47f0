#!/bin/bash
# round 2, GPU run W: final-state evidence: GPU suite, default bench (both arms), ncu launch list of the bench command,
# ncu --set full of one steady-state batch (raw CSV only; the report stays on the box), source-level capture of the
# conv kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2w_pytest.log
timeout 900 python bench.py > gpurun_out/r2w_bench_default.json 2> gpurun_out/r2w_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err; echo "reference rc=$?"
B="python bench.py --steps 1 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 3 --in-flight 1 --no-cpu-baseline --min-seconds 0.01 > gpurun_out/r2w_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none -s 260 -c 66 -f -o /tmp/r2w_batch $B > gpurun_out/r2w_ncu.log 2>&1; echo "ncu batch rc=$?"
ncu -i /tmp/r2w_batch.ncu-rep --page raw --csv > gpurun_out/r2w_batch_raw.csv 2>/dev/null; ls -la /tmp/r2w_batch.ncu-rep gpurun_out/r2w_batch_raw.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo_kernel -s 5 -c 1 -f -o gpurun_out/r2w_conv_halo $B > gpurun_out/r2w_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
ncu -i gpurun_out/r2w_conv_halo.ncu-rep --page details --csv > gpurun_out/r2w_conv_halo_details.csv 2>/dev/null
python - <<'PY'
import json
for n in ("default","reference"):
    try:
        d=json.load(open(f"gpurun_out/r2w_bench_{n}.json"))
        print(n, round(d["value"],2), d.get("e2e"), d.get("ms_per_step"), d.get("roofline"), d.get("cpu_baseline"), d.get("parity"), d.get("clocks"))
    except Exception as e: print(n, "failed", e)
PY
