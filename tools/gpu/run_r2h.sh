#!/bin/bash
# round 2, GPU run H: conv epilogue with 4-lane transposed stores (single + pair), beam search canonical ids
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2h_pytest.log
OCRS_B200_CONV_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01 > gpurun_out/r2h_dbg.json 2> gpurun_out/r2h_dbg.err; echo "dbg rc=$?"
grep -m5 "conv dbg" gpurun_out/r2h_dbg.err
timeout 400 python bench.py --steps 20 --warmup 3 --in-flight 3 --no-cpu-baseline > gpurun_out/r2h_bench_if3.json 2> gpurun_out/r2h_bench_if3.err; echo "bench if3 rc=$?"
OCRS_B200_CONV_PAIR=1 OCRS_B200_CONV_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01 > gpurun_out/r2h_dbg_pair.json 2> gpurun_out/r2h_dbg_pair.err; grep -m5 "conv dbg" gpurun_out/r2h_dbg_pair.err
OCRS_B200_CONV_PAIR=1 timeout 400 python bench.py --steps 20 --warmup 3 --in-flight 3 --no-cpu-baseline > gpurun_out/r2h_bench_if2.json 2> gpurun_out/r2h_bench_if2.err; echo "bench pair rc=$?"
#timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01 > gpurun_out/r2h_ncu_bench.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for n in ["if3","if2"]:
    try:
        d=json.load(open(f"gpurun_out/r2h_bench_{n}.json"))
        print(n, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["op_ms_per_step"], d["stage_ms_per_step"], d["parity"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2h_bench_{n}.err").read()[-1500:])
PY
