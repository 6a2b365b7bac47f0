#!/bin/bash
# round 2, GPU run I: ncu --set full of every kernel of one steady-state batch; the big report stays on the box,
# only its raw CSV comes back; source-level captures of the conv and GRU kernels are kept as (small) reports
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --in-flight 1 --no-cpu-baseline --min-seconds 0.01"
timeout 1200 ncu --set full --clock-control none -s 260 -c 66 -f -o /tmp/r2i_batch $B > gpurun_out/r2i_ncu.log 2>&1; echo "ncu batch rc=$?"
ncu -i /tmp/r2i_batch.ncu-rep --page raw --csv > gpurun_out/r2i_batch_raw.csv 2>/dev/null; ls -la /tmp/r2i_batch.ncu-rep gpurun_out/r2i_batch_raw.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel -s 7 -c 2 -f -o gpurun_out/r2i_conv $B > gpurun_out/r2i_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
ncu -i gpurun_out/r2i_conv.ncu-rep --page source --csv > gpurun_out/r2i_conv_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_cluster3 -s 2 -c 1 -f -o gpurun_out/r2i_gru $B > gpurun_out/r2i_ncu_gru.log 2>&1; echo "ncu gru rc=$?"
ncu -i gpurun_out/r2i_gru.ncu-rep --page source --csv > gpurun_out/r2i_gru_source.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
OCRS_B200_CONV_DEBUG=1 timeout 300 $B > gpurun_out/r2i_dbg.json 2> gpurun_out/r2i_dbg.err; grep -m5 "conv dbg" gpurun_out/r2i_dbg.err
timeout 400 python bench.py --steps 20 --warmup 3 --in-flight 3 --no-cpu-baseline > gpurun_out/r2i_bench_if3.json 2> gpurun_out/r2i_bench_if3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2i_bench_if3.json"))
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],2), d["op_ms_per_step"], d["stage_ms_per_step"])
PY
