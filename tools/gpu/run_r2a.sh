#!/bin/bash
# round 2, GPU run A: parity of the ragged conv path, micro-benchmarks, phase timers, bench at 1/2/3 in flight
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" 
timeout 120 tools/ubench/ubench_dsmem > gpurun_out/r2a_ubench_dsmem.log 2>&1; echo "dsmem rc=$?"
timeout 120 tools/ubench/ubench_mma > gpurun_out/r2a_ubench_mma.log 2>&1; echo "mma rc=$?"
OCRS_B200_CONV_DEBUG=1 OCRS_B200_GRU_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --in-flight 1 --no-cpu-baseline > gpurun_out/r2a_dbg.json 2> gpurun_out/r2a_dbg.err; echo "dbg rc=$?"
for f in 1 2 3; do
  timeout 300 python bench.py --steps 20 --warmup 3 --in-flight $f --no-cpu-baseline > gpurun_out/r2a_bench_if$f.json 2> gpurun_out/r2a_bench_if$f.err; echo "bench if$f rc=$?"
done
tail -3 gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_ubench_dsmem.log gpurun_out/r2a_ubench_mma.log
