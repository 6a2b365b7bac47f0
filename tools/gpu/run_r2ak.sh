#!/bin/bash
# round 2, GPU run AK: the conv-variant test
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_conv_variants.py -m gpu -x -q > gpurun_out/r2ak_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2ak_pytest.log | cut -c1-400
