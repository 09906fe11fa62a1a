#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q -x > gpurun_out/r02ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02ac_pytest.log
for v in 2 1; do
  echo "GODE_GEMM_V=$v"
  GODE_GEMM_V=$v BANDS=16 timeout 300 python tools/gemm_tc_bench.py 2>&1 | grep "product\|Error"
done
timeout 300 python tools/qc_profile.py 2>&1 | tail -1
timeout 300 python tools/gat_profile.py 2>&1 | tail -1
