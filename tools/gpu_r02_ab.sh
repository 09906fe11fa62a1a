#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_gemm_tc.py -q -x > gpurun_out/r02ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02ab_pytest.log
for pf in 0 4 8; do
  echo "prefetch $pf"
  GODE_GEMM_PREFETCH=$pf BANDS=16 python tools/gemm_tc_bench.py 2>&1 | grep product
done
python tools/qc_profile.py 2>&1 | tail -1
