#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_gemm_tc.py -q -x > gpurun_out/r02aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02aa_pytest.log
BANDS=16 python tools/gemm_tc_bench.py > gpurun_out/r02aa_gemm.jsonl 2> gpurun_out/r02aa_gemm.err; cat gpurun_out/r02aa_gemm.jsonl; tail -3 gpurun_out/r02aa_gemm.err
python tools/qc_profile.py 2>&1 | tail -1
python tools/gat_profile.py 2>&1 | tail -1
