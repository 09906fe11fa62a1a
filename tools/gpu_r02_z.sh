#!/bin/bash
mkdir -p gpurun_out
export BANDS=16
python tools/gemm_tc_bench.py > gpurun_out/r02z_gemm.jsonl 2>&1; cat gpurun_out/r02z_gemm.jsonl
ncu --set full --import-source on --clock-control none --kernel-name k_gemm_tc --launch-skip 1 --launch-count 1 -o gpurun_out/r02z_gemm_tc python tools/gemm_tc_bench.py > gpurun_out/r02z_ncu.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/r02z_gemm_tc.ncu-rep
