#!/bin/bash
# Round 2, call C: GPU tests after the accuracy work (wgrad / gemm_tc hierarchical accumulation), then the gather variants at N = 10 M.
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q -s -k "not two_gpus" ) > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|relu-regime|weight gradient|gemm_tc M|^FAILED|max err|relative L2" gpurun_out/r02c_pytest.log | cut -c1-300 | head -60
for cfg in "0 8 4 1" "6 8 4 1" "6 8 4 0" "6 6 4 1" "6 6 8 1" "6 4 8 1"; do
  set -- $cfg
  echo "== variant=$1 minb=$2 unr=$3 rowval=$4"
  GODE_SPMM_VARIANT=$1 GODE_SPMM_MINB=$2 GODE_SPMM_UNR=$3 GODE_SPMM_ROWVAL=$4 timeout 300 python tools/spmm_10m.py 2>&1 | grep -v Warning | tee -a gpurun_out/r02c_spmm.log
done
