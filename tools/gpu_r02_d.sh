#!/bin/bash
# Round 2, call D: all GPU tests incl. the new model fixtures; tile-staged gather variant at N = 10 M.
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -s -k "not two_gpus" ) > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|max err|relative L2| ok: worst|Error" gpurun_out/r02d_pytest.log | cut -c1-260 | head -90
for cfg in "7 8 4 1" "7 8 4 0" "7 6 4 1" "7 6 8 1" "7 8 8 1" "7 4 8 1"; do
  set -- $cfg
  echo "== variant=$1 minb=$2 unr=$3 rowval=$4"
  GODE_SPMM_VARIANT=$1 GODE_SPMM_MINB=$2 GODE_SPMM_UNR=$3 GODE_SPMM_ROWVAL=$4 timeout 300 python tools/spmm_10m.py 2>&1 | grep -v Warning | tee -a gpurun_out/r02d_spmm.log
done
