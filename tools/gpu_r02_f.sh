#!/bin/bash
# Round 2, call F: lean tile-staged gather (variant 8) at N = 10 M.
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/spmm_10m.py 2>&1 | grep -v Warning | grep -v "copy" | tee -a gpurun_out/r02f_spmm.log; }
run GODE_SPMM_VARIANT=8 GODE_SPMM_MINB=6
run GODE_SPMM_VARIANT=8 GODE_SPMM_MINB=8
run GODE_SPMM_VARIANT=8 GODE_SPMM_MINB=5
run GODE_SPMM_VARIANT=8 GODE_SPMM_MINB=4
run GODE_SPMM_VARIANT=8 GODE_SPMM_MINB=6 GODE_SPMM_TS_ROWS=32
run GODE_SPMM_VARIANT=8 GODE_SPMM_MINB=8 GODE_SPMM_TS_ROWS=32
run GODE_SPMM_VARIANT=8 GODE_SPMM_MINB=6 GODE_SPMM_TS_ROWS=128
run GODE_SPMM_VARIANT=8 GODE_SPMM_MINB=6 GODE_SPMM_PREFETCH=0
