#!/bin/bash
# Round 2, call E: GPU tests (model fixtures with conditioning-aware bars), tile-staged gather sweep, ncu of the gather.
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -s -k "not two_gpus" ) > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|max err|relative L2|Error" gpurun_out/r02e_pytest.log | cut -c1-260 | head -60
run() { echo "== $*"; env "$@" timeout 300 python tools/spmm_10m.py 2>&1 | grep -v Warning | grep -v "copy\|fp64" | tee -a gpurun_out/r02e_spmm.log; }
run GODE_SPMM_VARIANT=7 GODE_SPMM_MINB=6 GODE_SPMM_TS_ROWS=16
run GODE_SPMM_VARIANT=7 GODE_SPMM_MINB=6 GODE_SPMM_TS_ROWS=32
run GODE_SPMM_VARIANT=7 GODE_SPMM_MINB=8 GODE_SPMM_TS_ROWS=32
run GODE_SPMM_VARIANT=7 GODE_SPMM_MINB=6 GODE_SPMM_TS_ROWS=128
run GODE_SPMM_VARIANT=7 GODE_SPMM_MINB=5 GODE_SPMM_TS_ROWS=64
run GODE_SPMM_VARIANT=7 GODE_SPMM_MINB=7 GODE_SPMM_TS_ROWS=64
run GODE_SPMM_VARIANT=7 GODE_SPMM_MINB=6 GODE_SPMM_TS_ROWS=64 GODE_SPMM_PFDIST=1000
run GODE_SPMM_VARIANT=7 GODE_SPMM_MINB=6 GODE_SPMM_TS_ROWS=64 GODE_SPMM_PFDIST=2000
run GODE_SPMM_VARIANT=7 GODE_SPMM_MINB=6 GODE_SPMM_TS_ROWS=64 GODE_SPMM_PFDIST=4000
run GODE_SPMM_VARIANT=7 GODE_SPMM_MINB=6 GODE_SPMM_TS_ROWS=32 GODE_SPMM_PFDIST=4000
GODE_SPMM_VARIANT=7 GODE_SPMM_MINB=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_ts -s 2 -c 2 -o gpurun_out/r02e_ts -f python tools/spmm_10m.py 10000000 0.9 0 1 > gpurun_out/r02e_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02e_ts.ncu-rep --page raw --csv > gpurun_out/r02e_ts_raw.csv 2>/dev/null
ncu -i gpurun_out/r02e_ts.ncu-rep --page details --csv > gpurun_out/r02e_ts_details.csv 2>/dev/null
rm -f gpurun_out/r02e_ts.ncu-rep
