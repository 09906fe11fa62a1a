#!/bin/bash
# Round 2, call J: gather variants t2 vs t3, all single-GPU tests, default bench, GAT / QC configs.
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/spmm_10m.py 2>&1 | grep -v Warning | grep -v "copy" | tee -a gpurun_out/r02j_spmm.log; }
run GODE_SPMM_T3=0
run GODE_SPMM_T3=1
run GODE_SPMM_T3=1 GODE_SPMM_MINB=6
(time timeout 1500 python -m pytest tests -m gpu -q -s -k "not two_gpus" ) > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/r02j_pytest.log | cut -c1-300 | head -30
timeout 900 python bench.py > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/r02j_bench.json; echo
timeout 900 python tools/bench_configs.py 3 5 > gpurun_out/r02j_configs.jsonl 2> gpurun_out/r02j_configs.err; cut -c1-330 gpurun_out/r02j_configs.jsonl
