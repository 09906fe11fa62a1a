#!/bin/bash
# Round 2, call H: dense-chain kernel times per accuracy configuration; QC + GAT configs with the tcgen05 GEMM routing; tests.
mkdir -p gpurun_out
for env in "GODE_TC_ACC=19 GODE_WGRAD_RND=1" "GODE_TC_ACC=3 GODE_WGRAD_RND=1" "GODE_TC_ACC=3 GODE_WGRAD_RND=0" "GODE_TC_ACC=1 GODE_WGRAD_RND=1"; do
  env $env timeout 300 python tools/dense_chain_time.py 2>&1 | grep -v Warn | tee -a gpurun_out/r02h_dense.log
done
timeout 900 python tools/bench_configs.py 3 5 > gpurun_out/r02h_configs.jsonl 2> gpurun_out/r02h_configs.err; cut -c1-400 gpurun_out/r02h_configs.jsonl
(time timeout 1500 python -m pytest tests -m gpu -q -s -k "not two_gpus" ) > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|max err|relative L2|Error" gpurun_out/r02h_pytest.log | cut -c1-300 | head -40
