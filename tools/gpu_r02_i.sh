#!/bin/bash
# Round 2, call I: dense-chain times with the templated accuracy configurations, wgrad accuracy with / without rounded
# residuals, all single-GPU tests, ncu launch lists of configs 3 (GAT) and 5 (QC).
mkdir -p gpurun_out
for env in "GODE_TC_ACC=19 GODE_WGRAD_RND=1" "GODE_TC_ACC=3 GODE_WGRAD_RND=1" "GODE_TC_ACC=3 GODE_WGRAD_RND=0"; do
  env $env timeout 300 python tools/dense_chain_time.py 2>&1 | grep -v Warn | tee -a gpurun_out/r02i_dense.log
done
GODE_WGRAD_RND=0 timeout 300 python -m pytest tests/test_gpu_gcn.py -q -s -k "long_reduction or relu_regime" 2>&1 | grep -E "weight gradient|relu-regime|passed|failed" | cut -c1-700 | tee gpurun_out/r02i_wgrad_rnd0.log
(time timeout 1500 python -m pytest tests -m gpu -q -s -k "not two_gpus" ) > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02i_pytest.log | cut -c1-300 | head -40
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02i_gat_launches.csv python tools/bench_configs.py 3 > gpurun_out/r02i_gat_ncu.log 2>&1; echo "ncu gat rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02i_qc_launches.csv python tools/bench_configs.py 5 > gpurun_out/r02i_qc_ncu.log 2>&1; echo "ncu qc rc=$?"
