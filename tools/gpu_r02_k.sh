#!/bin/bash
# Round 2, call K: tests incl. the fused / graphed epoch; Cora epoch times eager vs graphed.
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -s -k "not two_gpus" ) > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/r02k_pytest.log | cut -c1-300 | head -30
for mode in off on; do for m in "res3" "ode3 --method rk4"; do
  echo "== fused-epoch $mode model $m"
  timeout 300 python -m graph_odenet_b200.GCN.train_res --model $m --npz tests/golden/planetoid_cora.npz --epochs 60 --fused-epoch $mode 2>&1 | grep -E "Epoch: 00(05|30|60)|Average time|Test set" | cut -c1-200
done; done 2>&1 | tee gpurun_out/r02k_epochs.log
