#!/bin/bash
# Round 2, call A: accuracy of the 3xTF32 configurations, first run of the general tcgen05 GEMM, L2 -> SM gather ceiling,
# regression run of the GPU tests with the new default configuration.
mkdir -p gpurun_out
timeout 900 python tests/transform_accuracy.py > gpurun_out/r02a_accuracy.jsonl 2> gpurun_out/r02a_accuracy.err; echo "accuracy rc=$?"
cut -c1-400 gpurun_out/r02a_accuracy.jsonl
GODE_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -x -q > gpurun_out/r02a_gemm_tc.log 2>&1; echo "gemm_tc rc=$?"; tail -5 gpurun_out/r02a_gemm_tc.log
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/l2_gather_bench tools/l2_gather_bench.cu && timeout 120 /tmp/l2_gather_bench > gpurun_out/r02a_l2_gather.jsonl 2>&1; echo "l2 rc=$?"; cat gpurun_out/r02a_l2_gather.jsonl
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02a_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02a_smoke.log 2>&1; tail -3 gpurun_out/r02a_smoke.log
