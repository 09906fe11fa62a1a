#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qc_collate.py tests/test_gpu_qc.py -x -q -m gpu > gpurun_out/r02af_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02af_pytest.log
