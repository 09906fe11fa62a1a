#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_gat.py tests/test_gpu_models_golden.py -q -x -k "gat" > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02w_pytest.log
python tools/gat_profile.py > gpurun_out/r02w_gat.log 2>&1; tail -2 gpurun_out/r02w_gat.log
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02w_gat_launches.csv python tools/gat_profile.py > gpurun_out/r02w_gat_ncu.log 2>&1; echo "ncu rc=$?"
