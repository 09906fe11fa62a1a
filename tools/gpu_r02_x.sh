#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_gat.py -q -x > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02x_pytest.log
python tools/gat_profile.py > gpurun_out/r02x_gat.log 2>&1; tail -1 gpurun_out/r02x_gat.log
python tools/qc_profile.py > gpurun_out/r02x_qc.log 2>&1; tail -1 gpurun_out/r02x_qc.log
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02x_qc_launches.csv python tools/qc_profile.py > gpurun_out/r02x_qc_ncu.log 2>&1; echo "ncu rc=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02x_gat_launches.csv python tools/gat_profile.py > gpurun_out/r02x_gat_ncu.log 2>&1; echo "ncu rc=$?"
