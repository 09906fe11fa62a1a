set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parallel.py::test_two_gpus_match_one > gpurun_out/pytest_gpu_3.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu_3.log
