#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -k "not two_gpus" > gpurun_out/r02ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02ad_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ad_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02ad_smoke.log
timeout 900 python tools/bench_configs.py 1 2 3 5 > gpurun_out/r02ad_configs.jsonl 2> gpurun_out/r02ad_configs.err; cut -c1-330 gpurun_out/r02ad_configs.jsonl
timeout 600 python bench.py --no-cpu-baseline --no-library-baseline > gpurun_out/r02ad_bench.json 2> gpurun_out/r02ad_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02ad_bench.json
