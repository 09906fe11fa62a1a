#!/bin/bash
# running final combination + unit transpose: GCN parity tests, smoke, bench
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_gcn.py tests/test_gpu_models_golden.py tests/test_gpu_train.py -m gpu -q -x -s) > gpurun_out/r02ba_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02ba_pytest.log | cut -c1-300 | head -20
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02ba_smoke.log 2>&1; tail -3 gpurun_out/r02ba_smoke.log | cut -c1-330
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r02ba_bench.json 2> gpurun_out/r02ba_bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/r02ba_bench.json; echo
GODE_RK_RUNNING=0 GODE_UNIT_T=0 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline --no-e2e > gpurun_out/r02ba_bench_plain.json 2> gpurun_out/r02ba_bench_plain.err; echo "bench plain rc=$?"; head -c 400 gpurun_out/r02ba_bench_plain.json; echo
