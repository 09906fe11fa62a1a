#!/bin/bash
# hub chunks scheduled inside the tile grid: parity, microbench, bench
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_gcn.py tests/test_gpu_models_golden.py -m gpu -q -x) > gpurun_out/r02bc_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02bc_pytest.log | cut -c1-300 | head -20
for s in 1 0; do
  GODE_SPMM_SCHED=$s timeout 300 python tools/spmm_10m.py > gpurun_out/r02bc_spmm_sched$s.log 2>&1; echo "spmm sched=$s rc=$?"; tail -6 gpurun_out/r02bc_spmm_sched$s.log | cut -c1-250
done
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline --no-e2e > gpurun_out/r02bc_bench.json 2> gpurun_out/r02bc_bench.err; echo "bench rc=$?"; head -c 330 gpurun_out/r02bc_bench.json; echo
GODE_SPMM_SCHED=0 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline --no-e2e > gpurun_out/r02bc_bench_nosched.json 2> gpurun_out/r02bc_bench_nosched.err; echo "bench nosched rc=$?"; head -c 330 gpurun_out/r02bc_bench_nosched.json; echo
