#!/bin/bash
# Round 2, call L (2 GPUs): P ranks == 1 rank parity tests for every exchange mode, 2-GPU bench, data-parallel QC step.
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests/test_gpu_parallel.py -q -s) > gpurun_out/r02l_par_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|world=" gpurun_out/r02l_par_pytest.log | cut -c1-400 | tail -24
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02l_bench_2g.json 2> gpurun_out/r02l_bench_2g.err; echo "bench2 rc=$?"; head -c 500 gpurun_out/r02l_bench_2g.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/bench_configs.py 5dp > gpurun_out/r02l_qc_dp2.jsonl 2> gpurun_out/r02l_qc_dp2.err; echo "qcdp rc=$?"; cat gpurun_out/r02l_qc_dp2.jsonl
