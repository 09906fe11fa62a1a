#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parallel.py -x -q -m gpu > gpurun_out/r02ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02ae_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r02ae_bench2.json 2> gpurun_out/r02ae_bench2.err; echo "bench rc=$?"; tail -1 gpurun_out/r02ae_bench2.json | cut -c1-300
