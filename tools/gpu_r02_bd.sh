#!/bin/bash
# y-push exchange at 2 GPUs: parity of the p2p-fused cases, bench with and without
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parallel.py -x -q -m gpu -k "p2p-fused or world1") > gpurun_out/r02bd_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02bd_pytest.log | cut -c1-400
for y in 1 0; do
GODE_PUSH_Y=$y timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-library-baseline --no-e2e > gpurun_out/r02bd_bench2_y$y.json 2> gpurun_out/r02bd_bench2_y$y.err; echo "bench y=$y rc=$?"; tail -1 gpurun_out/r02bd_bench2_y$y.json | cut -c1-300
done
