set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_4.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu_4.log
grep -E "Error|error" gpurun_out/pytest_gpu_4.log | grep -v "^    " | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2g_overlap.json 2> gpurun_out/bench_2g_overlap.err; echo "bench2 rc=$?"
cat gpurun_out/bench_2g_overlap.json | cut -c1-1200; tail -5 gpurun_out/bench_2g_overlap.err
