set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_2g.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_2g.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1g.json 2> gpurun_out/bench_1g.err; echo "bench1 rc=$?"
cat gpurun_out/bench_1g.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2g.json 2> gpurun_out/bench_2g.err; echo "bench2 rc=$?"
cat gpurun_out/bench_2g.json; tail -5 gpurun_out/bench_2g.err
