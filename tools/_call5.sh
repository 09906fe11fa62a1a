set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_5.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu_5.log
grep -E "Error|error|world=" gpurun_out/pytest_gpu_5.log | grep -v "^    " | head -30
for mode in sync async; do
GODE_HALO_MODE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2g_$mode.json 2> gpurun_out/bench_2g_$mode.err; echo "bench2 $mode rc=$?"
cut -c1-330 gpurun_out/bench_2g_$mode.json; tail -3 gpurun_out/bench_2g_$mode.err
done
