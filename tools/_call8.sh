mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_8.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu_8.log
for r in 20 0; do
GODE_RESERVE_SMS=$r GODE_HALO_MODE=async timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/trace_step.py 10000000 2>&1 | grep -v -i "warn\|OMP_NUM\|\*\*\*" | head -8 | cut -c1-300
done
