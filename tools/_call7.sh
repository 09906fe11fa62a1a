mkdir -p gpurun_out
for mode in async sync; do
GODE_HALO_MODE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/trace_step.py 10000000 2>&1 | grep -v -i "warn\|OMP\|\*\*\*" | tee gpurun_out/trace_summary_2g_$mode.txt
done
ls -la gpurun_out/trace_rank0*; gzip -f gpurun_out/trace_rank0_*.json
