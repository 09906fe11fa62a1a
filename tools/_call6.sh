set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for cfg in "8 async" "8 sync" "4 async"; do
set -- $cfg
GODE_HALO_MODE=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $1 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_$1g_$2.json 2> gpurun_out/bench_$1g_$2.err; echo "bench $1 $2 rc=$?"
cut -c1-330 gpurun_out/bench_$1g_$2.json; tail -3 gpurun_out/bench_$1g_$2.err
done
