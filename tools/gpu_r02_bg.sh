#!/bin/bash
# after the read-stamp fix: dopri5 / p2p-fused repeatability, p2p-fused parity subset, 2-GPU bench
mkdir -p gpurun_out
run() { name=$1; shift
  env GODE_HALO_MODE=p2p-fused "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/_parallel_worker.py dopri5 6000 128 smooth > gpurun_out/r02bg_$name.log 2>&1
  echo "$name rc=$? $(grep -h 'world=' gpurun_out/r02bg_$name.log | sed 's/.*nfe/nfe/' | cut -c1-100) $(grep -h 'world=' gpurun_out/r02bg_$name.log | sed 's/.*stats/stats/' | cut -c1-200)"
}
for i in 1 2 3 4; do run default$i; done
(time timeout 900 python -m pytest tests/test_gpu_parallel.py -q -m gpu -k "p2p-fused or world1") > gpurun_out/r02bg_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02bg_pytest.log | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r02bg_bench2.json 2> gpurun_out/r02bg_bench2.err; echo "bench rc=$?"; tail -1 gpurun_out/r02bg_bench2.json | cut -c1-300
