#!/bin/bash
# dopri5 / p2p-fused at 2 GPUs: repeatability per configuration
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env GODE_HALO_MODE=p2p-fused "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/_parallel_worker.py dopri5 6000 128 smooth > gpurun_out/r02bf_$name.log 2>&1
  echo "$name rc=$? $(grep -h 'world=' gpurun_out/r02bf_$name.log | sed 's/.*nfe/nfe/' | cut -c1-120) $(grep -h 'world=' gpurun_out/r02bf_$name.log | sed 's/.*stats/stats/' | cut -c1-200)"
}
for i in 1 2 3; do run default$i; done
for i in 1 2 3; do run nosched$i GODE_SPMM_SCHED=0; done
for i in 1 2; do run none$i GODE_UNIT_T=0 GODE_SPMM_SCHED=0 GODE_PUSH_Y=0 GODE_RK_RUNNING=0; done
for i in 1 2; do run blocking$i CUDA_LAUNCH_BLOCKING=1; done
for i in 1 2; do run async$i GODE_HALO_MODE=async; done
