#!/bin/bash
# bisect the dopri5 / p2p-fused difference at 2 GPUs
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env GODE_HALO_MODE=p2p-fused "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/_parallel_worker.py dopri5 6000 128 smooth > gpurun_out/r02be_$name.log 2>&1
  echo "$name rc=$? $(grep -h 'world=' gpurun_out/r02be_$name.log | cut -c1-700)"
}
run default
run nounit GODE_UNIT_T=0
run nosched GODE_SPMM_SCHED=0
run noy GODE_PUSH_Y=0
run none GODE_UNIT_T=0 GODE_SPMM_SCHED=0 GODE_PUSH_Y=0 GODE_RK_RUNNING=0
