#!/bin/bash
# N B200 of one box (gpurun --gpus N): row-partitioned bench at N ranks; with N = 2 also the P ranks == 1 rank parity tests of
# every exchange mode and the repeatability of the adaptive solver in the fused mode; QC data-parallel step.
#   gpurun --gpus 2 --timeout 2400 -- 'bash tools/gpu_multi.sh 2 [tag]'
N=${1:-2}; T=${2:-multi}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ] && [ -z "$SKIP_PARITY" ]; then
  (time timeout 1800 python -m pytest tests/test_gpu_parallel.py -q) > gpurun_out/${T}_par_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_par_pytest.log | cut -c1-300
  for i in 1 2 3; do
    GODE_HALO_MODE=p2p-fused timeout 300 $TR --master-port 29611 tests/_parallel_worker.py dopri5 6000 128 smooth > gpurun_out/${T}_dopri5_$i.log 2>&1
    echo "dopri5 run $i rc=$? $(grep -h 'world=' gpurun_out/${T}_dopri5_$i.log | sed 's/.*nfe/nfe/' | cut -c1-100)"
  done
fi
timeout 600 $TR --master-port 29541 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/${T}_bench_${N}g.json 2> gpurun_out/${T}_bench_${N}g.err; echo "bench rc=$?"; head -c 300 gpurun_out/${T}_bench_${N}g.json; echo
[ -n "$WITH_NOY" ] && { GODE_PUSH_Y=0 timeout 600 $TR --master-port 29545 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-e2e > gpurun_out/${T}_bench_${N}g_noY.json 2> gpurun_out/${T}_bench_${N}g_noY.err; echo "bench noY rc=$?"; head -c 300 gpurun_out/${T}_bench_${N}g_noY.json; echo; }
[ -n "$WITH_QC" ] && { timeout 600 $TR --master-port 29542 tools/bench_configs.py 5dp > gpurun_out/${T}_qc_dp${N}.jsonl 2> gpurun_out/${T}_qc_dp${N}.err; echo "qcdp rc=$?"; cat gpurun_out/${T}_qc_dp${N}.jsonl; }
exit 0
