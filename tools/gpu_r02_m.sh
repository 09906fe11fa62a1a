#!/bin/bash
# Round 2, call M (8 GPUs): the bench at 8 ranks (default graph; locality 0 with and without the degree relabelling) and the
# data-parallel QC step.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02m_bench_8g.json 2> gpurun_out/r02m_bench_8g.err; echo "bench8 rc=$?"; head -c 300 gpurun_out/r02m_bench_8g.json; echo
timeout 600 $TR --master-port 29542 tools/bench_configs.py 5dp > gpurun_out/r02m_qc_dp8.jsonl 2> gpurun_out/r02m_qc_dp8.err; echo "qcdp rc=$?"; cat gpurun_out/r02m_qc_dp8.jsonl
timeout 600 $TR --master-port 29543 bench.py --gpus 8 --steps 5 --warmup 3 --locality 0 > gpurun_out/r02m_bench_8g_loc0.json 2> gpurun_out/r02m_bench_8g_loc0.err; echo "loc0 rc=$?"; head -c 300 gpurun_out/r02m_bench_8g_loc0.json; echo
timeout 600 $TR --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 --locality 0 --reorder degree > gpurun_out/r02m_bench_8g_loc0_degree.json 2> gpurun_out/r02m_bench_8g_loc0_degree.err; echo "loc0 degree rc=$?"; head -c 300 gpurun_out/r02m_bench_8g_loc0_degree.json; echo
