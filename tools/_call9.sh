mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_final1.json 2> gpurun_out/bench_r01_final1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r01_final1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_10m.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmm_vec -s 39 -c 13 -f -o gpurun_out/spmm_vec_10m_bench_r01 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full.log
