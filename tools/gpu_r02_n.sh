#!/bin/bash
# Round 2, call N: validation of the shipped state on one B200 -- tests, smoke, default bench (+ reference arm, + locality 0),
# ncu launch list, ncu --set full of the gather and of the tcgen05 kernels.
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -s -k "not two_gpus" ) > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/r02n_pytest.log | cut -c1-300 | head -20
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02n_smoke.log 2>&1; tail -3 gpurun_out/r02n_smoke.log | cut -c1-330
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/r02n_bench.json; echo
timeout 900 python bench.py --steps 5 --warmup 3 --locality 0 --no-cpu-baseline --no-library-baseline > gpurun_out/r02n_bench_loc0.json 2> gpurun_out/r02n_bench_loc0.err; echo "bench loc0 rc=$?"; head -c 300 gpurun_out/r02n_bench_loc0.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02n_bench_ref.json 2> gpurun_out/r02n_bench_ref.err; echo "ref rc=$?"; head -c 400 gpurun_out/r02n_bench_ref.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02n_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline > gpurun_out/r02n_ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_spmm_t2|k_spmm_heavy2" -s 54 -c 6 -o gpurun_out/r02n_gather -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline > gpurun_out/r02n_ncu_gather.log 2>&1; echo "ncu gather rc=$?"
ncu -i gpurun_out/r02n_gather.ncu-rep --page raw --csv > gpurun_out/r02n_gather_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"k_rows_ws|k_wgrad_tc|k_gn_bwd" -s 48 -c 6 -o gpurun_out/r02n_tc -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline > gpurun_out/r02n_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
ncu -i gpurun_out/r02n_tc.ncu-rep --page raw --csv > gpurun_out/r02n_tc_raw.csv 2>/dev/null
rm -f gpurun_out/r02n_gather.ncu-rep gpurun_out/r02n_tc.ncu-rep
