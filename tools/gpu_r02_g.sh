#!/bin/bash
# Round 2, call G: all GPU tests, smoke, the default bench, ncu launch list and a --set full capture of the new gather kernel.
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -s -k "not two_gpus" ) > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|max err|relative L2|Error" gpurun_out/r02g_pytest.log | cut -c1-260 | head -40
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02g_smoke.log 2>&1; tail -3 gpurun_out/r02g_smoke.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/r02g_bench.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02g_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline > gpurun_out/r02g_ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_spmm_t2 -s 27 -c 3 -o gpurun_out/r02g_t2 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline > gpurun_out/r02g_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r02g_t2.ncu-rep --page raw --csv > gpurun_out/r02g_t2_raw.csv 2>/dev/null
rm -f gpurun_out/r02g_t2.ncu-rep
