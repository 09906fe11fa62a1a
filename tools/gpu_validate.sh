#!/bin/bash
# One B200: the validation pass whose outputs profiles/r02_*.md summarise -- GPU tests, smoke, default bench (+ reference arm),
# ncu launch list of one bench step, ncu --set full of the gather and of the dense chain.
#   gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh [tag]'      (outputs: gpurun_out/<tag>_*)
T=${1:-val}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -k "not two_gpus") > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/${T}_pytest.log | cut -c1-300 | head -20
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log | cut -c1-330
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/${T}_bench.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"; head -c 300 gpurun_out/${T}_bench_ref.json; echo
[ -n "$SKIP_NCU" ] && exit 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline > gpurun_out/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_spmm_t2" -s 40 -c 5 -o gpurun_out/${T}_gather -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline > gpurun_out/${T}_ncu_gather.log 2>&1; echo "ncu gather rc=$?"
ncu -i gpurun_out/${T}_gather.ncu-rep --page raw --csv > gpurun_out/${T}_gather_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"k_rows_ws|k_wgrad_tc|k_gn_bwd" -s 40 -c 5 -o gpurun_out/${T}_tc -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline > gpurun_out/${T}_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
ncu -i gpurun_out/${T}_tc.ncu-rep --page raw --csv > gpurun_out/${T}_tc_raw.csv 2>/dev/null
rm -f gpurun_out/${T}_gather.ncu-rep gpurun_out/${T}_tc.ncu-rep
