#!/bin/bash
# Round 2, call Q: final single-GPU validation of the committed state + configs 1, 2, 3 (Cora / Pubmed / GAT timings).
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -s -k "not two_gpus" ) > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/r02q_pytest.log | cut -c1-300 | head -20
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02q_smoke.log 2>&1; tail -1 gpurun_out/r02q_smoke.log
timeout 900 python tools/bench_configs.py 1 2 3 > gpurun_out/r02q_configs.jsonl 2> gpurun_out/r02q_configs.err; cut -c1-260 gpurun_out/r02q_configs.jsonl
timeout 600 python bench.py --no-cpu-baseline --no-library-baseline > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"; head -c 260 gpurun_out/r02q_bench.json; echo
