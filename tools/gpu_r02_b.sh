#!/bin/bash
# Round 2, call B: GPU tests with every tolerance at the 1e-5 bar (no -x: list everything that misses it).
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q -s -k "not two_gpus" ) > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|relu-regime|weight gradient|^FAILED|max err|relative L2" gpurun_out/r02b_pytest.log | cut -c1-400 | head -80
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02b_smoke.log 2>&1; tail -3 gpurun_out/r02b_smoke.log
