#!/bin/bash
# Round 2, call O: hub rows gathered inside the tile kernel vs the separate hub kernels.
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/spmm_10m.py 2>&1 | grep -v Warning | grep -v "copy" | tee -a gpurun_out/r02o_spmm.log; }
run GODE_SPMM_HUBS_INLINE=0
run GODE_SPMM_HUBS_INLINE=1
run GODE_SPMM_HUBS_INLINE=1 GODE_SPMM_MINB=6
timeout 600 python -m pytest tests/test_gpu_gcn.py -q -x 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline --no-library-baseline > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; echo "bench rc=$?"; head -c 260 gpurun_out/r02o_bench.json; echo
