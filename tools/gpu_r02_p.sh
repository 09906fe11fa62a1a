#!/bin/bash
# Round 2, call P (2 GPUs): parity of the pipelined forward stage (GODE_PIPE_G) and its effect on the 2-GPU step.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parallel.py -q -s -k "g2 or g3 or world1") > gpurun_out/r02p_par_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|world=|Error" gpurun_out/r02p_par_pytest.log | cut -c1-400 | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for g in 0 2 4; do
  GODE_PIPE_G=$g timeout 600 $TR --master-port 2955$g bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r02p_bench_2g_pipeG$g.json 2> gpurun_out/r02p_bench_2g_pipeG$g.err; echo "pipeG=$g rc=$?"; python - <<PY
import json
b=json.loads(open('gpurun_out/r02p_bench_2g_pipeG$g.json').read().strip().splitlines()[-1])
print('ms/step', round(b['ms_per_step'],2), {k:round(v['ms_total']/b['steps'],2) for k,v in b['roofline']['kernel_classes_ms'].items()})
PY
done
