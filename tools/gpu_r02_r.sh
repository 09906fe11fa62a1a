#!/bin/bash
mkdir -p gpurun_out
for tc in default 0; do
  for d in 64 128; do
    echo "=== d=$d GODE_TC=$tc" >> gpurun_out/r02r_noise.log
    if [ $tc = default ]; then python tests/d64_noise.py $d gpu >> gpurun_out/r02r_noise.log 2>&1
    else GODE_TC=$tc python tests/d64_noise.py $d gpu >> gpurun_out/r02r_noise.log 2>&1; fi
  done
done
grep -v Warning gpurun_out/r02r_noise.log | tail -150
