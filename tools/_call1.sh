set -x
mkdir -p gpurun_out
for v in 5 4 0 3; do GODE_SPMM_VARIANT=$v timeout 300 python tools/spmm_10m.py 10000000 0.9 0 5 2>&1 | grep -v Warn; done > gpurun_out/spmm10m_variants.log 2>&1
GODE_SPMM_BULK=1 timeout 300 python tools/spmm_10m.py 10000000 0.9 0 5 2>&1 | grep -v Warn >> gpurun_out/spmm10m_variants.log
for w in 2048 256; do timeout 300 python tools/spmm_10m.py 10000000 0.9 $w 5 2>&1 | grep -v Warn; done >> gpurun_out/spmm10m_variants.log 2>&1
timeout 300 python tools/spmm_10m.py 10000000 0.0 0 5 2>&1 | grep -v Warn >> gpurun_out/spmm10m_variants.log
cat gpurun_out/spmm10m_variants.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_pw -s 3 -c 1 -f -o gpurun_out/spmm_pw_10m_r01 python tools/spmm_10m.py 10000000 0.9 0 2 > gpurun_out/ncu_spmm10m.log 2>&1
tail -3 gpurun_out/ncu_spmm10m.log
