#!/bin/bash
# ncu --set full of the L2-blocked SpMV kernels on the full-size C5 matrix (one mid-matrix column block each)
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
for v in 0 1; do
  SPMV_AB_CHILD=1 LKB_CSR_BLOCKED_VARIANT=$v timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_csr_blocked -s 33 -c 2 \
     -o gpurun_out/prof_csrblk${v}_r02b -f python profiles/spmv_ab.py 1.0 48 > gpurun_out/ncu_csrblk$v.log 2>&1
  tail -3 gpurun_out/ncu_csrblk$v.log
done
ls -la gpurun_out/*.ncu-rep | tail -3
