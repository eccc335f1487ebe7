#!/bin/bash
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
for v in 2 3; do echo "pytest csr, variant $v"; LKB_CSR_BLOCKED_VARIANT=$v python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "csr" 2>&1 | tail -3; done
timeout 1200 python profiles/spmv_ab.py 1.0 48 64 > gpurun_out/r02_spmv_ab4.jsonl 2> gpurun_out/r02_spmv_ab.err
cut -c1-220 gpurun_out/r02_spmv_ab4.jsonl; tail -3 gpurun_out/r02_spmv_ab.err
