#!/bin/bash
# 3-D stencil CTA-width fix A/B + stencil parity tests + compute-sanitizer (memcheck, racecheck) on the round-2 kernels
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stencil or lanczos or cg_vs" 2>&1 | tail -3
python profiles/stencil_ab.py > gpurun_out/r02_stencil_ab4.txt 2>&1; cat gpurun_out/r02_stencil_ab4.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python profiles/sanitizer_workload.py > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python profiles/sanitizer_workload.py > gpurun_out/r02_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_racecheck.log
