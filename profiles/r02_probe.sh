#!/bin/bash
# round-2 first measurement pass: per-GPU-share overhead experiment (1 GPU at 1/8, 1/4, 1/2 of C2) and ncu captures
# of the kernels VERDICT r01 lists as furthest below the roof (k_csr, 3-D stencil, CG kernels, basis GEMM)
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
for ny in 512 1024 2048; do
  python bench.py --ny $ny --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_share_ny$ny.json 2> gpurun_out/r02_share_ny$ny.err
done
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_csr -s 22 -c 3 -o gpurun_out/prof_csr_r02a python profiles/spmv_profile.py 0.1 > gpurun_out/ncu_csr.log 2>&1
$NCU -k regex:k_stencil -s 6 -c 1 -o gpurun_out/prof_stencil3d_r02a python profiles/ncu_targets.py stencil3d > gpurun_out/ncu_st3.log 2>&1
$NCU -k regex:"k_cg_update|k_dot2|k_cg_direction" -s 9 -c 3 -o gpurun_out/prof_cg_r02a python profiles/ncu_targets.py cg > gpurun_out/ncu_cg.log 2>&1
$NCU -k regex:k_basis_gemm -c 1 -o gpurun_out/prof_gemm_r02a python profiles/ncu_targets.py gemm > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/ncu_*.log
cat gpurun_out/r02_share_ny*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['partition'], d['value'], {k: round(v['ms_total'],2) for k, v in d['kernels'].items()})
"
