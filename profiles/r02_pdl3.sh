#!/bin/bash
# why are PDL-launched LDG kernels (multi-dot, final multi-axpy) ~18 % slower?  L1 carve-out / L1 allocation hypotheses
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-profile-pass --ny 512"
run() { "$@" $B 2>> gpurun_out/r02_pdl3.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'], 1))"; }
for m in 0 2 8 255; do
  echo "mask $m default            $(run env LKB_PDL_MASK=$m)"
  echo "mask $m carveout L1 (0)    $(run env LKB_PDL_MASK=$m LKB_CARVEOUT_L1=0)"
  echo "mask $m carveout smem(100) $(run env LKB_PDL_MASK=$m LKB_CARVEOUT_L1=100)"
  echo "mask $m no_allocate build  $(run env LKB_PDL_MASK=$m LKB_SO=$PWD/lightkrylov_b200/csrc/liblkb_alt.so)"
done
tail -3 gpurun_out/r02_pdl3.err
