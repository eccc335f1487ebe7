#!/bin/bash
# which programmatic-launch link helps / hurts: LKB_PDL_MASK bits 1 matvec, 2 multi-dot, 4 fused, 8 multi-axpy(fin), 16 scale
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "programmatic" 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-profile-pass"
for ny in 512 4096; do
for m in 0 1 2 4 8 16 6 14 255; do
  LKB_PDL_MASK=$m $B --ny $ny 2>> gpurun_out/r02_pdl2.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ny', $ny, 'mask', $m, round(d['value'], 1))"
done
done
tail -3 gpurun_out/r02_pdl2.err
