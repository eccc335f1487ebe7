#!/bin/bash
# serpentine sweeps A/B (L2 reuse between the Gram-Schmidt kernels of a step): parity, then bench at the N = 8 share / N = 4 share / full size
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
run() { local ny=$1; shift; python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-profile-pass --ny $ny "$@" 2>> gpurun_out/r02_serp.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'], 1), (d.get('parity') or {}).get('ok'))"; }
for ny in 512 1024 4096; do
echo "$ny serpentine=1   $(run $ny)"
echo "$ny serpentine=0   $(run $ny --opt serpentine=0)"
done
echo "512 serpentine=1 again  $(run 512)"
tail -3 gpurun_out/r02_serp.err
