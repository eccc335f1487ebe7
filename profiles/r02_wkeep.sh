#!/bin/bash
# w kept L2-resident (evict_last) for small slices: parity + A/B through LKB_W_KEEP_MB (0 = off, default 48)
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -2
run() { local ny=$1; shift; "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --ny $ny 2>> gpurun_out/r02_wkeep.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'], 1), {k: round(v['GBps']) for k, v in d['kernels'].items()})"; }
for ny in 512 1024 2048 4096; do
echo "$ny keep 48 MB (default) $(run $ny env)"
echo "$ny keep off             $(run $ny env LKB_W_KEEP_MB=0)"
done
echo "2048 keep forced (80 MB)  $(run 2048 env LKB_W_KEEP_MB=80)"
tail -2 gpurun_out/r02_wkeep.err
