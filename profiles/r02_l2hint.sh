#!/bin/bash
# A-B of an alternative build (liblkb_alt.so) vs the default one
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
ALT=$PWD/lightkrylov_b200/csrc/liblkb_alt.so
run() { local ny=$1; shift; "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --ny $ny 2>> gpurun_out/r02_l2hint.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'], 1), {k: round(v['GBps']) for k, v in d['kernels'].items()})"; }
for ny in 512 4096; do
echo "$ny default  $(run $ny env)"
echo "$ny alt      $(run $ny env LKB_SO=$ALT)"
done
echo "512 default  $(run 512 env)"
tail -2 gpurun_out/r02_l2hint.err
