#!/bin/bash
# L1::no_allocate streaming loads (default build) vs allocating loads (liblkb_alt.so = -DLKB_L1_ALLOC), with / without PDL and explicit carve-out
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "programmatic or graph_equivalence or fused or full_size or lanczos or gmres or vector_tbps or dgs" 2>&1 | tail -3
ALT=$PWD/lightkrylov_b200/csrc/liblkb_alt.so
run() { local ny=$1; shift; "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-profile-pass --ny $ny 2>> gpurun_out/r02_pdl4.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'], 1))"; }
echo "512 new m0        $(run 512 env LKB_PDL_MASK=0)"
echo "512 old m0        $(run 512 env LKB_PDL_MASK=0 LKB_SO=$ALT)"
echo "512 new m4        $(run 512 env LKB_PDL_MASK=4)"
echo "512 new m255      $(run 512 env LKB_PDL_MASK=255)"
echo "512 new m0 c0     $(run 512 env LKB_PDL_MASK=0 LKB_CARVEOUT_L1=0)"
echo "512 new m4 c0     $(run 512 env LKB_PDL_MASK=4 LKB_CARVEOUT_L1=0)"
echo "512 new m255 c0   $(run 512 env LKB_PDL_MASK=255 LKB_CARVEOUT_L1=0)"
echo "512 new m0 again  $(run 512 env LKB_PDL_MASK=0)"
echo "4096 new m0       $(run 4096 env LKB_PDL_MASK=0)"
echo "4096 old m0       $(run 4096 env LKB_PDL_MASK=0 LKB_SO=$ALT)"
echo "4096 new m4       $(run 4096 env LKB_PDL_MASK=4)"
echo "4096 new m255 c0  $(run 4096 env LKB_PDL_MASK=255 LKB_CARVEOUT_L1=0)"
tail -3 gpurun_out/r02_pdl4.err
