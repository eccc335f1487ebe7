#!/bin/bash
# round-2 A/B of the new step kernels on one GPU at the full C2 size and at the 1/8 share (= the per-GPU work at N = 8)
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputests_a.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_gputests_a.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for ny in 512 4096; do
  $B --ny $ny > gpurun_out/r02_ab_new_ny$ny.json 2> gpurun_out/r02_ab.err
  $B --ny $ny --opt fin=0 > gpurun_out/r02_ab_nofin_ny$ny.json 2>> gpurun_out/r02_ab.err
  LKB_MULTIDOT_VARIANT=0 $B --ny $ny > gpurun_out/r02_ab_oldmd_ny$ny.json 2>> gpurun_out/r02_ab.err
done
for f in gpurun_out/r02_ab_*.json; do python - "$f" <<'PY'
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], round(d['value'], 1), {k: (round(v['ms_total'], 2), v['launches']) for k, v in d['kernels'].items()})
PY
done
