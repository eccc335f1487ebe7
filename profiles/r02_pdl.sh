#!/bin/bash
# PDL (programmatic dependent launch) A/B on one GPU: parity first, then the bench at the full size and at the N = 8 share
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "programmatic or graph_equivalence or breakdown or full_size or lanczos or gmres" 2>&1 | tail -5
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for ny in 512 4096; do
  $B --ny $ny > gpurun_out/r02_pdl1_ny$ny.json 2> gpurun_out/r02_pdl.err
  $B --ny $ny --opt pdl=0 > gpurun_out/r02_pdl0_ny$ny.json 2>> gpurun_out/r02_pdl.err
done
tail -3 gpurun_out/r02_pdl.err
for f in gpurun_out/r02_pdl*.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'], 1), (d.get('parity') or {}).get('ok'), {k: (round(v['ms_total'], 2), v['launches']) for k, v in d['kernels'].items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
python profiles/stencil_ab.py
