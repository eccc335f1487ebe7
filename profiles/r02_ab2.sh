#!/bin/bash
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python profiles/ktime_probe.py 16 64 128 > gpurun_out/r02_ktime.jsonl 2> gpurun_out/r02_ktime.err; tail -3 gpurun_out/r02_ktime.err; cat gpurun_out/r02_ktime.jsonl
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_final or csr_random or csr_create or bidiag or csr_matvec" 2>&1 | tail -5
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for ny in 512 4096; do
  $B --ny $ny > gpurun_out/r02_ab2_ny$ny.json 2> gpurun_out/r02_ab2.err
done
for f in gpurun_out/r02_ab2_*.json; do python - "$f" <<'PY'
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], round(d['value'], 1), {k: (round(v['ms_total'], 2), v['launches']) for k, v in d['kernels'].items()})
PY
done
