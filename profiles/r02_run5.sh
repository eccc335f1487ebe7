#!/bin/bash
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python profiles/stencil_ab.py > gpurun_out/r02_stencil_ab.txt 2>&1; cat gpurun_out/r02_stencil_ab.txt
timeout 900 python bench_configs.py --full --only c5 > gpurun_out/r02_c5_full_blocked.jsonl 2> gpurun_out/r02_c5_full_blocked.err; cat gpurun_out/r02_c5_full_blocked.jsonl | cut -c1-700; tail -3 gpurun_out/r02_c5_full_blocked.err
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
for ny in 512 4096; do
  $B --ny $ny > gpurun_out/r02_run5_ny$ny.json 2> gpurun_out/r02_run5.err
done
for f in gpurun_out/r02_run5_*.json; do python - "$f" <<'PY'
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], round(d['value'], 1), round(d['e2e']['value'], 1), (d.get('parity') or {}).get('ok'), {k: (round(v['ms_total'], 2), v['launches']) for k, v in d['kernels'].items()})
PY
done
python bench_configs.py --full --only c4 2>&1 | cut -c1-300
