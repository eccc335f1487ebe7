#!/bin/bash
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python profiles/ktime_probe.py 64 128 > gpurun_out/r02_ktime3.jsonl 2> gpurun_out/r02_ktime3.err; tail -3 gpurun_out/r02_ktime3.err; cat gpurun_out/r02_ktime3.jsonl
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
for ny in 512 4096; do
  $B --ny $ny > gpurun_out/r02_ab4_ny$ny.json 2> gpurun_out/r02_ab4.err
done
for f in gpurun_out/r02_ab4_*.json; do python - "$f" <<'PY'
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], round(d['value'], 1), round(d['e2e']['value'], 1), d.get('parity'), {k: (round(v['ms_total'], 2), v['launches']) for k, v in d['kernels'].items()})
PY
done
python bench_configs.py --full --only c2,c4 > gpurun_out/r02_configs_a.jsonl 2> gpurun_out/r02_configs_a.err; cat gpurun_out/r02_configs_a.jsonl | cut -c1-400
