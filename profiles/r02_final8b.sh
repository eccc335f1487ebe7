#!/bin/bash
# final multi-GPU lines of round 2 (code incl. the L2-resident work vector): world-8 parity worker, bench at N = 8 / 4 / 1 on one box
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -x -q -m gpu -k "8" 2>&1 | tail -2
run() { local N=$1; local P=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@"; }
run 8 29561 > gpurun_out/bench_r02b_8gpu.json 2> gpurun_out/r02_final8b.err
run 4 29562 > gpurun_out/bench_r02b_4gpu.json 2>> gpurun_out/r02_final8b.err
run 2 29563 --no-e2e > gpurun_out/bench_r02b_2gpu.json 2>> gpurun_out/r02_final8b.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02b_1gpu_8box.json 2>> gpurun_out/r02_final8b.err
tail -3 gpurun_out/r02_final8b.err
for f in gpurun_out/bench_r02b_*.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'], 1), round((d.get('e2e') or {}).get('value') or 0, 1), (d.get('parity') or {}).get('ok'), {k: round(v['ms_total'], 2) for k, v in d['kernels'].items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
