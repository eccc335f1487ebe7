#!/bin/bash
# 8-GPU validation of the round-2 multi-GPU path: parity worker at world 8 and 4, bench at N = 8 / 4 (+ NCCL A/B at 8)
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -x -q -m gpu -k "8 or 4" 2>&1 | tail -5
run() { # N port extra...
  local N=$1; local P=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@"
}
run 8 29521 > gpurun_out/r02_8gpu.json 2> gpurun_out/r02_8gpu.err
run 8 29522 --no-e2e --opt fin=0 > gpurun_out/r02_8gpu_nofin.json 2>> gpurun_out/r02_8gpu.err
run 8 29523 --no-e2e --no-p2p > gpurun_out/r02_8gpu_nccl.json 2>> gpurun_out/r02_8gpu.err
run 4 29524 > gpurun_out/r02_4gpu.json 2>> gpurun_out/r02_8gpu.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_8box_1gpu.json 2>> gpurun_out/r02_8gpu.err
tail -5 gpurun_out/r02_8gpu.err
for f in gpurun_out/r02_8gpu*.json gpurun_out/r02_4gpu.json gpurun_out/r02_8box_1gpu.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'], 1), round((d.get('e2e') or {}).get('value') or 0, 1), (d.get('parity') or {}).get('ok'), {k: (round(v['ms_total'], 2), v['launches']) for k, v in d['kernels'].items()}, json.dumps(d.get('sync_profile'))[:600])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
