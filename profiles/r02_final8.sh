#!/bin/bash
# Round-2 multi-GPU evidence run on one 8-GPU box: parity worker at world 2 / 4 / 8, bench at N = 8 / 4 / 2 / 1 (+ A/B of the
# programmatic launch and the serpentine sweeps at N = 8), C2 gmres and C3 eigs row-sharded over 8 GPUs.
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
run() { # N port extra...
  local N=$1; local P=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@"
}
run 8 29531 > gpurun_out/bench_r02_8gpu.json 2> gpurun_out/r02_final8.err
LKB_PDL_MASK=0 run 8 29532 --no-e2e --no-profile-pass > gpurun_out/r02_8gpu_nopdl.json 2>> gpurun_out/r02_final8.err
run 8 29533 --no-e2e --no-profile-pass --opt serpentine=0 > gpurun_out/r02_8gpu_noserp.json 2>> gpurun_out/r02_final8.err
run 4 29534 > gpurun_out/bench_r02_4gpu.json 2>> gpurun_out/r02_final8.err
run 2 29535 > gpurun_out/bench_r02_2gpu.json 2>> gpurun_out/r02_final8.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_1gpu_8box.json 2>> gpurun_out/r02_final8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29536 bench_configs.py --full --only c2,c3 > gpurun_out/r02_configs_8gpu.jsonl 2>> gpurun_out/r02_final8.err
tail -5 gpurun_out/r02_final8.err
for f in gpurun_out/bench_r02_8gpu.json gpurun_out/r02_8gpu_nopdl.json gpurun_out/r02_8gpu_noserp.json gpurun_out/bench_r02_4gpu.json gpurun_out/bench_r02_2gpu.json gpurun_out/bench_r02_1gpu_8box.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'], 1), round((d.get('e2e') or {}).get('value') or 0, 1), (d.get('parity') or {}).get('ok'), {k: (round(v['ms_total'], 2), v['launches']) for k, v in d['kernels'].items()}, json.dumps(d.get('sync_profile'))[-260:])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
cut -c1-600 gpurun_out/r02_configs_8gpu.jsonl
