#!/bin/bash
# 2-GPU validation of the round-2 multi-GPU path + per-GPU-share scaling probe (each GPU owns 512 grid rows = the N = 8 share)
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -15
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
B="bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline"
$T $B --ny 1024 > gpurun_out/r02_2gpu_ny1024.json 2> gpurun_out/r02_2gpu.err
$T $B --ny 1024 --opt fused_halo=0 > gpurun_out/r02_2gpu_ny1024_halokernel.json 2>> gpurun_out/r02_2gpu.err
$T $B --ny 1024 --opt fin=0 > gpurun_out/r02_2gpu_ny1024_nofin.json 2>> gpurun_out/r02_2gpu.err
$T $B > gpurun_out/r02_2gpu_full.json 2>> gpurun_out/r02_2gpu.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --ny 512 > gpurun_out/r02_2gpu_ref1_ny512.json 2>> gpurun_out/r02_2gpu.err
tail -5 gpurun_out/r02_2gpu.err
for f in gpurun_out/r02_2gpu_*.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'], 1), round(d.get('e2e', {}).get('value') or 0, 1), (d.get('parity') or {}).get('ok'), {k: (round(v['ms_total'], 2), v['launches']) for k, v in d['kernels'].items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
