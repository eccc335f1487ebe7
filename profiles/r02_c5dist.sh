#!/bin/bash
# row-sharded C5: parity worker at world 2, reduced-size config on 2 GPUs, block-Arnoldi measurement on 1 GPU
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -x -q -m gpu -k "2" 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench_configs.py --only c5 > gpurun_out/r02_c5_2gpu_reduced.jsonl 2> gpurun_out/r02_c5dist.err
cut -c1-700 gpurun_out/r02_c5_2gpu_reduced.jsonl; tail -3 gpurun_out/r02_c5dist.err
python bench_configs.py --full --only c2b > gpurun_out/r02_c2_block.jsonl 2>> gpurun_out/r02_c5dist.err; cut -c1-600 gpurun_out/r02_c2_block.jsonl
