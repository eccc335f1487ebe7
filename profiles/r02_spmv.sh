#!/bin/bash
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
timeout 1200 python profiles/spmv_ab.py 1.0 48 64 80 > gpurun_out/r02_spmv_ab3.jsonl 2> gpurun_out/r02_spmv_ab.err
cut -c1-220 gpurun_out/r02_spmv_ab3.jsonl; grep lkb gpurun_out/r02_spmv_ab.err | sort | uniq -c | head; tail -3 gpurun_out/r02_spmv_ab.err
