#!/bin/bash
# Round-2 single-GPU evidence run: full GPU test suite, bench (own arm + reference arm), ncu launch list + --set full captures,
# the other BASELINE configs at full size.  Outputs in gpurun_out/, summarised into profiles/ by summarize_ncu.py r02.
cd "$(dirname "$0")/.." ; mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench_r02.err; tail -2 gpurun_out/bench_r02.err
python bench.py --impl reference > gpurun_out/bench_r02_ref.json 2>> gpurun_out/bench_r02.err
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile-pass --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r02.csv $B > /dev/null 2> gpurun_out/ncu_r02.err
for k in multidot axpy_dot multiaxpy_fin stencil_smem; do
  ncu --set full --clock-control none --import-source on -k regex:k_$k -s 100 -c 1 -o gpurun_out/prof_${k}_r02 -f $B > /dev/null 2>> gpurun_out/ncu_r02.err
done
for t in stencil3d cg gemm; do
  case $t in stencil3d) kk=k_stencil; s=4;; cg) kk="k_cg_update|k_cg_direction|k_dot2"; s=12;; gemm) kk=k_basis_gemm; s=0;; esac
  ncu --set full --clock-control none --import-source on -k "regex:$kk" -s $s -c 3 -o gpurun_out/prof_${t}_r02 -f python profiles/ncu_targets.py $t > /dev/null 2>> gpurun_out/ncu_r02.err
done
tail -3 gpurun_out/ncu_r02.err
timeout 1500 python bench_configs.py --full --only c2,c4,c5 > gpurun_out/r02_configs_full.jsonl 2> gpurun_out/r02_configs_full.err; tail -2 gpurun_out/r02_configs_full.err
python profiles/stencil_ab.py > gpurun_out/r02_stencil_ab2.txt 2>&1
python - <<'PY'
import json
for f in ("bench_r02_1gpu", "bench_r02_ref"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"], 2), d.get("e2e", {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("parity") or {}).get("ok"), d.get("cpu_baseline"))
    if "kernels" in d: print({k: (round(v["ms_total"], 2), round(v["GBps"])) for k, v in d["kernels"].items()})
for l in open("gpurun_out/r02_configs_full.jsonl"):
    print(l[:420])
PY
ls -la gpurun_out/*_r02.ncu-rep
