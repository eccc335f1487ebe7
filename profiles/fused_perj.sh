B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile-pass --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_axpy_dot -c 128 --csv --log-file gpurun_out/fused_launches.csv $B > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/fused_launches.csv')) if len(r)>10]
hdr=rows[0]; vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
n=4096*4096*8
for k,r in enumerate(rows[1:]):
    j=k+1
    t=float(r[vi].replace(',',''))*{'ns':1e-9,'us':1e-6,'ms':1e-3}[r[ui]]
    if j in (1,16,32,33,48,64,65,80,96,112,128): print(j, f"{t*1e3:.3f} ms", f"alg {(j+2)*n/t/1e9:.0f} GB/s")
PY
