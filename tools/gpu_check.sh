#!/bin/bash
# Development loop on the GPU box: parity tests, device-resident timing, per-kernel time + instruction counts.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/microbench.py --inverse 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:k_ -s 12 -c 6 --csv --log-file gpurun_out/quick.csv python tools/microbench.py --inverse --iters 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/quick.csv')) if len(r)>8]
hdr=None
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; rows=rows[i+1:]; break
ki,mi,vi=hdr.index('Kernel Name'),hdr.index('Metric Name'),hdr.index('Metric Value')
idx=hdr.index('ID')
d={}
for r in rows: d.setdefault((r[idx],r[ki][:40]),{})[r[mi]]=r[vi]
for (i,k),m in d.items():
    print(k, ' | '.join(f"{a.split('.')[0][-24:]}={b}" for a,b in m.items()))
PY
