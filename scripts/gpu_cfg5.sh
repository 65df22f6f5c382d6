#!/bin/bash
python scripts/prof_wave.py 65536 3 5 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg5.csv python scripts/prof_wave.py 65536 2 5 > gpurun_out/prof_cfg5.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_cfg5.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
cols=rows[hdr]
ki=cols.index('Kernel Name'); vi=cols.index('Metric Value')
r2=[r for r in rows[hdr+1:] if len(r)>vi]
half=len(r2)//2
print(" ".join("%s:%.2f"%(r[ki][19:24], float(r[vi])/1e6) for r in r2[half:]))
PY
