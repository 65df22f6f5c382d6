#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_wave.csv python scripts/prof_wave.py 65536 2 > gpurun_out/prof_wave.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_wave.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
cols=rows[hdr]
ki=cols.index('Kernel Name'); vi=cols.index('Metric Value')
r2=[r for r in rows[hdr+1:] if len(r)>vi]
half=len(r2)//2
print(" ".join("%s:%.2f"%(r[ki][19:29], float(r[vi])/1e6) for r in r2[half:] if float(r[vi])>20000))
PY
ncu --set full --clock-control none --import-source on -k regex:vlr_wave_round_warp_kernel -s 12 -c 1 -o gpurun_out/ncu_wave_roundw -f python scripts/prof_wave.py 65536 2 > gpurun_out/ncu_round.log 2>&1
