#!/bin/bash
mkdir -p gpurun_out
for S in 8 16; do
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_s$S.csv python tools/small_s_launches.py $S > gpurun_out/small_s_l.log 2>&1
python - <<P
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_s$S.csv')) if len(r)>10 and r[0].isdigit()]
n=len(rows)//2
rows=rows[n:]
tot=sum(float(r[-1]) for r in rows)
print('S=$S', len(rows),'launches in the last pass, sum us', round(tot/1e3,1))
for r in rows:
    print(r[4][:56].ljust(56), r[5], r[-1])
P
done
