#!/bin/bash
# 2^22 (and 2^21) witness-map pass schedules under ncu (kernel times in isolation): bash tools/ntt_big.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
run() { name=$1; logn=$2; shift 2; env "$@" timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:ntt_pass -c 12 --csv --log-file gpurun_out/${tag}_${name}.csv python tools/prof_ntt.py $logn 1 > gpurun_out/${tag}_${name}.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/${tag}_${name}.csv') if l.startswith('"'))]
t=[float(r['Metric Value'].replace(',',''))/1e6 for r in rows if r['Metric Name']=='gpu__time_duration.sum']
print('${name}', 'launches', len(t), 'ms', [round(x,3) for x in t], 'sum', round(sum(t),3))
PY
}
run d22 22 B2G_X=0
run d22_2pass 22 B2G_NTT_TL=11 B2G_NTT_MAXK=11
run d22_k4 22 B2G_NTT_MAXK=4
run d22_radix2 22 B2G_NTT_RADIX2=1
run d21 21 B2G_X=0
run d21_2pass 21 B2G_NTT_TL=11 B2G_NTT_MAXK=11
