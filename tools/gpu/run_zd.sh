timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:acm_scan -c 8 --csv --log-file gpurun_out/scan_zd.csv python tools/profile_run.py --streams 10000 --runs 1 --workload config3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/scan_zd.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[h]
for r in rows[h+1:]:
    if len(r)>len(H)-1: print(r[H.index('Grid Size')], r[H.index('Metric Name')], r[H.index('Metric Value')])
PY
