# end-of-round evidence, one GPU: full GPU suite, headline bench, DRAM traffic of the fused kernel, config 3 bench, launch lists
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/bench_r02d_n1.json 2> gpurun_out/bench_r02d_n1.err
tail -c 300 gpurun_out/bench_r02d_n1.err
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
timeout 600 ncu $M --clock-control none -k regex:acm_decode_fast2 -s 2 -c 1 python tools/profile_run.py --streams 10000 --runs 3 2>&1 | grep -E "dram__|inst_exec|duration|hit_rate"
timeout 900 python bench.py --workload config3 --no-config4 > gpurun_out/bench_r02e_config3.json 2> gpurun_out/bench_r02e_config3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02d_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --config4-streams 200000 > /dev/null 2>&1
wc -l gpurun_out/launches_r02d_bench.csv
python - <<'PY'
import json
for f in ('gpurun_out/bench_r02d_n1.json', 'gpurun_out/bench_r02e_config3.json'):
    t=open(f).read()
    j=json.loads([l for l in t.splitlines() if l.startswith('{')][-1])
    print(f, j['value'], j['ms_per_step'], j['roofline']['frac'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], j['e2e']['copy_floor']['d2h_only_ms'], 'config4', (j.get('config4') or {}).get('gsamples_s'), (j.get('config4') or {}).get('ms_per_step'), (j.get('config4') or {}).get('oracle_failures'), 'parity', j['parity_gate']['oracle_failures'], j['parity_gate']['timed_output_equals_checked_output'], j['streaming']['gpu'])
PY
