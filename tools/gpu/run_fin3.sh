# end-of-round evidence (final code), one GPU
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_r02e_n1.json 2> gpurun_out/bench_r02e_n1.err
timeout 900 python bench.py --workload config3 --no-config4 > gpurun_out/bench_r02f_config3.json 2> gpurun_out/bench_r02f_config3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02e_reference.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02e_config3.csv python bench.py --workload config3 --steps 2 --warmup 3 --no-cpu --no-config4 --no-e2e --no-streaming > /dev/null 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/bench_r02e_n1.json', 'gpurun_out/bench_r02f_config3.json', 'gpurun_out/bench_r02e_reference.json'):
    t=open(f).read()
    j=json.loads([l for l in t.splitlines() if l.startswith('{')][-1])
    print(f, j.get('impl'), j['value'], j.get('ms_per_step'), (j.get('roofline') or {}).get('frac'), 'e2e', j['e2e']['value'], j['e2e'].get('ms_per_step'), (j['e2e'].get('copy_floor') or {}).get('d2h_only_ms'), 'config4', (j.get('config4') or {}).get('gsamples_s'), (j.get('config4') or {}).get('ms_per_step'), (j.get('config4') or {}).get('oracle_failures'), 'parity', (j.get('parity_gate') or {}).get('oracle_failures'), (j.get('parity_gate') or {}).get('timed_output_equals_checked_output'), 'launches', j.get('gpu_launches'), (j.get('streaming') or {}).get('gpu'))
PY
