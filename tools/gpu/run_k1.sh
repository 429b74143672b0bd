timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | tail -2
timeout 1500 python bench.py > gpurun_out/bench_r02c_n1.json 2> gpurun_out/bench_r02c_n1.err
python - <<'PY'
import json
t=open('gpurun_out/bench_r02c_n1.json').read()
j=json.loads([l for l in t.splitlines() if l.startswith('{')][-1])
print(j['value'], j['ms_per_step'], j['roofline']['frac'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], j['e2e']['copy_floor']['d2h_only_ms'], 'config4', j['config4']['gsamples_s'], j['config4']['ms_per_step'], j['config4']['oracle_failures'], 'parity', j['parity_gate']['oracle_failures'], j['parity_gate']['timed_output_equals_checked_output'], j['streaming']['gpu'])
PY
