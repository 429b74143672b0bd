for r in 0 1 0 1; do
echo "== seg ramp $r"
ACM_B200_SEG_RAMP=$r timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-config4 --no-streaming 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('value ms', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], j['e2e']['value'], 'floor', j['e2e']['copy_floor']['d2h_only_ms'], 'parity', j['parity_gate']['oracle_failures'])"
done
timeout 600 python -m pytest tests/test_gpu_batch.py -x -q -m gpu -k "host_path or stress_corpus_host or unaligned" 2>&1 | tail -2
