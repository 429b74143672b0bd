timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu -k "fallout or config2_full or more_streams or truncations or fast_shape or unaligned or healthy or host_path or formats or decode_twice" 2>&1 | tail -2
for r in 0 1 0 1; do
echo "== sparse $r"
ACM_B200_SCAN_SPARSE=$r timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-config4 --no-streaming 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('value ms', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], j['e2e']['value'], 'floor', j['e2e']['copy_floor']['d2h_only_ms'], 'parity', j['parity_gate']['oracle_failures'])"
done
for n in 300 1000 2500; do
for r in 0 1; do
echo "== $n streams resident, sparse $r"
ACM_B200_SCAN_SPARSE=$r timeout 300 python tools/profile_run.py --streams $n --runs 3 2>&1 | tail -2 | head -1
done
done
ACM_B200_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-config4 --no-streaming 2>&1 >/dev/null | grep "acm trace" | tail -13
