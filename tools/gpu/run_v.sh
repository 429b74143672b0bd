# grouped general path: parity test, then config 3 with 1 / 3 / 4 / 6 / 8 groups
timeout 600 python -m pytest tests/test_gpu_batch.py -x -q -m gpu -k "general_path_stream_groups or routes or levels_11 or stress_corpus" 2>&1 | tail -3
for g in 1 3 4 6 8 1; do
echo "== groups $g"
ACM_B200_GEN_GROUPS=$g timeout 300 python tools/profile_run.py --streams 10000 --runs 4 --workload config3 2>&1 | tail -3
done
