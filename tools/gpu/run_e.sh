set -x
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/profile_run.py --streams 10000 --runs 4 2>&1 | tail -3
timeout 600 python tools/profile_run.py --streams 125000 --runs 3 --workload config4 2>&1 | tail -2
for v in prof carry2; do
echo "== variant $v"
F2_PROF=1 ACM_B200_LIB=libacm_b200/_lib/var/$v/libacm_b200.so timeout 600 python tools/profile_run.py --streams 125000 --runs 2 --workload config4 2>&1 | tail -6
done
