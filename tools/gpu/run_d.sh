set -x
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu -k "twice or exactly" 2>&1 | tail -4
F2_PROF=1 ACM_B200_LIB=libacm_b200/_lib/var/prof/libacm_b200.so timeout 600 python tools/profile_run.py --streams 125000 --runs 2 --workload config4 2>&1 | tail -6
