set -x
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python tools/profile_run.py --streams 10000 --runs 4 2>&1 | tail -6
timeout 600 python tools/profile_run.py --streams 125000 --runs 3 --workload config4 2>&1 | tail -5
