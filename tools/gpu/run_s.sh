set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python tools/prof_aux.py --streams 20000 2>&1 | tail -2
timeout 300 python tools/profile_run.py --streams 10000 --runs 4 2>&1 | tail -2
timeout 600 python tools/profile_run.py --streams 125000 --runs 3 --workload config4 2>&1 | tail -1
