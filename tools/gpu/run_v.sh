set -x
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python tools/profile_run.py --streams 10000 --runs 3 --kernel 2 2>&1 | tail -3
timeout 600 python tools/profile_run.py --streams 125000 --runs 3 --workload config4 --kernel 2 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"acm_walk|acm_unpack|acm_lift|acm_finish" python tools/profile_run.py --streams 10000 --runs 1 --kernel 2 2>&1 | grep -E "acm_|gpu__time" | head -12
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"acm_walk|acm_unpack|acm_lift|acm_finish" python tools/profile_run.py --streams 125000 --runs 1 --workload config4 --kernel 2 2>&1 | grep -E "acm_|gpu__time" | head -12
