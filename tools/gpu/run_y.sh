set -x
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu -k "2 or kernel" 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"acm_walk|acm_unpack|acm_lift" python tools/profile_run.py --streams 10000 --runs 1 --kernel 2 2>&1 | grep -E "gpu__time|inst_exec" | head -12
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"acm_walk|acm_unpack|acm_lift" python tools/profile_run.py --streams 125000 --runs 1 --workload config4 --kernel 2 2>&1 | grep -E "gpu__time|inst_exec" | head -12
