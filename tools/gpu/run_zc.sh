timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
echo "== config 3"
timeout 300 python tools/profile_run.py --streams 10000 --runs 4 --workload config3 2>&1 | tail -3
echo "== general path forced, 20000 config-4 streams"
timeout 600 python tools/prof_aux.py --streams 20000 --kernel 1 2>&1 | tail -1
