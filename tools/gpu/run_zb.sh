timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
echo "== config 3"
timeout 300 python tools/profile_run.py --streams 10000 --runs 4 --workload config3 2>&1 | tail -3
echo "== split path (kernel 2), config 2"
timeout 300 python tools/profile_run.py --streams 10000 --runs 3 --kernel 2 2>&1 | tail -2
echo "== general path forced, 20000 config-4 streams"
timeout 600 python tools/prof_aux.py --streams 20000 --kernel 1 2>&1 | tail -1
timeout 300 python tools/stream_time.py 2>&1 | tail -3
