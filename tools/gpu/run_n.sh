set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/prof_aux.py --streams 20000 2>&1 | tail -2
timeout 300 python tools/stream_time.py 2>&1 | tail -3
